"""Kernel timeline of one captured optimiser step (not a test): CUPTI trace through torch.profiler of a
CUDA-graph replay -> gpurun_out/timeline.csv (name, stream, start_us, dur_us).  Analyse with
profiles/timeline_report.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from fusiondepth_b200 import _lib, synth, training

_lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
models = training.build_models(bench.NUM_LAYERS, dev)
step = training.TrainStep(models, lr=1e-4, accumulate=bench.ACCUM)
cpu_batches, cpu_noises = bench.synthetic_step_inputs(100, dev)
batches = [synth.to_device(b, dev) for b in cpu_batches]
noises = [{s: t.to(dev) for s, t in n.items()} for n in cpu_noises]
step.capture(batches, noises, warmup=1)
for _ in range(3):
    step.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step.replay()
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "timeline.csv")
os.makedirs(os.path.dirname(out), exist_ok=True)
n = 0
with open(out, "w") as f:
    f.write("name,stream,start_us,dur_us\n")
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace(",", ";")
            stream = getattr(ev, "stream", None)
            if stream is None:
                stream = ev.device_resource_id if hasattr(ev, "device_resource_id") else -1
            f.write("%s,%s,%.3f,%.3f\n" % (name[:100], stream, ev.time_range.start, ev.time_range.end - ev.time_range.start))
            n += 1
print("wrote", n, "device events to", out)
