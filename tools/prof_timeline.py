"""Kernel timeline of one captured optimiser step (not a test): CUPTI trace through torch.profiler of a
CUDA-graph replay -> gpurun_out/timeline.csv (name, stream, start_us, dur_us).  Analyse with
profiles/timeline_report.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from fusiondepth_b200 import _lib, synth, training

if len(sys.argv) > 2:
    bench.select_workload(sys.argv[2])          # r18 (default) / r50
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
import json, tempfile
tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
n = 0
with open(out, "w") as f:
    f.write("name,stream,start_us,dur_us,ctas,threads,smem,regs\n")
    for ev in json.load(open(tmp))["traceEvents"]:
        if ev.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in ev:
            a = ev.get("args", {})
            g, b = a.get("grid", [0, 0, 0]), a.get("block", [0, 0, 0])
            f.write("%s,%s,%.3f,%.3f,%d,%d,%d,%d\n" % (
                ev["name"].replace(",", ";")[:100], a.get("stream", -1), ev["ts"], ev["dur"],
                g[0] * g[1] * g[2], b[0] * b[1] * b[2], a.get("shared memory", 0), a.get("registers per thread", 0)))
            n += 1
print("wrote", n, "device events to", out)
