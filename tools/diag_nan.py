"""Diagnostic (not a test): per-step losses of the bench workload for a given input seed, to see when / why the
si-loss mask empties (NaN, like the reference: SURVEY.md Appendix E.5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fusiondepth_b200 import synth, training, ops

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 101
dev = torch.device("cuda", 0)
torch.manual_seed(0)
models = training.build_models(18, dev)
step = training.TrainStep(models, lr=1.5e-4, accumulate=2)
cb, cn = bench.synthetic_step_inputs(seed, dev)
b = [synth.to_device(x, dev) for x in cb]
n = [{s: t.to(dev) for s, t in x.items()} for x in cn]
for it in range(40):
    outs = []
    for mb in range(2):
        with torch.no_grad():
            o, l = training.process_batch(models, b[mb], n[mb])
        outs.append({k: float(v) for k, v in l.items()})
    print(it, " ".join("%s=%.4f" % (k.replace("loss/", ""), outs[0][k]) for k in outs[0]), "| mb1 loss=%.4f si0=%.4f" % (outs[1]["loss"], outs[1]["loss/si_loss0"]))
    loss = step.step(b, n)
    if not torch.isfinite(loss):
        print("step", it, "non-finite")
        break
