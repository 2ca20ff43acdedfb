"""Diagnostic (not a test): a few optimiser steps at the bench configuration, printing every loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fusiondepth_b200 import synth, training, ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
models = training.build_models(18, dev)
par = "--serial" not in sys.argv
step = training.TrainStep(models, lr=1e-4 * 12 / 8, accumulate=2, parallel_trunks=par)
cb, cn = bench.synthetic_step_inputs(100, dev)
batches = [synth.to_device(b, dev) for b in cb]
noises = [{s: t.to(dev) for s, t in n.items()} for n in cn]
fb = batches[0]["4beam"]
print("4beam nonzero frac", float((fb > 0).float().mean()), "range", float(fb[fb > 0].min()) * 100, float(fb.max()) * 100)
for it in range(12):
    step.flat.zero_grad()
    tot = 0.0
    rows = []
    for inputs, noise in zip(batches, noises):
        outputs, losses = training.process_batch(models, inputs, noise, None, streams=step.streams)
        (losses["loss"] / 2).backward()
        rows.append({k: round(float(v), 5) for k, v in losses.items()})
        d0 = outputs[("disp", 0)]
    gn = float(step.flat.grad.norm())
    ops.adam_step(step.flat.data, step.flat.grad, step.exp_avg, step.exp_avg_sq, step.adam_state, step.lr)
    print(it, "gradnorm %.4g" % gn, "disp0 mean %.4f min %.4f max %.4f" % (float(d0.mean()), float(d0.min()), float(d0.max())), rows[0])
