"""Diagnostic (not a test): per-parameter difference between the captured-graph step's gradients and the eager
step's, on the small problem of tests/test_gpu_step.py::test_cuda_graph_replay_matches_eager.
usage: python tools/diag_graph_eager.py [n_repeats] [serial] [nodirect] [nocache] [full] [mb2] [mid]
(full: the bench configuration, 2 micro-batches of 6 x 192 x 640)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusiondepth_b200 import training, synth
from tests._util import rel_err
from tests.test_gpu_step import _load

torch.manual_seed(0)
models = training.build_models(18, "cuda")
_load(models, 5)
flags = sys.argv[2:]
NB, (B, H, W) = (2, (6, 192, 640)) if "full" in flags else (1, (2, 64, 96))
if "mb2" in flags:          # two concurrent micro-batches of the small problem
    NB = 2
if "mid" in flags:
    B, H, W = 3, 96, 160
batches = [synth.to_device(synth.make_batch(B, H, W, seed=20 + i, with_noise=False), "cuda") for i in range(NB)]
noises = [{s: torch.randn(B, 2, H, W, device="cuda") for s in range(4)} for i in range(NB)]
step = training.TrainStep(models, lr=1e-4, accumulate=NB, parallel_trunks="serial" not in flags,
                          direct_grad="nodirect" not in flags, cache_weight_prep="nocache" not in flags)
print("flags", flags)
step.capture(batches, noises)
w0 = step.flat.data.clone()
bn0 = step._bn_state()
names = {}
for mname, m in models.items():
    for n, p in m.named_parameters():
        names[p] = mname + "." + n


def reset():
    step.flat.data.copy_(w0)
    step._bn_state(bn0)
    step.exp_avg.zero_(); step.exp_avg_sq.zero_(); step.adam_state[:3].zero_()


runs = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    reset()
    l = float(step.replay())
    runs.append(("graph%d" % i, l, step.flat.grad.clone()))
    reset()
    l = float(step.step(batches, noises))
    runs.append(("eager%d" % i, l, step.flat.grad.clone()))
ref = runs[0]
def per_param(a, b, top=12):
    worst = []
    for p in step.flat.params:
        off, k = step.flat.offsets[p], p.numel()
        if float(b[off:off + k].abs().max()) > 0:
            worst.append((rel_err(a[off:off + k].cpu(), b[off:off + k].cpu()), names.get(p, "?"), k))
    worst.sort(reverse=True)
    for e, n, k in worst[:top]:
        print("      %.3e  %-60s %d" % (e, n, k))
    print("      parameters above 1e-3: %d of %d" % (sum(1 for e, _, _ in worst if e > 1e-3), len(worst)))
# reference = eager0 (runs[1])
for name, l, g in runs:
    e = rel_err(g.cpu(), runs[1][2].cpu())
    print("%-8s loss %.8f  rel_err vs eager0 %.3e" % (name, l, e))
    if e > 1e-3:
        per_param(g, runs[1][2])
worst = []
a, b = runs[0][2], runs[1][2]
for p in step.flat.params:
    off, k = step.flat.offsets[p], p.numel()
    if float(b[off:off + k].abs().max()) > 0:
        worst.append((rel_err(a[off:off + k].cpu(), b[off:off + k].cpu()), names.get(p, "?"), k))
worst.sort(reverse=True)
print("graph0 vs eager0, worst parameters:")
for e, n, k in worst[:25]:
    print("  %.3e  %-60s %d" % (e, n, k))
print("parameters above 1e-3: %d of %d" % (sum(1 for e, _, _ in worst if e > 1e-3), len(worst)))
