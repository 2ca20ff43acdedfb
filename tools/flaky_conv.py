"""Repeats one conv forward many times and reports the spread of its error vs an fp64 reference (ordering experiments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from fusiondepth_b200 import ops
CL = torch.channels_last
def run(B, Cin, H, W, Cout, k, p, n=200):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, 1, p)
    xc = x.cuda().contiguous(memory_format=CL); wc = w.cuda().contiguous(memory_format=CL)
    errs = []; outs = set()
    with torch.no_grad():
        for i in range(n):
            y = ops.conv2d(xc, wc, None, 1, p, "none")
            e = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
            errs.append(e); outs.add(hash(y.cpu().numpy().tobytes()))
    errs.sort()
    print("conv %s: min err %.2e median %.2e max %.2e, distinct outputs %d of %d" % ((B, Cin, H, W, Cout, k, p), errs[0], errs[n // 2], errs[-1], len(outs), n))
for shape in [(2, 48, 10, 14, 40, 3, 1), (2, 64, 16, 24, 64, 3, 1), (3, 128, 9, 13, 256, 3, 1), (6, 64, 48, 160, 64, 3, 1), (2, 512, 6, 20, 512, 3, 1), (2, 160, 12, 40, 128, 3, 1)]:
    run(*shape, n=int(os.environ.get("N", "200")))
