"""Runs a few convolutions through the tensor-core path (for `ncu -k regex:conv_`): iterations are the
outer loop, so the last 9 matching launches are fwd / dgrad / wgrad of the three shapes, warm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusiondepth_b200 import ops
CL = torch.channels_last
shapes = [(6, 128, 24, 80, 128, 3, 1, 1), (6, 64, 48, 160, 64, 3, 1, 1), (6, 512, 6, 20, 512, 3, 1, 1)]
ts = []
for B, Cin, H, W, Cout, k, s, p in shapes:
    x = torch.randn(B, Cin, H, W, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    w = torch.randn(Cout, Cin, k, k, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    ts.append((x, w, s, p))
for _ in range(3):
    for x, w, s, p in ts:
        y = ops.conv2d(x, w, None, s, p, "none")
        gx, gw = torch.autograd.grad(y, (x, w), torch.ones_like(y))
torch.cuda.synchronize()
