"""Runs a few convolutions through the tensor-core path (for `ncu -k regex:conv_tc`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusiondepth_b200 import ops
CL = torch.channels_last
shapes = [(6, 128, 24, 80, 128, 3, 1, 1), (6, 64, 48, 160, 64, 3, 1, 1), (6, 512, 6, 20, 512, 3, 1, 1)]
for B, Cin, H, W, Cout, k, s, p in shapes:
    x = torch.randn(B, Cin, H, W, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    w = torch.randn(Cout, Cin, k, k, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    for _ in range(3):
        y = ops.conv2d(x, w, None, s, p, "none")
        gx, gw = torch.autograd.grad(y, (x, w), torch.ones_like(y))
torch.cuda.synchronize()
