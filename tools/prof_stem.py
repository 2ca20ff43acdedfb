"""Runs the stem path (im2col + GEMM + BN + maxpool, forward and backward) a few times for `ncu -k regex:...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusiondepth_b200 import networks
enc = networks.ResnetEncoder(18, False).cuda().train()
x = torch.rand(6, 3, 192, 640, device="cuda")
for _ in range(3):
    f = enc(x)
    (f[0].sum() + f[1].sum()).backward()
torch.cuda.synchronize()
