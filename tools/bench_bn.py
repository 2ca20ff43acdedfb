"""Micro-benchmark (not a test): BatchNorm forward / backward kernel time per trunk layer shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusiondepth_b200 import ops

CL = torch.channels_last


def timeit(fn, n=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (3 * n)


for B, C, H, W in [(6, 64, 96, 320), (6, 64, 48, 160), (6, 128, 24, 80), (6, 256, 12, 40), (6, 512, 6, 20)]:
    x = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    r = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    g_, b_ = torch.ones(C, device="cuda", requires_grad=True), torch.zeros(C, device="cuda", requires_grad=True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    gy = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=CL)
    mb = x.numel() * 4 / 1e6
    with torch.no_grad():
        tf = timeit(lambda: ops.batch_norm(x, g_, b_, rm, rv, r, True, 0.1, 1e-5, True))

    def fb():
        y = ops.batch_norm(x, g_, b_, rm, rv, r, True, 0.1, 1e-5, True)
        torch.autograd.grad(y, (x, r, g_, b_), gy)
    tb = timeit(fb) - tf
    print("BN %3d ch %3dx%3d (%.1f MB/tensor): fwd %.1f us (%.0f GB/s of 4 passes), bwd %.1f us (%.0f GB/s of 7 passes)"
          % (C, H, W, mb, tf, 4 * mb / tf * 1e3 / 1e3, tb, 7 * mb / tb * 1e3 / 1e3))
