"""Micro-benchmark (not a test): per-layer conv time, tensor-core vs CUDA-core path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusiondepth_b200 import ops

CL = torch.channels_last
SHAPES = [  # B, Cin, H, W, Cout, k, s, p, name
    (6, 64, 48, 160, 64, 3, 1, 1, "layer1 3x3"),
    (6, 64, 48, 160, 128, 3, 2, 1, "layer2.0 3x3/2"),
    (6, 128, 24, 80, 128, 3, 1, 1, "layer2 3x3"),
    (6, 128, 24, 80, 256, 3, 2, 1, "layer3.0 3x3/2"),
    (6, 256, 12, 40, 256, 3, 1, 1, "layer3 3x3"),
    (6, 256, 12, 40, 512, 3, 2, 1, "layer4.0 3x3/2"),
    (6, 512, 6, 20, 512, 3, 1, 1, "layer4 3x3"),
    (6, 512, 8, 22, 256, 3, 1, 0, "dec upconv(4,0)"),
    (6, 512, 14, 42, 256, 3, 1, 0, "dec upconv(4,1)"),
    (6, 256, 26, 82, 128, 3, 1, 0, "dec upconv(3,1)"),
    (6, 128, 50, 162, 64, 3, 1, 0, "dec upconv(2,1)"),
    (6, 96, 98, 322, 32, 3, 1, 0, "dec upconv(1,1)"),
    (6, 32, 98, 322, 16, 3, 1, 0, "dec upconv(0,0)"),
    (6, 16, 194, 642, 16, 3, 1, 0, "dec upconv(0,1)"),
    (6, 3, 192, 640, 64, 7, 2, 3, "stem 7x7/2"),
]


def timeit(fn, n=10):
    """kernel-only time: the calls are captured in a CUDA graph so Python/launch overhead vanishes"""
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
    return e0.elapsed_time(e1) / (3 * n)


for B, Cin, H, W, Cout, k, s, p, name in SHAPES:
    x = torch.randn(B, Cin, H, W, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    w = torch.randn(Cout, Cin, k, k, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    gy = torch.randn(B, Cout, Ho, Wo, device="cuda").contiguous(memory_format=CL)
    flop = 2.0 * B * Ho * Wo * Cout * Cin * k * k
    row = "%-18s M=%6d N=%3d K=%4d " % (name, B * Ho * Wo, Cout, Cin * k * k)
    for backend in (("tc",) if os.environ.get("FD_BENCH_TC_ONLY") else ("tc", "cudacore")):
        ops.CONV_BACKEND = backend
        with torch.no_grad():
            tf = timeit(lambda: ops.conv2d(x, w, None, s, p, "none"))
        def fb():
            y = ops.conv2d(x, w, None, s, p, "none")
            torch.autograd.grad(y, (x, w), gy)
        tb = timeit(fb) - tf
        row += "| %s fwd %.3f ms (%.1f TF) bwd %.3f ms " % (backend, tf, flop / tf / 1e9, tb)
    print(row)
