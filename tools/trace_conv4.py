"""Timing experiment (not a test): clock64() trace of CTA 0 of one persistent conv_tc4 launch.
usage: python tools/trace_conv4.py  ->  per-k-block stamps (k-blocks counted over all tiles of the CTA) of A splitter
group 0 and the two MMA issuers, per-tile stamps of the epilogue warps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from fusiondepth_b200 import ops, _lib

lib = _lib.load()
CL = torch.channels_last
KB = 256
for B, Cin, H, W, Cout, pad in [(6, 64, 48, 160, 64, 1), (6, 128, 50, 162, 64, 0)]:
    x = torch.randn(B, Cin, H, W, device="cuda").contiguous(memory_format=CL)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda").contiguous(memory_format=CL)
    trace = torch.zeros(3 * KB * 4 + 8, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            ops.conv2d(x, w, None, 1, pad, "none")
        torch.cuda.synchronize()
        lib.fd_debug_set_conv_trace(ctypes.c_void_p(trace.data_ptr()))
        ops.conv2d(x, w, None, 1, pad, "none")
        torch.cuda.synchronize()
        lib.fd_debug_set_conv_trace(None)
    t = trace.cpu()
    ph = t[3 * KB * 4:3 * KB * 4 + 5].tolist()
    nk = Cin * 9 // 32
    r = t[:3 * KB * 4].view(3, KB, 4)
    t0 = ph[0]
    print("shape", (B, Cin, H, W, Cout), "nk per tile", nk, "CTA exit at", ph[4] - t0)
    names = ["A split g0 [start, loaded+split, post tfree wait, arrived]", "mma        [start, post wready, post tfull+turn, issued]",
             "epilogue   per tile [start, accumulators complete, set released, tile stored]"]
    for role in range(3):
        print(" ", names[role])
        for kb in range(KB):
            v = r[role, kb].tolist()
            if any(z > 0 for z in v):
                mark = "  <- tile boundary" if role < 2 and kb % nk == 0 else ""
                print("    %s %3d: %s%s" % ("tile" if role == 2 else "g", kb, " ".join("%7d" % (z - t0 if z else 0) for z in v), mark))
