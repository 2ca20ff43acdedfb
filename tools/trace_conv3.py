"""Timing experiment (not a test): per-role clock64() trace of CTA (0,0) of one conv_tc3 launch.
usage: python tools/trace_conv3.py  ->  kernel phases and per-k-block intervals of the A splitters and the MMA warp"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from fusiondepth_b200 import ops, _lib

lib = _lib.load()
CL = torch.channels_last
KB = 256
for B, Cin, H, W, Cout in [(6, 64, 48, 160, 64), (6, 128, 24, 80, 128), (6, 512, 6, 20, 512)]:
    x = torch.randn(B, Cin, H, W, device="cuda").contiguous(memory_format=CL)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda").contiguous(memory_format=CL)
    trace = torch.zeros(3 * KB * 4 + 8, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            ops.conv2d(x, w, None, 1, 1, "none")
        torch.cuda.synchronize()
        lib.fd_debug_set_conv_trace(ctypes.c_void_p(trace.data_ptr()))
        ops.conv2d(x, w, None, 1, 1, "none")
        torch.cuda.synchronize()
        lib.fd_debug_set_conv_trace(None)
    t = trace.cpu()
    ph = t[3 * KB * 4:3 * KB * 4 + 8].tolist()
    nk = Cin * 9 // 32
    r = t[:3 * KB * 4].view(3, KB, 4)[:, :nk]
    t0 = ph[0]
    print("shape", (B, Cin, H, W, Cout), "nk", nk)
    print("  kernel phases (clk since CTA start): prologue done %d, epilogue start %d, epilogue done %d, exit barrier %d"
          % (ph[1] - t0, ph[2] - t0, ph[3] - t0, ph[4] - t0))
    print("  inside the prologue: barriers initialised %d, tensor memory allocated %d, row table written %d"
          % (ph[5] - t0, ph[6] - t0, ph[7] - t0))
    names = ["A split g0 [start, loaded+split, post tfree wait, arrived]", "mma        [start, post wready, post tfull, issued]",
             "W path     [TMA warp at kb, stage free -> TMA issued, landed (splitter woke), W_lo ready]"]
    show = list(range(0, min(nk, 12))) + list(range(max(12, nk - 4), nk))
    for role in range(3):
        print(" ", names[role])
        for kb in show:
            v = (r[role, kb] - t0).tolist()
            if any(z > 0 for z in v):
                print("    kb %3d: %s" % (kb, " ".join("%7d" % z for z in v)))
    mma = r[1]
    if nk > 8:
        steady = (mma[nk - 1, 3] - mma[4, 3]).item() / (nk - 5)
        print("  steady-state clk per k-block (MMA issue to issue): %.0f" % steady)
