"""Timing experiment (not a test): per-role clock64() trace of CTA (0,0) of one tensor-core conv launch.
usage: FD_TC2_FLAGS=<flags|128> python tools/trace_conv.py  ->  per-k-block intervals of loader / splitter / MMA"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from fusiondepth_b200 import ops, _lib

lib = _lib.load()
CL = torch.channels_last
KB = 256
for B, Cin, H, W, Cout in [(6, 128, 24, 80, 128), (6, 64, 48, 160, 64)]:
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
    ph = t[3 * KB * 4:3 * KB * 4 + 5].tolist()
    nk = Cin * 9 // 32
    r = t[:3 * KB * 4].view(3, KB, 4)[:, :nk]
    t0 = ph[0]
    print("shape", (B, Cin, H, W, Cout), "flags", os.environ.get("FD_TC2_FLAGS"), "nk", nk)
    print("  kernel phases (clk since CTA start): prologue done %d, epilogue start %d, epilogue done %d, exit barrier %d"
          % (ph[1] - t0, ph[2] - t0, ph[3] - t0, ph[4] - t0))
    names = ["loader  [start, pre-wait, post-wait(sfree), issued]", "splitter[start, post-wait(landed), pre-store(post tfree), arrived]",
             "mma     [start, post-wait(tfull), post-fence, issued]"]
    for role in range(3):
        print(" ", names[role])
        for kb in list(range(0, min(nk, 10))) + list(range(max(10, nk - 3), nk)):
            v = (r[role, kb] - t0).tolist()
            print("    kb %3d: %s" % (kb, " ".join("%7d" % z for z in v)))
