"""Timing experiment (not a test): where does the launch-to-launch time of the persistent layer-1 conv go?
Six back-to-back conv_tc4 launches inside one CUDA graph, each stamping CTA 0's start / exit with clock64() and
%globaltimer into its own trace buffer: CTA lifetime in cycles and ns (-> SM clock), gap between one launch's
exit and the next one's start, against the event-timed average per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from fusiondepth_b200 import _lib

lib = _lib.load()
KB = 256
B, C, H, W = 6, 64, 48, 160
x = torch.randn(B, H, W, C, device="cuda")
w = torch.randn(C, 3, 3, C, device="cuda")
wlo = torch.empty_like(w)
y = torch.empty(B, H, W, C, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
N = 8
traces = [torch.zeros(3 * KB * 4 + 8, dtype=torch.int64, device="cuda") for _ in range(N)]
s = torch.cuda.Stream()


def launch():
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.fd_conv2d_fwd_tc(P(x), P(w), P(wlo), None, P(y), B, H, W, C, C, 3, 3, 1, 1, 0, st), "conv")


with torch.cuda.stream(s):
    lib.fd_tf32_split(P(w), P(wlo), w.numel(), ctypes.c_void_p(s.cuda_stream))
    for _ in range(3):
        launch()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(N):
        lib.fd_debug_set_conv_trace(P(traces[i]))
        launch()
    lib.fd_debug_set_conv_trace(None)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
print("event-timed: %.2f us per launch (traced build of the launch: stamps only)" % (1e3 * e0.elapsed_time(e1) / N))
rows = []
for t in traces:
    ph = t[3 * KB * 4:3 * KB * 4 + 8].tolist()
    rows.append((ph[0], ph[4], ph[5], ph[6]))
for i, (c0, c1, g0, g1) in enumerate(rows):
    life_clk, life_ns = c1 - c0, g1 - g0
    gap = (g0 - rows[i - 1][3]) if i else 0
    print("launch %d: CTA 0 lifetime %6d clk = %6d ns (%.2f GHz); gap since previous exit %5d ns; start-to-start %6d ns"
          % (i, life_clk, life_ns, life_clk / max(life_ns, 1), gap, (g0 - rows[i - 1][2]) if i else 0))
