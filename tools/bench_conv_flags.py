"""Timing experiment (not a test): forward conv time of three layer shapes under FD_TC2_FLAGS variants that
switch off one pipeline role at a time (results are numerically meaningless for flags & ~1)."""
import os, sys, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    from fusiondepth_b200 import ops

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
        return e0.elapsed_time(e1) / (3 * n)
    CL = torch.channels_last
    out = []
    for B, Cin, H, W, Cout in [(6, 64, 48, 160, 64), (6, 128, 24, 80, 128), (6, 512, 6, 20, 512), (6, 256, 26, 82, 128)]:
        x = torch.randn(B, Cin, H, W, device="cuda").contiguous(memory_format=CL)
        w = torch.randn(Cout, Cin, 3, 3, device="cuda").contiguous(memory_format=CL)
        with torch.no_grad():
            out.append("%.1f" % (1e3 * timeit(lambda: ops.conv2d(x, w, None, 1, 1, "none"))))
    print("flags=%s us: %s" % (os.environ.get("FD_TC2_FLAGS"), " ".join(out)))
else:
    for f in [int(x) for x in os.environ.get('FD_FLAG_LIST', '1,33,31,63,95,127').split(',')]:
        env = dict(os.environ, FD_TC2_FLAGS=str(f), FD_BENCH_TC_ONLY="1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=120)
        print((r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1])
