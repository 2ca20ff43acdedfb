import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from fusiondepth_b200 import ops
CL = torch.channels_last
B, Cin, H, W, Cout, k, s, p = 2, 64, 16, 24, 64, 3, 1, 1
g = torch.Generator().manual_seed(0)
x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g) * 0.05
gy = torch.randn(B, Cout, H, W, generator=g)
wr = w.clone().requires_grad_(True)
F.conv2d(x, wr, None, s, p).backward(gy)
xc = x.cuda().contiguous(memory_format=CL).requires_grad_(True)
wc = w.cuda().contiguous(memory_format=CL).requires_grad_(True)
y = ops.conv2d(xc, wc, None, s, p, "none")
torch.cuda.synchronize(); t0 = time.time()
y.backward(gy.cuda())
torch.cuda.synchronize(); print("bwd time", time.time() - t0)
got, ref = wc.grad.cpu(), wr.grad
print("ref[0,:4,0,0]", ref[0, :4, 0, 0], "\ngot", got[0, :4, 0, 0])
print("ref[5,30:34,1,2]", ref[5, 30:34, 1, 2], "\ngot", got[5, 30:34, 1, 2])
print("nonzero frac got", float((got != 0).float().mean()), "ratio mean", float((got / ref).median()))
print("got[0,:8,0,0]", got[0,:8,0,0], "got[:8,0,0,0]", got[:8,0,0,0]); print("got[0,0]", got[0,0]); print("ref[0,0]", ref[0,0])
print("max abs got", float(got.abs().max()), "ref", float(ref.abs().max()))
