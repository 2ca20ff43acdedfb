"""Diagnostic (not a test): per-tensor gradient comparison CUDA vs oracle for one micro-batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._util import clone_sd, synth_weights, rel_err
from fusiondepth_b200 import synth, training
from oracle import step_oracle as SO

torch.manual_seed(0)
models = training.build_models(18, "cuda")
sds = {}
for i, (name, m) in enumerate(sorted(models.items())):
    sds[name] = synth_weights(m.state_dict(), i)
    m.load_state_dict(sds[name]); m.train()
osd = {k: clone_sd(v, requires_grad=True) for k, v in sds.items()}
mode = sys.argv[1] if len(sys.argv) > 1 else "uniform"
inputs = synth.make_batch(3, 96, 160, seed=1, mode=mode, lidar_density=0.25)
noise = inputs.pop("noise")
oo, ol = SO.process_batch(osd, inputs, noise, 18, True)
for f in (-1, 1):
    oo[("cam_T_cam", 0, f)].retain_grad(); oo[("axisangle", 0, f)].retain_grad()
for s in range(4):
    oo[("disp", s)].retain_grad()
ol["loss"].backward()
ci = synth.to_device(inputs, "cuda"); cn = {s: t.cuda() for s, t in noise.items()}
co, cl = training.process_batch(models, ci, cn, None, True)
for f in (-1, 1):
    co[("cam_T_cam", 0, f)].retain_grad(); co[("axisangle", 0, f)].retain_grad()
for s in range(4):
    co[("disp", s)].retain_grad()
cl["loss"].backward()
print("loss", float(cl["loss"]), float(ol["loss"]))
for f in (-1, 1):
    print("gT", f, rel_err(co[("cam_T_cam", 0, f)].grad.cpu(), oo[("cam_T_cam", 0, f)].grad))
    print(co[("cam_T_cam", 0, f)].grad.cpu()[0], "\n", oo[("cam_T_cam", 0, f)].grad[0])
    print("gaa", f, rel_err(co[("axisangle", 0, f)].grad.cpu(), oo[("axisangle", 0, f)].grad))
for s in range(4):
    print("gdisp", s, rel_err(co[("disp", s)].grad.cpu(), oo[("disp", s)].grad))
for name in sorted(models):
    worst = (0, None)
    for k, p in models[name].named_parameters():
        og = osd[name][k].grad
        if og is None: continue
        e = rel_err(p.grad.cpu(), og)
        if e > worst[0]: worst = (e, k)
    print(name, "worst grad rel err", worst)
