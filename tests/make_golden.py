"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python tests/make_golden.py
The fixtures pin the oracle (oracle/) -- the reference ships no golden vectors of its own
 and cannot travel to the GPU box.  Inputs are regenerated from seeds
by fusiondepth_b200.synth / tests._util.synth_weights; only small inputs that cannot be
regenerated bit-exactly are stored.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _refharness as RH            # noqa: E402
from tests._util import GOLDEN, coarse_disparity, synth_weights   # noqa: E402
from fusiondepth_b200 import synth              # noqa: E402

torch.set_num_threads(8)


def lidar_cases():
    return {
        "ring4": synth.make_scan(3),
        "piled": synth.make_scan(5, piled=600),
        "dense": synth.make_dense_scan(7, n=30000),
        "edge": synth.make_dense_scan(9, n=20000, edge_heavy=True),
    }


def gen_lidar(ns):
    out = {}
    with tempfile.TemporaryDirectory() as d:
        synth.write_calib_files(d)
        for name, pts in lidar_cases().items():
            fn = os.path.join(d, name + ".bin")
            p = pts.copy()
            p.tofile(fn)
            out[name + "/points"] = pts
            dm = ns.kitti_utils.generate_depth_map(d, fn, 2, shape=[384, 1280])
            out[name + "/depth384"] = dm
            raw = ns.kitti_utils.generate_depth_map(d, fn, 2)
            out[name + "/depth_raw"] = raw
            if name in ("ring4", "edge"):
                out[name + "/depth_vel375"] = ns.kitti_utils.generate_depth_map(
                    d, fn, 3, vel_depth=True, shape=[375, 1242])
            # kitti_dataset.py:105-107 + mono_dataset.py:194-198
            fb = F.max_pool2d(torch.tensor(dm).unsqueeze(0), 2, ceil_mode=True).squeeze().numpy()
            fb = torch.from_numpy(np.expand_dims(fb, 0).astype(np.float32)) / 100.0
            out[name + "/4beam"] = fb[0].numpy()
            e, c = ns.get_4beam_2channel(fb[0])
            out[name + "/2channel"] = torch.stack([e, c]).numpy()
            print("lidar", name, pts.shape, int((dm > 0).sum()), int((fb > 0).sum()))
    # random-density inputs for get_4beam_2channel
    for dens in (0.02, 0.15, 0.5):
        g = torch.Generator().manual_seed(int(dens * 1000))
        fb = (torch.rand(192, 640, generator=g) < dens).float() * (0.02 + torch.rand(192, 640, generator=g))
        e, c = ns.get_4beam_2channel(fb)
        out["rand%03d/4beam" % int(dens * 100)] = fb.numpy()
        out["rand%03d/2channel" % int(dens * 100)] = torch.stack([e, c]).numpy()
        print("2channel density", dens)
    np.savez_compressed(os.path.join(GOLDEN, "lidar.npz"), **out)


def _load_weights(models, seed):
    for i, (name, m) in enumerate(sorted(models.items())):
        m.load_state_dict(synth_weights(m.state_dict(), seed * 100 + i))


def gen_step(ns, B=3, H=96, W=160):
    """Full Trainer.process_batch + backward (config 2 semantics at a small size)."""
    models = RH.make_models(ns, 18)
    _load_weights(models, 0)
    for m in models.values():
        m.train()
    tr = RH.make_trainer(ns, models, B, H, W)
    inputs = synth.make_batch(B, H, W, seed=1, mode="coherent", lidar_density=0.25)
    noise = inputs.pop("noise")
    with RH.FixedNoise([noise[s] for s in range(4)]):
        outputs, losses = tr.process_batch(dict(inputs))
    losses["loss"].backward()
    out = {}
    for s in range(4):
        out["disp%d" % s] = outputs[("disp", s)].detach().numpy()
        out["identity_selection%d" % s] = outputs["identity_selection/%d" % s].numpy()
        out["depth%d" % s] = outputs[("depth", 0, s)].detach().numpy()
    for f in (-1, 1):
        out["cam_T_cam%d" % f] = outputs[("cam_T_cam", 0, f)].detach().numpy()
        out["axisangle%d" % f] = outputs[("axisangle", 0, f)].detach().numpy()
        out["translation%d" % f] = outputs[("translation", 0, f)].detach().numpy()
        out["color%d_0" % f] = outputs[("color", f, 0)].detach().numpy()
    for k, v in losses.items():
        out["loss:" + k] = v.detach().numpy()
    for name, m in sorted(models.items()):
        for k, p in m.named_parameters():
            if p.grad is not None:
                out["gnorm:%s/%s" % (name, k)] = p.grad.double().norm().numpy()
        for k, b in m.named_buffers():
            if k.endswith("running_mean") or k.endswith("running_var"):
                out["buf:%s/%s" % (name, k)] = b.double().norm().numpy()
        out["grad:%s/last" % name] = list(m.parameters())[-1].grad.numpy() \
            if list(m.parameters())[-1].grad is not None else np.zeros(1)
    out["grad:encoder/conv1"] = models["encoder"].encoder.conv1.weight.grad.numpy()
    out["grad:depth/decoder.0"] = models["depth"].decoder[0].conv.conv.weight.grad.numpy()[:8]
    np.savez_compressed(os.path.join(GOLDEN, "step_r18.npz"), **out)
    print("step_r18", {k: float(v) for k, v in losses.items()})


def gen_completor(B=2, H=96, W=160):
    """Completor.process_batch + backward (completor.py:268-321, 546-728: the completion driver's step; the same
    networks and layers.* surface as the trainer, its own copy of the loss code)."""
    ns = RH.load(with_completor=True)
    models = RH.make_models(ns, 18)
    _load_weights(models, 3)
    for m in models.values():
        m.train()
    cp = RH.make_completor(ns, models, B, H, W)
    inputs = synth.make_batch(B, H, W, seed=6, mode="coherent", lidar_density=0.25)
    noise = inputs.pop("noise")
    with RH.FixedNoise([noise[s] for s in range(4)]):
        outputs, losses = cp.process_batch(dict(inputs))
    losses["loss"].backward()
    out = {}
    for s in range(4):
        out["disp%d" % s] = outputs[("disp", s)].detach().numpy()
        out["identity_selection%d" % s] = outputs["identity_selection/%d" % s].numpy()
        out["depth%d" % s] = outputs[("depth", 0, s)].detach().numpy()
    for f in (-1, 1):
        out["cam_T_cam%d" % f] = outputs[("cam_T_cam", 0, f)].detach().numpy()
        out["color%d_0" % f] = outputs[("color", f, 0)].detach().numpy()
    for k, v in losses.items():
        out["loss:" + k] = v.detach().numpy()
    for name, m in sorted(models.items()):
        for k, p in m.named_parameters():
            if p.grad is not None:
                out["gnorm:%s/%s" % (name, k)] = p.grad.double().norm().numpy()
    np.savez_compressed(os.path.join(GOLDEN, "step_completor.npz"), **out)
    print("step_completor", {k: float(v) for k, v in losses.items()})


def gdc_scene(H=120, W=400, seed=0):
    """Synthetic road scene for GDC: ground plane 1.65 m below the camera, a wall at 40 m, two boxes; the
    prediction is the truth with a smooth ~4 % error, the LiDAR ground truth four sparse scan lines of the truth."""
    rng = np.random.RandomState(seed)
    f, cu, cv = 290.0, W / 2.0 - 0.5, H * 0.45
    v, u = np.mgrid[0:H, 0:W].astype(np.float64)
    with np.errstate(divide="ignore"):
        ground = np.where(v > cv + 1, 1.65 * f / np.maximum(v - cv, 1e-6), 1e9)
    true = np.minimum(ground, 40.0)
    for (u0, u1, z) in ((60, 110, 12.0), (250, 300, 22.0)):
        box = (u >= u0) & (u < u1) & (ground > z) & (v > cv - 20)
        true = np.where(box, z, true)
    err = 1.0 + 0.04 * np.sin(u / 37.0) * np.cos(v / 23.0) + 0.003 * rng.randn(H, W)
    pred = true * err
    gt = np.full((H, W), -1.0)
    for row in (int(cv) + 6, int(cv) + 11, int(cv) + 19, int(cv) + 33):
        cols = np.arange(3, W - 3, 3) + rng.randint(0, 2, size=len(np.arange(3, W - 3, 3)))
        gt[row, cols] = true[row, cols] * (1.0 + 0.002 * rng.randn(len(cols)))
    calib_txt = ("R_rect_00: 1 0 0 0 1 0 0 0 1\nP_rect_02: %r 0 %r 13.0 0 %r %r 0.5 0 0 1 0.003\n"
                 "P_rect_03: %r 0 %r -100.0 0 %r %r 0.5 0 0 1 0.003\n" % (f, cu, f, cv, f, cu, f, cv))
    return pred, gt, calib_txt


def gen_gdc():
    """The reference's own GDC (gdc_old.py:74-250) on two synthetic scenes, run under two third-party API shims
    (none edits a reference file): pykdtree.kdtree.KDTree -> scipy.spatial.cKDTree (pykdtree is not installed;
    both are exact kNN), the `tol=` keyword of scipy.sparse.linalg.cg, removed in SciPy 1.14 -> `rtol=`, and
    NumPy 1.x's stacked-vector reading of np.linalg.solve(As, bs)."""
    import tempfile
    import types
    import scipy.sparse.linalg as spl
    from scipy.spatial import cKDTree
    kd = types.ModuleType("pykdtree.kdtree")
    kd.KDTree = cKDTree
    sys.modules.setdefault("pykdtree", types.ModuleType("pykdtree"))
    sys.modules["pykdtree.kdtree"] = kd
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    ref = RH.make_ref.root()
    sys.path.insert(0, ref)
    try:
        import importlib
        gdc_old = importlib.import_module("gdc_old")
        pse = importlib.import_module("kitti_util_from_pse")
    finally:
        sys.path.remove(ref)
    gdc_old.cg = lambda A, b, x0=None, tol=1e-5: spl.cg(A, b, x0=x0, rtol=tol)
    # third shim: NumPy 2 reads a (N, M) right-hand side next to (N, M, M) matrices as ONE matrix; the reference
    # was written for NumPy 1.x, where it is a stack of N vectors (gdc_old.py:186)
    real_solve = np.linalg.solve

    def solve_np1(a, b):
        return real_solve(a, b[..., None])[..., 0] if b.ndim == a.ndim - 1 else real_solve(a, b)

    gdc_old.np = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    gdc_old.np.linalg = types.SimpleNamespace(**{k: getattr(np.linalg, k) for k in dir(np.linalg) if not k.startswith("__")})
    gdc_old.np.linalg.solve = solve_np1
    out = {}
    for i, (seed, rng_deg) in enumerate(((0, (-0.1, 4.0)), (1, (-1.5, 9.0)))):
        pred, gt, calib_txt = gdc_scene(seed=seed)
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
            fh.write(calib_txt)
        calib = pse.Calibration(fh.name)
        os.unlink(fh.name)
        corrected = gdc_old.GDC(pred.copy(), gt.copy(), calib, W_tol=3e-5, recon_tol=5e-4, k=10, method="cg",
                                verbose=False, consider_range=rng_deg)          # inf_gdc.py:81
        out["pred%d" % i], out["gt%d" % i], out["corrected%d" % i] = pred, gt, corrected
        out["calib%d" % i] = np.array([calib.c_u, calib.c_v, calib.f_u, calib.f_v, calib.b_x, calib.b_y])
        out["range%d" % i] = np.array(rng_deg)
        print("gdc scene %d: changed pixels %d, mean |corr - pred| %.4f" % (
            i, int((corrected != pred).sum()), float(np.abs(corrected - pred).mean())))
    np.savez_compressed(os.path.join(GOLDEN, "gdc.npz"), **out)


def gen_forward_variants(ns, H=64, W=96):
    """Config 1 (enc+beam enc+decoder forward, eval mode) for R18 and R50, plus the stage-2
    refine2d decoder (road,catxy,deep) and PoseCNN forward."""
    N = ns.networks
    out = {}
    g = torch.Generator().manual_seed(11)
    rgb = torch.rand(1, 3, H, W, generator=g)
    two = torch.rand(1, 2, H, W, generator=g) * (torch.rand(1, 1, H, W, generator=g) < 0.1)
    out["rgb"], out["two"] = rgb.numpy(), two.numpy()
    for nl in (18, 50):
        enc = N.ResnetEncoder(nl, False)
        benc = N.ResnetEncoder(nl, False, beam_encoder=True)
        dec = N.DepthDecoder(enc.num_ch_enc, [0, 1, 2, 3])
        for i, m in enumerate((enc, benc, dec)):
            m.load_state_dict(synth_weights(m.state_dict(), 1000 + nl * 10 + i))
            m.eval()
        with torch.no_grad():
            feats, bf = enc(rgb), benc(two)
            feats = [f.clone() for f in feats]
            d = dec(feats, beam_features=bf)
        for s in range(4):
            out["r%d/disp%d" % (nl, s)] = d[("disp", s)].numpy()
        out["r%d/feat4" % nl] = feats[4].numpy()
        if nl == 18:
            ref2d = N.DepthDecoder(enc.num_ch_enc, [0, 1, 2, 3], road=True, catxy=True, deep=True)
            ref2d.load_state_dict(synth_weights(ref2d.state_dict(), 2000))
            ref2d.eval()
            dm = {("disp", s): torch.rand(1, 6, H >> s, W >> s, generator=g) for s in range(4)}
            with torch.no_grad():
                r = ref2d(feats, beam_features=bf, depth_maps=dm, tanh=False)
            for s in range(4):
                out["refine/dm%d" % s] = dm[("disp", s)].numpy()
                out["refine/disp%d" % s] = r[("disp", s)].numpy()
    pc = N.PoseCNN(2)
    pc.load_state_dict(synth_weights(pc.state_dict(), 3000))
    with torch.no_grad():
        aa, tt = pc(torch.cat([rgb, rgb.flip(3)], 1))
    out["posecnn/aa"], out["posecnn/t"] = aa.numpy(), tt.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "forward_variants.npz"), **out)
    print("forward variants done")


def gen_loss_chain(ns, B=2, H=96, W=160):
    """generate_images_pred + compute_losses alone on coherent frames, with grads wrt the
    four disparities and the two poses."""
    models = {}
    tr = RH.make_trainer(ns, models, B, H, W)
    inputs = synth.make_batch(B, H, W, seed=5, mode="coherent")
    noise = inputs.pop("noise")
    g = torch.Generator().manual_seed(77)
    outputs = {}
    leaves = {}
    for s in range(4):
        low = torch.rand(B, 1, (H >> s) // 4 + 2, (W >> s) // 4 + 2, generator=g)
        d = torch.sigmoid(3 * (F.interpolate(low, (H >> s, W >> s), mode="bilinear",
                                             align_corners=False) - 0.5) - 1.0)
        d = d + 0.01 * torch.rand(B, 1, H >> s, W >> s, generator=g)
        leaves["disp%d" % s] = d.clone().requires_grad_(True)
        outputs[("disp", s)] = leaves["disp%d" % s]
    for f in (-1, 1):
        aa = (0.01 * torch.randn(B, 1, 3, generator=g)).requires_grad_(True)
        tt = (0.05 * torch.randn(B, 1, 3, generator=g)).requires_grad_(True)
        leaves["aa%d" % f], leaves["tt%d" % f] = aa, tt
        T = ns.layers.transformation_from_parameters(aa, tt, invert=(f < 0))
        T.retain_grad()
        leaves["T%d" % f] = T
        outputs[("cam_T_cam", 0, f)] = T
    # beam map consistent with depth*26 so the si-loss mask is populated
    with torch.no_grad():
        up = F.interpolate(leaves["disp0"], [H, W], mode="bilinear", align_corners=False)
        dep = 26.0 / (0.01 + 9.99 * up)
        mask = (torch.rand(B, 1, H, W, generator=g) < 0.05).float()
        inputs["4beam"] = mask * (dep + (torch.rand(B, 1, H, W, generator=g) - 0.5) * 3.0) / 100.0
    tr.generate_images_pred(inputs, outputs, [0, -1, 1])
    with RH.FixedNoise([noise[s] for s in range(4)]):
        losses = tr.compute_losses(inputs, outputs)
    losses["loss"].backward()
    out = {"4beam": inputs["4beam"].numpy()}
    for k, v in leaves.items():
        out["in:" + k] = v.detach().numpy()
        out["grad:" + k] = v.grad.numpy()
    for k, v in losses.items():
        out["loss:" + k] = v.detach().numpy()
    for s in range(4):
        out["identity_selection%d" % s] = outputs["identity_selection/%d" % s].numpy()
        out["depth%d" % s] = outputs[("depth", 0, s)].detach().numpy()
        if s in (0, 3):
            for f in (-1, 1):
                out["color%d_%d" % (f, s)] = outputs[("color", f, s)].detach().numpy()
    np.savez_compressed(os.path.join(GOLDEN, "loss_chain.npz"), **out)
    print("loss_chain", {k: float(v) for k, v in losses.items()})


def gen_refiner(B=2, H=192, W=640):
    """Refiner.process_batch + backward (BASELINE config 5 semantics; the reference hard-codes 192x640 in
    the GDC term and the 78:190 x 23:617 crop, so the fixture runs at that size with B=2)."""
    ns = RH.load(with_refiner=True)
    models = RH.make_models(ns, 18)
    models["refine2d_decoder"] = RH.make_refine_decoder(ns, models["encoder"].num_ch_enc)
    _load_weights(models, 4)
    for m in models.values():
        m.train()                                    # refiner.py:268 set_train() flips the frozen nets too
    rf = RH.make_refiner(ns, models, B, H, W)
    inputs = synth.make_refiner_batch(B, H, W, seed=6)
    noise = inputs.pop("noise")
    with RH.FixedNoise([noise[s] for s in range(4)]):
        outputs, losses = rf.process_batch(dict(inputs))
    losses["loss"].backward()
    out = {}
    for s in range(4):
        d = outputs[("disp", s)].detach().numpy()
        out["disp%d" % s] = d if s else d[:, :, ::4, ::4]
        out["identity_selection%d_mean" % s] = outputs["identity_selection/%d" % s].mean().numpy()
    for f in (-1, 1):
        out["cam_T_cam%d" % f] = outputs[("cam_T_cam", 0, f)].detach().numpy()
    for k, v in losses.items():
        out["loss:" + k] = v.detach().numpy()
    for k, p in models["refine2d_decoder"].named_parameters():
        out["gnorm:" + k] = p.grad.double().norm().numpy()
    out["grad:decoder.0.0"] = models["refine2d_decoder"].decoder[0][0].conv.conv.weight.grad.numpy()[:4]
    out["grad:decoder.13"] = models["refine2d_decoder"].decoder[13].conv.weight.grad.numpy()
    for k, b in models["encoder"].named_buffers():
        if k.endswith("running_mean"):
            out["buf:encoder/" + k] = b.double().norm().numpy()
    np.savez_compressed(os.path.join(GOLDEN, "refiner.npz"), **out)
    print("refiner", {k: float(v) for k, v in losses.items()})

    # the pseudo-3D pack alone (refiner.py:316-346) on a synthetic coarse disparity: stored whole
    Bp = 2
    inputs = synth.make_refiner_batch(Bp, H, W, seed=8)
    coarse = coarse_disparity(Bp, H, W, seed=3)
    rf2 = RH.make_refiner(ns, models, Bp, H, W)

    class _Frozen:                                        # stands in for the three frozen networks
        def __init__(self, ret):
            self.ret = ret

        def __call__(self, *a, **k):
            return self.ret
    cap = {}

    class _Capture:
        def __call__(self, features, beam_features=None, depth_maps=None, tanh=False):
            for s in range(4):
                cap[s] = depth_maps[("disp", s)].clone()
            raise StopIteration
    rf2.models = dict(models)
    rf2.models["encoder"] = _Frozen(None)
    rf2.models["beam_encoder"] = _Frozen(None)
    rf2.models["depth"] = _Frozen({("disp", 0): coarse})
    rf2.models["refine2d_decoder"] = _Capture()
    rf2.use_pose_net = False
    try:
        rf2.process_batch(dict((k, v) for k, v in inputs.items() if k != "noise"))
    except StopIteration:
        pass
    pk = {}
    for s in range(4):
        pk["pack%d" % s] = cap[s].numpy() if s else cap[s].numpy()[:, :, ::2, ::2]
    np.savez_compressed(os.path.join(GOLDEN, "refiner_pack.npz"), **pk)
    print("refiner pack", [tuple(cap[s].shape) for s in range(4)])


def gen_r50_train(B=2, H=96, W=160):
    """ResNet-50 (Bottleneck) encoder + beam encoder + decoder in TRAIN mode with backward: a scalar loss
    over the four disparities, feature maps, parameter-gradient norms and BN running statistics."""
    ns = RH.load()
    N = ns.networks
    enc, benc = N.ResnetEncoder(50, False), N.ResnetEncoder(50, False, beam_encoder=True)
    dec = N.DepthDecoder(enc.num_ch_enc, [0, 1, 2, 3])
    mods = {"enc": enc, "benc": benc, "dec": dec}
    for i, (k, m) in enumerate(sorted(mods.items())):
        m.load_state_dict(synth_weights(m.state_dict(), 5000 + i))
        m.train()
    g = torch.Generator().manual_seed(21)
    rgb = torch.rand(B, 3, H, W, generator=g)
    two = torch.rand(B, 2, H, W, generator=g) * (torch.rand(B, 1, H, W, generator=g) < 0.1)
    wts = [torch.randn(B, 1, H >> s, W >> s, generator=g) for s in range(4)]
    feats = enc(rgb)
    feats = [f for f in feats]
    d = dec(feats, beam_features=benc(two))
    loss = sum((d[("disp", s)] * wts[s]).mean() for s in range(4))
    loss.backward()
    out = {"rgb": rgb.numpy(), "two": two.numpy(), "loss": loss.detach().numpy()}
    for s in range(4):
        out["w%d" % s] = wts[s].numpy()
        out["disp%d" % s] = d[("disp", s)].detach().numpy()
    out["feat4"] = feats[4].detach().numpy()
    for name, m in mods.items():
        for k, p in m.named_parameters():
            if p.grad is not None:
                out["gnorm:%s/%s" % (name, k)] = p.grad.double().norm().numpy()
        for k, b in m.named_buffers():
            if k.endswith("running_mean") or k.endswith("running_var"):
                out["buf:%s/%s" % (name, k)] = b.double().norm().numpy()
    out["grad:enc/layer1.0.conv1"] = enc.encoder.layer1[0].conv1.weight.grad.numpy()
    out["grad:enc/layer4.2.conv3"] = enc.encoder.layer4[2].conv3.weight.grad.numpy()[:16]
    np.savez_compressed(os.path.join(GOLDEN, "r50_train.npz"), **out)
    print("r50 train", float(loss))


def gen_state_dict_keys(ns):
    """key -> shape of every network variant the drivers build, from the REFERENCE modules."""
    import json
    N = ns.networks
    ch = np.array([64, 64, 128, 256, 512])
    built = {
        "enc18": N.ResnetEncoder(18, False), "enc50": N.ResnetEncoder(50, False),
        "beam18": N.ResnetEncoder(18, False, beam_encoder=True),
        "pose18": N.ResnetEncoder(18, False, num_input_images=2),
        "beampose18": N.ResnetEncoder(18, False, num_input_images=2, beam_encoder=True),
        "refineenc18": N.ResnetEncoder(18, False, refine_encoder=True),
        "depth": N.DepthDecoder(ch, [0, 1, 2, 3]),
        "refine2d": N.DepthDecoder(ch, [0, 1, 2, 3], road=True, catxy=True, deep=True),
        "cat2end": N.DepthDecoder(ch, [0, 1, 2, 3], cat2end=True),
        "pose": N.PoseDecoder(ch, 1, 2), "posecnn": N.PoseCNN(2),
    }
    out = {k: [[kk, list(v.shape)] for kk, v in m.state_dict().items()] for k, m in built.items()}
    json.dump(out, open(os.path.join(GOLDEN, "state_dict_keys.json"), "w"))


if __name__ == "__main__":
    assert RH.available(), "needs /root/reference (build container only)"
    os.makedirs(GOLDEN, exist_ok=True)
    ns = RH.load()
    which = sys.argv[1:] or ["lidar", "step", "fwd", "loss", "keys", "refiner", "r50", "completor", "gdc"]
    if "refiner" in which:
        gen_refiner()
    if "completor" in which:
        gen_completor()
    if "gdc" in which:
        gen_gdc()
    if "r50" in which:
        gen_r50_train()
    if "keys" in which:
        gen_state_dict_keys(ns)
    if "lidar" in which:
        gen_lidar(ns)
    if "step" in which:
        gen_step(ns)
    if "fwd" in which:
        gen_forward_variants(ns)
    if "loss" in which:
        gen_loss_chain(ns)

