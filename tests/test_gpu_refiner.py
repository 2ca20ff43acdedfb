"""Stage 2 (refiner.py, BASELINE config 5): the pseudo-3D pack, the refine2d decoder on the channel-padded
tensor-core path, the GDC-clone si-loss, and the whole step -- against fixtures produced by the UNMODIFIED
reference's Refiner.process_batch (tests/make_golden.py gen_refiner) and against the CPU oracle."""
import numpy as np
import pytest
import torch

from tests._util import GOLDEN, clone_sd, coarse_disparity, rel_err, synth_weights
from fusiondepth_b200 import synth
from oracle import step_oracle as SO

pytestmark = pytest.mark.gpu


def _gtol(key):
    """Gradient-norm tolerance.  The biases of the four 1-channel disparity heads get the plain SUM of
    dL/dlogit over all pixels -- signed terms that cancel to ~1e-3 of their absolute sum -- so fp32
    summation-order differences of 1e-6 show up as ~1 % there.  Their weights see the si-loss term of their scale
    directly, and that term averages over the few LiDAR points passing `|depth - beam| < thresh`: one point at the
    threshold flipping (fp32 rounding; varies run to run with the atomics upstream) moved decoder.12.conv.weight's
    norm by 0.8 % in one of five runs.  Heads: 2e-2; every other tensor holds 5e-3."""
    return 2e-2 if key.split(".")[1] in ("10", "11", "12", "13") and ".conv." in key else 5e-3


def _load(models, seed):
    sds = {}
    for i, (name, m) in enumerate(sorted(models.items())):
        sds[name] = synth_weights(m.state_dict(), seed * 100 + i)
        m.load_state_dict(sds[name])
    return sds


def test_refine_pack_vs_reference_fixture(cuda):
    """refiner.py:316-346 on a synthetic coarse disparity: the reference's own 6-channel maps."""
    from fusiondepth_b200 import ops
    g = np.load(GOLDEN + "/refiner_pack.npz")
    inputs = synth.make_refiner_batch(2, 192, 640, seed=8)
    coarse = coarse_disparity(2, 192, 640, seed=3)
    packed, ratios = ops.refine_pack(coarse.cuda(), inputs["4beam"].cuda(), inputs["2channel"].cuda(),
                                     [inputs[("inv_K", s)].cuda() for s in range(4)])
    _, oratios = SO.pseudo3d_pack({("disp", 0): coarse}, inputs)
    for s in range(4):
        assert float(ratios[s]) == float(oratios[s]), (s, float(ratios[s]), float(oratios[s]))   # exact medians
        got = packed[s].cpu()
        assert got.shape == (2, 6, 192 >> s, 640 >> s)
        got = got if s else got[:, :, ::2, ::2]
        want = torch.from_numpy(g["pack%d" % s])
        assert torch.equal(got[:, 4:6], want[:, 4:6]), s                      # max-pooled 2-channel map
        assert rel_err(got[:, 0], want[:, 0]) < 1e-5, s                       # scaled disparity
        for c in (1, 2, 3):
            assert rel_err(got[:, c], want[:, c]) < 1e-5, (s, c)              # x/30, y/2, (z-40)/40


def test_refiner_step_vs_reference_fixture(cuda):
    """Refiner.process_batch + backward at 2x192x640: every loss, the refined disparities, the poses, all
    48 gradient norms of the refine2d decoder and two full gradient tensors."""
    from fusiondepth_b200 import refine
    g = np.load(GOLDEN + "/refiner.npz")
    models = refine.build_refiner_models(18, "cuda")
    _load(models, 4)
    for m in models.values():
        m.train()
    inputs = synth.make_refiner_batch(2, 192, 640, seed=6)
    noise = {s: t.cuda() for s, t in inputs.pop("noise").items()}
    outputs, losses = refine.process_batch(models, synth.to_device(inputs, "cuda"), noise, None, materialize=True)
    losses["loss"].backward()
    torch.cuda.synchronize()
    for k in losses:
        assert rel_err(losses[k].detach().cpu(), g["loss:" + k]) < 1e-4, (k, float(losses[k]), float(g["loss:" + k]))
    for s in range(4):
        d = outputs[("disp", s)].detach().cpu()
        assert rel_err(d if s else d[:, :, ::4, ::4], g["disp%d" % s]) < 1e-4, s
        sel = outputs["identity_selection/%d" % s].mean()
        assert abs(float(sel) - float(g["identity_selection%d_mean" % s])) < 1e-3
    for f in (-1, 1):
        assert rel_err(outputs[("cam_T_cam", 0, f)].detach().cpu(), g["cam_T_cam%d" % f]) < 1e-5
    bad, n = [], 0
    dec = dict(models["refine2d_decoder"].named_parameters())
    for key in g.files:
        if key.startswith("gnorm:"):
            got, want = float(dec[key[6:]].grad.double().norm()), float(g[key])
            n += 1
            if abs(got - want) > _gtol(key) * want + 1e-9:
                bad.append((key, got, want))
        elif key.startswith("buf:encoder/"):
            b = dict(models["encoder"].named_buffers())[key[12:]]
            assert abs(float(b.double().norm()) - float(g[key])) < 1e-4 * float(g[key]), key
    assert n == 48 and not bad, bad[:8]
    assert rel_err(dec["decoder.0.0.conv.conv.weight"].grad.cpu()[:4], g["grad:decoder.0.0"]) < 1e-2
    assert rel_err(dec["decoder.13.conv.weight"].grad.cpu(), g["grad:decoder.13"]) < 1e-2
    # the frozen networks carry no gradient
    for name in ("encoder", "beam_encoder", "depth", "pose_encoder", "pose", "beam_encoder_pose"):
        assert all(p.grad is None for p in models[name].parameters()), name


def test_refine_step_graph_vs_oracle(cuda):
    """RefineStep (captured CUDA graph, stream-parallel frozen trunks) at the bench size 6x192x640: loss of
    the step and the Adam-updated refine2d weights against the CPU oracle."""
    from fusiondepth_b200 import refine
    B, H, W = 6, 192, 640
    models = refine.build_refiner_models(18, "cuda")
    sds = _load(models, 9)
    osd = {k: clone_sd(v, requires_grad=(k == "refine2d_decoder")) for k, v in sds.items()}
    inputs = synth.make_refiner_batch(B, H, W, seed=12)
    noise = inputs.pop("noise")
    _, ol = SO.refiner_process_batch(osd, inputs, noise, 18, training=True)
    ol["loss"].backward()
    step = refine.RefineStep(models, lr=1e-4)
    cb = [synth.to_device(inputs, "cuda")]
    cn = [{s: t.cuda() for s, t in noise.items()}]
    step.capture(cb, cn, warmup=1)
    w0 = step.flat.data.clone()
    loss = float(step.replay())
    torch.cuda.synchronize()
    assert abs(loss - float(ol["loss"])) < 1e-4 * abs(float(ol["loss"])), (loss, float(ol["loss"]))
    dec = dict(models["refine2d_decoder"].named_parameters())
    for key in ("decoder.0.0.conv.conv.weight", "decoder.3.1.conv.conv.weight", "decoder.7.0.conv.conv.weight",
                "decoder.9.1.conv.conv.weight", "decoder.10.conv.weight"):
        assert rel_err(dec[key].grad.cpu(), osd["refine2d_decoder"][key].grad) < 5e-3, key
    # only the refine2d decoder moved
    nt = step.flat.n_train
    assert float((step.flat.data[nt:] - w0[nt:]).abs().max()) == 0.0
    moved = (step.flat.data[:nt] - w0[:nt]).abs()
    assert float(moved.max()) > 0.5e-4 and float(moved.max()) <= 1.01e-4


def test_r50_train_vs_reference_fixture(cuda):
    """Bottleneck (ResNet-50) encoders + decoder in TRAIN mode with backward (BASELINE config 3's networks):
    disparities, features, every parameter-gradient norm and BN running statistic vs the reference."""
    from fusiondepth_b200 import networks
    g = np.load(GOLDEN + "/r50_train.npz")
    enc, benc = networks.ResnetEncoder(50, False), networks.ResnetEncoder(50, False, beam_encoder=True)
    dec = networks.DepthDecoder(enc.num_ch_enc, [0, 1, 2, 3])
    mods = {"enc": enc, "benc": benc, "dec": dec}
    for i, (k, m) in enumerate(sorted(mods.items())):
        m.load_state_dict(synth_weights(m.state_dict(), 5000 + i))
        m.cuda().train()
    rgb, two = torch.from_numpy(g["rgb"]).cuda(), torch.from_numpy(g["two"]).cuda()
    feats = list(enc(rgb))
    d = dec(feats, beam_features=benc(two))
    loss = sum((d[("disp", s)] * torch.from_numpy(g["w%d" % s]).cuda()).mean() for s in range(4))
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(loss.detach().cpu(), g["loss"]) < 1e-4
    for s in range(4):
        assert rel_err(d[("disp", s)].detach().cpu(), g["disp%d" % s]) < 1e-4, s
    # 2048-channel layer-4 features after 16 Bottlenecks whose BatchNorms see only 30 samples per channel at this
    # size: an intermediate tensor, not one of the depth / disp / loss tensors the 1e-4 bar is stated for
    assert rel_err(feats[4].detach().cpu(), g["feat4"]) < 5e-4
    bad, n = [], 0
    for key in g.files:
        if key.startswith("gnorm:"):
            name, pk = key[6:].split("/", 1)
            got, want = float(dict(mods[name].named_parameters())[pk].grad.double().norm()), float(g[key])
            n += 1
            # 1e-2: the beam encoder sees a 90 %-empty input and B*H*W is small here, so its BatchNorm
            # gradients are sums with heavy cancellation (trunk conv weights < 0.2 % off).  The beam encoder's
            # BatchNorm scale / shift gradients move from run to run with the order of the fp32 atomic adds in the
            # weight-gradient and statistic kernels (0.3 ... 1.1 % observed over repeated runs): 2e-2 for those.
            tol = 2e-2 if (name == "benc" and (".bn" in pk or "downsample.1" in pk)) else 1e-2
            if abs(got - want) > tol * want + 1e-9:
                bad.append((key, got, want))
        elif key.startswith("buf:"):
            name, pk = key[4:].split("/", 1)
            got, want = float(dict(mods[name].named_buffers())[pk].double().norm()), float(g[key])
            if abs(got - want) > 1e-4 * want + 1e-9:
                bad.append((key, got, want))
    assert n > 300 and not bad, bad[:8]
    # element-wise (max |diff| / max |ref|): the layer-4 BatchNorms normalise over 30 samples per channel here, which
    # makes every upstream gradient ill-conditioned (the CPU reference itself moves by ~1 % between BLAS builds)
    assert rel_err(enc.encoder.layer1[0].conv1.weight.grad.cpu(), g["grad:enc/layer1.0.conv1"]) < 3e-2
    assert rel_err(enc.encoder.layer4[2].conv3.weight.grad.cpu()[:16], g["grad:enc/layer4.2.conv3"]) < 3e-2
