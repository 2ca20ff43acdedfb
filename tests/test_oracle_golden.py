"""Pins the CPU oracle (oracle/) against fixtures produced by the UNMODIFIED reference
(tests/make_golden.py, run in the build container).  Runs everywhere, no GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests._util import GOLDEN, clone_sd, rel_err, synth_weights
from fusiondepth_b200 import networks, synth
from oracle import lidar_oracle as LO
from oracle import step_oracle as SO

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


def _g(name):
    return np.load("%s/%s.npz" % (GOLDEN, name))


# ------------------------------------------------------------------------------------------ lidar
@pytest.mark.parametrize("case", ["ring4", "piled", "dense", "edge"])
def test_lidar_oracle_bit_exact(case):
    g = _g("lidar")
    P2 = synth.velo_to_image_matrix(synth.parse_roundtrip(), 2)
    pts = g[case + "/points"]
    d = LO.depth_map(pts, P2, 1242, 375, shape=(384, 1280))
    assert np.array_equal(d, g[case + "/depth384"])
    assert np.array_equal(LO.depth_map(pts, P2, 1242, 375), g[case + "/depth_raw"])
    if case + "/depth_vel375" in g:
        P3 = synth.velo_to_image_matrix(synth.parse_roundtrip(), 3)
        assert np.array_equal(LO.depth_map(pts, P3, 1242, 375, vel_depth=True, shape=(375, 1242)),
                              g[case + "/depth_vel375"])
    fb = LO.pool_scale(d)
    assert np.array_equal(fb, g[case + "/4beam"])
    assert np.array_equal(LO.two_channel(fb), g[case + "/2channel"])


@pytest.mark.parametrize("case", ["rand002", "rand015", "rand050"])
def test_two_channel_oracle_bit_exact(case):
    g = _g("lidar")
    assert np.array_equal(LO.two_channel(g[case + "/4beam"]), g[case + "/2channel"])


def test_lidar_scans_regenerate():
    """the fixtures' inputs come from seeds: the generator must be stable"""
    g = _g("lidar")
    assert np.array_equal(synth.make_scan(3), g["ring4/points"])
    assert np.array_equal(synth.make_scan(5, piled=600), g["piled/points"])


def test_lidar_empty_and_degenerate():
    P2 = synth.velo_to_image_matrix(synth.parse_roundtrip(), 2)
    z = LO.depth_map(np.zeros((0, 4), np.float32), P2, 1242, 375, shape=(384, 1280))
    assert z.shape == (384, 1280) and not z.any()
    behind = np.array([[-1.0, 0, 0, 0], [np.nan, 0, 0, 0]], np.float32)
    assert not LO.depth_map(behind, P2, 1242, 375).any()
    assert not LO.two_channel(np.zeros((192, 640), np.float32)).any()


# ------------------------------------------------------------------------------------------ nets
def _models_sd(seed=0, num_layers=18):
    from fusiondepth_b200.training import MODEL_NAMES  # noqa: F401
    tmpl = {
        "encoder": networks.ResnetEncoder(num_layers, False),
        "beam_encoder": networks.ResnetEncoder(num_layers, False, beam_encoder=True),
        "beam_encoder_pose": networks.ResnetEncoder(num_layers, False, num_input_images=2, beam_encoder=True),
        "depth": networks.DepthDecoder(np.array([64, 64, 128, 256, 512]), [0, 1, 2, 3]),
        "pose_encoder": networks.ResnetEncoder(num_layers, False, num_input_images=2),
        "pose": networks.PoseDecoder(np.array([64, 64, 128, 256, 512]), 1, 2),
    }
    return {name: synth_weights(m.state_dict(), seed * 100 + i)
            for i, (name, m) in enumerate(sorted(tmpl.items()))}


def test_step_oracle_matches_reference_fixture():
    g = _g("step_r18")
    sds = {k: clone_sd(v, requires_grad=True) for k, v in _models_sd(0).items()}
    inputs = synth.make_batch(3, 96, 160, seed=1, mode="coherent", lidar_density=0.25)
    noise = inputs.pop("noise")
    outputs, losses = SO.process_batch(sds, inputs, noise, 18, training=True)
    losses["loss"].backward()
    for k in losses:
        assert rel_err(losses[k].detach(), g["loss:" + k]) < 2e-5, k
    for s in range(4):
        assert rel_err(outputs[("disp", s)].detach(), g["disp%d" % s]) < 1e-5
        assert rel_err(outputs[("depth", 0, s)].detach(), g["depth%d" % s]) < 1e-5
        assert (outputs["identity_selection/%d" % s].numpy() != g["identity_selection%d" % s]).mean() < 1e-3
    for f in (-1, 1):
        assert rel_err(outputs[("cam_T_cam", 0, f)].detach(), g["cam_T_cam%d" % f]) < 1e-6
        assert rel_err(outputs[("color", f, 0)].detach(), g["color%d_0" % f]) < 1e-4
    checked = 0
    for key in g.files:
        if key.startswith("gnorm:"):
            name, pk = key[6:].split("/", 1)
            gr = sds[name][pk].grad
            assert gr is not None, key
            assert abs(float(gr.double().norm()) - float(g[key])) <= 2e-4 * float(g[key]) + 1e-9, key
            checked += 1
        elif key.startswith("buf:"):
            name, pk = key[4:].split("/", 1)
            assert abs(float(sds[name][pk].double().norm()) - float(g[key])) <= 1e-5 * float(g[key]) + 1e-9, key
    assert checked > 200
    assert rel_err(sds["encoder"]["encoder.conv1.weight"].grad, g["grad:encoder/conv1"]) < 1e-3


def test_forward_variants_oracle():
    g = _g("forward_variants")
    rgb, two = torch.from_numpy(g["rgb"]), torch.from_numpy(g["two"])
    for nl in (18, 50):
        enc = networks.ResnetEncoder(nl, False)
        benc = networks.ResnetEncoder(nl, False, beam_encoder=True)
        dec = networks.DepthDecoder(enc.num_ch_enc, [0, 1, 2, 3])
        sds = [synth_weights(m.state_dict(), 1000 + nl * 10 + i) for i, m in enumerate((enc, benc, dec))]
        with torch.no_grad():
            f = SO.resnet_encoder(sds[0], rgb, nl, training=False)
            b = SO.resnet_encoder(sds[1], two, nl, training=False)
            d = SO.depth_decoder(sds[2], f, beam_feats=b)
        assert rel_err(f[4], g["r%d/feat4" % nl]) < 1e-5
        for s in range(4):
            assert rel_err(d[("disp", s)], g["r%d/disp%d" % (nl, s)]) < 1e-5
        if nl == 18:
            r2d = networks.DepthDecoder(enc.num_ch_enc, [0, 1, 2, 3], road=True, catxy=True, deep=True)
            sd = synth_weights(r2d.state_dict(), 2000)
            dm = {("disp", s): torch.from_numpy(g["refine/dm%d" % s]) for s in range(4)}
            with torch.no_grad():
                r = SO.depth_decoder(sd, f, beam_feats=b, depth_maps=dm, deep=True)
            for s in range(4):
                assert rel_err(r[("disp", s)], g["refine/disp%d" % s]) < 1e-5


def loss_chain_inputs(g):
    inputs = synth.make_batch(2, 96, 160, seed=5, mode="coherent")
    noise = inputs.pop("noise")
    inputs["4beam"] = torch.from_numpy(g["4beam"])
    return inputs, noise


def test_loss_chain_oracle():
    g = _g("loss_chain")
    inputs, noise = loss_chain_inputs(g)
    disps = {("disp", s): torch.from_numpy(g["in:disp%d" % s]).requires_grad_(True) for s in range(4)}
    Ts = {f: torch.from_numpy(g["in:T%d" % f]).requires_grad_(True) for f in (-1, 1)}
    losses, outs = SO.photometric_chain(inputs, disps, Ts, noise)
    losses["loss"].backward()
    for k in losses:
        assert rel_err(losses[k].detach(), g["loss:" + k]) < 1e-5, k
    for s in range(4):
        assert rel_err(disps[("disp", s)].grad, g["grad:disp%d" % s]) < 1e-4
        assert np.array_equal(outs["identity_selection/%d" % s].numpy(), g["identity_selection%d" % s])
    for f in (-1, 1):
        assert rel_err(Ts[f].grad, g["grad:T%d" % f]) < 1e-4
        # pose matrix restatement
        T = SO.pose_matrix(torch.from_numpy(g["in:aa%d" % f]), torch.from_numpy(g["in:tt%d" % f]), f < 0)
        assert rel_err(T, g["in:T%d" % f]) < 1e-6


def test_adam_oracle_matches_torch():
    g = torch.Generator().manual_seed(0)
    p = torch.randn(1000, generator=g)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], 1e-3)
    m, v = torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 4):
        gr = torch.randn(1000, generator=g)
        ref.grad = gr.clone()
        opt.step()
        SO.adam_step([p], [gr], [m], [v], step, 1e-3)
    assert rel_err(p, ref.detach()) < 1e-6


# ------------------------------------------------------------------------------------------ refiner
def _refiner_sds(seed=4):
    tmpl = {
        "encoder": networks.ResnetEncoder(18, False),
        "beam_encoder": networks.ResnetEncoder(18, False, beam_encoder=True),
        "beam_encoder_pose": networks.ResnetEncoder(18, False, num_input_images=2, beam_encoder=True),
        "depth": networks.DepthDecoder(np.array([64, 64, 128, 256, 512]), [0, 1, 2, 3]),
        "pose_encoder": networks.ResnetEncoder(18, False, num_input_images=2),
        "pose": networks.PoseDecoder(np.array([64, 64, 128, 256, 512]), 1, 2),
        "refine2d_decoder": networks.DepthDecoder(np.array([64, 64, 128, 256, 512]), [0, 1, 2, 3],
                                                  road=True, catxy=True, deep=True),
    }
    return {name: synth_weights(m.state_dict(), seed * 100 + i)
            for i, (name, m) in enumerate(sorted(tmpl.items()))}


def test_refiner_pack_oracle_matches_reference_fixture():
    """refiner.py:316-346 (median rescale, cumulative max-pools, Cat_xy) on a synthetic coarse disparity."""
    from tests._util import coarse_disparity
    g = _g("refiner_pack")
    inputs = synth.make_refiner_batch(2, 192, 640, seed=8)
    packed, _ = SO.pseudo3d_pack({("disp", 0): coarse_disparity(2, 192, 640, seed=3)}, inputs)
    for s in range(4):
        got = packed[("disp", s)] if s else packed[("disp", s)][:, :, ::2, ::2]
        assert torch.equal(got, torch.from_numpy(g["pack%d" % s])), s


def test_refiner_oracle_matches_reference_fixture():
    """Refiner.process_batch + backward at 2x192x640 (BASELINE config 5 semantics)."""
    g = _g("refiner")
    sds = {k: clone_sd(v, requires_grad=(k == "refine2d_decoder")) for k, v in _refiner_sds().items()}
    inputs = synth.make_refiner_batch(2, 192, 640, seed=6)
    noise = inputs.pop("noise")
    outputs, losses = SO.refiner_process_batch(sds, inputs, noise, 18, training=True)
    losses["loss"].backward()
    for k in losses:
        assert rel_err(losses[k].detach(), g["loss:" + k]) < 1e-5, (k, float(losses[k]), float(g["loss:" + k]))
    for s in range(4):
        d = outputs[("disp", s)].detach()
        assert rel_err(d if s else d[:, :, ::4, ::4], g["disp%d" % s]) < 1e-5, s
        assert abs(float(outputs["identity_selection/%d" % s].mean()) - float(g["identity_selection%d_mean" % s])) < 1e-4
    for f in (-1, 1):
        assert rel_err(outputs[("cam_T_cam", 0, f)].detach(), g["cam_T_cam%d" % f]) < 1e-5
    n = 0
    for key in g.files:
        if key.startswith("gnorm:"):
            got = float(sds["refine2d_decoder"][key[6:]].grad.double().norm())
            assert abs(got - float(g[key])) < 1e-3 * float(g[key]) + 1e-12, key
            n += 1
        elif key.startswith("buf:encoder/"):
            got = float(sds["encoder"][key[12:]].double().norm())
            assert abs(got - float(g[key])) < 1e-5 * float(g[key]), key
    assert n == 48
    assert rel_err(sds["refine2d_decoder"]["decoder.0.0.conv.conv.weight"].grad[:4], g["grad:decoder.0.0"]) < 1e-3
    assert rel_err(sds["refine2d_decoder"]["decoder.13.conv.weight"].grad, g["grad:decoder.13"]) < 1e-3


def test_r50_train_oracle_matches_reference_fixture():
    """Bottleneck (ResNet-50) train-mode forward + backward vs the reference fixture."""
    g = _g("r50_train")
    ch = np.array([64, 256, 512, 1024, 2048])
    tmpl = {"enc": networks.ResnetEncoder(50, False), "benc": networks.ResnetEncoder(50, False, beam_encoder=True),
            "dec": networks.DepthDecoder(ch, [0, 1, 2, 3])}
    sds = {k: clone_sd(synth_weights(m.state_dict(), 5000 + i), requires_grad=True)
           for i, (k, m) in enumerate(sorted(tmpl.items()))}
    rgb, two = torch.from_numpy(g["rgb"]), torch.from_numpy(g["two"])
    feats = SO.resnet_encoder(sds["enc"], rgb, 50, True)
    d = SO.depth_decoder(sds["dec"], feats, beam_feats=SO.resnet_encoder(sds["benc"], two, 50, True))
    loss = sum((d[("disp", s)] * torch.from_numpy(g["w%d" % s])).mean() for s in range(4))
    loss.backward()
    assert rel_err(loss.detach(), g["loss"]) < 1e-5
    for s in range(4):
        assert rel_err(d[("disp", s)].detach(), g["disp%d" % s]) < 1e-5
    assert rel_err(feats[4].detach(), g["feat4"]) < 1e-5
    n = 0
    for key in g.files:
        if key.startswith("gnorm:"):
            name, pk = key[6:].split("/", 1)
            got = float(sds[name][pk].grad.double().norm())
            assert abs(got - float(g[key])) < 1e-3 * float(g[key]) + 1e-12, key
            n += 1
    assert n > 300
