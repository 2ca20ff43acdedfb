"""The reference-facing module surface (names, signatures, state-dict keys) and the unfused
geometry layers vs the oracle -- CPU only."""
import json
import os
import sys

import numpy as np
import torch

from tests._util import GOLDEN, ROOT, rel_err
from oracle import step_oracle as SO


def test_layers_exports_what_the_drivers_use():
    from fusiondepth_b200 import layers
    for name in ("disp_to_depth", "transformation_from_parameters", "rot_from_axisangle",
                 "get_translation_matrix", "ConvBlock", "Conv3x3", "BackprojectDepth", "Cat_xy",
                 "Project3D", "upsample", "get_smooth_loss", "SSIM", "compute_depth_errors",
                 "F", "nn", "np", "torch"):
        assert hasattr(layers, name), name


def test_dropin_shims_resolve():
    sys.path.insert(0, os.path.join(ROOT, "fusiondepth_b200", "dropin"))
    try:
        for m in ("layers", "networks"):
            sys.modules.pop(m, None)
        import layers as L
        import networks as N
        assert L.F.conv2d is torch.nn.functional.conv2d and L.nn is torch.nn
        assert N.ResnetEncoder.__module__.startswith("fusiondepth_b200")
        assert {"ResnetEncoder", "DepthDecoder", "PoseDecoder", "PoseCNN"} <= set(dir(N))
    finally:
        sys.path.pop(0)
        for m in ("layers", "networks"):
            sys.modules.pop(m, None)


def test_state_dict_contract():
    """key -> shape of every network variant the drivers build, against the keys recorded from the
    reference modules (tests/golden/state_dict_keys.json, written by make_golden.py)."""
    from fusiondepth_b200 import networks as N
    want = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
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
    for name, m in built.items():
        got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert got == want[name], name
    with __import__("pytest").raises(ValueError):
        N.ResnetEncoder(19, False)


def test_torch_composed_layers_match_oracle():
    """The functions of the surface that are plain tensor compositions (no kernel needed) on CPU; the
    modules with kernels behind them are checked on the GPU (tests/test_gpu_geometry.py)."""
    from fusiondepth_b200 import layers as L
    g = torch.Generator().manual_seed(0)
    B, H, W = 2, 16, 24
    depth = 0.5 + torch.rand(B, 1, H, W, generator=g) * 10
    aa, tt = 0.02 * torch.randn(B, 1, 3, generator=g), 0.1 * torch.randn(B, 1, 3, generator=g)
    for inv in (False, True):
        assert rel_err(L.transformation_from_parameters(aa, tt, inv), SO.pose_matrix(aa, tt, inv)) < 1e-6
    x = torch.rand(B, 3, H, W, generator=g)
    disp = torch.rand(B, 1, H, W, generator=g)
    assert rel_err(L.get_smooth_loss(disp, x), SO.smooth_loss(disp, x)) < 1e-6
    sd, d = L.disp_to_depth(disp, 0.1, 100.0)
    osd, od = SO.disp_to_depth(disp, 0.1, 100.0)
    assert torch.equal(sd, osd) and torch.equal(d, od)
    errs = L.compute_depth_errors(depth, depth * 1.1)
    assert abs(float(errs[0]) - 0.1) < 1e-5 and float(errs[4]) == 1.0
    # modules with kernels behind them refuse CPU tensors (no CPU fallback)
    from fusiondepth_b200 import _lib
    K = torch.eye(4).repeat(B, 1, 1)
    with __import__("pytest").raises(_lib.FusionDepthLibraryError):
        L.BackprojectDepth(B, H, W)(depth, K)
    with __import__("pytest").raises(_lib.FusionDepthLibraryError):
        L.SSIM()(x, x)
    # layers.F is torch.nn.functional for everything off the hot path
    assert L.F.relu is torch.nn.functional.relu and L.F.max_pool2d is torch.nn.functional.max_pool2d
    y = L.F.interpolate(disp, [2 * H, 2 * W], mode="bilinear", align_corners=False)      # CPU: torch's own
    assert torch.equal(y, torch.nn.functional.interpolate(disp, [2 * H, 2 * W], mode="bilinear", align_corners=False))


def test_pretrained_loads_local_imagenet_checkpoint(tmp_path, monkeypatch):
    """resnet_encoder.py:45-49, 71-87: ImageNet weights (here a stand-in state dict with torchvision's keys),
    conv1 tiled over the stacked frames and divided by their number; the 2/4/5/6-channel variants keep a fresh
    conv1.  No file -> a RuntimeError naming the directories (the package never downloads)."""
    import fusiondepth_b200.networks as N
    from fusiondepth_b200.networks.resnet_encoder import ResNetTrunk
    monkeypatch.setenv("FD_PRETRAINED_DIR", str(tmp_path))
    monkeypatch.setattr(torch.hub, "get_dir", lambda: str(tmp_path / "nohub"))
    try:
        N.ResnetEncoder(18, True)
        assert False, "expected RuntimeError"
    except RuntimeError as e:
        assert "resnet18-*.pth" in str(e) and str(tmp_path) in str(e)
    torch.manual_seed(3)
    src = ResNetTrunk(18, 3)
    for p in src.parameters():
        p.data.normal_()
    sd = {k: v.clone().contiguous() for k, v in src.state_dict().items()}
    torch.save(sd, tmp_path / "resnet18-0badc0de.pth")
    one = N.ResnetEncoder(18, True)
    pair = N.ResnetEncoder(18, True, num_input_images=2)
    beam = N.ResnetEncoder(18, True, beam_encoder=True)
    for enc in (one, pair, beam):
        got = enc.encoder.state_dict()
        assert set(got) == set(sd)
        for k in sd:
            if k != "conv1.weight":
                assert torch.equal(got[k], sd[k]), k
        for m in enc.encoder.modules():
            if isinstance(m, torch.nn.Conv2d):
                assert m.weight.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(one.encoder.conv1.weight, sd["conv1.weight"])
    assert torch.equal(pair.encoder.conv1.weight, torch.cat([sd["conv1.weight"]] * 2, 1) / 2)
    assert tuple(beam.encoder.conv1.weight.shape) == (64, 2, 7, 7)
