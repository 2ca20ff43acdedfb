"""Imports the *unmodified* reference from /root/reference on CPU (build container only).

Used by tests/make_golden.py (fixture generation) and tests/test_oracle_vs_reference.py.
Nothing here runs on the GPU box: /root/reference does not exist there.

Shims: stub modules tensorboardX / skimage; WANDB_MODE=disabled;
argv fixed before `import trainer` (options are parsed at import, trainer.py:751-754);
Trainer built with object.__new__ because __init__ needs CUDA, W&B and KITTI files;
torch.Tensor.cuda -> identity because compute_losses hard-codes .cuda() (trainer.py:541,551).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(REF)


_loaded = {}


def load():
    """Returns a namespace with the reference's layers, networks, trainer modules."""
    if _loaded:
        return _loaded["ns"]
    os.environ.setdefault("WANDB_MODE", "disabled")
    for name in ("tensorboardX", "skimage", "skimage.transform"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "tensorboardX":
                m.SummaryWriter = object
            sys.modules[name] = m
    if not hasattr(np, "int"):
        np.int = int                      # kitti_utils.py:80 uses the removed alias
    sys.path.insert(0, REF)
    argv = sys.argv
    sys.argv = ["trainer.py", "--num_layers", "18", "--weights_init", "scratch"]
    try:
        import layers as ref_layers
        import networks as ref_networks
        import kitti_utils as ref_kitti
        import trainer as ref_trainer
    finally:
        sys.argv = argv
    torch.Tensor.cuda = lambda self, *a, **k: self
    src = open(os.path.join(REF, "gen2channel.py")).read().split("\n")
    env = {"torch": torch}
    exec("\n".join(src[59:117]), env)     # get_4beam_2channel only (gen2channel.py:60-117)
    ns = types.SimpleNamespace(layers=ref_layers, networks=ref_networks, kitti_utils=ref_kitti,
                               trainer=ref_trainer, get_4beam_2channel=env["get_4beam_2channel"])
    _loaded["ns"] = ns
    return ns


def make_models(ns, num_layers=18):
    """The six networks of Trainer.__init__ (trainer.py:66-115), default flags."""
    N = ns.networks
    m = {}
    m["encoder"] = N.ResnetEncoder(num_layers, False)
    m["beam_encoder"] = N.ResnetEncoder(num_layers, False, beam_encoder=True)
    m["beam_encoder_pose"] = N.ResnetEncoder(num_layers, False, num_input_images=2, beam_encoder=True)
    m["depth"] = N.DepthDecoder(m["encoder"].num_ch_enc, [0, 1, 2, 3])
    m["pose_encoder"] = N.ResnetEncoder(num_layers, False, num_input_images=2)
    m["pose"] = N.PoseDecoder(m["pose_encoder"].num_ch_enc, num_input_features=1,
                              num_frames_to_predict_for=2)
    return m


def make_trainer(ns, models, B, H, W):
    """object.__new__(Trainer) wired by hand (trainer.py:24-205 minus CUDA/W&B/KITTI)."""
    T = ns.trainer
    tr = object.__new__(T.Trainer)
    import copy
    opt = copy.deepcopy(T.opts)
    opt.height, opt.width, opt.batch_size = H, W, B
    tr.opt = opt
    tr.device = torch.device("cpu")
    tr.batch_size = B
    tr.num_scales = len(opt.scales)
    tr.num_input_frames = len(opt.frame_ids)
    tr.num_pose_frames = 2
    tr.use_pose_net = True
    tr.models = models
    tr.ssim = ns.layers.SSIM()
    tr.backproject_depth, tr.project_3d = {}, {}
    for s in opt.scales:
        tr.backproject_depth[s] = ns.layers.BackprojectDepth(B, H >> s, W >> s)
        tr.project_3d[s] = ns.layers.Project3D(B, H >> s, W >> s)
    return tr


class FixedNoise:
    """Context manager: makes torch.randn return the supplied tensors in call order so the
    reference's trainer.py:551 draws exactly the noise the oracle is given."""

    def __init__(self, tensors):
        self.tensors = list(tensors)

    def __enter__(self):
        self._orig = torch.randn
        it = iter(self.tensors)

        def fake(*a, **k):
            return next(it).clone()

        torch.randn = fake
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig
