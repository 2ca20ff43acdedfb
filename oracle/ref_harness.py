"""Imports and drives the UNMODIFIED reference (trainer.py / refiner.py / layers.py / networks/*).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/make_ref.py for where the sources come from: /root/reference
in the build container, the staged byte-for-byte copy oracle/_ref on the GPU box).  Used by
tests/make_golden.py (fixture generation), the drop-in tests, and bench.py's `--impl reference` /
`--impl pytorch-gpu` arms.  Nothing under fusiondepth_b200/ imports this.

Shims -- none of them edits a reference file:
  * stub modules tensorboardX / skimage (logging / dataset-only imports); WANDB_MODE=disabled;
  * argv fixed before `import trainer` / `import refiner` (options are parsed at import time);
  * Trainer / Refiner built with object.__new__ and wired by hand: __init__ needs W&B, KITTI files and (for
    the refiner) stage-1 checkpoints on disk;
  * CPU runs only: torch.Tensor.cuda -> identity, because compute_losses hard-codes .cuda().

`load(dropin=True)` resolves `import layers` / `import networks` to fusiondepth_b200/dropin (this repo's
CUDA modules) while trainer.py / refiner.py stay the reference's unchanged files.
"""
from __future__ import annotations

import copy
import importlib
import os
import sys
import types

import numpy as np
import torch

from oracle import make_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "fusiondepth_b200", "dropin")
_REF_MODULES = ("layers", "networks", "trainer", "refiner", "utils", "kitti_utils", "datasets", "options")


def available() -> bool:
    return make_ref.root() is not None


def _purge():
    for name in list(sys.modules):
        if name in _REF_MODULES or name.split(".")[0] in ("networks", "datasets"):
            del sys.modules[name]


_cache = {}


def load(dropin: bool = False, cpu_shim: bool = True, with_refiner: bool = False, with_completor: bool = False):
    """Namespace with the reference's modules.  `dropin`: layers / networks come from this repo."""
    key = (dropin, with_refiner, with_completor)
    if key in _cache:
        return _cache[key]
    ref = make_ref.root()
    assert ref is not None, "reference sources not found (neither /root/reference nor oracle/_ref)"
    os.environ.setdefault("WANDB_MODE", "disabled")
    for name in ("tensorboardX", "skimage", "skimage.transform"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "tensorboardX":
                m.SummaryWriter = object
            sys.modules[name] = m
    if not hasattr(np, "int"):
        np.int = int                      # kitti_utils.py:80 uses the removed alias
    _purge()
    paths = ([DROPIN] if dropin else []) + [ref]
    old_path, argv = list(sys.path), sys.argv
    sys.path[:0] = paths
    sys.argv = ["trainer.py", "--num_layers", "18", "--weights_init", "scratch"]
    try:
        ns = types.SimpleNamespace(root=ref, dropin=dropin)
        ns.layers = importlib.import_module("layers")
        ns.networks = importlib.import_module("networks")
        ns.kitti_utils = importlib.import_module("kitti_utils")
        ns.trainer = importlib.import_module("trainer")
        if with_refiner:
            ns.refiner = importlib.import_module("refiner")
        if with_completor:
            ns.completor = importlib.import_module("completor")
    finally:
        sys.argv = argv
        sys.path[:] = old_path
        _purge()                          # the namespace keeps the module objects alive
    if cpu_shim and not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    src = open(os.path.join(ref, "gen2channel.py")).read().split("\n")
    env = {"torch": torch}
    exec("\n".join(src[59:117]), env)     # get_4beam_2channel only (gen2channel.py:60-117)
    ns.get_4beam_2channel = env["get_4beam_2channel"]
    _cache[key] = ns
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


def make_refine_decoder(ns, num_ch_enc):
    """refiner.py:146-158: the stage-2 decoder (refine_2d, road, catxy='true', refine2d_deep='true')."""
    return ns.networks.DepthDecoder(num_ch_enc, [0, 1, 2, 3], road=True, catxy=True, deep=True)


def _wire(obj, opt, models, B, H, W, device, ns):
    opt.height, opt.width = H, W
    obj.opt = opt
    obj.device = torch.device(device)
    obj.batch_size = B
    obj.num_scales = len(opt.scales)
    obj.num_input_frames = len(opt.frame_ids)
    obj.num_pose_frames = 2
    obj.use_pose_net = True
    obj.models = models
    obj.ssim = ns.layers.SSIM().to(device)
    obj.backproject_depth, obj.project_3d = {}, {}
    for s in opt.scales:
        obj.backproject_depth[s] = ns.layers.BackprojectDepth(B, H >> s, W >> s).to(device)
        obj.project_3d[s] = ns.layers.Project3D(B, H >> s, W >> s).to(device)


def make_trainer(ns, models, B, H, W, device="cpu", batch_size_flag=None):
    """object.__new__(Trainer) wired by hand (trainer.py:24-205 minus W&B / KITTI files)."""
    T = ns.trainer
    tr = object.__new__(T.Trainer)
    opt = copy.deepcopy(T.opts)
    opt.batch_size = batch_size_flag if batch_size_flag is not None else B
    _wire(tr, opt, models, B, H, W, device, ns)
    return tr


def make_refiner(ns, models, B, H, W, device="cpu"):
    """object.__new__(Refiner) wired by hand (refiner.py:24-258 minus W&B / KITTI / checkpoint files).
    `models` = the six stage-1 networks + "refine2d_decoder"."""
    R = ns.refiner
    rf = object.__new__(R.Refiner)
    opt = copy.deepcopy(R.opts)
    opt.batch_size = B
    opt.clone_gdc, opt.refine_2d = True, True           # refiner.py:29-30
    _wire(rf, opt, models, B, H, W, device, ns)
    rf.eval_scales = opt.scales                          # refiner.py:46-47
    rf.catxy = {}
    for s in opt.scales:                                 # refiner.py:217-227
        for val in ("False", "True"):
            rf.catxy[val, s] = ns.layers.Cat_xy(B, H >> s, W >> s).to(device)
    return rf


def make_completor(ns, models, B, H, W, device="cpu"):
    """object.__new__(Completor) wired by hand (completor.py:28-186 minus W&B / KITTI files and the forced
    1216x352 size, completor.py:31-34: the drop-in test runs it at a size the CPU reference finishes in seconds).
    Default flags: beam encoder, separate_resnet pose network, automasking, completion_siloss on scale 0."""
    C = ns.completor
    cp = object.__new__(C.Completor)
    opt = copy.deepcopy(C.opts)
    opt.batch_size = B
    _wire(cp, opt, models, B, H, W, device, ns)
    return cp


class FixedNoise:
    """Context manager: makes torch.randn return the supplied tensors in call order so the
    reference's trainer.py:551 draws exactly the noise the oracle is given.  `cpu=True` (a CPU run of the
    reference on a machine that has a GPU): Tensor.cuda() is the identity inside the context, because
    compute_losses hard-codes `.cuda()` on that noise (trainer.py:551-552, refiner.py:652-653)."""

    def __init__(self, tensors, cpu: bool = False):
        self.tensors = list(tensors)
        self.cpu = cpu

    def __enter__(self):
        self._orig = torch.randn
        it = iter(self.tensors)

        def fake(*a, **k):
            return next(it).clone()

        torch.randn = fake
        if self.cpu:
            self._cuda = torch.Tensor.cuda
            torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig
        if self.cpu:
            torch.Tensor.cuda = self._cuda
