"""The per-step training hot path: six networks + fused photometric loss + Adam (+ DP all-reduce).

Mirrors the call sequence of ``Trainer.process_batch`` / ``predict_poses`` /
``generate_images_pred`` / ``compute_losses`` (reference trainer.py:268-596) with the reference's
default flags (separate_resnet pose net on image pairs, beam encoders on, automasking, SSIM,
si-loss on all scales), and of ``Trainer.run_epoch``'s accumulate-then-Adam loop
(trainer.py:237-248).

Two ways in:
  * ``patch_trainer(trainer)`` -- keeps the reference's unchanged ``Trainer`` object and swaps
    its ``generate_images_pred`` + ``compute_losses`` for the fused kernel pair;
  * ``TrainStep`` -- the whole optimiser step (micro-batches, backward, gradient all-reduce, Adam
    on flat buffers), optionally captured in one CUDA graph.  This is what bench.py times.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import networks, ops
from .layers import transformation_from_parameters

MODEL_NAMES = ("encoder", "beam_encoder", "beam_encoder_pose", "depth", "pose_encoder", "pose")


def build_models(num_layers: int = 18, device="cuda", scales=(0, 1, 2, 3)) -> Dict[str, nn.Module]:
    """The six networks of Trainer.__init__ (reference trainer.py:66-115), weights_init=scratch."""
    m: Dict[str, nn.Module] = {}
    m["encoder"] = networks.ResnetEncoder(num_layers, False)
    m["beam_encoder"] = networks.ResnetEncoder(num_layers, False, beam_encoder=True)
    m["beam_encoder_pose"] = networks.ResnetEncoder(num_layers, False, num_input_images=2,
                                                    beam_encoder=True)
    m["depth"] = networks.DepthDecoder(m["encoder"].num_ch_enc, list(scales))
    m["pose_encoder"] = networks.ResnetEncoder(num_layers, False, num_input_images=2)
    m["pose"] = networks.PoseDecoder(m["pose_encoder"].num_ch_enc, num_input_features=1,
                                     num_frames_to_predict_for=2)
    for k in m:
        m[k].to(device)
    return m


def predict_poses(models, inputs, frame_ids=(0, -1, 1)):
    """trainer.py:321-388, num_pose_frames == 2, separate_resnet, beam encoder on."""
    outputs = {}
    for f_i in frame_ids[1:]:
        pair = (f_i, 0) if f_i < 0 else (0, f_i)          # temporal order (trainer.py:339-346)
        img = torch.cat([inputs[("color_aug", i, 0)] for i in pair], 1)
        two = torch.cat([inputs[("2channel", i, 0)] for i in pair], 1)
        pf = [models["pose_encoder"](img)]
        bf = [models["beam_encoder_pose"](two)]
        axisangle, translation = models["pose"](pf, beam_inputs=bf)
        outputs[("axisangle", 0, f_i)] = axisangle
        outputs[("translation", 0, f_i)] = translation
        outputs[("cam_T_cam", 0, f_i)] = transformation_from_parameters(
            axisangle[:, 0], translation[:, 0], invert=(f_i < 0))
    return outputs


def loss_static(inputs, noise, frame_ids=(0, -1, 1), si_target="4beam"):
    """Non-differentiable operands of the fused loss, gathered from the loader's dict.  `si_target`: the
    map the scale-invariant term compares against ("4beam" in trainer.py:577-589, "inf_gdc" in
    refiner.py:678-688)."""
    color = {}
    for s in range(4):
        color[(0, s)] = inputs[("color", 0, s)]
    for f in frame_ids[1:]:
        color[(f, 0)] = inputs[("color", f, 0)]
    return {"color": color, "K": inputs[("K", 0)], "inv_K": inputs[("inv_K", 0)],
            "beam": inputs[si_target], "noise": noise}


def fused_losses(inputs, outputs, noise, opts: Optional[Dict] = None, materialize: bool = False,
                 frame_ids=(0, -1, 1), si_target="4beam"):
    """generate_images_pred + compute_losses (trainer.py:425-596) in two kernel launches.
    Returns the reference's ``losses`` dict; with ``materialize`` the per-scale
    ("depth",0,s), ("color",f,s), identity_selection/s tensors are written into ``outputs``."""
    outs = {} if materialize else None
    disps = [outputs[("disp", s)] for s in range(4)]
    lv = ops.photoloss(disps, outputs[("cam_T_cam", 0, frame_ids[1])],
                       outputs[("cam_T_cam", 0, frame_ids[2])],
                       loss_static(inputs, noise, frame_ids, si_target), opts, outs)
    losses = {name: lv[i] for i, name in enumerate(ops.LOSS_NAMES)}
    if materialize:
        sel = outs.pop("sel")
        for s in range(4):
            outputs["identity_selection/%d" % s] = (sel[s] > 1).float()
        outputs.update(outs)
        for f in frame_ids[1:]:
            for s in range(4):
                outputs[("color_identity", f, s)] = inputs[("color", f, 0)]
    return losses


def draw_noise(B, H, W, device):
    """The tie-break noise exactly as the reference draws it: CPU generator, one [B,2,H,W] randn
    per scale, then copied to the device (trainer.py:551-552)."""
    return {s: torch.randn(B, 2, H, W).to(device) for s in range(4)}


class TrunkStreams:
    """Side streams for the six independent ResNet trunks of one micro-batch.  The trunks share
    no data until the decoder / pose heads, so they are issued on separate CUDA streams (forked
    from and joined back to the current stream with events); inside a captured CUDA graph this
    becomes six parallel branches, which keeps all 148 SMs busy through the small layer3/layer4
    GEMMs.  Autograd replays each branch's backward on the stream its forward ran on."""

    def __init__(self, device, n: int = 5):
        self.side = [torch.cuda.Stream(device=device) for _ in range(n)]

    def fork(self):
        cur = torch.cuda.current_stream()
        for s in self.side:
            s.wait_stream(cur)

    def join(self, idx):
        cur = torch.cuda.current_stream()
        for i in idx:
            cur.wait_stream(self.side[i])


def process_batch(models, inputs, noise=None, opts: Optional[Dict] = None, materialize=False,
                  frame_ids=(0, -1, 1), streams: Optional[TrunkStreams] = None):
    """Trainer.process_batch (trainer.py:268-319), default flags."""
    if streams is None:
        feats = models["encoder"](inputs[("color_aug", 0, 0)])
        beam = models["beam_encoder"](inputs["2channel"])
        outputs = dict(models["depth"](feats, beam_features=beam))
        outputs.update(predict_poses(models, inputs, frame_ids))
    else:
        # same call sequence, trunks on parallel streams
        streams.fork()
        side = streams.side
        with torch.cuda.stream(side[0]):
            beam = models["beam_encoder"](inputs["2channel"])
        pose_feats, pose_beam = {}, {}
        for n, f_i in enumerate(frame_ids[1:]):
            pair = (f_i, 0) if f_i < 0 else (0, f_i)
            with torch.cuda.stream(side[1 + 2 * n]):
                img = torch.cat([inputs[("color_aug", i, 0)] for i in pair], 1)
                pose_feats[f_i] = [models["pose_encoder"](img)]
            with torch.cuda.stream(side[2 + 2 * n]):
                two = torch.cat([inputs[("2channel", i, 0)] for i in pair], 1)
                pose_beam[f_i] = [models["beam_encoder_pose"](two)]
        feats = models["encoder"](inputs[("color_aug", 0, 0)])
        streams.join([0])
        outputs = dict(models["depth"](feats, beam_features=beam))
        for n, f_i in enumerate(frame_ids[1:]):
            sp, sb = side[1 + 2 * n], side[2 + 2 * n]
            sp.wait_stream(sb)
            with torch.cuda.stream(sp):
                axisangle, translation = models["pose"](pose_feats[f_i], beam_inputs=pose_beam[f_i])
                outputs[("axisangle", 0, f_i)] = axisangle
                outputs[("translation", 0, f_i)] = translation
                outputs[("cam_T_cam", 0, f_i)] = transformation_from_parameters(
                    axisangle[:, 0], translation[:, 0], invert=(f_i < 0))
        streams.join([1, 2, 3, 4])
    if noise is None:
        B, _, H, W = inputs[("color", 0, 0)].shape
        noise = draw_noise(B, H, W, inputs[("color", 0, 0)].device)
    losses = fused_losses(inputs, outputs, noise, opts, materialize, frame_ids)
    return outputs, losses


def _default_loss_config(opt, who):
    if (opt.v1_multiscale or opt.disable_automasking or opt.avg_reprojection or opt.no_ssim
            or opt.predictive_mask or opt.use_stereo or opt.pose_model_type == "posecnn"
            or list(opt.scales) != [0, 1, 2, 3] or list(opt.frame_ids) != [0, -1, 1]):
        raise NotImplementedError("%s covers the reference's default loss configuration" % who)


def patch_trainer(trainer, materialize: bool = True):
    """Swap the reference Trainer's warp + loss methods for the fused kernels, in place."""
    opt = trainer.opt
    _default_loss_config(opt, "patch_trainer")
    opts = {"min_depth": opt.min_depth, "max_depth": opt.max_depth,
            "smoothness": opt.disparity_smoothness, "si_thresh": opt.gdc_loss_threshold,
            "si_var": opt.si_var, "use_si": opt.trainer_siloss == "true",
            # trainer.py:578: every scale with --trainer_siloss_all_scale (default on), else scale 0 only
            "si_scales": 0xF if opt.trainer_siloss_all_scale else 0x1}
    return _patch_loss_methods(trainer, opt, opts, materialize)


def patch_completor(completor, materialize: bool = True):
    """The same for the completion driver (completor.py:426-474, 546-728): its loss code is a copy of the
    trainer's with its own flags -- the si-loss (x26 depth, 4-beam target, 0.1 sqrt(mean d^2 - si_var mean(d)^2),
    completor.py:694-715) on scale 0, or on every scale with --completion_siloss_all_scale true."""
    opt = completor.opt
    _default_loss_config(opt, "patch_completor")
    if opt.completion_l1loss and not opt.completion_siloss:
        raise NotImplementedError("patch_completor: --completion_l1loss is not part of the fused loss")
    opts = {"min_depth": opt.min_depth, "max_depth": opt.max_depth,
            "smoothness": opt.disparity_smoothness, "si_thresh": opt.gdc_loss_threshold,
            "si_var": opt.si_var, "use_si": bool(opt.completion_siloss),
            "si_scales": 0xF if opt.completion_siloss_all_scale == "true" else 0x1}
    return _patch_loss_methods(completor, opt, opts, materialize)


def _patch_loss_methods(trainer, opt, opts, materialize):

    def generate_images_pred(inputs, outputs, frame_ids):
        if len(frame_ids) > 1:
            return None                               # done inside the fused loss launch
        # validation (Trainer.val -> process_batch(val=True), trainer.py:305-307): no source frames and no
        # compute_losses afterwards, but compute_depth_losses reads ("depth", 0, s) (trainer.py:604)
        H, W = opt.height, opt.width
        for s in opt.scales:
            up = ops.upsample_bilinear(outputs[("disp", s)], H, W)
            outputs[("depth", 0, s)] = 1.0 / (1.0 / opt.max_depth + (1.0 / opt.min_depth - 1.0 / opt.max_depth) * up)
        return None

    def compute_losses(inputs, outputs):
        B, _, H, W = inputs[("color", 0, 0)].shape
        return fused_losses(inputs, outputs, draw_noise(B, H, W, inputs[("color", 0, 0)].device),
                            opts, materialize)

    trainer.generate_images_pred = generate_images_pred
    trainer.compute_losses = compute_losses
    return trainer


# ------------------------------------------------------------------------------------------------
class FlatParams:
    """All trainable parameters (and their gradients) as views of two flat fp32 buffers, so that
    the data-parallel gradient exchange is one all-reduce and Adam is one kernel.  Conv weights
    keep their channels-last storage order inside the buffer."""

    ALIGN = 64      # floats: every parameter starts on a 256 B boundary (float4 / TMA loads)

    def __init__(self, models: Dict[str, nn.Module], order=None):
        """`order(name, param_name) -> sortable`: optional bucket key; parameters are laid out in ascending
        key order (the gradient exchange sends contiguous bucket slices).  Parameters with requires_grad
        False (the refiner's frozen stage-1 networks) come last: they share the weight buffer (and its
        derived W_lo / W^T copies) but have no gradient slot and are not touched by Adam."""
        named = []
        for name in models:
            for k, p in models[name].named_parameters():
                named.append(((0 if p.requires_grad else 1, order(name, k) if order else 0, len(named)), name, k, p))
        named.sort(key=lambda t: t[0])
        self.names = [(name, k) for _, name, k, _ in named]
        self.params = [p for _, _, _, p in named]
        self.keys = [key[1] for key, _, _, _ in named]
        A = self.ALIGN
        pad = lambda k: (k + A - 1) // A * A
        n = sum(pad(p.numel()) for p in self.params)
        self.n_train = sum(pad(p.numel()) for p in self.params if p.requires_grad)
        dev = self.params[0].device
        self.data = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(self.n_train, device=dev, dtype=torch.float32)
        self.offsets = {}
        off = 0
        for p in self.params:
            k = p.numel()
            if p.dim() == 4 and not p.is_contiguous(memory_format=torch.channels_last):
                # the kernels read/write conv weights as [Cout,KH,KW,Cin]
                p.data = p.data.contiguous(memory_format=torch.channels_last)
            dv = self._view(self.data[off:off + k], p)
            dv.copy_(p.data)
            p.data = dv
            if p.requires_grad:
                gv = self._view(self.grad[off:off + k], p)
                p.grad = gv
                p._fd_grad = gv          # ops.DIRECT_GRAD: backward kernels add into this view
            self.offsets[p] = off
            off += pad(k)
        self.numel = n
        # derived conv-weight tensors (W_lo, W^T, W^T_lo) live in flat buffers with the same layout and
        # are refreshed by two launches per optimiser step (prepare_weights)
        self.lo = None
        self.t = None
        self.t_lo = None
        self.t_desc = None
        self.t_total = 0

    def prepare_weights(self, cache: Dict):
        """W_lo = W - tf32(W) for the whole buffer and the transposed copies (data-gradient operand) of
        every tensor-core-eligible conv weight: two launches, then `cache` (ops.WEIGHT_CACHE) entries
        keyed like ops.Conv2dFn's per-tensor path."""
        from . import _lib
        from ctypes import c_void_p
        lib = _lib.load()
        st = c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.lo is None:
            self.lo = torch.empty_like(self.data)
            rows, start = [], 0
            self.t_params = []
            for p in self.params:
                if p.dim() == 4 and p.shape[0] % 32 == 0 and p.shape[1] % 16 == 0:
                    o, i, kh, kw = p.shape
                    rows.append([start, self.offsets[p], o, kh * kw, i])
                    start += p.numel()
                    self.t_params.append(p)
            self.t_total = start
            if rows:
                self.t = torch.empty_like(self.data)
                self.t_lo = torch.empty_like(self.data)
                self.t_desc = torch.tensor(rows, dtype=torch.int64, device=self.data.device)
        _lib.check(lib.fd_tf32_split(c_void_p(self.data.data_ptr()), c_void_p(self.lo.data_ptr()), self.numel, st),
                   "fd_tf32_split")
        if self.t_desc is not None:
            _lib.check(lib.fd_weight_transpose_split_batched(
                c_void_p(self.data.data_ptr()), c_void_p(self.t.data_ptr()), c_void_p(self.t_lo.data_ptr()),
                c_void_p(self.t_desc.data_ptr()), self.t_desc.shape[0], self.t_total, st),
                "fd_weight_transpose_split_batched")
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        for p in self.params:
            if p.dim() == 4:
                off, k = self.offsets[p], p.numel()
                cache[(p.data_ptr(), "lo")] = ((self.lo[off:off + k],), ev, cur)
        for p in getattr(self, "t_params", []):
            off, k = self.offsets[p], p.numel()
            cache[(p.data_ptr(), "t")] = ((self.t[off:off + k], self.t_lo[off:off + k]), ev, cur)

    @staticmethod
    def _view(flat, p):
        if p.dim() == 4:
            o, i, kh, kw = p.shape
            return flat.view(o, kh, kw, i).permute(0, 3, 1, 2)
        return flat.view(p.shape)

    def bucket_slices(self):
        """[(key, start, end)] over the trainable part of the buffer, one entry per distinct order key."""
        out, start, cur = [], 0, None
        A = self.ALIGN
        off = 0
        for p, key in zip(self.params, self.keys):
            if not p.requires_grad:
                break
            if cur is None:
                cur = key
            if key != cur:
                out.append((cur, start, off))
                start, cur = off, key
            off += (p.numel() + A - 1) // A * A
        if cur is not None:
            out.append((cur, start, off))
        return out

    def zero_grad(self):
        self.grad.zero_()


def reduce_gradients(flat: "FlatParams", world: int, group=None, lo: int = 0, hi: Optional[int] = None):
    """The data-parallel exchange step: a sum all-reduce of (a bucket slice of) the flat gradient buffer
    (NCCL over NVLink on GPUs; the 1/world average is folded into the Adam kernel's grad_scale).  Samples
    are independent through the whole step and BatchNorm statistics stay per rank, as in the
    single-process reference, so this is the only collective on the path."""
    if world > 1:
        buf = flat.grad if (lo == 0 and hi is None) else flat.grad[lo:hi]
        torch.distributed.all_reduce(buf, op=torch.distributed.ReduceOp.SUM, group=group)
    return flat.grad


# Gradient buckets in the order their gradients complete during the backward pass:
# 0 = both decoders' heads + every trunk's layer4 (72 % of a ResNet-18's parameters, finished first),
# 1 = layer3, 2 = the rest of the trunks.  The never-used ImageNet `fc` heads (resnet_encoder.py:74,
# grad=None in the reference) are placed last and excluded from the exchange.
N_BUCKETS = 3


def bucket_of(model_name: str, param_name: str) -> int:
    if ".fc." in param_name:
        return N_BUCKETS
    if model_name in ("depth", "pose", "refine2d_decoder") or ".layer4." in param_name:
        return 0
    if ".layer3." in param_name:
        return 1
    return 2


class TrainStep:
    """One optimiser step of trainer.py:237-248: ``accumulate`` micro-batches, each
    process_batch -> (loss/accumulate).backward(), then Adam.  With world_size > 1 the flat
    gradient buffer is averaged across ranks with one NCCL all-reduce before Adam."""

    def __init__(self, models, lr: float = 1e-4, accumulate: int = 1, opts: Optional[Dict] = None,
                 process_group=None, parallel_trunks: bool = True, direct_grad: bool = True,
                 cache_weight_prep: bool = True, concurrent_microbatches: bool = True,
                 bucketed_allreduce: bool = True):
        self.models = models
        self.accumulate = accumulate
        self.opts = opts
        self.flat = FlatParams(models, order=bucket_of)
        dev = self.flat.data.device
        nt = self.flat.n_train
        self.exp_avg = torch.zeros(nt, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(nt, device=dev, dtype=torch.float32)
        # word 0: step counter, words 1-2: bias-correction scalars, word 3: learning rate (float bits) --
        # all on the device so that a captured graph follows set_lr() / the step count
        self.adam_state = torch.zeros(4, device=dev, dtype=torch.int32)
        self.set_lr(lr)
        self.lr0 = float(lr)
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        if self.world > 1:
            # replicas must start from the same weights and BatchNorm buffers (rank 0's)
            torch.distributed.broadcast(self.flat.data, 0, group=process_group)
            for m in models.values():
                for b in m.buffers():
                    torch.distributed.broadcast(b, 0, group=process_group)
        self.buckets = [(k, lo, hi) for k, lo, hi in self.flat.bucket_slices() if k < N_BUCKETS]
        self.bucketed = bool(bucketed_allreduce) and self.world > 1 and len(self.buckets) > 1 \
            and os.environ.get("FD_BUCKETED_ALLREDUCE", "1") != "0"
        self.comm_stream = torch.cuda.Stream(device=dev) if self.bucketed else None
        self._ready = None
        # Opt-in experiment (FD_WGRAD_STREAM=1): weight-gradient kernels on side streams, off the backward pass's
        # dependency chain (ops.WGRAD_STREAMS).  Measured SLOWER on B200 (359 vs 385 images/s): the step is bound by
        # SM time, not by dependencies, and the extra concurrency only makes the tensor-core kernels share SMs.
        self.wgrad_streams = {} if (direct_grad and os.environ.get("FD_WGRAD_STREAM", "0") == "1") else None
        self.direct_grad, self.cache_weight_prep = direct_grad, cache_weight_prep
        # The micro-batches of a step are independent given the weights (gradients accumulate with
        # atomics, BatchNorm running statistics through ops.BNSchedule), so each gets its own stream
        # and trunk streams: while one micro-batch walks its serial decoder -> loss -> decoder
        # backward chain, the other one's trunks keep the SMs busy.
        self.concurrent = (bool(concurrent_microbatches) and accumulate > 1
                           and os.environ.get("FD_CONCURRENT_MB", "1") != "0")
        nsets = accumulate if self.concurrent else 1
        self.trunks = [TrunkStreams(dev) if parallel_trunks else None for _ in range(nsets)]
        self.mb_streams = [torch.cuda.Stream(device=dev) for _ in range(nsets)] if self.concurrent else []
        self.streams = self.trunks[0]
        self.bn_schedule = self._make_bn_schedule() if self.concurrent or parallel_trunks else None
        self.stats_arena = None
        self.stream = torch.cuda.Stream(device=dev)      # warm-up and capture share one stream
        self.graph = None
        self.static_inputs = None
        self.static_noise = None
        self.loss_out = None
        for m in models.values():
            m.train()

    def set_lr(self, lr: float):
        """Learning rate of the following steps (StepLR: trainer.py:131-132, 266).  Lives on the device, so
        it also reaches a captured graph's replays."""
        self.lr = float(lr)
        self.adam_state[3:4].copy_(torch.tensor([self.lr], dtype=torch.float32).view(torch.int32))

    def step_lr_schedule(self, epoch: int, step_size: int, gamma: float = 0.1, base_lr: Optional[float] = None):
        """optim.lr_scheduler.StepLR(step_size, gamma) evaluated at `epoch` (trainer.py:131-132)."""
        base = self.lr0 if base_lr is None else base_lr
        self.set_lr(base * gamma ** (epoch // step_size))

    def _grad_ready(self, bucket: int):
        """Backward hook (ops.GRAD_READY): the gradients of `bucket` of this trunk call are on the stream."""
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        self._ready[bucket].append(ev)
        side = self.wgrad_streams.get(cur) if self.wgrad_streams is not None else None
        if side is not None and side in ops.WGRAD_USED:
            ev2 = torch.cuda.Event()
            ev2.record(side)                              # the bucket's weight-gradient kernels run there
            self._ready[bucket].append(ev2)

    def _make_bn_schedule(self):
        """Calls per optimiser step of every BatchNorm layer: the pose trunks see both image pairs
        of a micro-batch (trainer.py:339-365), every other trunk one batch."""
        from .networks.resnet_encoder import BatchNorm2d
        calls = {}
        for name, m in self.models.items():
            per_mb = 2 if name in ("pose_encoder", "beam_encoder_pose") else 1
            for mod in m.modules():
                if isinstance(mod, BatchNorm2d) and mod.track_running_stats and mod.momentum is not None:
                    calls[mod] = per_mb * self.accumulate
        return ops.BNSchedule(calls)

    # -- eager -------------------------------------------------------------------------------
    def _run(self, batches: Sequence[Dict], noises: Sequence[Dict]):
        self.flat.zero_grad()
        ops.DIRECT_GRAD = self.direct_grad
        ops.WEIGHT_CACHE = {} if self.cache_weight_prep else None
        ops.BN_SCHEDULE = self.bn_schedule
        ops.WGRAD_STREAMS = self.wgrad_streams
        if self.stats_arena is None:
            self.stats_arena = ops.StatsArena(self.flat.data.device)
        ops.STATS_ARENA = self.stats_arena
        self.stats_arena.begin_step()                 # one memset for every conv + BN pair of the step
        if self.bucketed:
            self._ready = [[] for _ in range(N_BUCKETS)]
            ops.GRAD_READY = self._grad_ready
        try:
            if self.bn_schedule is not None:
                self.bn_schedule.begin_step()
            if self.cache_weight_prep and ops.CONV_BACKEND == "tc":
                self.flat.prepare_weights(ops.WEIGHT_CACHE)
            out = self._run_inner(batches, noises)
            if self.bn_schedule is not None:
                self.bn_schedule.end_step()
            return out
        finally:
            ops.DIRECT_GRAD = False
            ops.WEIGHT_CACHE = None
            ops.BN_SCHEDULE = None
            ops.STATS_ARENA = None
            ops.GRAD_READY = None
            ops.WGRAD_STREAMS = None

    def _micro_batch(self, inputs, noise, trunks):
        _, losses = process_batch(self.models, inputs, noise, self.opts, streams=trunks)
        loss = losses["loss"] / self.accumulate
        loss.backward()
        if trunks is not None:
            # direct gradient accumulation bypasses AccumulateGrad, so autograd has no leaf
            # streams to join: wait for every trunk stream's backward kernels explicitly
            trunks.join(range(len(trunks.side)))
        return loss.detach()

    def _run_inner(self, batches: Sequence[Dict], noises: Sequence[Dict]):
        total = None
        if self.concurrent:
            cur = torch.cuda.current_stream()
            parts = []
            for i, (inputs, noise) in enumerate(zip(batches, noises)):
                ms = self.mb_streams[i]
                ms.wait_stream(cur)
                with torch.cuda.stream(ms):
                    parts.append(self._micro_batch(inputs, noise, self.trunks[i]))
            for ms in self.mb_streams:
                cur.wait_stream(ms)
            for part in parts:
                total = part if total is None else total + part
        else:
            for inputs, noise in zip(batches, noises):
                part = self._micro_batch(inputs, noise, self.trunks[0])
                total = part if total is None else total + part
        self._exchange_and_update()
        return total

    def _exchange_and_update(self):
        """Gradient exchange + Adam.  Bucketed: the all-reduce of a bucket is issued on a communication
        stream that waits only for that bucket's backward kernels (events recorded by ops.GRAD_READY hooks
        inside every trunk), so inside the captured graph it runs under the rest of the backward pass;
        only the last bucket (layers 1-2 and the stems) is exposed."""
        nt = self.flat.n_train
        ops.join_wgrad_streams()
        if self.world > 1:
            end = self.buckets[-1][2] if self.buckets else nt       # the unused fc heads are not exchanged
            if self.bucketed:
                cur = torch.cuda.current_stream()
                cs = self.comm_stream
                for k, lo, hi in self.buckets[:-1]:
                    for ev in self._ready[k]:
                        cs.wait_event(ev)
                    with torch.cuda.stream(cs):
                        reduce_gradients(self.flat, self.world, self.pg, lo, hi)
                cs.wait_stream(cur)
                with torch.cuda.stream(cs):
                    k, lo, hi = self.buckets[-1]
                    reduce_gradients(self.flat, self.world, self.pg, lo, hi)
                cur.wait_stream(cs)
            else:
                reduce_gradients(self.flat, self.world, self.pg, 0, end)
        ops.adam_step(self.flat.data[:nt], self.flat.grad, self.exp_avg, self.exp_avg_sq, self.adam_state,
                      -1.0, grad_scale=1.0 / self.world)

    def step(self, batches: Sequence[Dict], noises: Sequence[Dict]):
        assert len(batches) == self.accumulate
        return self._run(batches, noises)

    # -- CUDA graph ----------------------------------------------------------------------------
    def capture(self, batches: Sequence[Dict], noises: Sequence[Dict], warmup: int = 2):
        """Captures the whole step over static copies of the inputs; ``replay`` then costs one
        graph launch.  Warm-up steps run on a side stream first (allocator + NCCL warm)."""
        self.static_inputs = [{k: v.clone() for k, v in b.items()} for b in batches]
        self.static_noise = [{k: v.clone() for k, v in n.items()} for n in noises]
        keep = (self.flat.data.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(),
                self.adam_state.clone(), self._bn_state())
        s = self.stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            # at least one eager pass: allocator warm-up, NCCL, and the one-time host->device tables
            # (FlatParams.prepare_weights) must not happen under capture
            for _ in range(max(int(warmup), 1)):
                self._run(self.static_inputs, self.static_noise)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=s):
            self.loss_out = self._run(self.static_inputs, self.static_noise)
        # undo the warm-up / capture side effects on the training state
        self.flat.data.copy_(keep[0]); self.exp_avg.copy_(keep[1]); self.exp_avg_sq.copy_(keep[2])
        self.adam_state.copy_(keep[3]); self._bn_state(keep[4])
        torch.cuda.synchronize()

    def load_inputs(self, batches: Sequence[Dict], noises: Sequence[Dict]):
        for dst, src in zip(self.static_inputs, batches):
            for k in dst:
                dst[k].copy_(src[k], non_blocking=True)
        for dst, src in zip(self.static_noise, noises):
            for k in dst:
                dst[k].copy_(src[k], non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.loss_out

    def feeder(self) -> "InputFeeder":
        """Double-buffered host->device input path for the captured step (see InputFeeder)."""
        return InputFeeder(self)

    def _bn_state(self, restore=None):
        bufs = []
        for m in self.models.values():
            bufs += [b for _, b in m.named_buffers()]
        if restore is None:
            return [b.clone() for b in bufs]
        for b, r in zip(bufs, restore):
            b.copy_(r)


class InputFeeder:
    """Host -> device input pipeline for a captured TrainStep (what a DataLoader with pinned memory and
    prefetch does in the reference's run_epoch, trainer.py:230-236): ``prefetch`` copies the next step's
    pinned host batch into a device staging set on a copy stream while the current step computes;
    ``commit`` moves the staging set into the graph's static inputs (device-to-device, a fraction of a
    millisecond) on the compute stream right before ``replay``."""

    def __init__(self, step: TrainStep):
        assert step.static_inputs is not None, "capture the step first"
        self.step = step
        self.copy_stream = torch.cuda.Stream(device=step.flat.data.device)
        self.staging_inputs = [{k: torch.empty_like(v) for k, v in b.items()} for b in step.static_inputs]
        self.staging_noise = [{k: torch.empty_like(v) for k, v in n.items()} for n in step.static_noise]
        self._pairs = None
        self.ready = torch.cuda.Event()
        self.consumed = torch.cuda.Event()
        self.consumed.record(torch.cuda.current_stream())

    def prefetch(self, batches: Sequence[Dict], noises: Sequence[Dict]):
        self.copy_stream.wait_event(self.consumed)           # staging is no longer being read
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.staging_inputs, batches):
                for k in dst:
                    dst[k].copy_(src[k], non_blocking=True)
            for dst, src in zip(self.staging_noise, noises):
                for k in dst:
                    dst[k].copy_(src[k], non_blocking=True)
            self.ready.record(self.copy_stream)

    def commit(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready)
        if self._pairs is None:
            dsts, srcs = [], []
            for dst, src in zip(list(self.step.static_inputs) + list(self.step.static_noise),
                                list(self.staging_inputs) + list(self.staging_noise)):
                for k in dst:
                    dsts.append(dst[k])
                    srcs.append(src[k])
            self._pairs = (dsts, srcs)
        # one multi-tensor copy instead of ~60 small launches between two graph replays
        torch._foreach_copy_(self._pairs[0], self._pairs[1])
        self.consumed.record(cur)
