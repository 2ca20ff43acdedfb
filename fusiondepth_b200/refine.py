"""Stage 2: the refiner's per-step hot path (reference refiner.py:299-378, 586-693; BASELINE config 5).

Frozen stage-1 encoder / beam encoder / depth decoder (no gradient, BatchNorm still in train mode because
Refiner.run_epoch calls set_train(), refiner.py:268) -> pseudo-3D pack (fd_refine_pack: masked median ratio,
cumulative max-pools, Cat_xy; refiner.py:316-346) -> pose networks -> the refine2d decoder
(DepthDecoder(road, catxy, deep), the only trainable network, refiner.py:146-158) -> the same fused warp +
photometric + smoothness chain as stage 1 with the GDC-clone si-loss on scale 0 (refiner.py:678-688) -> Adam.

One deliberate difference in WORK, not in results: the reference runs predict_poses with autograd enabled
although the pose networks are not in the optimiser (only with --train_entire_net), so it back-propagates
into them and throws the gradients away.  Here the pose networks run without a tape; every tensor the
optimiser or the caller sees is identical.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from . import networks, ops
from .training import (TrainStep, TrunkStreams, build_models, draw_noise, fused_losses, predict_poses)

CROP = (78, 190, 23, 617)          # refiner.py:330
TRAINABLE = ("refine2d_decoder",)


def build_refiner_models(num_layers: int = 18, device="cuda") -> Dict[str, nn.Module]:
    """The stage-1 networks (loaded from checkpoints by Refiner.__init__, refiner.py:79-140) plus the
    stage-2 decoder (refiner.py:146-158).  Only the latter is trainable."""
    m = build_models(num_layers, device)
    m["refine2d_decoder"] = networks.DepthDecoder(m["encoder"].num_ch_enc, [0, 1, 2, 3], road=True, catxy=True,
                                                  deep=True).to(device)
    for name, mod in m.items():
        if name not in TRAINABLE:
            for p in mod.parameters():
                p.requires_grad_(False)
    return m


def loss_opts(opt=None) -> Dict:
    """fd_photoloss options for refiner.py:557-563, 678-688: siloss(depth, inf_gdc) * 10 * gdc_loss_weight
    * 4 on scale 0 only (gdc_loss_only_on_scale_0 is a store_false flag: on by default)."""
    g = (lambda k, d: getattr(opt, k, d)) if opt is not None else (lambda k, d: d)
    only0 = g("gdc_loss_only_on_scale_0", True)
    return {"min_depth": g("min_depth", 0.1), "max_depth": g("max_depth", 100.0),
            "smoothness": g("disparity_smoothness", 1e-3), "si_thresh": g("gdc_loss_threshold", 2.0),
            "si_var": g("si_var", 0.3), "use_si": True, "si_scales": 0x1 if only0 else 0xF,
            "si_pred_mul": 1.0, "si_tgt_mul": 1.0, "si_lo": 1e-3,
            "si_weight": 10.0 * g("gdc_loss_weight", 0.008) * (4.0 if only0 else 1.0)}


def rename_losses(losses: Dict, gama: float = 1.0) -> Dict:
    """The reference's keys (refiner.py:676, 688, 692)."""
    out = {"loss/gama%s_scale%d" % (gama, s): losses["loss/%d" % s] for s in range(4)}
    out["loss/gdc_scale0"] = losses["loss/si_loss0"]
    out["loss"] = losses["loss"]
    return out


def process_batch(models, inputs, noise=None, opts: Optional[Dict] = None, materialize=False,
                  frame_ids=(0, -1, 1), streams: Optional[TrunkStreams] = None, crop=CROP):
    """Refiner.process_batch (refiner.py:299-378), default flags (refine_a0, catxy, refine2d_deep 'true';
    refine_depthnet_with_beam 'false'; refine_iter 1; refine_offset off)."""
    opts = opts if opts is not None else loss_opts()
    with torch.no_grad():
        if streams is not None:
            streams.fork()
            side = streams.side
            with torch.cuda.stream(side[0]):
                beam = models["beam_encoder"](inputs["2channel"])
            pose_out = {}
            with torch.cuda.stream(side[1]):
                pose_out.update(predict_poses(models, inputs, frame_ids[:2]))
            with torch.cuda.stream(side[2]):
                pose_out.update(predict_poses(models, inputs, (frame_ids[0], frame_ids[2])))
        feats = models["encoder"](inputs[("color_aug", 0, 0)])
        feats = list(feats)
        coarse = models["depth"](feats)                                   # no beam features (refiner.py:313)
        packed, ratios = ops.refine_pack(coarse[("disp", 0)], inputs["4beam"], inputs["2channel"],
                                         [inputs[("inv_K", s)] for s in range(4)], crop,
                                         opts.get("min_depth", 0.1), opts.get("max_depth", 100.0))
        if streams is None:
            beam = models["beam_encoder"](inputs["2channel"])
            pose_out = predict_poses(models, inputs, frame_ids)
        else:
            streams.join([0, 1, 2])
        beam = list(beam)
    outputs = {("pseudo3d", s): packed[s] for s in range(4)}
    outputs["ratios"] = ratios
    outputs.update(pose_out)
    depth_maps = {("disp", s): packed[s] for s in range(4)}
    outputs.update(models["refine2d_decoder"](feats, beam_features=beam, depth_maps=depth_maps, tanh=False))
    if noise is None:
        B, _, H, W = inputs[("color", 0, 0)].shape
        noise = draw_noise(B, H, W, inputs[("color", 0, 0)].device)
    losses = fused_losses(inputs, outputs, noise, opts, materialize, frame_ids, si_target="inf_gdc")
    return outputs, rename_losses(losses)


class RefineStep(TrainStep):
    """One optimiser step of refiner.py:270-278: process_batch -> backward -> Adam on the refine2d decoder.
    The reference does not accumulate in stage 2 (it halves the batch for --batch_size > 8 but steps after
    every micro-batch, refiner.py:34-45, 270-278), so `accumulate` stays 1: --batch_size 12 is one optimiser
    step per 6 images."""

    def __init__(self, models, lr: float = 1e-4, opts: Optional[Dict] = None, process_group=None,
                 parallel_trunks: bool = True):
        super().__init__(models, lr=lr, accumulate=1, opts=opts if opts is not None else loss_opts(),
                         process_group=process_group, parallel_trunks=parallel_trunks,
                         concurrent_microbatches=False)
        if parallel_trunks:
            # three side branches: beam encoder, and the two pose pairs (pose + beam-pose encoders each)
            self.trunks = [TrunkStreams(self.flat.data.device, 3)]
            self.streams = self.trunks[0]

    def _micro_batch(self, inputs, noise, trunks):
        _, losses = process_batch(self.models, inputs, noise, self.opts, streams=trunks)
        loss = losses["loss"]
        loss.backward()
        return loss.detach()


def patch_refiner(refiner, materialize: bool = True):
    """Swap the reference Refiner's process_batch (training mode) for this module's, in place: the pack, the
    warp and the loss run on the fused kernels, the networks are whatever `refiner.models` holds (this
    package's modules when the driver imported `networks` from fusiondepth_b200/dropin)."""
    opt = refiner.opt
    if (opt.v1_multiscale or opt.disable_automasking or opt.avg_reprojection or opt.no_ssim or opt.predictive_mask
            or opt.use_stereo or opt.pose_model_type != "separate_resnet" or opt.train_entire_net
            or opt.refine_iter != 1 or opt.refine_offset or opt.refine_a0 != "true" or opt.catxy != "true"
            or opt.refine_depthnet_with_beam == "true" or list(opt.scales) != [0, 1, 2, 3]
            or list(opt.frame_ids) != [0, -1, 1]):
        raise NotImplementedError("patch_refiner covers the reference's default stage-2 configuration")
    opts = loss_opts(opt)
    reference_process_batch = refiner.process_batch

    def patched(inputs, val=False):
        if val:
            return reference_process_batch(inputs, val=True)
        for key, ipt in inputs.items():
            if torch.is_tensor(ipt):
                inputs[key] = ipt.to(refiner.device)
        return process_batch(refiner.models, inputs, None, opts, materialize)

    refiner.process_batch = patched
    return refiner
