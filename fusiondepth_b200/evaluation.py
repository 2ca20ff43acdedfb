"""Inference / evaluation path.

  * ``EvalRunner`` -- encoder + beam encoder + depth decoder (+ the stage-2 pack and refine2d decoder) forward in
    eval mode, the way evaluate_depth.py:162-237 and inf_depth_map.py:166-170 wire them, with every BatchNorm
    FOLDED into its convolution (w' = w * gamma * rstd, b' = beta - mean * gamma * rstd: conv + bias + ReLU is
    one tensor-core kernel, the residual join one add+ReLU kernel) and the whole forward captured in a CUDA
    graph for batch-1 latency.
  * ``evaluate_frame`` / ``evaluate_split`` -- evaluate_depth.py:344-378, 470-488: resize the disparity to the
    ground-truth size, 1/disp, Eigen mask + Garg crop, median scaling (numpy.median), clamp to [1e-3, 80], the
    seven metrics of compute_errors (evaluate_depth.py:42-60), mean over frames -- all on the device
    (fd_upsample_bilinear_fwd + fd_depth_errors).
  * ``compute_depth_losses`` -- Trainer.compute_depth_losses (trainer.py:598-630), the validation metric the
    trainer computes every 250 steps over the 697 Eigen test frames (torch.median semantics, 153:371 x 44:1197
    crop); ``patch_validation`` swaps it into an unchanged Trainer.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .networks.resnet_encoder import BasicBlock, Bottleneck, BatchNorm2d, Conv2d

MIN_DEPTH, MAX_DEPTH = 1e-3, 80.0                 # evaluate_depth.py:29-30
DEPTH_METRIC_NAMES = ["de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"]


# ------------------------------------------------------------------------------------------------
def fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """(weight', bias') of conv followed by eval-mode BatchNorm, channels-last, fp32 (computed in fp64)."""
    w = conv.weight.detach().double()
    scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    b = bn.bias.detach().double() - bn.running_mean.detach().double() * scale
    if conv.bias is not None:
        b = b + conv.bias.detach().double() * scale
    w = (w * scale.view(-1, 1, 1, 1)).float().contiguous(memory_format=torch.channels_last)
    return w, b.float().contiguous()


class FoldedTrunk:
    """A ResNet trunk (networks.ResnetEncoder) with its BatchNorms folded away, for inference."""

    def __init__(self, encoder):
        enc = encoder.encoder
        self.stem_w, self.stem_b = fold_conv_bn(enc.conv1, enc.bn1)
        self.blocks: List = []
        for layer in (enc.layer1, enc.layer2, enc.layer3, enc.layer4):
            stage = []
            for blk in layer:
                convs = [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
                if isinstance(blk, Bottleneck):
                    convs.append((blk.conv3, blk.bn3))
                folded = [(c.stride[0], c.padding[0]) + fold_conv_bn(c, b) for c, b in convs]
                down = None
                if blk.downsample is not None:
                    dc, db = blk.downsample[0], blk.downsample[1]
                    down = (dc.stride[0], dc.padding[0]) + fold_conv_bn(dc, db)
                stage.append((folded, down))
            self.blocks.append(stage)

    def __call__(self, image):
        # (x - 0.45) / 0.225 + 7x7/2 conv through the normalised-im2col GEMM, folded bias + ReLU in its epilogue
        x = ops.stem_conv(image, self.stem_w, self.stem_b, "relu")
        feats = [x]
        x = ops.maxpool3x3s2(x)
        for stage in self.blocks:
            for folded, down in stage:
                out = x
                for i, (stride, pad, w, b) in enumerate(folded):
                    out = ops.conv2d(out, w, b, stride, pad, "relu" if i + 1 < len(folded) else "none")
                idt = x if down is None else ops.conv2d(x, down[2], down[3], down[0], down[1], "none")
                x = ops.add_relu(out, idt)
            feats.append(x)
        return feats


class EvalRunner:
    """evaluate_depth.py:173-237 on this library: returns the scaled disparity [B,H,W] evaluate_depth stores
    (disp_to_depth(...)[0] of ("disp", 0) brought to HxW).  ``refine``: the stage-2 models dict entry
    "refine2d_decoder" is applied after the pseudo-3D pack (evaluate_depth.py:197-233)."""

    def __init__(self, models: Dict[str, nn.Module], batch: int, height: int, width: int, refine: bool = False,
                 fold_bn: bool = True, graph: bool = True, min_depth: float = 0.1, max_depth: float = 100.0,
                 depth_with_beam: Optional[bool] = None):
        self.models, self.refine = models, refine
        self.B, self.H, self.W = batch, height, width
        self.min_depth, self.max_depth = min_depth, max_depth
        for m in models.values():
            m.eval()
        dev = next(models["encoder"].parameters()).device
        self.enc = FoldedTrunk(models["encoder"]) if fold_bn else models["encoder"]
        self.benc = FoldedTrunk(models["beam_encoder"]) if fold_bn else models["beam_encoder"]
        # evaluate_depth.py:183-186: with a refiner the stage-1 decoder runs without beam features unless
        # --refine_depthnet_with_beam true
        self.depth_with_beam = (not refine) if depth_with_beam is None else depth_with_beam
        self.static = {"color": torch.zeros(batch, 3, height, width, device=dev),
                       "2channel": torch.zeros(batch, 2, height, width, device=dev)}
        if refine:
            self.static["4beam"] = torch.zeros(batch, 1, height, width, device=dev)
            for s in range(4):
                self.static[("inv_K", s)] = torch.zeros(batch, 4, 4, device=dev)
        self.graph = None
        self.out = None
        if graph:
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), torch.no_grad():
                self._forward()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=s), torch.no_grad():
                self.out = self._forward()

    def _forward(self):
        st = self.static
        feats = list(self.enc(st["color"]))                       # evaluate_depth.py:177: color, not color_aug
        beam = list(self.benc(st["2channel"]))
        dec = self.models["depth"]
        out = dec(feats, beam_features=beam) if self.depth_with_beam else dec(feats)
        if self.refine:
            packed, _ = ops.refine_pack(out[("disp", 0)], st["4beam"], st["2channel"],
                                        [st[("inv_K", s)] for s in range(4)], min_depth=self.min_depth,
                                        max_depth=self.max_depth)
            out = self.models["refine2d_decoder"](feats, beam_features=beam,
                                                  depth_maps={("disp", s): packed[s] for s in range(4)}, tanh=False)
        disp = out[("disp", 0)]
        if disp.shape[-2:] != (self.H, self.W):
            disp = ops.upsample_bilinear(disp, self.H, self.W)
        return (1.0 / self.max_depth + (1.0 / self.min_depth - 1.0 / self.max_depth) * disp)[:, 0]

    @torch.no_grad()
    def __call__(self, inputs: Dict) -> torch.Tensor:
        self.static["color"].copy_(inputs[("color", 0, 0)], non_blocking=True)
        self.static["2channel"].copy_(inputs["2channel"], non_blocking=True)
        if self.refine:
            self.static["4beam"].copy_(inputs["4beam"], non_blocking=True)
            for s in range(4):
                self.static[("inv_K", s)].copy_(inputs[("inv_K", s)], non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
            return self.out
        return self._forward()


# ------------------------------------------------------------------------------------------------
def garg_crop(gt_height: int, gt_width: int):
    """evaluate_depth.py:361-362."""
    c = np.array([0.40810811 * gt_height, 0.99189189 * gt_height,
                  0.03594771 * gt_width, 0.96405229 * gt_width]).astype(np.int32)
    return int(c[0]), int(c[1]), int(c[2]), int(c[3])


def evaluate_frame(pred_disp: torch.Tensor, gt_depth: torch.Tensor, eigen: bool = True,
                   median_scaling: bool = True) -> torch.Tensor:
    """evaluate_depth.py:344-378 + 470-478 for one frame.  pred_disp [h,w] (the scaled disparity the eval loop
    stores), gt_depth [H,W]; returns the device tensor [abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, n, ratio]."""
    gt_h, gt_w = gt_depth.shape[-2:]
    d = pred_disp.reshape(1, 1, *pred_disp.shape[-2:])
    if d.shape[-2:] != (gt_h, gt_w):
        d = ops.upsample_bilinear(d, gt_h, gt_w)                  # cv2.resize(INTER_LINEAR) == half-pixel bilinear
    if eigen:
        window, lo, hi = garg_crop(gt_h, gt_w), MIN_DEPTH, MAX_DEPTH
    else:
        window, lo, hi = (0, gt_h, 0, gt_w), 0.0, float("inf")
    return ops.depth_errors(gt_depth.reshape(1, 1, gt_h, gt_w), d, window, lo, hi, pred_is_disp=True,
                            median_scaling=median_scaling, numpy_median=True, clamp=(MIN_DEPTH, MAX_DEPTH))


def evaluate_split(pred_disps: Sequence[torch.Tensor], gt_depths: Sequence[torch.Tensor], eigen: bool = True):
    """Mean of the per-frame metrics (evaluate_depth.py:485-488) and the scaling ratios (480-483)."""
    rows = torch.stack([evaluate_frame(p, g, eigen) for p, g in zip(pred_disps, gt_depths)])
    mean = rows[:, :7].mean(0)
    return {k: float(v) for k, v in zip(ops.DEPTH_METRIC_NAMES, mean)}, rows[:, 8]


def compute_depth_losses(inputs: Dict, outputs: Dict, losses: Dict, accumulate: bool = False,
                         crop=(153, 371, 44, 1197)):
    """Trainer.compute_depth_losses (trainer.py:598-630): ("depth", 0, 0) resized to the ground truth, clamped
    to [1e-3, 80], Garg/Eigen crop of the 375x1242 frame, scaled by torch.median(gt) / torch.median(pred),
    clamped again, layers.compute_depth_errors."""
    gt = inputs["depth_gt"]
    gh, gw = gt.shape[-2:]
    pred = outputs[("depth", 0, 0)].detach()
    if pred.shape[-2:] != (gh, gw):
        pred = ops.upsample_bilinear(pred, gh, gw)
    m = ops.depth_errors(gt, pred, crop, 0.0, float("inf"), pred_is_disp=False, pre_clamp=(1e-3, 80.0),
                         median_scaling=True, numpy_median=False, clamp=(1e-3, 80.0)).cpu()
    for i, name in enumerate(DEPTH_METRIC_NAMES):
        v = np.array(m[i])
        losses[name] = losses[name] + v if accumulate else v
    return losses


def patch_validation(trainer):
    """Swap the reference Trainer's compute_depth_losses for the device-side one, in place."""
    trainer.compute_depth_losses = lambda inputs, outputs, losses, accumulate=False: compute_depth_losses(
        inputs, outputs, losses, accumulate)
    return trainer
