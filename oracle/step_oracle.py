"""CPU oracle for the FusionDepth per-step training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fusiondepth_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the
reported CPU baseline -- never as the thing shipped.

What it is: a functional, plain-PyTorch fp32 restatement of the reference's
algorithm for the path SURVEY.md section 8 scopes (the reference is pure Python;
``/root/reference`` cannot travel to the GPU box, so the algorithm is restated
here over plain ``state_dict`` tensors).  Each function cites the reference
``file:line`` it follows.

Parity pin: the reference ships no golden vectors or tests, so
this oracle is pinned against *outputs of the reference itself run in the build
container*: ``tests/make_golden.py`` imports ``/root/reference`` and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this module
against those fixtures everywhere, and ``tests/test_oracle_vs_reference.py``
checks it against the live reference whenever ``/root/reference`` exists.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

RESNET_BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3),
                 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


# --------------------------------------------------------------------------
# a1  ResnetEncoder.forward            networks/resnet_encoder.py:92-103
# --------------------------------------------------------------------------
def _bn(sd, name, x, training, momentum=0.1, eps=1e-5):
    # torchvision BasicBlock/Bottleneck use nn.BatchNorm2d defaults (eps 1e-5, momentum 0.1).
    rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
    if training and (name + ".num_batches_tracked") in sd:
        sd[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"],
                        training, momentum, eps)


def _basic_block(sd, p, x, stride, training):
    # torchvision.models.resnet.BasicBlock.forward (third-party; call site resnet_encoder.py:62-74)
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)
    out = F.relu(_bn(sd, p + ".bn1", out, training))
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
    out = _bn(sd, p + ".bn2", out, training)
    if (p + ".downsample.0.weight") in sd:
        idt = F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0)
        idt = _bn(sd, p + ".downsample.1", idt, training)
    else:
        idt = x
    return F.relu(out + idt)


def _bottleneck(sd, p, x, stride, training):
    # torchvision.models.resnet.Bottleneck.forward (v1.5: stride on the 3x3)
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"]), training))
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], None, stride, 1), training))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]), training)
    if (p + ".downsample.0.weight") in sd:
        idt = F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0)
        idt = _bn(sd, p + ".downsample.1", idt, training)
    else:
        idt = x
    return F.relu(out + idt)


def resnet_encoder(sd: Dict[str, Tensor], image: Tensor, num_layers: int = 18,
                   training: bool = True) -> List[Tensor]:
    """Five feature maps of ``ResnetEncoder.forward`` (resnet_encoder.py:92-103).

    ``sd`` is the module's state_dict (keys ``encoder.conv1.weight`` ...); running
    statistics in it are updated in place in training mode, like nn.BatchNorm2d.
    Input normalisation (x-0.45)/0.225 is applied to every channel, LiDAR ones
    included (resnet_encoder.py:94).
    """
    block = _basic_block if num_layers < 50 else _bottleneck
    feats = []
    x = (image - 0.45) / 0.225
    x = F.conv2d(x, sd["encoder.conv1.weight"], None, 2, 3)
    x = F.relu(_bn(sd, "encoder.bn1", x, training))
    feats.append(x)
    x = F.max_pool2d(x, 3, 2, 1)
    for li, nblk in enumerate(RESNET_BLOCKS[num_layers], start=1):
        for bi in range(nblk):
            stride = 2 if (li > 1 and bi == 0) else 1
            x = block(sd, "encoder.layer%d.%d" % (li, bi), x, stride, training)
        feats.append(x)
    return feats


# --------------------------------------------------------------------------
# a2  DepthDecoder.forward             networks/depth_decoder.py:63-96
# --------------------------------------------------------------------------
def _conv3x3_refl(x, w, b):
    # layers.py:115-130  Conv3x3 = ReflectionPad2d(1) + Conv2d(3)
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b)


def _convblock(sd, key, x, deep):
    # layers.py:100-112 ConvBlock = Conv3x3 + ELU ; `deep` doubles it (depth_decoder.py:27-33)
    if deep:
        for j in (0, 1):
            x = F.elu(_conv3x3_refl(x, sd["%s.%d.conv.conv.weight" % (key, j)],
                                    sd["%s.%d.conv.conv.bias" % (key, j)]))
        return x
    return F.elu(_conv3x3_refl(x, sd[key + ".conv.conv.weight"], sd[key + ".conv.conv.bias"]))


def depth_decoder(sd: Dict[str, Tensor], feats: Sequence[Tensor],
                  beam_feats: Optional[Sequence[Tensor]] = None,
                  depth_maps: Optional[Dict] = None, two_channel: Optional[Tensor] = None,
                  scales: Sequence[int] = (0, 1, 2, 3), deep: bool = False,
                  cat2end: bool = False, tanh: bool = False, use_skips: bool = True):
    """``DepthDecoder.forward`` (depth_decoder.py:63-96).  Module-list order
    (depth_decoder.py:22-58): decoder.{0..9} = upconv(4,0),(4,1),...,(0,1);
    decoder.{10+k} = dispconv of the k-th entry of ``scales``."""
    out = {}
    x = feats[-1] + beam_feats[-1] if beam_feats is not None else feats[-1]
    scales = list(scales)
    for i in range(4, -1, -1):
        k0 = "decoder.%d" % (2 * (4 - i))
        k1 = "decoder.%d" % (2 * (4 - i) + 1)
        x = _convblock(sd, k0, x, deep)
        xs = [F.interpolate(x, scale_factor=2, mode="nearest")]          # layers.py:229-232
        if use_skips and i > 0:
            xs.append(feats[i - 1] + beam_feats[i - 1] if beam_feats is not None else feats[i - 1])
        if depth_maps is not None and i in scales and use_skips:
            xs.append(depth_maps[("disp", i)])
        x = _convblock(sd, k1, torch.cat(xs, 1), deep)
        if i in scales:
            kd = "decoder.%d" % (10 + scales.index(i))
            xin = torch.cat((x, two_channel), 1) if (i == 0 and cat2end) else x
            y = _conv3x3_refl(xin, sd[kd + ".conv.weight"], sd[kd + ".conv.bias"])
            out[("disp", i)] = torch.tanh(y) if (tanh and not (i == 0 and cat2end)) else torch.sigmoid(y)
    return out


# --------------------------------------------------------------------------
# a3  PoseDecoder.forward              networks/pose_decoder.py:29-51
# --------------------------------------------------------------------------
def pose_decoder(sd: Dict[str, Tensor], last_feature: Tensor, beam_last: Optional[Tensor] = None,
                 num_frames_to_predict_for: int = 2):
    """Single-input-feature PoseDecoder (num_input_features=1, trainer.py:101-104)."""
    f = last_feature + beam_last if beam_last is not None else last_feature
    x = F.relu(F.conv2d(f, sd["net.0.weight"], sd["net.0.bias"]))
    x = F.relu(F.conv2d(x, sd["net.1.weight"], sd["net.1.bias"], 1, 1))
    x = F.relu(F.conv2d(x, sd["net.2.weight"], sd["net.2.bias"], 1, 1))
    x = F.conv2d(x, sd["net.3.weight"], sd["net.3.bias"])
    x = x.mean(3).mean(2)
    x = 0.01 * x.view(-1, num_frames_to_predict_for, 1, 6)
    return x[..., :3], x[..., 3:]


# --------------------------------------------------------------------------
# a4  transformation_from_parameters   layers.py:23-97
# --------------------------------------------------------------------------
def pose_matrix(axisangle: Tensor, translation: Tensor, invert: bool = False) -> Tensor:
    """axisangle, translation: [B,1,3] -> [B,4,4].  Rodrigues with angle=|v|,
    axis=v/(angle+1e-7) (layers.py:59-97); M = T@R, or R^T @ T(-t) if invert (23-40)."""
    vec = axisangle
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    B = vec.shape[0]
    rows = [x * xC + ca, xyC - zs, zxC + ys,
            xyC + zs, y * yC + ca, yzC - xs,
            zxC - ys, yzC + xs, z * zC + ca]
    R3 = torch.stack([r.reshape(B) for r in rows], 1).view(B, 3, 3)
    R = torch.zeros(B, 4, 4, dtype=vec.dtype)
    R[:, :3, :3] = R3
    R[:, 3, 3] = 1
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = torch.zeros(B, 4, 4, dtype=vec.dtype)
    T[:, 0, 0] = 1
    T[:, 1, 1] = 1
    T[:, 2, 2] = 1
    T[:, 3, 3] = 1
    T[:, :3, 3, None] = t.contiguous().view(-1, 3, 1)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


# --------------------------------------------------------------------------
# a5-a13  the photometric loss chain   trainer.py:425-596, layers.py:11-20,133-162,204-281
# --------------------------------------------------------------------------
def disp_to_depth(disp, min_depth=0.1, max_depth=100.0):
    # layers.py:11-20
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled = min_disp + (max_disp - min_disp) * disp
    return scaled, 1 / scaled


def backproject(depth: Tensor, inv_K: Tensor) -> Tensor:
    # layers.py:133-162 ; depth [B,1,H,W] -> cam points [B,4,H*W]
    B, _, H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32),
                            torch.arange(W, dtype=torch.float32), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W)], 0)
    pix = pix.unsqueeze(0).repeat(B, 1, 1)
    cam = torch.matmul(inv_K[:, :3, :3], pix)
    cam = depth.view(B, 1, -1) * cam
    return torch.cat([cam, torch.ones(B, 1, H * W)], 1)


def project(points: Tensor, K: Tensor, T: Tensor, H: int, W: int, eps: float = 1e-7) -> Tensor:
    # layers.py:204-226 ; returns the sampling grid [B,H,W,2] normalised by (W-1),(H-1)
    B = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, points)
    pix = cam[:, :2, :] / (cam[:, 2, :].unsqueeze(1) + eps)
    pix = pix.view(B, 2, H, W).permute(0, 2, 3, 1)
    pix = torch.stack([pix[..., 0] / (W - 1), pix[..., 1] / (H - 1)], -1)
    return (pix - 0.5) * 2


def ssim(x: Tensor, y: Tensor) -> Tensor:
    # layers.py:251-281
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sigma_x + sigma_y + C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def reprojection_loss(pred: Tensor, target: Tensor) -> Tensor:
    # trainer.py:476-488
    l1 = torch.abs(target - pred).mean(1, True)
    return 0.85 * ssim(pred, target).mean(1, True) + 0.15 * l1


def smooth_loss(disp: Tensor, img: Tensor) -> Tensor:
    # layers.py:235-248
    gdx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
    gdy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
    gix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
    giy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
    return (gdx * torch.exp(-gix)).mean() + (gdy * torch.exp(-giy)).mean()


def photometric_chain(inputs: Dict, disps: Dict, cam_T: Dict, noise: Dict,
                      frame_ids=(0, -1, 1), scales=(0, 1, 2, 3),
                      min_depth=0.1, max_depth=100.0, smoothness=1e-3,
                      siloss=True, siloss_all_scale=True, si_var=0.3, si_thresh=2.0):
    """generate_images_pred + compute_losses (trainer.py:425-474, 490-596) with the
    reference's default flags (automasking on, SSIM on, v1_multiscale off,
    avg_reprojection off, trainer_siloss on for all scales).

    ``noise[s]`` is the [B,2,H,W] tensor the reference draws with
    ``torch.randn(...)`` at trainer.py:551 (passed in so both sides see the same
    tie-break noise).  Returns (losses, outputs) keyed like the reference.
    """
    H, W = inputs[("color", 0, 0)].shape[-2:]
    outputs, losses = {}, {}
    total = 0
    target = inputs[("color", 0, 0)]
    for s in scales:
        disp = disps[("disp", s)]
        up = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
        _, depth = disp_to_depth(up, min_depth, max_depth)
        outputs[("depth", 0, s)] = depth
        reproj = []
        for f in frame_ids[1:]:
            cam = backproject(depth, inputs[("inv_K", 0)])
            grid = project(cam, inputs[("K", 0)], cam_T[f], H, W)
            outputs[("sample", f, s)] = grid
            warped = F.grid_sample(inputs[("color", f, 0)], grid, padding_mode="border",
                                   align_corners=False)
            outputs[("color", f, s)] = warped
            reproj.append(reprojection_loss(warped, target))
        reproj = torch.cat(reproj, 1)
        ident = torch.cat([reprojection_loss(inputs[("color", f, 0)], target)
                           for f in frame_ids[1:]], 1)
        ident = ident + noise[s] * 0.00001
        combined = torch.cat((ident, reproj), 1)
        to_opt, idxs = torch.min(combined, dim=1)
        outputs["identity_selection/%d" % s] = (idxs > ident.shape[1] - 1).float()
        outputs["to_optimise/%d" % s] = to_opt
        loss = to_opt.mean()
        mean_disp = disp.mean(2, True).mean(3, True)
        norm_disp = disp / (mean_disp + 1e-7)
        loss = loss + smoothness * smooth_loss(norm_disp, inputs[("color", 0, s)]) / (2 ** s)
        total = total + loss
        losses["loss/%d" % s] = loss
        if siloss and (siloss_all_scale or s == 0):
            # trainer.py:577-589
            d_si = disp_to_depth(up, min_depth, max_depth)[1] * 26.0
            beam = inputs["4beam"] * 100.0
            valid = ((beam > 1) * (d_si < 80) * (d_si > 1) * (abs(d_si - beam) < si_thresh)).detach()
            d = torch.log(d_si[valid]) - torch.log(beam[valid])
            si = torch.sqrt((d ** 2).mean() - si_var * (d.mean() ** 2)) * 0.1
            total = total + si
            losses["loss/si_loss%d" % s] = si
    losses["loss"] = total / len(scales)
    return losses, outputs


# --------------------------------------------------------------------------
# a15  process_batch / predict_poses   trainer.py:268-388 (default flag set)
# --------------------------------------------------------------------------
def process_batch(models: Dict[str, Dict[str, Tensor]], inputs: Dict, noise: Dict,
                  num_layers: int = 18, training: bool = True, frame_ids=(0, -1, 1),
                  scales=(0, 1, 2, 3)):
    """One micro-batch through the six trunks, decoder, two pose heads and the loss
    chain.  ``models`` maps the reference's model names (trainer.py:66-115) to
    state_dicts: encoder, beam_encoder, beam_encoder_pose, depth, pose_encoder, pose."""
    feats = resnet_encoder(models["encoder"], inputs[("color_aug", 0, 0)], num_layers, training)
    beam = resnet_encoder(models["beam_encoder"], inputs["2channel"], num_layers, training)
    disps = depth_decoder(models["depth"], feats, beam_feats=beam, scales=scales)
    outputs = dict(disps)
    cam_T = {}
    for f in frame_ids[1:]:
        # temporal order (trainer.py:339-346)
        pair = (f, 0) if f < 0 else (0, f)
        img = torch.cat([inputs[("color_aug", i, 0)] for i in pair], 1)
        two = torch.cat([inputs[("2channel", i, 0)] for i in pair], 1)
        pf = resnet_encoder(models["pose_encoder"], img, num_layers, training)
        bf = resnet_encoder(models["beam_encoder_pose"], two, num_layers, training)
        aa, tr = pose_decoder(models["pose"], pf[-1], bf[-1])
        outputs[("axisangle", 0, f)], outputs[("translation", 0, f)] = aa, tr
        cam_T[f] = pose_matrix(aa[:, 0], tr[:, 0], invert=(f < 0))
        outputs[("cam_T_cam", 0, f)] = cam_T[f]
    losses, o2 = photometric_chain(inputs, disps, cam_T, noise, frame_ids, scales)
    outputs.update(o2)
    return outputs, losses


def adam_step(params: List[Tensor], grads: List[Tensor], exp_avg: List[Tensor],
              exp_avg_sq: List[Tensor], step: int, lr: float,
              beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update rule (trainer.py:129, defaults)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))


# --------------------------------------------------------------------------
# a19 refiner-only pieces              layers.py:165-201, refiner.py:316-346,557-563
# --------------------------------------------------------------------------
def cat_xy(depth: Tensor, inv_K: Tensor) -> Tensor:
    # layers.py:165-201
    B, _, H, W = depth.shape
    cam = backproject(depth, inv_K)[:, :3].view(B, 3, H, W).clone()
    cam[:, 0] = cam[:, 0] / 30.0
    cam[:, 1] = cam[:, 1] / 2.0
    cam[:, 2] = (cam[:, 2] - 40) / 40.0
    return cam


def siloss(pred: Tensor, gt: Tensor, thresh: float = 2.0, si_var: float = 0.3) -> Tensor:
    # refiner.py:557-563 (called with depth, inf_gdc at 678-688)
    valid = ((gt > 1e-3) * (pred < 80) * (pred > 1e-3) * (abs(pred - gt) < thresh)).detach()
    d = torch.log(pred[valid]) - torch.log(gt[valid])
    return torch.sqrt((d ** 2).mean() - si_var * (d.mean() ** 2)) * 10.0


CROP = (78, 190, 23, 617)      # refiner.py:330 (the 375x1242 Garg-like crop at 192x640)


def pseudo3d_pack(disps: Dict, inputs: Dict, scales=(0, 1, 2, 3), min_depth=0.1, max_depth=100.0):
    """refiner.py:316-346 with the default flags (refine_a0='true', catxy='true'): the 6-channel
    pseudo-3D maps [scaled_disp(1), xyz(3), two_cha(2)] per scale, from the frozen stage-1 disparity
    at scale 0.  Returns ({("disp", s): [B,6,h,w]}, ratios[s])."""
    H, W = inputs[("color", 0, 0)].shape[-2:]
    beam = inputs["4beam"]
    two_cha = inputs["2channel"]
    disp_0 = disps[("disp", 0)]
    out, ratios = {}, []
    for s in scales:
        disp = disp_0
        disp_0 = F.max_pool2d(disp_0, 2, ceil_mode=True)
        disp640 = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
        depth = disp_to_depth(disp640, min_depth, max_depth)[1]
        mask = beam > 0
        crop = torch.zeros_like(mask)
        crop[:, :, CROP[0]:CROP[1], CROP[2]:CROP[3]] = 1
        mask = mask * crop
        ratio = torch.median(beam[mask] * 100.0) / torch.median(depth[mask]).detach()
        ratios.append(ratio)
        depth = depth * ratio
        scaled_disp = (F.interpolate(1 / depth, disp.shape[2:], mode="bilinear",
                                     align_corners=False) - 0.01) / 9.9
        if s != 0:
            two_cha = F.max_pool2d(two_cha, 2, ceil_mode=True)
        for _ in range(s):
            depth = F.max_pool2d(depth, 2, ceil_mode=True)
        xyz = cat_xy(depth, inputs[("inv_K", s)])
        out[("disp", s)] = torch.cat([scaled_disp, xyz, two_cha], 1)
    return out, ratios


def refiner_process_batch(models: Dict[str, Dict[str, Tensor]], inputs: Dict, noise: Dict,
                          num_layers: int = 18, training: bool = True, frame_ids=(0, -1, 1),
                          scales=(0, 1, 2, 3), gdc_weight=0.008, gdc_thresh=2.0, si_var=0.3,
                          smoothness=1e-3):
    """Refiner.process_batch + compute_losses (refiner.py:299-378, 586-693), default flags: frozen
    stage-1 encoder / beam encoder / depth decoder under no_grad (BatchNorm still in train mode --
    run_epoch calls set_train(), refiner.py:268 -- and `depth(features)` WITHOUT beam features,
    refine_depthnet_with_beam='false'), pseudo-3D pack, pose nets, the refine2d decoder
    (road, catxy, deep; sigmoid since refine_offset is off), the same warp + photometric + smoothness
    chain as stage 1, and the GDC-clone si-loss on scale 0 (x gdc_loss_weight x 4)."""
    with torch.no_grad():
        feats = resnet_encoder(models["encoder"], inputs[("color_aug", 0, 0)], num_layers, training)
        beam = resnet_encoder(models["beam_encoder"], inputs["2channel"], num_layers, training)
        coarse = depth_decoder(models["depth"], feats, beam_feats=None, scales=scales)
        packed, ratios = pseudo3d_pack(coarse, inputs, scales)
    outputs = {("pseudo3d", s): packed[("disp", s)] for s in scales}
    outputs["ratios"] = torch.stack(ratios)
    cam_T = {}
    for f in frame_ids[1:]:
        pair = (f, 0) if f < 0 else (0, f)
        img = torch.cat([inputs[("color_aug", i, 0)] for i in pair], 1)
        two = torch.cat([inputs[("2channel", i, 0)] for i in pair], 1)
        pf = resnet_encoder(models["pose_encoder"], img, num_layers, training)
        bf = resnet_encoder(models["beam_encoder_pose"], two, num_layers, training)
        aa, tr = pose_decoder(models["pose"], pf[-1], bf[-1])
        cam_T[f] = pose_matrix(aa[:, 0], tr[:, 0], invert=(f < 0))
        outputs[("cam_T_cam", 0, f)] = cam_T[f]
    disps = depth_decoder(models["refine2d_decoder"], feats, beam_feats=beam, depth_maps=packed,
                          scales=scales, deep=True)
    outputs.update(disps)
    pl, o2 = photometric_chain(inputs, disps, cam_T, noise, frame_ids, scales, smoothness=smoothness,
                               siloss=False)
    outputs.update(o2)
    H, W = inputs[("color", 0, 0)].shape[-2:]
    losses = {"loss/gama1.0_scale%d" % s: pl["loss/%d" % s] for s in scales}
    total = sum(pl["loss/%d" % s] for s in scales)
    # refiner.py:678-688 (gdc_loss_only_on_scale_0 defaults to True: store_false flag)
    d0 = F.interpolate(disps[("disp", 0)], [H, W], mode="bilinear", align_corners=False).squeeze()
    depth0 = disp_to_depth(d0, 0.1, 100.0)[1]
    gdc = siloss(depth0, inputs["inf_gdc"].squeeze(), gdc_thresh, si_var) * gdc_weight * 4.0
    losses["loss/gdc_scale0"] = gdc
    losses["loss"] = (total + gdc) / len(scales)
    return outputs, losses


# --------------------------------------------------------------------------
# evaluate_depth.py:42-60 compute_errors + 344-378 median scaling (AbsRel oracle)
# --------------------------------------------------------------------------
def eval_abs_rel(pred_disp, gt_depth, min_depth=1e-3, max_depth=80.0):
    import numpy as np
    gt_h, gt_w = gt_depth.shape
    pd = torch.from_numpy(pred_disp)[None, None]
    # cv2.resize(bilinear) == F.interpolate(align_corners=False) for upscaling
    pd = F.interpolate(pd, (gt_h, gt_w), mode="bilinear", align_corners=False)[0, 0].numpy()
    pred = 1 / pd
    mask = np.logical_and(gt_depth > min_depth, gt_depth < max_depth)
    crop = np.array([0.40810811 * gt_h, 0.99189189 * gt_h,
                     0.03594771 * gt_w, 0.96405229 * gt_w]).astype(np.int32)
    cm = np.zeros(mask.shape)
    cm[crop[0]:crop[1], crop[2]:crop[3]] = 1
    mask = np.logical_and(mask, cm)
    pred, gt = pred[mask], gt_depth[mask]
    pred = pred * (np.median(gt) / np.median(pred))
    pred = np.clip(pred, min_depth, max_depth)
    return float(np.mean(np.abs(gt - pred) / gt))
