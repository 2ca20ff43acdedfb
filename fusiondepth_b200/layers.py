"""Drop-in mirror of the reference's ``layers`` module surface (reference layers.py:1-302).

Same public names, signatures and results, so ``from layers import *`` in the unchanged
trainer.py / refiner.py / completor.py / evaluate_depth.py resolves here (the drivers also pick
up ``F``, ``nn``, ``np`` and ``torch`` from this namespace -- SURVEY.md section 0).

``ConvBlock`` / ``Conv3x3`` run on the package's CUDA kernels.  The unfused geometry modules
(``BackprojectDepth``, ``Project3D``, ``SSIM``, ``get_smooth_loss`` ...) are kept for API
compatibility as thin tensor compositions; the product path does not call them -- the whole
chain they form runs fused in ``fd_photoloss_fwd/bwd`` (see fusiondepth_b200.training).
"""
from __future__ import absolute_import, division, print_function

import numpy as np

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops as _ops


def disp_to_depth(disp, min_depth, max_depth):
    """Sigmoid output -> (scaled disparity, depth).  Reference layers.py:11-20."""
    lo, hi = 1 / max_depth, 1 / min_depth
    scaled_disp = lo + (hi - lo) * disp
    return scaled_disp, 1 / scaled_disp


def rot_from_axisangle(vec):
    """[B,1,3] axis-angle -> [B,4,4] rotation (Rodrigues; angle=|v|, axis=v/(angle+1e-7)).
    Reference layers.py:59-97."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    B = vec.shape[0]
    zero = torch.zeros(B, device=vec.device, dtype=vec.dtype)
    one = torch.ones(B, device=vec.device, dtype=vec.dtype)
    e = [x * xC + ca, xyC - zs, zxC + ys, zero,
         xyC + zs, y * yC + ca, yzC - xs, zero,
         zxC - ys, yzC + xs, z * zC + ca, zero,
         zero, zero, zero, one]
    return torch.stack([t.reshape(B) for t in e], 1).view(B, 4, 4)


def get_translation_matrix(translation_vector):
    """[B,1,3] -> [B,4,4] homogeneous translation.  Reference layers.py:43-56."""
    t = translation_vector.contiguous().view(-1, 3)
    T = torch.eye(4, device=t.device, dtype=t.dtype).unsqueeze(0).repeat(t.shape[0], 1, 1)
    T[:, :3, 3] = t
    return T


def transformation_from_parameters(axisangle, translation, invert=False):
    """(axisangle, translation) -> 4x4; M = T@R, or R^T@T(-t) when inverting.
    Reference layers.py:23-40."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        return torch.matmul(R.transpose(1, 2), get_translation_matrix(t * -1))
    return torch.matmul(get_translation_matrix(t), R)


class Conv3x3(nn.Module):
    """ReflectionPad2d(1) (or zero pad) + 3x3 conv.  Reference layers.py:115-130.
    State-dict keys: ``conv.weight``, ``conv.bias``."""

    def __init__(self, in_channels, out_channels, use_refl=True):
        super(Conv3x3, self).__init__()
        self.use_refl = bool(use_refl)
        self.pad = nn.ReflectionPad2d(1) if use_refl else nn.ZeroPad2d(1)
        self.conv = nn.Conv2d(int(in_channels), int(out_channels), 3)
        self.conv.weight.data = self.conv.weight.data.contiguous(memory_format=torch.channels_last)

    def forward(self, x, act="none", segments=None):
        """``segments`` (internal): un-assembled inputs [(a, b|None, upsample)], fused with the pad."""
        if self.use_refl:
            xp = _ops.assemble(segments if segments is not None else [(x, None, False)], pad=1)
            return _ops.conv2d(xp, self.conv.weight, self.conv.bias, 1, 0, act)
        if segments is not None:
            x = _ops.assemble(segments, pad=0)
        return _ops.conv2d(x, self.conv.weight, self.conv.bias, 1, 1, act)


class ConvBlock(nn.Module):
    """Conv3x3 + ELU (fused in the conv epilogue).  Reference layers.py:100-112."""

    def __init__(self, in_channels, out_channels):
        super(ConvBlock, self).__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.nonlin = nn.ELU(inplace=True)

    def forward(self, x, segments=None):
        return self.conv(x, act="elu", segments=segments)


def _pixel_grid(batch_size, height, width):
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32),
                         indexing="ij")
    id_coords = np.stack([xs, ys], 0)
    ones = torch.ones(batch_size, 1, height * width)
    pix = torch.from_numpy(id_coords.reshape(2, -1)).unsqueeze(0).repeat(batch_size, 1, 1)
    return torch.from_numpy(id_coords), ones, torch.cat([pix, ones], 1)


class BackprojectDepth(nn.Module):
    """Depth image -> homogeneous camera points [B,4,HW].  Reference layers.py:133-162.
    ``id_coords``, ``ones``, ``pix_coords`` are frozen Parameters sized by the ctor batch size."""

    def __init__(self, batch_size, height, width):
        super(BackprojectDepth, self).__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        idc, ones, pix = _pixel_grid(batch_size, height, width)
        self.id_coords = nn.Parameter(idc, requires_grad=False)
        self.ones = nn.Parameter(ones, requires_grad=False)
        self.pix_coords = nn.Parameter(pix, requires_grad=False)

    def forward(self, depth, inv_K):
        rays = torch.matmul(inv_K[:, :3, :3], self.pix_coords)
        pts = depth.view(self.batch_size, 1, -1) * rays
        return torch.cat([pts, self.ones], 1)


class Cat_xy(nn.Module):
    """Pseudo-3D xyz map: x/30, y/2, (z-40)/40.  Reference layers.py:165-201."""

    def __init__(self, batch_size, height, width):
        super(Cat_xy, self).__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        idc, ones, pix = _pixel_grid(batch_size, height, width)
        self.id_coords = nn.Parameter(idc, requires_grad=False)
        self.ones = nn.Parameter(ones, requires_grad=False)
        self.pix_coords = nn.Parameter(pix, requires_grad=False)

    def forward(self, depth, inv_K):
        rays = torch.matmul(inv_K[:, :3, :3], self.pix_coords)
        pts = (depth.view(self.batch_size, 1, -1) * rays).view(self.batch_size, 3, self.height, self.width)
        x, y, z = pts[:, 0:1] / 30.0, pts[:, 1:2] / 2.0, (pts[:, 2:3] - 40) / 40.0
        return torch.cat([x, y, z], 1)


class Project3D(nn.Module):
    """Camera points -> sampling grid normalised by (W-1),(H-1).  Reference layers.py:204-226."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super(Project3D, self).__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T):
        P = torch.matmul(K, T)[:, :3, :]
        cam = torch.matmul(P, points)
        uv = cam[:, :2, :] / (cam[:, 2, :].unsqueeze(1) + self.eps)
        uv = uv.view(self.batch_size, 2, self.height, self.width).permute(0, 2, 3, 1)
        scale = uv.new_tensor([self.width - 1, self.height - 1])
        return (uv / scale - 0.5) * 2


def upsample(x):
    """Nearest x2.  Reference layers.py:229-232."""
    return F.interpolate(x, scale_factor=2, mode="nearest")


def get_smooth_loss(disp, img):
    """Edge-aware first-order smoothness.  Reference layers.py:235-248."""
    dx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs()
    dy = (disp[:, :, :-1, :] - disp[:, :, 1:, :]).abs()
    wx = torch.exp(-(img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True))
    wy = torch.exp(-(img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, keepdim=True))
    return (dx * wx).mean() + (dy * wy).mean()


class SSIM(nn.Module):
    """3x3 SSIM loss over reflect-padded images.  Reference layers.py:251-281."""

    def __init__(self):
        super(SSIM, self).__init__()
        self.mu_x_pool = nn.AvgPool2d(3, 1)
        self.mu_y_pool = nn.AvgPool2d(3, 1)
        self.sig_x_pool = nn.AvgPool2d(3, 1)
        self.sig_y_pool = nn.AvgPool2d(3, 1)
        self.sig_xy_pool = nn.AvgPool2d(3, 1)
        self.refl = nn.ReflectionPad2d(1)
        self.C1, self.C2 = 0.01 ** 2, 0.03 ** 2

    def forward(self, x, y):
        x, y = self.refl(x), self.refl(y)
        mx, my = self.mu_x_pool(x), self.mu_y_pool(y)
        vx = self.sig_x_pool(x ** 2) - mx ** 2
        vy = self.sig_y_pool(y ** 2) - my ** 2
        vxy = self.sig_xy_pool(x * y) - mx * my
        num = (2 * mx * my + self.C1) * (2 * vxy + self.C2)
        den = (mx ** 2 + my ** 2 + self.C1) * (vx + vy + self.C2)
        return torch.clamp((1 - num / den) / 2, 0, 1)


def compute_depth_errors(gt, pred):
    """abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3.  Reference layers.py:284-302."""
    ratio = torch.max(gt / pred, pred / gt)
    a1, a2, a3 = ((ratio < 1.25 ** k).float().mean() for k in (1, 2, 3))
    err = gt - pred
    rmse = torch.sqrt((err ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = (err.abs() / gt).mean()
    sq_rel = (err ** 2 / gt).mean()
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3
