"""Drop-in mirror of the reference's ``layers`` module surface (reference layers.py:1-302).

Same public names, signatures and results, so ``from layers import *`` in the unchanged
trainer.py / refiner.py / completor.py / evaluate_depth.py resolves here (the drivers also pick
up ``F``, ``nn``, ``np`` and ``torch`` from this namespace -- SURVEY.md section 0).

Everything with arithmetic in it runs on the package's CUDA kernels: ``ConvBlock`` / ``Conv3x3`` (conv
kernels), ``transformation_from_parameters`` (fd_pose_matrix_*), ``BackprojectDepth`` / ``Project3D`` /
``Cat_xy`` / ``SSIM`` (csrc/geometry.cu), and -- through the ``F`` proxy below -- the drivers' own
``F.interpolate(bilinear)`` and ``F.grid_sample(border)`` calls.  So an UNPATCHED reference driver runs its
whole loss chain on this library, op by op; ``training.patch_trainer`` swaps that chain for the fused
``fd_photoloss_fwd/bwd`` pair, which is what the benchmark times.  There is no CPU fallback: the kernels
reject CPU tensors.
"""
from __future__ import absolute_import, division, print_function

import types as _types

import numpy as np

import torch
import torch.nn as nn
import torch.nn.functional as _torch_F

from . import ops as _ops


class _FunctionalProxy(_types.ModuleType):
    """``layers.F``: torch.nn.functional with the hot-path signatures routed to this library's kernels.

    The drivers do ``from layers import *`` after their own ``import torch.nn.functional as F`` and
    layers.py has no ``__all__``, so the ``F`` they end up using is *this* name.
    Only the exact call shapes of the per-step path are intercepted -- bilinear ``interpolate`` with
    ``align_corners=False`` to an explicit size (trainer.py:434, 579; refiner.py:325, 335, 680) and
    ``grid_sample(padding_mode="border")`` (trainer.py:467) on 4-D fp32 CUDA tensors; every other
    attribute is torch.nn.functional's own."""

    def __getattr__(self, name):
        return getattr(_torch_F, name)

    @staticmethod
    def _ours(t):
        return isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4

    def interpolate(self, input, size=None, scale_factor=None, mode="nearest", align_corners=None, **kw):
        if (mode == "bilinear" and not align_corners and size is not None and scale_factor is None
                and not kw and self._ours(input)):
            size = tuple(int(v) for v in (size if isinstance(size, (list, tuple, torch.Size)) else (size, size)))
            return _ops.upsample_bilinear(input, size[0], size[1])
        return _torch_F.interpolate(input, size=size, scale_factor=scale_factor, mode=mode,
                                    align_corners=align_corners, **kw)

    def grid_sample(self, input, grid, mode="bilinear", padding_mode="zeros", align_corners=None):
        if mode == "bilinear" and padding_mode == "border" and not align_corners and self._ours(input) \
                and self._ours(grid):
            return _ops.grid_sample_border(input, grid)
        return _torch_F.grid_sample(input, grid, mode=mode, padding_mode=padding_mode,
                                    align_corners=align_corners)


F = _FunctionalProxy("fusiondepth_b200.layers.F")


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
    if axisangle.is_cuda:
        return _ops.pose_matrix(axisangle, translation, invert)       # one kernel each way
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

    def forward(self, x, act="none", segments=None, pad_out=False):
        """``segments`` (internal): un-assembled inputs [(a, b|None, upsample)], fused with the pad.

        Channel counts that are not multiples of 32 (the refine2d decoder's 262 / 134 / 102 / 22) are
        zero-padded up to the next multiple so that the layer runs on the tensor-core kernels: the input by
        an all-zero segment of the same gather, the weight by ops.pad_conv_params.  With ``pad_out`` the
        OUTPUT keeps its zero-padded channels (they are exactly act(0) = 0 for ELU / ReLU) and the next
        Conv3x3 consumes them as already-padded input."""
        segs = list(segments) if segments is not None else [(x, None, False)]
        w, b = self.conv.weight, self.conv.bias
        cout, cin_w = w.shape[0], w.shape[1]
        cin = sum(int(a.shape[1]) for a, _, _ in segs)
        cin_p = cin
        if cout != 1 and cin > 16 and cin % 32 != 0 and _ops.CONV_BACKEND == "tc":
            cin_p = (cin + 31) // 32 * 32
            a0, _, up0 = segs[0]
            B = a0.shape[0]
            H, W = a0.shape[2] * (2 if up0 else 1), a0.shape[3] * (2 if up0 else 1)
            segs.append((_ops.zero_segment(B, cin_p - cin, H, W, a0.device), None, False))
        cout_p = cout
        if pad_out and cout > 16 and cout % 32 != 0 and _ops.CONV_BACKEND == "tc":
            if act not in ("elu", "relu", "none"):
                raise ValueError("pad_out needs an activation with act(0) == 0")
            cout_p = (cout + 31) // 32 * 32
        if cin_p != cin_w or cout_p != cout:
            if cin_p < cin_w:
                raise RuntimeError("Conv3x3: input has %d channels, weight expects %d" % (cin, cin_w))
            w, b = _ops.pad_conv_params(w, b, cout_p, cin_p)
        if self.use_refl:
            xp = _ops.assemble(segs, pad=1)
            return _ops.conv2d(xp, w, b, 1, 0, act)
        if segments is not None or len(segs) > 1:
            x = _ops.assemble(segs, pad=0)
        return _ops.conv2d(x, w, b, 1, 1, act)


class ConvBlock(nn.Module):
    """Conv3x3 + ELU (fused in the conv epilogue).  Reference layers.py:100-112."""

    def __init__(self, in_channels, out_channels):
        super(ConvBlock, self).__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.nonlin = nn.ELU(inplace=True)

    def forward(self, x, segments=None, pad_out=False):
        return self.conv(x, act="elu", segments=segments, pad_out=pad_out)


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
        if depth.shape[0] != self.batch_size:
            raise RuntimeError("BackprojectDepth was built for batch %d, got %d" % (self.batch_size, depth.shape[0]))
        return _ops.backproject(depth, inv_K, self.height, self.width)


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
        if depth.shape[0] != self.batch_size:
            raise RuntimeError("Cat_xy was built for batch %d, got %d" % (self.batch_size, depth.shape[0]))
        return _ops.cat_xy(depth, inv_K, self.height, self.width)


class Project3D(nn.Module):
    """Camera points -> sampling grid normalised by (W-1),(H-1).  Reference layers.py:204-226."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super(Project3D, self).__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T):
        if points.shape[0] != self.batch_size:
            raise RuntimeError("Project3D was built for batch %d, got %d" % (self.batch_size, points.shape[0]))
        return _ops.project3d(points, K, T, self.height, self.width, self.eps)


def upsample(x):
    """Nearest x2.  Reference layers.py:229-232.  (The decoder of this package never calls it: the upsample is
    folded into the next convolution's gather, fd_assemble_fwd.)"""
    return _torch_F.interpolate(x, scale_factor=2, mode="nearest")


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
        return _ops.ssim(x, y)


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
