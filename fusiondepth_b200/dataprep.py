"""On-device colour side of the reference's data producer.

Mirrors what ``MonoDataset.__getitem__`` / ``preprocess`` (datasets/mono_dataset.py:85-104, 156-206) do per
frame with PIL on the host -- flip of the native image, the ``Image.ANTIALIAS`` pyramid in which every scale is
resized from the previous one, ``ColorJitter``, ``ToTensor`` -- for a whole batch of uint8 frames resident in
HBM, bit-identical to Pillow / torchvision (kernels: csrc/dataprep.cu).  The LiDAR side of the producer is
``fusiondepth_b200.lidar`` (generate_depth_map -> max-pool -> 2-channel maps).

    pyr = ColorPyramid(192, 640)
    out = pyr(native_u8, flip=flip_mask, jitter=pyr.sample_jitter(B))   # native_u8 [B,Hn,Wn,3] uint8 cuda
    out[("color", 0)], out[("color_aug", 3)], ...                       # [B,3,h,w] float32, the loader's keys

Random parameters are drawn on the host exactly like torchvision's ``ColorJitter.get_params`` (one permutation
and four factors per call); the reference calls the transform once per frame and scale (mono_dataset.py:104),
so ``sample_jitter`` draws one set per (scale, image).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_cache: Dict[Tuple[int, int, str], Tuple[int, torch.Tensor, torch.Tensor]] = {}


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def lanczos_tables(in_size: int, out_size: int, device) -> Tuple[int, torch.Tensor, torch.Tensor]:
    """(ksize, bounds [out,2] int32, weights [out,ksize] int32) on `device`, cached per size pair."""
    key = (in_size, out_size, str(device))
    if key not in _cache:
        lib = _lib.load()
        ksize = lib.fd_lanczos_ksize(in_size, out_size)
        bounds = np.zeros((out_size, 2), np.int32)
        kk = np.zeros((out_size, ksize), np.int32)
        _lib.check(lib.fd_lanczos_coeffs(in_size, out_size, bounds.ctypes.data_as(ctypes.c_void_p),
                                         kk.ctypes.data_as(ctypes.c_void_p)), "fd_lanczos_coeffs")
        _cache[key] = (ksize, torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device))
    return _cache[key]


def _check_images(x: torch.Tensor):
    if not (x.is_cuda and x.dtype == torch.uint8 and x.dim() == 4 and x.shape[-1] == 3 and x.is_contiguous()):
        raise ValueError("expected a contiguous CUDA uint8 tensor [B,H,W,3]")


def resize_lanczos(x: torch.Tensor, height: int, width: int, flip: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PIL ``Image.resize((width, height), Image.LANCZOS)`` of every image of x [B,H,W,3] uint8 (after an optional
    per-image left-right flip, flip [B] bool/uint8)."""
    _check_images(x)
    lib = _lib.load()
    B, H, W, _ = x.shape
    kw, bw, cw = lanczos_tables(W, width, x.device)
    kh, bh, ch = lanczos_tables(H, height, x.device)
    tmp = torch.empty(B, H, width, 3, dtype=torch.uint8, device=x.device)
    out = torch.empty(B, height, width, 3, dtype=torch.uint8, device=x.device)
    fl = None
    if flip is not None:
        fl = flip.to(device=x.device, dtype=torch.uint8).contiguous()
        if width == W:
            raise ValueError("flip rides on the horizontal pass: flip at the first (size-changing) resize")
    _lib.check(lib.fd_resize_lanczos_u8(_ptr(x), _ptr(tmp), _ptr(out), B, H, W, height, width, _ptr(bw), _ptr(cw), kw,
                                        _ptr(bh), _ptr(ch), kh, _ptr(fl), _stream()), "fd_resize_lanczos_u8")
    return out


def color_jitter(x: torch.Tensor, order: torch.Tensor, factors: torch.Tensor) -> torch.Tensor:
    """torchvision ``ColorJitter`` with explicit parameters on x [B,H,W,3] uint8: order [B,4] int (permutation of
    0 brightness, 1 contrast, 2 saturation, 3 hue), factors [B,4] float indexed by op (hue as the hue factor in
    [-0.5, 0.5]; NaN = skip).  Returns a new tensor."""
    _check_images(x)
    lib = _lib.load()
    B, H, W, _ = x.shape
    f = factors.detach().to("cpu", torch.float64).clone()
    hue = f[:, 3]
    shift = torch.where(torch.isnan(hue), hue, ((hue * 255).to(torch.int32) & 255).to(torch.float64))
    f[:, 3] = shift                                   # np.int32(hue * 255).astype(np.uint8), computed in double
    out = x.clone()
    o = order.to(device=x.device, dtype=torch.int32).contiguous()
    fd = f.to(device=x.device, dtype=torch.float32).contiguous()
    ws = torch.zeros(B, dtype=torch.int64, device=x.device)
    _lib.check(lib.fd_color_jitter_u8(_ptr(out), B, H, W, _ptr(o), _ptr(fd), _ptr(ws), _stream()),
               "fd_color_jitter_u8")
    return out


def to_tensor(x: torch.Tensor) -> torch.Tensor:
    """``transforms.ToTensor``: [B,H,W,3] uint8 -> [B,3,H,W] float32 in [0,1]."""
    _check_images(x)
    B, H, W, _ = x.shape
    out = torch.empty(B, 3, H, W, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().fd_image_to_tensor(_ptr(x), _ptr(out), B, H, W, _stream()), "fd_image_to_tensor")
    return out


class ColorPyramid:
    """``MonoDataset.preprocess`` for a batch of frames (mono_dataset.py:85-104)."""

    def __init__(self, height: int, width: int, num_scales: int = 4, brightness=(0.8, 1.2), contrast=(0.8, 1.2),
                 saturation=(0.8, 1.2), hue=(-0.1, 0.1)):
        self.height, self.width, self.num_scales = height, width, num_scales
        self.ranges = (brightness, contrast, saturation, hue)        # mono_dataset.py:64-70

    def sample_jitter(self, B: int, generator: Optional[torch.Generator] = None):
        """One ColorJitter.get_params draw per (scale, image): (order [S,B,4] int32, factors [S,B,4] float64)."""
        S = self.num_scales
        order = torch.stack([torch.stack([torch.randperm(4, generator=generator) for _ in range(B)]) for _ in range(S)])
        lo = torch.tensor([r[0] for r in self.ranges], dtype=torch.float64)
        hi = torch.tensor([r[1] for r in self.ranges], dtype=torch.float64)
        u = torch.rand(S, B, 4, generator=generator, dtype=torch.float64)
        return order.to(torch.int32), lo + u * (hi - lo)

    def __call__(self, native: torch.Tensor, flip: Optional[torch.Tensor] = None,
                 jitter: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Dict:
        out = {}
        cur = native
        for s in range(self.num_scales):
            cur = resize_lanczos(cur, self.height >> s, self.width >> s, flip if s == 0 else None)
            out[("color", s)] = to_tensor(cur)
            if jitter is None:
                out[("color_aug", s)] = out[("color", s)]
            else:
                out[("color_aug", s)] = to_tensor(color_jitter(cur, jitter[0][s], jitter[1][s]))
        return out
