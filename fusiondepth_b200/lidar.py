"""Sparse-LiDAR leg of the hot path on the GPU (bit-exact with the reference's numpy/Python).

Mirrors
  * ``kitti_utils.generate_depth_map(calib_dir, velo_filename, cam, vel_depth, shape)``
    (reference kitti_utils.py:40-102) -- same arguments, returns the same float64 array;
  * ``KITTIRAWDataset.get_4beam`` + the ``/100`` of ``MonoDataset.__getitem__``
    (kitti_dataset.py:93-117, mono_dataset.py:194-198) -> ``four_beam``;
  * ``get_4beam_2channel`` (gen2channel.py:60-117) -> ``two_channel``.
Batched device entry points (``depth_maps``, ``four_beam``, ``two_channel``,
``lidar_inputs``) take many frames per launch.
"""
from __future__ import annotations

import os
from ctypes import c_void_p
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .synth import lidar_window


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def read_calib_file(path: str) -> Dict[str, np.ndarray]:
    """KITTI calibration text file -> dict (same parsing rule as reference kitti_utils.py:14-30)."""
    out = {}
    ok = set("0123456789.e+- ")
    with open(path, "r") as f:
        for line in f:
            if ":" not in line:
                continue
            key, value = line.split(":", 1)
            value = value.strip()
            out[key] = value
            if ok.issuperset(value):
                try:
                    out[key] = np.array([float(v) for v in value.split(" ")])
                except ValueError:
                    pass
    return out


def projection_from_calib(calib_dir: str, cam: int = 2):
    """(P_velo2im [3,4] float64, (W_im, H_im)) from calib_cam_to_cam.txt / calib_velo_to_cam.txt
    (reference kitti_utils.py:43-56)."""
    cam2cam = read_calib_file(os.path.join(calib_dir, "calib_cam_to_cam.txt"))
    velo2cam = read_calib_file(os.path.join(calib_dir, "calib_velo_to_cam.txt"))
    v2c = np.hstack((velo2cam["R"].reshape(3, 3), velo2cam["T"][..., np.newaxis]))
    v2c = np.vstack((v2c, np.array([0, 0, 0, 1.0])))
    im_shape = cam2cam["S_rect_02"][::-1].astype(np.int32)
    R = np.eye(4)
    R[:3, :3] = cam2cam["R_rect_00"].reshape(3, 3)
    P_rect = cam2cam["P_rect_0" + str(cam)].reshape(3, 4)
    return np.dot(np.dot(P_rect, R), v2c), (int(im_shape[1]), int(im_shape[0]))


def _out_shape(H_im, W_im, shape):
    if shape is None:
        return H_im, W_im
    sh, sw = int(shape[0]), int(shape[1])
    return H_im + abs(sh - H_im) - (2 if sh < H_im else 0), sw


def depth_maps(points: Sequence, P: Sequence[np.ndarray], W_im: int, H_im: int,
               vel_depth: bool = False, shape=None, device="cuda") -> torch.Tensor:
    """Batched generate_depth_map.  points: list of [n_i,4] float32 (numpy or tensor);
    P: list of [3,4] float64.  Returns [F,out_h,out_w] float64 on ``device``."""
    lib = _lib.load()
    F_ = len(points)
    pts = [torch.as_tensor(p, dtype=torch.float32).reshape(-1, 4) for p in points]
    offs = np.zeros(F_ + 1, dtype=np.int32)
    offs[1:] = np.cumsum([p.shape[0] for p in pts])
    allp = (torch.cat(pts, 0) if offs[-1] > 0 else torch.zeros(1, 4)).to(device).contiguous()
    offs_d = torch.from_numpy(offs).to(device)
    P_d = torch.from_numpy(np.ascontiguousarray(np.stack(P, 0), dtype=np.float64).reshape(F_, 12)).to(device)
    oh, ow = _out_shape(H_im, W_im, shape)
    out = torch.empty((F_, oh, ow), device=device, dtype=torch.float64)
    ws = torch.empty(lib.fd_lidar_workspace_bytes(F_, W_im, H_im), device=device, dtype=torch.uint8)
    sh, sw = (0, 0) if shape is None else (int(shape[0]), int(shape[1]))
    _lib.check(lib.fd_lidar_depth_map(_p(allp), _p(offs_d), F_, int(offs[-1]), _p(P_d), W_im, H_im,
                                      int(vel_depth), sh, sw, _p(out), oh, ow, _p(ws), _stream()),
               "fd_lidar_depth_map")
    return out


def generate_depth_map(calib_dir, velo_filename, cam=2, vel_depth=False, shape=None):
    """Drop-in for kitti_utils.generate_depth_map: returns a float64 numpy array."""
    P, (W_im, H_im) = projection_from_calib(calib_dir, cam)
    pts = np.fromfile(velo_filename, dtype=np.float32).reshape(-1, 4)
    return depth_maps([pts], [P], W_im, H_im, vel_depth, shape)[0].cpu().numpy()


def four_beam(depth: torch.Tensor) -> torch.Tensor:
    """[F,H,W] float64 -> [F,ceil(H/2),ceil(W/2)] float32: 2x2 ceil max-pool, cast, /100."""
    lib = _lib.load()
    depth = depth.contiguous()
    F_, H, W = depth.shape
    out = torch.empty((F_, (H + 1) // 2, (W + 1) // 2), device=depth.device, dtype=torch.float32)
    _lib.check(lib.fd_lidar_pool_scale(_p(depth), F_, H, W, _p(out), _stream()), "fd_lidar_pool_scale")
    return out


def two_channel(fourbeam: torch.Tensor, window=None) -> torch.Tensor:
    """[F,H,W] float32 -> [F,2,H,W] (expanded depth, confidence)."""
    lib = _lib.load()
    fb = fourbeam.contiguous()
    F_, H, W = fb.shape
    r0, r1, c0, c1 = window if window is not None else lidar_window(H, W)
    out = torch.empty((F_, 2, H, W), device=fb.device, dtype=torch.float32)
    _lib.check(lib.fd_two_channel(_p(fb), F_, H, W, r0, r1, c0, c1, _p(out), _stream()), "fd_two_channel")
    return out


def get_4beam_2channel(fourbeam, height=192, width=640, expand=2):
    """Drop-in for gen2channel.get_4beam_2channel: [H,W] tensor -> (expanded_depth, confidence)."""
    if expand != 2:
        raise NotImplementedError("the reference only ever uses expand=2")
    fb = torch.as_tensor(fourbeam, dtype=torch.float32).reshape(1, height, width).cuda()
    out = two_channel(fb)[0].cpu()
    return out[0], out[1]


def lidar_inputs(points: Sequence, P: Sequence[np.ndarray], W_im: int = 1242, H_im: int = 375,
                 shape=(384, 1280), device="cuda"):
    """points -> (4beam [F,1,h,w], 2channel [F,2,h,w]) in three launches (+ init)."""
    fb = four_beam(depth_maps(points, P, W_im, H_im, False, shape, device))
    return fb.unsqueeze(1), two_channel(fb)
