"""ctypes front-end of the C LiDAR oracle (oracle/lidar_oracle.c).

TEST INFRASTRUCTURE ONLY -- never imported by fusiondepth_b200/.  Builds
``oracle/_build/liblidar_oracle.so`` with gcc on first use.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblidar_oracle.so")
_lib = None


def build() -> str:
    src = os.path.join(_HERE, "lidar_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/liblidar_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.fdo_depth_map.restype = ctypes.c_long
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def depth_map(points: np.ndarray, P: np.ndarray, W_im: int, H_im: int, vel_depth: bool = False,
              shape=None) -> np.ndarray:
    """generate_depth_map (kitti_utils.py:40-102) -> float64 map."""
    lib = _load()
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 4)
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
    d = np.empty((H_im, W_im), dtype=np.float64)
    lib.fdo_depth_map(_p(pts, ctypes.c_float), ctypes.c_long(pts.shape[0]), _p(P, ctypes.c_double),
                      int(W_im), int(H_im), int(vel_depth), _p(d, ctypes.c_double))
    if shape is None:
        return d
    sh, sw = int(shape[0]), int(shape[1])
    out_h = H_im + abs(sh - H_im) - (2 if sh < H_im else 0)
    out = np.empty((out_h, sw), dtype=np.float64)
    lib.fdo_pad(_p(d, ctypes.c_double), H_im, W_im, sh, sw, _p(out, ctypes.c_double))
    return out


def pool_scale(depth: np.ndarray) -> np.ndarray:
    """max_pool2d(2, ceil) -> float32 -> /100 (kitti_dataset.py:105-107, mono_dataset.py:196-198)."""
    lib = _load()
    d = np.ascontiguousarray(depth, dtype=np.float64)
    H, W = d.shape
    out = np.empty(((H + 1) // 2, (W + 1) // 2), dtype=np.float32)
    lib.fdo_pool_scale(_p(d, ctypes.c_double), H, W, _p(out, ctypes.c_float))
    return out


def two_channel(fourbeam: np.ndarray, window=(76, 190, 2, 638)) -> np.ndarray:
    """get_4beam_2channel (gen2channel.py:60-117) -> [2,H,W] float32."""
    lib = _load()
    fb = np.ascontiguousarray(fourbeam, dtype=np.float32)
    H, W = fb.shape
    out = np.empty((2, H, W), dtype=np.float32)
    lib.fdo_two_channel(_p(fb, ctypes.c_float), H, W, *[int(x) for x in window],
                        _p(out[0], ctypes.c_float), _p(out[1], ctypes.c_float))
    return out
