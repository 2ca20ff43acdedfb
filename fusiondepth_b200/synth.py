"""Seeded synthetic inputs with the shapes/dtypes of the reference's data contract.

The reference's loader (datasets/mono_dataset.py:109-228) yields a dict keyed
("color",f,s), ("color_aug",f,s), ("K",s), ("inv_K",s), "4beam", "2channel",
("2channel",f,0).  There is no KITTI data (and no network) in the build image, so the
benchmark and the parity tests run on batches produced here.

Everything is generated on the CPU from a seeded generator so the oracle and the CUDA
path see identical bits; the LiDAR maps are produced by whichever implementation the
caller passes as ``lidar_fn`` (the CUDA kernels in the product, the C oracle in tests).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# KITTI 2011_09_26-like rectified calibration (public dataset constants)
CALIB = {
    "S_rect_02": np.array([1242.0, 375.0]),
    "P_rect_02": np.array([7.215377e+02, 0.0, 6.095593e+02, 4.485728e+01,
                           0.0, 7.215377e+02, 1.728540e+02, 2.163791e-01,
                           0.0, 0.0, 1.0, 2.745884e-03]),
    "P_rect_03": np.array([7.215377e+02, 0.0, 6.095593e+02, -3.395242e+02,
                           0.0, 7.215377e+02, 1.728540e+02, 2.199936e+00,
                           0.0, 0.0, 1.0, 2.729905e-03]),
    "R_rect_00": np.array([9.999239e-01, 9.837760e-03, -7.445048e-03,
                           -9.869795e-03, 9.999421e-01, -4.278459e-03,
                           7.402527e-03, 4.351614e-03, 9.999631e-01]),
    "R": np.array([7.533745e-03, -9.999714e-01, -6.166020e-04,
                   1.480249e-02, 7.280733e-04, -9.998902e-01,
                   9.998621e-01, 7.523790e-03, 1.480755e-02]),
    "T": np.array([-4.069766e-03, -7.631618e-02, -2.717806e-01]),
}


def velo_to_image_matrix(calib: Dict[str, np.ndarray] = CALIB, cam: int = 2) -> np.ndarray:
    """P_velo2im = P_rect_0{cam} @ R_rect_00(4x4) @ [R|T;0 0 0 1] in fp64
    (reference kitti_utils.py:44-56).  Returns [3,4] float64."""
    velo2cam = np.hstack((calib["R"].reshape(3, 3), calib["T"][..., np.newaxis]))
    velo2cam = np.vstack((velo2cam, np.array([0, 0, 0, 1.0])))
    R_cam2rect = np.eye(4)
    R_cam2rect[:3, :3] = calib["R_rect_00"].reshape(3, 3)
    P_rect = calib["P_rect_0" + str(cam)].reshape(3, 4)
    return np.dot(np.dot(P_rect, R_cam2rect), velo2cam)


def write_calib_files(calib_dir: str, calib: Dict[str, np.ndarray] = CALIB) -> None:
    """calib_cam_to_cam.txt / calib_velo_to_cam.txt in KITTI's text format (for running the
    reference's generate_depth_map when fixtures are generated)."""
    import os

    def line(k, v):
        return k + ": " + " ".join("%.12e" % x for x in np.asarray(v).ravel()) + "\n"

    with open(os.path.join(calib_dir, "calib_cam_to_cam.txt"), "w") as f:
        f.write("calib_time: 09-Jan-2012 13:57:47\n")
        for k in ("S_rect_02", "P_rect_02", "P_rect_03", "R_rect_00"):
            f.write(line(k, calib[k]))
    with open(os.path.join(calib_dir, "calib_velo_to_cam.txt"), "w") as f:
        f.write("calib_time: 15-Mar-2012 11:37:16\n")
        f.write(line("R", calib["R"]))
        f.write(line("T", calib["T"]))


def parse_roundtrip(calib: Dict[str, np.ndarray] = CALIB) -> Dict[str, np.ndarray]:
    """The calibration as the reference would read it back from the %.12e text files."""
    return {k: np.array([float("%.12e" % x) for x in np.asarray(v).ravel()]) for k, v in calib.items()}


def make_scan(seed: int, n_az: int = 1024,
              ring_deg: Sequence[float] = (1.0, -1.0, -3.0, -4.6),
              fov_deg: float = 45.0, piled: int = 0, rmin: float = 5.0, rmax: float = 60.0) -> np.ndarray:
    """A sparsified 4-ring velodyne scan [n,4] float32 (x fwd, y left, z up, reflectance),
    following sparsify.py `--H 64 --W 1024 --line_spec 2 7 12 16` (ring centres
    +1,-1,-3,-4.6 deg) and its x/y/z pre-filter (sparsify.py:98-104).  ``piled`` extra
    points are stacked on random existing azimuth bins to exercise duplicate handling."""
    rng = np.random.default_rng(seed)
    az = np.deg2rad(np.linspace(-fov_deg, fov_deg, n_az, endpoint=False))
    pts = []
    for e in ring_deg:
        r = rng.uniform(rmin, rmax, n_az)
        er = np.deg2rad(e + rng.uniform(-0.15, 0.15, n_az))
        pts.append(np.stack([r * np.cos(er) * np.cos(az), r * np.cos(er) * np.sin(az),
                             r * np.sin(er), rng.uniform(0, 1, n_az)], 1))
    p = np.concatenate(pts, 0)
    if piled:
        idx = rng.integers(0, p.shape[0], piled)
        dup = p[idx].copy()
        dup[:, :3] *= rng.uniform(0.7, 1.3, (piled, 1))
        p = np.concatenate([p, dup], 0)
    p = p[rng.permutation(p.shape[0])]
    keep = (p[:, 0] >= 0) & (p[:, 0] < 120) & (p[:, 1] >= -50) & (p[:, 1] < 50) & \
           (p[:, 2] >= -2.5) & (p[:, 2] < 1.5)
    return np.ascontiguousarray(p[keep].astype(np.float32))


def make_dense_scan(seed: int, n: int = 120000, edge_heavy: bool = False) -> np.ndarray:
    """A 64-beam-like dense scan (many same-pixel duplicates and image-edge hits)."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(4.0, 70.0, n)
    if edge_heavy:
        az = np.deg2rad(np.where(rng.uniform(size=n) < 0.5, rng.normal(-40.6, 0.4, n),
                                 rng.normal(40.3, 0.4, n)))
    else:
        az = np.deg2rad(rng.uniform(-50, 50, n))
    el = np.deg2rad(rng.uniform(-14.0, 3.0, n))
    p = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el),
                  rng.uniform(0, 1, n)], 1)
    return np.ascontiguousarray(p.astype(np.float32))


def intrinsics(h: int, w: int):
    """K of kitti_dataset.py:36-39 scaled like mono_dataset.py:166-175; inv_K = pinv(K)."""
    K = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    K[0, :] *= w
    K[1, :] *= h
    return torch.from_numpy(K), torch.from_numpy(np.linalg.pinv(K))


def _coherent_frames(B, H, W, g):
    """Smooth target + sources that are sub-pixel shifted copies with a little noise, so the
    warp gathers are spatially coherent as with real video."""
    low = torch.rand(B, 3, H // 8 + 3, W // 8 + 3, generator=g)
    big = F.interpolate(low, (H + 16, W + 16), mode="bicubic", align_corners=False).clamp(0, 1)
    tex = 0.08 * (torch.rand(B, 3, H + 16, W + 16, generator=g) - 0.5)
    big = (big + tex).clamp(0, 1)
    out = {}
    for f, (dy, dx) in {0: (8, 8), -1: (8, 5), 1: (9, 12)}.items():
        img = big[:, :, dy:dy + H, dx:dx + W]
        if f != 0:
            img = img + 0.01 * (torch.rand(B, 3, H, W, generator=g) - 0.5)
        out[f] = img.clamp(0, 1).contiguous()
    return out


def make_batch(B: int, H: int, W: int, seed: int = 1,
               lidar_fn: Optional[Callable[[np.ndarray], Dict[str, torch.Tensor]]] = None,
               frame_ids=(0, -1, 1), scales=(0, 1, 2, 3), mode: str = "uniform",
               with_noise: bool = True, lidar_density: float = 0.02, scan_range=(5.0, 60.0)) -> Dict:
    """One micro-batch keyed like the reference loader's output, all CPU fp32.

    ``lidar_fn(points[n,4] float32) -> {"4beam": [1,H,W], "2channel": [2,H,W]}``; when it
    is None a Bernoulli(lidar_density) stand-in is used (speed-only runs).
    ``noise[s]`` ([B,2,H,W] standard normal from the CPU generator) is what the reference
    draws at trainer.py:551 for the auto-mask tie-break.
    """
    g = torch.Generator().manual_seed(seed)
    inputs: Dict = {}
    if mode == "coherent":
        frames = _coherent_frames(B, H, W, g)
    else:
        frames = {f: torch.rand(B, 3, H, W, generator=g) for f in frame_ids}
    for f in frame_ids:
        for s in scales:
            img = frames[f] if s == 0 else F.avg_pool2d(frames[f], 2 ** s)
            inputs[("color", f, s)] = img.contiguous()
            inputs[("color_aug", f, s)] = img.clone()
    for s in scales:
        K, invK = intrinsics(H >> s, W >> s)
        inputs[("K", s)] = K.unsqueeze(0).repeat(B, 1, 1)
        inputs[("inv_K", s)] = invK.unsqueeze(0).repeat(B, 1, 1)
    four, two = [], {f: [] for f in frame_ids}
    for b in range(B):
        for f in frame_ids:
            if lidar_fn is not None:
                m = lidar_fn(make_scan(seed * 1000 + b * 7 + (f + 1), rmin=scan_range[0], rmax=scan_range[1]))
                fb, tc = m["4beam"], m["2channel"]
            else:
                mask = (torch.rand(1, H, W, generator=g) < lidar_density).float()
                fb = mask * (0.03 + 0.05 * torch.rand(1, H, W, generator=g))
                tc = torch.cat([fb, mask], 0)
            two[f].append(tc.float().cpu())
            if f == 0:
                four.append(fb.float().cpu().view(1, H, W))
    inputs["4beam"] = torch.stack(four, 0)
    inputs["2channel"] = torch.stack(two[0], 0)
    for f in frame_ids:
        inputs[("2channel", f, 0)] = torch.stack(two[f], 0)
    if with_noise:
        inputs["noise"] = {s: torch.randn(B, 2, H, W, generator=g) for s in scales}
    return inputs


def make_refiner_batch(B: int, H: int, W: int, seed: int = 1, **kw) -> Dict:
    """A stage-2 (refiner.py) micro-batch: the stage-1 batch + ``inf_gdc`` [B,H,W], the GDC-corrected
    depth the refiner clones (refiner.py:678-688).  With random-init weights the refined depth is
    ~0.2..2 m (no x26 in the refiner), so inf_gdc ~ U[0.05, 2.0] keeps the GDC mask populated
    (SURVEY.md section 8(c)/(d)); ~30 % of it is zero like real GDC output (pixels without a value)."""
    kw.setdefault("mode", "coherent")
    kw.setdefault("lidar_density", 0.05)
    inputs = make_batch(B, H, W, seed=seed, **kw)
    g = torch.Generator().manual_seed(seed + 17)
    gdc = 0.05 + 1.95 * torch.rand(B, H, W, generator=g)
    inputs["inf_gdc"] = gdc * (torch.rand(B, H, W, generator=g) > 0.3).float()
    return inputs


def to_device(inputs: Dict, device) -> Dict:
    out = {}
    for k, v in inputs.items():
        if isinstance(v, dict):
            out[k] = {kk: vv.to(device) for kk, vv in v.items()}
        else:
            out[k] = v.to(device)
    return out


def lidar_window(H: int, W: int):
    """Source window of get_4beam_2channel (gen2channel.py:64-65: rows 76..189, cols 2..637
    at 192x640).  The reference is hard-wired to 192x640; for other sizes (BASELINE config
    3, 320x1024) the window is scaled proportionally -- a documented extension."""
    if (H, W) == (192, 640):
        return 76, 190, 2, 638
    return int(round(76 * H / 192)), H - 2, 2, W - 2
