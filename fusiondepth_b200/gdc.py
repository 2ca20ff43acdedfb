"""Graph-based Depth Correction on the device.

Drop-in for ``gdc_old.GDC`` (gdc_old.py:74-250) as the reference's offline stage calls it between stage 1 and
stage 2 (inf_gdc.py:81: ``GDC(pred_depth, gtd, calib, W_tol=3e-5, recon_tol=5e-4, k=10, method='cg',
consider_range=...)``): the predicted depth map is corrected towards the sparse LiDAR returns by (i) selecting the
pseudo-LiDAR points in range, (ii) a 10-nearest-neighbour graph over them and the LiDAR anchors, (iii) local
reconstruction weights from an (k+2)x(k+2) solve per point, (iv) conjugate gradients on the normal equations of
``[I - W_PLPL ; W_PLL] z = [W_LPL gt ; gt - W_LL gt]``.  Kernels: csrc/gdc.cu (fp64, like the reference).  PyTorch
here is plumbing: device buffers, index compaction / sort of the static graph, the host side of the iteration.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def calib_tuple(calib) -> Sequence[float]:
    """(c_u, c_v, f_u, f_v, b_x, b_y) from a kitti_util_from_pse.Calibration-like object or a 6-sequence."""
    if hasattr(calib, "c_u"):
        return (calib.c_u, calib.c_v, calib.f_u, calib.f_v, calib.b_x, calib.b_y)
    return tuple(float(v) for v in calib)


class GDCSystem:
    """The static part of one frame's correction: points, graph, weights, operator."""

    def __init__(self, pred: torch.Tensor, gt: torch.Tensor, calib, k: int, W_tol: float, consider_range):
        lib = _lib.load()
        H, W = pred.shape
        dev = pred.device
        self.k = k
        cal = (ctypes.c_double * 6)(*calib_tuple(calib))
        cls = torch.empty(H * W, dtype=torch.uint8, device=dev)
        pts = torch.empty(H * W, 3, dtype=torch.float64, device=dev)
        _lib.check(lib.fd_gdc_select(_p(pred), _p(gt), H, W, ctypes.cast(cal, ctypes.c_void_p),
                                     math.radians(consider_range[0]), math.radians(consider_range[1]), _p(cls), _p(pts),
                                     _st()), "fd_gdc_select")
        self.idx_pl = torch.nonzero(cls == 1).squeeze(1)          # raster order (gdc_old.py:163-168)
        self.idx_l = torch.nonzero(cls == 2).squeeze(1)
        self.n_pl, self.n_l = int(self.idx_pl.numel()), int(self.idx_l.numel())
        self.n = self.n_pl + self.n_l
        if self.n_pl == 0 or self.n <= k:
            raise ValueError("GDC: %d pseudo-LiDAR points and %d anchors in range" % (self.n_pl, self.n_l))
        order = torch.cat([self.idx_pl, self.idx_l])
        self.points = pts[order].contiguous()
        self.x_info = pred.reshape(-1)[order].contiguous()
        self.gt_info = gt.reshape(-1)[self.idx_l].contiguous()
        self.nbr = torch.empty(self.n, k, dtype=torch.int32, device=dev)
        _lib.check(lib.fd_gdc_knn(_p(self.points), self.n, k, _p(self.nbr), _st()), "fd_gdc_knn")
        self.W = torch.empty(self.n, k, dtype=torch.float64, device=dev)
        _lib.check(lib.fd_gdc_weights(_p(self.x_info), _p(self.nbr), self.n, k, float(W_tol), _p(self.W), _st()),
                   "fd_gdc_weights")
        # column-major entry list of A^T: (row, slot) pairs whose neighbour is a pseudo-LiDAR point
        flat = self.nbr.reshape(-1).long()
        ent = torch.nonzero(flat < self.n_pl).squeeze(1)
        cols, perm = torch.sort(flat[ent], stable=True)
        self.entries = ent[perm].contiguous()
        self.col_ptr = torch.searchsorted(cols, torch.arange(self.n_pl + 1, device=dev)).contiguous()
        self.b = torch.empty(self.n, dtype=torch.float64, device=dev)
        _lib.check(lib.fd_gdc_rhs(_p(self.W), _p(self.nbr), self.n, self.n_pl, k, _p(self.gt_info), _p(self.b), _st()),
                   "fd_gdc_rhs")
        self._t = torch.empty(self.n, dtype=torch.float64, device=dev)

    def apply(self, x, out=None):
        """A x"""
        out = self._t if out is None else out
        _lib.check(_lib.load().fd_gdc_apply(_p(self.W), _p(self.nbr), self.n, self.n_pl, self.k, _p(x), _p(out), _st()),
                   "fd_gdc_apply")
        return out

    def apply_t(self, y, out):
        """A^T y"""
        _lib.check(_lib.load().fd_gdc_apply_t(_p(self.W), _p(self.entries), _p(self.col_ptr), self.n_pl, self.k, _p(y),
                                              _p(out), _st()), "fd_gdc_apply_t")
        return out

    def normal(self, v, out):
        """A^T A v"""
        return self.apply_t(self.apply(v), out)

    def solve(self, recon_tol: float, maxiter: Optional[int] = None):
        """scipy.sparse.linalg.cg(A^T A, A^T b, x0 = predicted depths, tol = recon_tol) (gdc_old.py:224-228)."""
        lib = _lib.load()
        n = self.n_pl
        dev = self.b.device
        rhs = self.apply_t(self.b, torch.empty(n, dtype=torch.float64, device=dev))
        x = self.x_info[:n].clone()
        r = rhs - self.normal(x, torch.empty(n, dtype=torch.float64, device=dev))
        q = torch.empty(n, dtype=torch.float64, device=dev)
        scal = torch.zeros(4, dtype=torch.float64, device=dev)
        _lib.check(lib.fd_gdc_dot(_p(rhs), _p(rhs), n, _p(scal[3:]), _st()), "fd_gdc_dot")
        _lib.check(lib.fd_gdc_dot(_p(r), _p(r), n, _p(scal), _st()), "fd_gdc_dot")
        s = scal.tolist()
        atol2 = (recon_tol ** 2) * s[3]
        rho = s[0]
        p = r.clone()
        maxiter = maxiter or 10 * n
        it = 0
        while it < maxiter and not rho < atol2:
            self.normal(p, q)
            _lib.check(lib.fd_gdc_dot(_p(p), _p(q), n, _p(scal[1:]), _st()), "fd_gdc_dot")
            _lib.check(lib.fd_gdc_cg_update(_p(x), _p(r), _p(p), _p(q), n, _p(scal), _st()), "fd_gdc_cg_update")
            rho = float(scal[2])                                  # the host decides when to stop, as scipy does
            _lib.check(lib.fd_gdc_cg_dir(_p(p), _p(r), n, _p(scal), _st()), "fd_gdc_cg_dir")
            it += 1
        self.iterations = it
        return x


def GDC(pred_depth, gt_depth, calib, k=10, W_tol=1e-5, recon_tol=1e-4, verbose=False, method="cg",
        consider_range=(-0.1, 3.0), subsample=False, idx=0, maxiter=None, return_system=False):
    """gdc_old.GDC: the depth map after Graph-based Depth Correction.  ``gt_depth``: LiDAR depth, <= 0 (-1) where
    there is no return.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out (float64)."""
    if method != "cg":
        raise NotImplementedError("GDC: the offline stage runs method='cg' (inf_gdc.py:81); gmres is not provided")
    if subsample:
        raise NotImplementedError("GDC: subsample=True (gdc_old.py:34-51, np.random permutation) is not provided")
    as_numpy = not torch.is_tensor(pred_depth)
    pred = torch.as_tensor(np.asarray(pred_depth, np.float64) if as_numpy else pred_depth).to("cuda", torch.float64).contiguous()
    gt = torch.as_tensor(np.asarray(gt_depth, np.float64) if not torch.is_tensor(gt_depth) else gt_depth).to(
        "cuda", torch.float64).contiguous()
    system = GDCSystem(pred, gt, calib, int(k), float(W_tol), consider_range)
    x_new = system.solve(float(recon_tol), maxiter)
    out = pred.clone()
    out.view(-1)[system.idx_pl] = x_new
    out = torch.where(gt > 0, gt, out)
    res = out.cpu().numpy() if as_numpy else out
    return (res, system) if return_system else res
