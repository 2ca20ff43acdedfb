"""CPU restatement of the reference's Graph-based Depth Correction (gdc_old.py:74-250, called from inf_gdc.py:81
with k=10, W_tol=3e-5, recon_tol=5e-4, method='cg').

TEST INFRASTRUCTURE ONLY -- nothing under fusiondepth_b200/ imports this; the product path is
fusiondepth_b200/gdc.py on csrc/gdc.cu.

The reference leans on two third-party pieces that are not under /root/reference: `pykdtree.kdtree.KDTree` (absent
from this image; exact k-nearest-neighbour search -- scipy's cKDTree returns the same neighbours) and
`scipy.sparse.linalg.cg` (SciPy 1.18 here; the reference's `tol=` keyword is the relative tolerance
||r|| <= tol ||b|| of the classic conjugate-gradient recurrence, restated below so that the iteration count is
explicit).  Pinned by tests/test_oracle_gdc.py against the reference's own GDC run under those two shims
(tests/make_golden.py gen_gdc).
"""
from __future__ import annotations

import numpy as np


def filter_mask(pc):
    """gdc_old.py:18-26"""
    return ((pc[:, 2] < 80) & (pc[:, 2] > 1) & (pc[:, 0] < 40) & (pc[:, 0] >= -40) & (pc[:, 1] < 2.5)
            & (pc[:, 1] >= -1))


def filter_theta_mask(pc, low, high):
    """gdc_old.py:54-62"""
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    d = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    theta = np.arcsin(y / d)
    return (theta >= low) & (theta < high)


def depth2ptc(depth, calib):
    """gdc_old.py:65-71 + kitti_util_from_pse.py:204-215.  calib: (c_u, c_v, f_u, f_v, b_x, b_y)."""
    c_u, c_v, f_u, f_v, b_x, b_y = calib
    rows, cols = depth.shape
    c, r = np.meshgrid(np.arange(cols), np.arange(rows))
    u, v, d = c.reshape(-1).astype(np.float64), r.reshape(-1).astype(np.float64), depth.reshape(-1).astype(np.float64)
    return np.stack([((u - c_u) * d) / f_u + b_x, ((v - c_v) * d) / f_v + b_y, d], 1)


def select_points(pred_depth, gt_depth, calib, consider_range=(-0.1, 3.0)):
    """gdc_old.py:114-168: the pseudo-LiDAR points to correct (pred_mask) and the anchors (gt_mask)."""
    ptc = depth2ptc(pred_depth, calib)
    ptc_gt = depth2ptc(gt_depth, calib)
    with np.errstate(invalid="ignore", divide="ignore"):
        consider_PL = (filter_mask(ptc) & filter_theta_mask(ptc, np.radians(consider_range[0]),
                                                            np.radians(consider_range[1]))).reshape(pred_depth.shape)
    consider_L = filter_mask(ptc_gt).reshape(gt_depth.shape)
    gt_mask = consider_L & consider_PL
    gt_mask[gt_mask] &= (np.abs(pred_depth[gt_mask] - gt_depth[gt_mask]) < 2)
    pred_mask = np.logical_not(gt_mask) & consider_PL
    return ptc, pred_mask, gt_mask


def knn(points, k):
    from scipy.spatial import cKDTree
    return cKDTree(points).query(points, k=k + 1)[1][:, 1:]


def local_weights(x_info, neighbors, k, W_tol):
    """gdc_old.py:174-184: per point, the constrained reconstruction weights of its depth from its neighbours'."""
    N = x_info.shape[0]
    As = np.zeros((N, k + 2, k + 2))
    bs = np.zeros((N, k + 2))
    As[:, :k, :k] = np.eye(k) * (1 + W_tol)
    As[:, k + 1, :k] = 1
    As[:, :k, k + 1] = 1
    bs[:, k + 1] = 1
    bs[:, k] = x_info
    As[:, k, :k] = x_info[neighbors]
    As[:, :k, k] = x_info[neighbors]
    return np.linalg.solve(As, bs[..., None])[:, :k, 0]


def build_system(W, neighbors, gt_info, N_PL, N_L):
    """gdc_old.py:196-222: A = [I - W_PLPL ; W_PLL], b = [W_LPL gt ; gt - W_LL gt] as scipy CSR blocks (the same
    storage and therefore the same floating-point summation order as the reference)."""
    from scipy.sparse import csr_matrix, eye as seye, vstack

    def block(rows, want_pl, ncol):
        nb, w = neighbors[rows], W[rows]
        idx = (nb < N_PL) if want_pl else (nb >= N_PL)
        indptr = np.concatenate(([0], np.cumsum(idx.sum(axis=1))))
        return csr_matrix((w[idx], nb[idx] - (0 if want_pl else N_PL), indptr), shape=(nb.shape[0], ncol))

    PL, L = slice(0, N_PL), slice(N_PL, N_PL + N_L)
    W_PLPL, W_LPL = block(PL, True, N_PL), block(PL, False, N_L)
    W_PLL, W_LL = block(L, True, N_PL), block(L, False, N_L)
    A = vstack((seye(N_PL) - W_PLPL, W_PLL)).tocsr()
    b = np.concatenate((W_LPL.dot(gt_info), gt_info - W_LL.dot(gt_info)))
    return A, b


def conjugate_gradient(matvec, b, x0, tol, maxiter=None):
    """scipy.sparse.linalg.cg without preconditioner: stop when ||r|| <= tol * ||b||."""
    x = x0.copy()
    r = b - matvec(x)
    atol = tol * np.linalg.norm(b)
    maxiter = maxiter or 10 * b.shape[0]
    p = None
    rho_prev = None
    for it in range(maxiter):
        if np.linalg.norm(r) < atol:
            return x, it
        rho = np.dot(r, r)
        if it > 0:
            p *= rho / rho_prev
            p += r
        else:
            p = r.copy()
        q = matvec(p)
        alpha = rho / np.dot(p, q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
    return x, maxiter


def GDC(pred_depth, gt_depth, calib, k=10, W_tol=1e-5, recon_tol=1e-4, consider_range=(-0.1, 3.0), details=False):
    """gdc_old.py:74-250 with method='cg', subsample=False.  gt_depth: -1 / <= 0 where there is no LiDAR return."""
    pred_depth = np.asarray(pred_depth, np.float64)
    gt_depth = np.asarray(gt_depth, np.float64)
    ptc, pred_mask, gt_mask = select_points(pred_depth, gt_depth, calib, consider_range)
    x_info = np.concatenate((pred_depth[pred_mask], pred_depth[gt_mask]))
    gt_info = gt_depth[gt_mask]
    N_PL, N_L = int(pred_mask.sum()), int(gt_mask.sum())
    pts = np.concatenate((ptc[pred_mask.reshape(-1)], ptc[gt_mask.reshape(-1)]))
    neighbors = knn(pts, k)
    W = local_weights(x_info, neighbors, k, W_tol)
    A, b = build_system(W, neighbors, gt_info, N_PL, N_L)
    x_new, iters = conjugate_gradient(lambda v: A.T.dot(A.dot(v)), A.T.dot(b), x_info[:N_PL], recon_tol)
    out = pred_depth.copy()
    out[pred_mask] = x_new
    out[gt_depth > 0] = gt_depth[gt_depth > 0]
    if details:
        return out, dict(N_PL=N_PL, N_L=N_L, neighbors=neighbors, W=W, iters=iters, pred_mask=pred_mask, gt_mask=gt_mask,
                         A=A, b=b, x_info=x_info, gt_info=gt_info, points=pts)
    return out
