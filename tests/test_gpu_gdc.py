"""Graph-based Depth Correction kernels (csrc/gdc.cu) against the CPU oracle and the fixture
produced by the reference's own gdc_old.GDC (tests/golden/gdc.npz)."""
import numpy as np
import pytest
import torch

from oracle import gdc_oracle as G
from tests._util import GOLDEN

pytestmark = pytest.mark.gpu


def _scene(i):
    g = np.load(GOLDEN + "/gdc.npz")
    return g["pred%d" % i], g["gt%d" % i], tuple(g["calib%d" % i]), tuple(g["range%d" % i]), g["corrected%d" % i]


@pytest.mark.parametrize("i", [0, 1])
def test_gdc_stages_vs_oracle(cuda, i):
    """Point selection (exact), neighbour graph (same sets), local weights, the sparse operator and its transpose."""
    from fusiondepth_b200 import gdc
    pred, gt, calib, rng, _ = _scene(i)
    _, info = G.GDC(pred, gt, calib, 10, 3e-5, 5e-4, rng, details=True)
    s = gdc.GDCSystem(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), calib, 10, 3e-5, rng)
    assert (s.n_pl, s.n_l) == (info["N_PL"], info["N_L"])
    assert np.array_equal(s.idx_pl.cpu().numpy(), np.flatnonzero(info["pred_mask"].reshape(-1)))
    assert np.array_equal(s.idx_l.cpu().numpy(), np.flatnonzero(info["gt_mask"].reshape(-1)))
    assert np.abs(s.points.cpu().numpy() - info["points"]).max() < 1e-12
    ours, ref = np.sort(s.nbr.cpu().numpy(), 1), np.sort(info["neighbors"], 1)
    rows = np.flatnonzero((ours != ref).any(1))
    if rows.size:                                   # exact distance ties may be broken differently: same distances
        P = info["points"]
        for r in rows:
            d_o = np.sort(((P[s.nbr[r].cpu().numpy()] - P[r]) ** 2).sum(1))
            d_r = np.sort(((P[info["neighbors"][r]] - P[r]) ** 2).sum(1))
            assert np.allclose(d_o, d_r, rtol=0, atol=1e-12), r
    assert rows.size <= 0.001 * s.n
    nbr = s.nbr.cpu().numpy().astype(np.int64)
    W_ref = G.local_weights(info["x_info"], nbr, 10, 3e-5)
    assert np.abs(s.W.cpu().numpy() - W_ref).max() < 1e-8 * np.abs(W_ref).max()
    if rows.size == 0:
        A, b = G.build_system(W_ref, nbr, info["gt_info"], s.n_pl, s.n_l)
        assert np.abs(s.b.cpu().numpy() - b).max() < 1e-9 * np.abs(b).max()
        rs = np.random.RandomState(1)
        x, y = rs.randn(s.n_pl), rs.randn(s.n)
        Ax = s.apply(torch.from_numpy(x).cuda()).cpu().numpy()
        ATy = s.apply_t(torch.from_numpy(y).cuda(), torch.empty(s.n_pl, dtype=torch.float64, device="cuda")).cpu().numpy()
        assert np.abs(Ax - A.dot(x)).max() < 1e-10 * np.abs(A.dot(x)).max()
        assert np.abs(ATy - A.T.dot(y)).max() < 1e-10 * np.abs(A.T.dot(y)).max()


@pytest.mark.parametrize("i", [0, 1])
def test_gdc_vs_reference_fixture(cuda, i):
    """The corrected depth map against the reference's own output.  Conjugate gradients on the normal equations
    amplify rounding differences: two fp64 implementations of the same recurrence agree to ~1e-7 when they stop
    after ~40 iterations on scene 0 and only to the solver tolerance on scene 1 (the CPU oracle with a different
    summation order shows the same 0.05 m there), so the comparison is (i) at the stopping tolerance, (ii) the
    stopping rule itself, (iii) tightly at a tolerance where both have converged."""
    from fusiondepth_b200 import gdc
    pred, gt, calib, rng, want = _scene(i)
    got, s = gdc.GDC(pred, gt, calib, k=10, W_tol=3e-5, recon_tol=5e-4, method="cg", consider_range=rng,
                     return_system=True)
    assert got.shape == want.shape and got.dtype == np.float64
    _, info = G.GDC(pred, gt, calib, 10, 3e-5, 5e-4, rng, details=True)
    untouched = ~info["pred_mask"] & (gt <= 0)
    assert np.array_equal(got[untouched], pred[untouched])
    assert np.array_equal(got[gt > 0], gt[gt > 0])
    diff = np.abs(got - want)
    assert diff.max() < (1e-5 if i == 0 else 0.15) and diff.mean() < (1e-7 if i == 0 else 2e-3), (diff.max(), diff.mean())
    # (ii) the returned solution satisfies scipy's stopping rule, measured with the oracle's operator
    A, b = info["A"], info["b"]
    x = got[info["pred_mask"]]
    res = A.T.dot(A.dot(x) - b)
    assert np.linalg.norm(res) < 5e-4 * np.linalg.norm(A.T.dot(b)) * 1.05
    assert abs(s.iterations - info["iters"]) <= 3
    # (iii) scene 0: tightly, at a tolerance where both have converged.  Scene 1's normal equations are so ill
    # conditioned that two fp64 solutions with ||r|| <= 1e-8 ||b|| still differ by 0.05 m (1.5e-4 m at 1e-11, after
    # thousands of iterations): there the least-squares objective both minimise is compared instead.
    if i == 0:
        tight = gdc.GDC(pred, gt, calib, k=10, W_tol=3e-5, recon_tol=1e-10, method="cg", consider_range=rng)
        ref_tight = G.GDC(pred, gt, calib, 10, 3e-5, 1e-10, rng)
        assert np.abs(tight - ref_tight).max() < 1e-5, np.abs(tight - ref_tight).max()
    f_ours = np.linalg.norm(A.dot(x) - b) ** 2
    f_ref = np.linalg.norm(A.dot(want[info["pred_mask"]]) - b) ** 2
    f_start = np.linalg.norm(A.dot(pred[info["pred_mask"]]) - b) ** 2
    # (stopping one iteration earlier or later moves the objective by a few per cent at recon_tol = 5e-4: scene 1 ends
    # at 1.72 here and 1.65 in the reference, from 299 at the start)
    assert abs(f_ours - f_ref) < 0.15 * f_ref and f_ours < 0.05 * f_start, (f_ours, f_ref, f_start)
