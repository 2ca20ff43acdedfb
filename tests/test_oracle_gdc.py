"""oracle/gdc_oracle.py against the reference's own GDC (gdc_old.py) run in the build container
(tests/make_golden.py gen_gdc -> tests/golden/gdc.npz)."""
import numpy as np

from oracle import gdc_oracle as G
from tests._util import GOLDEN


def test_gdc_oracle_matches_reference_fixture():
    g = np.load(GOLDEN + "/gdc.npz")
    for i in range(2):
        pred, gt, want = g["pred%d" % i], g["gt%d" % i], g["corrected%d" % i]
        got, info = G.GDC(pred, gt, tuple(g["calib%d" % i]), k=10, W_tol=3e-5, recon_tol=5e-4,
                          consider_range=tuple(g["range%d" % i]), details=True)
        assert info["N_PL"] > 5000 and info["N_L"] > 100
        changed = want != pred
        # same kNN graph, same CSR blocks, same conjugate-gradient recurrence and stopping rule: the same iterates
        assert np.abs(got - want).max() < 1e-9 * np.abs(want).max(), np.abs(got - want).max()
        assert np.abs(want - pred)[changed].mean() > 1e-2         # the correction is not a no-op
