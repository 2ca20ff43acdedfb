#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_gdc.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25 | cut -c1-300
python - <<'PY'
import numpy as np, torch, time
from fusiondepth_b200 import gdc
import sys; sys.path.insert(0, 'tests')
from tests.make_golden import gdc_scene
pred, gt, calib_txt = gdc_scene(375, 1242, seed=2)
calib = (620.5, 168.25, 290.0 * 3.1, 290.0 * 3.1, 13.0 / (-290.0 * 3.1), 0.5 / (-290.0 * 3.1))
p, g = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, s = gdc.GDC(p, g, calib, k=10, W_tol=3e-5, recon_tol=5e-4, consider_range=(-0.1, 4.0), return_system=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print("GDC 375x1242: %d + %d points, %d CG iterations, %.1f ms" % (s.n_pl, s.n_l, s.iterations, 1e3 * (t1 - t0)))
PY
