"""Diagnostic (not a test): repeats the conv op tests in-process to expose run-to-run variation."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import test_gpu_ops as T
fails = {}
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    for i, case in enumerate(T.CONV_CASES):
        try:
            T.test_conv2d_fwd_bwd(None, case)
        except Exception as e:  # noqa
            fails.setdefault(i, []).append(repr(e)[:300])
    for i, case in enumerate(T.TC_CASES):
        try:
            T.test_conv_tensor_core_vs_exact_fp32(None, case)
        except Exception as e:  # noqa
            fails.setdefault(100 + i, []).append(repr(e)[:300])
print("failures:", {k: (len(v), v[0]) for k, v in fails.items()})
