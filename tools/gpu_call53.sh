#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "test_conv2d_fwd_bwd" > gpurun_out/c53_memcheck.log 2>&1
grep -E "passed|failed" gpurun_out/c53_memcheck.log | tail -1
grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/c53_memcheck.log | sort | uniq -c | head
