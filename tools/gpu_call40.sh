#!/bin/bash
mkdir -p gpurun_out
for k in case18 case19 case20 case21; do
  timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "test_conv2d_fwd_bwd and $k" > gpurun_out/c40_$k.log 2>&1
  echo "$k: $(tail -1 gpurun_out/c40_$k.log)"
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "test_conv2d_fwd_bwd and case20" > gpurun_out/c40_sanitizer.log 2>&1
grep -m1 -A12 "Invalid\|Error:" gpurun_out/c40_sanitizer.log | head -40
