#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/c44_stress.log
for i in $(seq 1 18); do
  timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/c44_run.log 2>&1
  tail -1 gpurun_out/c44_run.log >> gpurun_out/c44_stress.log
  if grep -q failed gpurun_out/c44_run.log; then grep -n "FAILED\|Error\|assert" gpurun_out/c44_run.log | head -8 >> gpurun_out/c44_stress.log; fi
done
cat gpurun_out/c44_stress.log
