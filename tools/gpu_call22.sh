#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_conv3.py > gpurun_out/c22_trace3.txt 2>&1
FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c22_bench_conv.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c22_tests.log 2>&1
tail -5 gpurun_out/c22_tests.log
grep -v "kb " gpurun_out/c22_trace3.txt
grep "mma " -A 14 gpurun_out/c22_trace3.txt | head -50
cat gpurun_out/c22_bench_conv.txt
