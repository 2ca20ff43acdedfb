#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x -s > gpurun_out/c25_tests.log 2>&1
tail -3 gpurun_out/c25_tests.log; grep "^.conv (6" gpurun_out/c25_tests.log
timeout 300 python tools/trace_conv4.py > gpurun_out/c25_trace4.txt 2>&1
grep -B1 -A58 "mma " gpurun_out/c25_trace4.txt | head -64
grep -A4 "epilogue" gpurun_out/c25_trace4.txt
FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c25_bench_conv.txt 2>&1
head -12 gpurun_out/c25_bench_conv.txt
