#!/bin/bash
# round-2 call 1: new parity tests, BN-stats epilogue validation, baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c1_tests.log
timeout 300 python bench.py --steps 10 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench rc=$?" >> gpurun_out/c1_bench.err
FD_BN_FUSE_STATS=1 timeout 300 python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c1_tests_bnfuse.log 2>&1
FD_BN_FUSE_STATS=1 timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/c1_bench_bnfuse.json 2> gpurun_out/c1_bench_bnfuse.err
tail -5 gpurun_out/c1_tests.log
