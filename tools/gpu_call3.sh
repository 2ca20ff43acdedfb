#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k conv2d > gpurun_out/c3_convtests.log 2>&1
echo "rc=$?" >> gpurun_out/c3_convtests.log
tail -3 gpurun_out/c3_convtests.log
FD_BENCH_TC_ONLY=1 timeout 300 python tools/bench_conv.py > gpurun_out/c3_bench_conv_tc3.txt 2>&1
FD_CONV_TC3=0 FD_BENCH_TC_ONLY=1 timeout 300 python tools/bench_conv.py > gpurun_out/c3_bench_conv_tc2.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c3_tests.log
tail -3 gpurun_out/c3_tests.log
FD_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 10 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench rc=$?" >> gpurun_out/c3_bench.err
FD_CONV_TC3=0 timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c3_bench_tc2.json 2> gpurun_out/c3_bench_tc2.err
FD_BENCH_VERBOSE=1 timeout 600 python bench.py --workload refiner --steps 10 --no-extras > gpurun_out/c3_bench_refiner.json 2> gpurun_out/c3_bench_refiner.err
echo "bench rc=$?" >> gpurun_out/c3_bench_refiner.err
FD_BENCH_VERBOSE=1 timeout 900 python bench.py --workload r50 --steps 5 --no-extras > gpurun_out/c3_bench_r50.json 2> gpurun_out/c3_bench_r50.err
echo "bench rc=$?" >> gpurun_out/c3_bench_r50.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c3_bench_ref.json 2> gpurun_out/c3_bench_ref.err
