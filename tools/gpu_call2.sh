#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c2_tests.log
FD_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 10 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
echo "bench rc=$?" >> gpurun_out/c2_bench.err
FD_BENCH_VERBOSE=1 timeout 600 python bench.py --workload refiner --steps 10 --no-extras > gpurun_out/c2_bench_refiner.json 2> gpurun_out/c2_bench_refiner.err
echo "bench rc=$?" >> gpurun_out/c2_bench_refiner.err
FD_BENCH_VERBOSE=1 timeout 900 python bench.py --workload r50 --steps 5 --no-extras > gpurun_out/c2_bench_r50.json 2> gpurun_out/c2_bench_r50.err
echo "bench rc=$?" >> gpurun_out/c2_bench_r50.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c2_bench_ref.json 2> gpurun_out/c2_bench_ref.err
FD_BENCH_TC_ONLY=1 timeout 300 python tools/bench_conv.py > gpurun_out/c2_bench_conv.txt 2>&1
tail -5 gpurun_out/c2_tests.log
