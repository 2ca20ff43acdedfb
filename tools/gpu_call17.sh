#!/bin/bash
# session-2 baseline: whole GPU suite, the default bench line (all legs, as the driver runs it), the other workloads
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/c17_tests.log 2>&1
tail -5 gpurun_out/c17_tests.log
( time timeout 900 python bench.py ) > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
tail -3 gpurun_out/c17_bench.err
( time timeout 600 python bench.py --workload r50 --steps 5 --no-extras --no-cpu-baseline ) > gpurun_out/c17_bench_r50.json 2> gpurun_out/c17_bench_r50.err
( time timeout 600 python bench.py --workload refiner --steps 5 --no-extras --no-cpu-baseline ) > gpurun_out/c17_bench_refiner.json 2> gpurun_out/c17_bench_refiner.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/c17_bench_ref.json 2> gpurun_out/c17_bench_ref.err
tail -4 gpurun_out/c17_bench_r50.err gpurun_out/c17_bench_refiner.err gpurun_out/c17_bench_ref.err
cat gpurun_out/c17_bench.json | cut -c1-600
