#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c57_tests.log 2>&1
tail -3 gpurun_out/c57_tests.log
FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c57_bench_conv.txt 2>&1
sed -n 1p gpurun_out/c57_bench_conv.txt; sed -n 11,15p gpurun_out/c57_bench_conv.txt
timeout 900 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c57_bench.json 2> gpurun_out/c57_bench.err
timeout 600 python bench.py --workload r50 --steps 5 --no-extras --no-cpu-baseline > gpurun_out/c57_bench_r50.json 2> gpurun_out/c57_bench_r50.err
python - <<'PY'
import json
for f in ('c57_bench','c57_bench_r50'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['avg_us'])
PY
