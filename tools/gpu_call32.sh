#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c32_tests.log 2>&1
tail -3 gpurun_out/c32_tests.log
timeout 300 python tools/bench_bn.py > gpurun_out/c32_bn.txt 2>&1; tail -20 gpurun_out/c32_bn.txt
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c32_bench.json 2> gpurun_out/c32_bench.err
python - <<'PY'
import json
for f in ('c32_bench',):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
