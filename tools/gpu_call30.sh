#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c30_tests.log 2>&1
tail -3 gpurun_out/c30_tests.log
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err
FD_WGRAD_WAVES=1 timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c30_bench_w1.json 2> gpurun_out/c30_bench_w1.err
python - <<'PY'
import json
for f in ('c30_bench','c30_bench_w1'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
