#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_photoloss.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c34_tests.log 2>&1
tail -3 gpurun_out/c34_tests.log
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c34_bench.json 2> gpurun_out/c34_bench.err
python - <<'PY'
import json
for f in ('c34_bench',):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_loss']['ms_per_step'], d['roofline_loss']['frac'], d['roofline']['frac'], d['roofline']['family']['frac'])
PY
tail -3 gpurun_out/c34_bench.err
