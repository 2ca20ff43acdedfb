#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c14_tests.log 2>&1
tail -8 gpurun_out/c14_tests.log
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
python - <<'PY'
import json
for f in ('c14_bench',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['inference'])
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
