#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/c35_tests.log 2>&1
tail -5 gpurun_out/c35_tests.log
( time timeout 600 python bench.py --workload r50 --steps 5 --no-extras --no-cpu-baseline ) > gpurun_out/c35_bench_r50.json 2> gpurun_out/c35_bench_r50.err
( time timeout 600 python bench.py --workload refiner --steps 5 --no-extras --no-cpu-baseline ) > gpurun_out/c35_bench_refiner.json 2> gpurun_out/c35_bench_refiner.err
python - <<'PY'
import json
for f in ('c35_bench_r50','c35_bench_refiner'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_us'])
    except Exception as e:
        print(f,'ERR',e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
