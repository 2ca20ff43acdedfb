#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/c52_bench.json 2> gpurun_out/c52_bench.err
( time timeout 600 python bench.py --workload r50 --steps 5 --no-extras --no-cpu-baseline ) > gpurun_out/c52_bench_r50.json 2> gpurun_out/c52_bench_r50.err
( time timeout 600 python bench.py --workload refiner --steps 5 --no-extras --no-cpu-baseline ) > gpurun_out/c52_bench_refiner.json 2> gpurun_out/c52_bench_refiner.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/c52_bench_ref.json 2> gpurun_out/c52_bench_ref.err
python - <<'PY'
import json
for f in ('c52_bench','c52_bench_r50','c52_bench_refiner','c52_bench_ref'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('avg_us'), d.get('inference'))
    except Exception as e:
        print(f,'ERR',e); print(open('gpurun_out/%s.err'%f).read()[-800:])
PY
grep real gpurun_out/c52_*.err
