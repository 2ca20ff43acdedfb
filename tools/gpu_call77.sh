#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload r50 > gpurun_out/c77_r50.json 2> gpurun_out/c77_r50.err
timeout 900 python bench.py --workload refiner > gpurun_out/c77_refiner.json 2> gpurun_out/c77_refiner.err
python - <<'PY'
import json
for n in ('r50','refiner'):
    d=json.loads(open('gpurun_out/c77_%s.json'%n).read().strip().split('\n')[-1])
    print(n, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d.get('parity'))
PY
