#!/bin/bash
mkdir -p gpurun_out
for p in 296 96 148 48; do
FD_BN_PARTS=$p timeout 900 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c51_bench_$p.json 2> gpurun_out/c51_bench_$p.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c51_bench_$p.json').read().strip().split('\n')[-1]); print('parts $p: ', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
PY
done
