#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
( time timeout 1200 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider ) > gpurun_out/c46_tests$i.log 2>&1
grep -E "passed|failed" gpurun_out/c46_tests$i.log | tail -1; grep -E "^FAILED|^/root.*Error" gpurun_out/c46_tests$i.log | head -5
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/c46_bench.json 2> gpurun_out/c46_bench.err
tail -4 gpurun_out/c46_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c46_bench.json').read().strip().split('\n')[-1])
print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['avg_us'], 'family', d['roofline']['family']['frac'], 'loss', d['roofline_loss']['frac'])
print('cpu', d['cpu_baseline']); print('parity', d['parity']); print('pytorch_gpu', d.get('pytorch_gpu')); print('fast', {k: d['fast_mode'][k] for k in ('value','e2e')}); print('inference', d['inference']); print('lidar', d['roofline_lidar']['frac'], d['roofline_lidar']['us_per_batch'])
PY
