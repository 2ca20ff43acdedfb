#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
( time timeout 1200 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider ) > gpurun_out/c75_tests$i.log 2>&1
grep -E "passed|failed" gpurun_out/c75_tests$i.log | tail -1; grep -E "^FAILED|^/root.*Error" gpurun_out/c75_tests$i.log | head -5; grep real gpurun_out/c75_tests$i.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py ) > gpurun_out/c75_bench.json 2> gpurun_out/c75_bench.err
grep real gpurun_out/c75_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c75_bench.json').read().strip().split('\n')[-1])
print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['avg_us'], 'parity', d['parity']['ok'], d['parity']['rel_err'])
PY
