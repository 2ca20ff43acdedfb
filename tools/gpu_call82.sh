#!/bin/bash
# flake census: the whole GPU suite five times on one box, then smoke + the default bench
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
timeout 600 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/c82_tests$i.log 2>&1
grep -E "passed|failed" gpurun_out/c82_tests$i.log | tail -1; grep -E "^FAILED|^/root.*Error|^E " gpurun_out/c82_tests$i.log | head -6
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/c82_bench.json 2> gpurun_out/c82_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c82_bench.json').read().strip().split('\n')[-1])
print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['avg_us'], 'parity', d['parity']['ok'], d['parity']['rel_err'])
PY
