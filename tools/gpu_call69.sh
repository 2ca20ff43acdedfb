#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c69_tests.log 2>&1
tail -4 gpurun_out/c69_tests.log
for i in 1 2 3; do
FD_BN_XMASK=1 timeout 300 python tools/diag_graph_eager.py 16 2>&1 | grep "^graph" | awk '{print $7}' | sort -g | tail -3 | tr '\n' ' '; echo
done
FD_STATS_ARENA=1 timeout 300 python tools/diag_graph_eager.py 16 2>&1 | grep "^graph" | awk '{print $7}' | sort -g | tail -3 | tr '\n' ' '; echo
timeout 600 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c69_bench.json 2> gpurun_out/c69_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c69_bench.json').read().strip().split('\n')[-1]); print('r18', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
PY
