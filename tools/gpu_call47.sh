#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c48_tests.log 2>&1
tail -4 gpurun_out/c48_tests.log; grep -n "^E  .*assert\|^FAILED" gpurun_out/c48_tests.log | head
FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c48_bench_conv.txt 2>&1
FD_DGRAD_S2=0 FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c48_bench_conv_off.txt 2>&1
paste <(awk '{print $1,$2, "bwd", $(NF-1)}' gpurun_out/c48_bench_conv.txt) <(awk '{print $(NF-1)}' gpurun_out/c48_bench_conv_off.txt) | head -8
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c48_tests2.log 2>&1
tail -3 gpurun_out/c48_tests2.log
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c48_bench.json 2> gpurun_out/c48_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c48_bench.json').read().strip().split('\n')[-1]); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
PY
