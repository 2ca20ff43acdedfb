#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c41_tests.log 2>&1
tail -3 gpurun_out/c41_tests.log
timeout 300 python tools/trace_conv4.py > gpurun_out/c41_trace4.txt 2>&1
grep "CTA exit" gpurun_out/c41_trace4.txt; grep -A3 "epilogue" gpurun_out/c41_trace4.txt | head -4
grep -A58 "mma " gpurun_out/c41_trace4.txt | awk '/^    g/ {print $2, $6}' | head -54 | tr '\n' ' '; echo
FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c41_bench_conv.txt 2>&1
head -12 gpurun_out/c41_bench_conv.txt
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c41_bench.json').read().strip().split('\n')[-1]); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['avg_us'], d['roofline']['frac'])
PY
tail -2 gpurun_out/c41_bench.err
