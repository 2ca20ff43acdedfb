#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x -s > gpurun_out/c29_tests.log 2>&1
tail -3 gpurun_out/c29_tests.log; grep "^.conv (" gpurun_out/c29_tests.log | cut -c1-200
FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c29_bench_conv.txt 2>&1
FD_WGRAD2=0 FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c29_bench_conv_v1.txt 2>&1
paste -d'|' <(cut -c1-75 gpurun_out/c29_bench_conv.txt) <(cut -c55-80 gpurun_out/c29_bench_conv_v1.txt)
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c29_tests2.log 2>&1
tail -3 gpurun_out/c29_tests2.log
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c29_bench.json').read().strip().split('\n')[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['dominant_kernel'])
PY
