#!/bin/bash
# BatchNorm backward with the ReLU mask rebuilt from x (FD_BN_XMASK, default on): tests + A/B bench on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c66_tests.log 2>&1
tail -3 gpurun_out/c66_tests.log
for rep in 1 2; do
for xm in 0 1; do
FD_BN_XMASK=$xm timeout 600 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c66_bench_x$xm.json 2> gpurun_out/c66_bench_x$xm.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c66_bench_x$xm.json').read().strip().split('\n')[-1]); print('r18 xmask=$xm', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
PY
done
done
for xm in 0 1; do
FD_BN_XMASK=$xm timeout 600 python bench.py --workload r50 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c66_r50_x$xm.json 2> gpurun_out/c66_r50_x$xm.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c66_r50_x$xm.json').read().strip().split('\n')[-1]); print('r50 xmask=$xm', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
PY
done
