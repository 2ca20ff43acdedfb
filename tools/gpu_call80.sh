#!/bin/bash
mkdir -p gpurun_out
FD_CONV_TC4=0 timeout 300 python tools/trace_conv3.py 2>&1 | grep -E "shape|phases|inside" | head -6
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c80_tests.log 2>&1
tail -2 gpurun_out/c80_tests.log
timeout 300 python - <<'PY'
import torch, bench
bench.select_workload("r18")
d = bench.dominant_kernel_leg(torch.device("cuda:0"))
print("dominant", round(d["avg_us"], 2), "us")
PY
for i in 1 2; do
timeout 600 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c80_bench.json 2> gpurun_out/c80_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c80_bench.json').read().strip().split('\n')[-1]); print('r18', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
PY
done
