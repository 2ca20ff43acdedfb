#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py tests/test_gpu_refiner.py tests/test_gpu_dropin.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c54_tests.log 2>&1
tail -3 gpurun_out/c54_tests.log
timeout 900 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c54_bench.json 2> gpurun_out/c54_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c54_bench.json').read().strip().split('\n')[-1]); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_step'])
PY
