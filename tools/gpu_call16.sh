#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_photoloss.py tests/test_gpu_fullsize.py tests/test_gpu_step.py tests/test_gpu_dropin.py tests/test_gpu_refiner.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c16_tests.log 2>&1
tail -6 gpurun_out/c16_tests.log
timeout 900 python bench.py --steps 10 --no-extras > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
python - <<'PY'
import json
for f in ('c16_bench',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_loss'], d['parity'])
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
