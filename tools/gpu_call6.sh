#!/bin/bash
mkdir -p gpurun_out
FD_FLAG_LIST=256,768,1280,1792 timeout 600 python tools/bench_conv_flags.py > gpurun_out/c6_flags.txt 2>&1
cat gpurun_out/c6_flags.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c6_tests.log 2>&1
tail -8 gpurun_out/c6_tests.log
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/c6_bench.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['roofline']['serial_step_ms'])"
