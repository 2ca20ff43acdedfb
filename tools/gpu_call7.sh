#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k conv2d > gpurun_out/c7_convtests.log 2>&1
tail -3 gpurun_out/c7_convtests.log
FD_CONV_TC3=0 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k conv2d > gpurun_out/c7_convtests_tc2.log 2>&1
tail -3 gpurun_out/c7_convtests_tc2.log
FD_BENCH_TC_ONLY=1 timeout 300 python tools/bench_conv.py > gpurun_out/c7_bench_conv_tc3.txt 2>&1
FD_CONV_TC3=0 FD_BENCH_TC_ONLY=1 timeout 300 python tools/bench_conv.py > gpurun_out/c7_bench_conv_tc2.txt 2>&1
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
FD_CONV_TC3=0 timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c7_bench_tc2.json 2> gpurun_out/c7_bench_tc2.err
python -c "
import json
for f in ('c7_bench','c7_bench_tc2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['serial_step_ms'], d['roofline']['dominant_kernel']['avg_us'])"
timeout 600 python tools/prof_timeline.py c7_timeline.csv > gpurun_out/c7_timeline.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c7_tests.log 2>&1
tail -6 gpurun_out/c7_tests.log
