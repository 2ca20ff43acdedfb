#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k conv2d 2>&1 | tail -2
FD_BENCH_TC_ONLY=1 timeout 300 python tools/bench_conv.py > gpurun_out/c9_bench_conv_tc3.txt 2>&1
FD_CONV_TC3=0 FD_BENCH_TC_ONLY=1 timeout 300 python tools/bench_conv.py > gpurun_out/c9_bench_conv_tc2.txt 2>&1
paste -d'\n' gpurun_out/c9_bench_conv_tc3.txt gpurun_out/c9_bench_conv_tc2.txt | head -24
FD_TC2_FLAGS=128 timeout 120 python tools/trace_conv.py 2>&1 | grep -A1 "^shape"
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
FD_CONV_TC3=0 timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c9_bench_tc2.json 2> gpurun_out/c9_bench_tc2.err
python -c "
import json
for f in ('c9_bench','c9_bench_tc2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['serial_step_ms'], d['roofline']['dominant_kernel']['avg_us'])"
