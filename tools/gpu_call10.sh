#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_bn.py 2>&1 | tee gpurun_out/c10_bn.txt
FD_BN_FUSE_STATS=1 timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c10_bench_fuse.json 2> gpurun_out/c10_bench_fuse.err
python -c "
import json
for f in ('c10_bench_fuse',):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['serial_step_ms'], d['launches_per_step'])"
