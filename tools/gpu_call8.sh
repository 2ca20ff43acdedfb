#!/bin/bash
mkdir -p gpurun_out
FD_TC2_FLAGS=128 timeout 120 python tools/trace_conv.py 2>&1 | grep -A2 "^shape" > gpurun_out/c8_trace.txt
cat gpurun_out/c8_trace.txt
FD_FLAG_LIST=256,768 timeout 600 python tools/bench_conv_flags.py 2>&1 | tee gpurun_out/c8_flags.txt
