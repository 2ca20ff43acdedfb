#!/bin/bash
mkdir -p gpurun_out
FD_TC2_FLAGS=128 timeout 120 python tools/trace_conv.py > gpurun_out/c5_trace.txt 2>&1
FD_FLAG_LIST=256,258,260,264,272,276,284,320,336 timeout 600 python tools/bench_conv_flags.py > gpurun_out/c5_flags.txt 2>&1
cat gpurun_out/c5_flags.txt
