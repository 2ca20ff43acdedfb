#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_conv3.py > gpurun_out/c18_trace3.txt 2>&1
FD_BENCH_TC_ONLY=1 timeout 600 python tools/bench_conv.py > gpurun_out/c18_bench_conv.txt 2>&1
cat gpurun_out/c18_trace3.txt | head -120
cat gpurun_out/c18_bench_conv.txt
