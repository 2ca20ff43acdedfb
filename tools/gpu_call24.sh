#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_conv4.py > gpurun_out/c24_trace4.txt 2>&1
cat gpurun_out/c24_trace4.txt | head -150
