#!/bin/bash
mkdir -p gpurun_out
for e in 0 1 2 3; do
FD_TC4_EXP=$e timeout 300 python tools/trace_conv4.py > gpurun_out/c26_trace4_exp$e.txt 2>&1
echo "== FD_TC4_EXP=$e"; grep "CTA exit" gpurun_out/c26_trace4_exp$e.txt; grep -A3 "epilogue" gpurun_out/c26_trace4_exp$e.txt | head -4
grep -A58 "mma " gpurun_out/c26_trace4_exp$e.txt | awk '/^    g/ {print $2, $6}' | head -54 | tr '\n' ' '; echo
done
