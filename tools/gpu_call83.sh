#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/c83_tests$i.log 2>&1
grep -E "passed|failed" gpurun_out/c83_tests$i.log | tail -1; grep -E "^FAILED|^E " gpurun_out/c83_tests$i.log | head -6
done
