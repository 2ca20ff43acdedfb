#!/bin/bash
for v in "" "FD_DGRAD_S2=0" "FD_BN_PARTS=296" "FD_CONV=cudacore"; do
echo "== $v"; env $v timeout 600 python -m pytest tests/test_gpu_refiner.py -m gpu -q --tb=line -p no:cacheprovider -k r50_train 2>&1 | grep -E "passed|failed|gnorm" | cut -c1-220 | tail -2
done
