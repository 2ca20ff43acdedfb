#!/bin/bash
mkdir -p gpurun_out
for xm in 1 0; do
echo "== FD_BN_XMASK=$xm"
FD_BN_XMASK=$xm timeout 300 python tools/diag_graph_eager.py 3 2>&1 | tail -42
done
echo "== isolated test"
timeout 300 python -m pytest tests/test_gpu_step.py -m gpu -q --tb=line -p no:cacheprovider -k graph_replay 2>&1 | tail -3
