#!/bin/bash
mkdir -p gpurun_out
for fl in "" "serial" "nodirect" "nocache"; do
echo "== flags: $fl"
FD_BN_XMASK=1 timeout 300 python tools/diag_graph_eager.py 12 $fl 2>&1 | grep -v "^graph0 vs\|^  [0-9]" | tail -60
done
