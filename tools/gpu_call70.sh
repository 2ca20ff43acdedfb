#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_graph_eager.py 16 full > gpurun_out/c70_diag_full.txt 2>&1
grep "^graph\|^eager" gpurun_out/c70_diag_full.txt | awk '{print $1, $3, $7}' | tr '\n' ';'; echo
grep -c "parameters above" gpurun_out/c70_diag_full.txt
tail -3 gpurun_out/c70_diag_full.txt
