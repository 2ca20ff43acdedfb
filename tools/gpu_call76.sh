#!/bin/bash
for fl in "mb2" "mb2 mid" "mid"; do
echo "== $fl"
timeout 300 python tools/diag_graph_eager.py 24 $fl 2>&1 | grep "^graph\|^eager" | awk '{print $1, $7}' | sort -k2 -g | tail -4 | tr '\n' ';'; echo
done
