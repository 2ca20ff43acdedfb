#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 600 python tools/flaky_conv.py 2>&1 | tail -8
echo "== FD_TC_FENCE=1"; FD_TC_FENCE=1 timeout 600 python tools/flaky_conv.py 2>&1 | tail -8
