#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -3
done
for i in 1 2 3; do
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "test_conv2d_fwd_bwd and case10" 2>&1 | tail -1
done
