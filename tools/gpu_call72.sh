#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - <<'PY'
import torch, bench
bench.select_workload("r18")
dev = torch.device("cuda:0")
for i in range(3):
    d = bench.dominant_kernel_leg(dev)
    print("dominant", round(d["avg_us"], 2), "us", round(d["achieved_tflops"], 1), "TF")
PY
