#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dataprep.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c38_tests.log 2>&1
tail -15 gpurun_out/c38_tests.log
python - <<'PY'
import torch, time
from fusiondepth_b200 import dataprep
x = torch.randint(0, 256, (36, 375, 1242, 3), dtype=torch.uint8, device="cuda")
pyr = dataprep.ColorPyramid(192, 640)
jit = pyr.sample_jitter(36)
flip = torch.rand(36) > 0.5
for _ in range(3): out = pyr(x, flip, jit)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = pyr(x, flip, jit)
e1.record(); torch.cuda.synchronize()
print("color pyramid + jitter + ToTensor for the 36 frames of one 12-image step: %.3f ms" % (e0.elapsed_time(e1) / 10))
PY
