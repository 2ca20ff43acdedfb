#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/c37_bench_2gpu.json 2> gpurun_out/c37_bench_2gpu.err
tail -3 gpurun_out/c37_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c37_bench_2gpu.json').read().strip().split('\n')[-1]); print('2gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp'))
PY
