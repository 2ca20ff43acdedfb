#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/c59_bench_${N}gpu.json 2> gpurun_out/c59_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c59_bench_${N}gpu.json').read().strip().split('\n')[-1]); print('${N}gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
