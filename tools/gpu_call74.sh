#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c74_bench2.json 2> gpurun_out/c74_bench2.err
tail -2 gpurun_out/c74_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c74_bench2.json').read().strip().split('\n')[-1])
print('2gpu', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('dp'))
PY
