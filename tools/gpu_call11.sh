#!/bin/bash
mkdir -p gpurun_out
export FD_BENCH_VERBOSE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c11_bench2.json 2> gpurun_out/c11_bench2.err
echo "rc=$?" >> gpurun_out/c11_bench2.err
FD_BUCKETED_ALLREDUCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c11_bench2_single.json 2> gpurun_out/c11_bench2_single.err
echo "rc=$?" >> gpurun_out/c11_bench2_single.err
timeout 600 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c11_bench1.json 2> gpurun_out/c11_bench1.err
python - <<'PY'
import json
for f in ('c11_bench2','c11_bench2_single','c11_bench1'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp'), d['loss_first'], d['loss'])
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
