#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/c50_bench.json 2> gpurun_out/c50_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c50_bench.json').read().strip().split('\n')[-1]); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
timeout 600 python tools/prof_timeline.py c50_timeline.csv > gpurun_out/c50_timeline.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches.csv python bench.py --profile-step > gpurun_out/c50_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip 18 -c 9 -o gpurun_out/r02c_conv_full -f python tools/prof_conv.py > gpurun_out/c50_ncu_conv.log 2>&1
ncu -i gpurun_out/r02c_conv_full.ncu-rep --page raw --csv > gpurun_out/r02c_conv_full_raw.csv 2> /dev/null
ls -la gpurun_out | tail -5
