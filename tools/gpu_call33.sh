#!/bin/bash
mkdir -p gpurun_out
# launch list of one eager optimiser step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02b_launches.csv python bench.py --profile-step > gpurun_out/c33_launches.log 2>&1
# full counters: the conv kernels on three layer shapes (fwd / dgrad / wgrad), warm third iteration
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip 18 -c 9 -o gpurun_out/r02b_conv_full -f python tools/prof_conv.py > gpurun_out/c33_ncu_conv.log 2>&1
ncu -i gpurun_out/r02b_conv_full.ncu-rep --page raw --csv > gpurun_out/r02b_conv_full_raw.csv 2> /dev/null
ls -la gpurun_out | tail -6
