#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_timeline.py c15_timeline.csv > gpurun_out/c15_timeline.log 2>&1
# launch list of one eager optimiser step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python bench.py --profile-step > gpurun_out/c15_launches.log 2>&1
# full counters: the conv kernels on three layer shapes (fwd / dgrad / wgrad), warm third iteration
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip 18 -c 9 -o gpurun_out/r02_conv_full -f python tools/prof_conv.py > gpurun_out/c15_ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:photoloss -c 2 -o gpurun_out/r02_photoloss_full -f python bench.py --profile-step > gpurun_out/c15_ncu_pl.log 2>&1
ls -la gpurun_out | tail -12
