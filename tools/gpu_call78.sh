#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_timeline.py c78_timeline.csv > gpurun_out/c78_timeline.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02d_launches.csv python bench.py --profile-step > gpurun_out/c78_launches.log 2>&1
ls -la gpurun_out | grep -E "c78|r02d"
