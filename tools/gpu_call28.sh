#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_timeline.py c31_timeline.csv > gpurun_out/c31_timeline.log 2>&1
tail -2 gpurun_out/c31_timeline.log
