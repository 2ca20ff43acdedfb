#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_timeline.py c36_timeline_r50.csv r50 > gpurun_out/c36_timeline.log 2>&1
tail -2 gpurun_out/c36_timeline.log
