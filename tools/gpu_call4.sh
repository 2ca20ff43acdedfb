#!/bin/bash
mkdir -p gpurun_out
# which kernels run, and the full counters of the patch kernel vs conv_tc2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip 18 -c 9 -o gpurun_out/c4_tc3 -f python tools/prof_conv.py > gpurun_out/c4_ncu_tc3.log 2>&1
FD_CONV_TC3=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip 18 -c 9 -o gpurun_out/c4_tc2 -f python tools/prof_conv.py > gpurun_out/c4_ncu_tc2.log 2>&1
ls -la gpurun_out/*.ncu-rep
