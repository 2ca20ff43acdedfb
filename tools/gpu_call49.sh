#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stem_im2col|maxpool|bn_apply|bn_bwd" --launch-skip 0 -c 12 -o gpurun_out/r02b_stem_full -f python tools/prof_stem.py > gpurun_out/c49_ncu.log 2>&1
ncu -i gpurun_out/r02b_stem_full.ncu-rep --page raw --csv > gpurun_out/r02b_stem_full_raw.csv 2>/dev/null
tail -3 gpurun_out/c49_ncu.log
