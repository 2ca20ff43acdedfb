#!/bin/bash
FD_CONV_TC4=0 timeout 300 python tools/trace_conv3.py 2>&1 | grep -E "shape|phases|inside" | head -12
