#!/bin/bash
timeout 300 python tools/trace_conv4_gaps.py 2>&1 | tail -12
