#!/bin/bash
timeout 900 python tools/flaky_case10.py 2>&1 | tail -8
