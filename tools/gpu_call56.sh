#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider -k config3 ) 2>&1 | tail -25 | cut -c1-300
