#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -15 | cut -c1-300
