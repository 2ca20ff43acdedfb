#!/bin/bash
run() { echo "== $*"; timeout 900 python -m pytest "$@" -m gpu -q --tb=line -p no:cacheprovider 2>&1 | grep -E "passed|failed|^/root|^/tmp" | tail -3; }
FD_STATS_ARENA=0 run tests/test_gpu_refiner.py tests/test_gpu_step.py
run tests/test_gpu_refiner.py::test_refine_step_graph_vs_oracle tests/test_gpu_step.py::test_cuda_graph_replay_matches_eager
run tests/test_gpu_refiner.py::test_refiner_step_vs_reference_fixture tests/test_gpu_step.py::test_cuda_graph_replay_matches_eager
run tests/test_gpu_refiner.py::test_r50_train_vs_reference_fixture tests/test_gpu_step.py::test_cuda_graph_replay_matches_eager
run tests/test_gpu_refiner.py::test_refine_pack_vs_reference_fixture tests/test_gpu_step.py::test_cuda_graph_replay_matches_eager
