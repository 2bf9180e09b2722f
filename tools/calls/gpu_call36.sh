#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_ring.py tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_ring.log 2>&1; tail -3 gpurun_out/pytest_ring.log
bash tools/calls/gpu_call25.sh
