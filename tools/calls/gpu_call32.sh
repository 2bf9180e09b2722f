#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-legs --no-cpu --no-other-build"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_sub.log 2>&1; tail -2 gpurun_out/pytest_sub.log
for k in 64 32; do $B --precision $k > gpurun_out/bench_p$k.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/bench_p$k.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],3), d['roofline']['phases']['by_kind_ms'])"; done
