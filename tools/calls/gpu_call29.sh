#!/bin/bash
# round 2: bulk-staged force kernel with the L2 prefetch: parity subset, three bench passes (noise), ncu --set full of it
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-legs --no-cpu --no-other-build"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ring.py tests/test_gpu_slab.py -m gpu -q > gpurun_out/pytest_sub.log 2>&1; tail -2 gpurun_out/pytest_sub.log
for k in 1 2 3; do $B > gpurun_out/bench_p$k.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/bench_p$k.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],3), d['roofline']['phases']['by_kind_ms'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_force_st' -s 3 -c 1 -o gpurun_out/r02_full_force_bulk python bench.py --steps 2 --warmup 3 --no-e2e --no-legs --no-cpu --no-other-build > gpurun_out/ncu_fb.log 2>&1
ls -la gpurun_out/*.ncu-rep
