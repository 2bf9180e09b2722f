#!/bin/bash
# round 2, call 3: list-reuse path after the null-pointer fix; per-step probe; ncu of the two new kernels
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_reuse.py -q --tb=short > gpurun_out/pytest_reuse.log 2>&1; echo "reuse rc=$?"
tail -25 gpurun_out/pytest_reuse.log
SPHB_REUSE_PERIOD=9 timeout 300 python tools/reuse_probe.py --steps 20 > gpurun_out/reuse_probe.txt 2>&1
SPHB_REUSE_PERIOD=9 timeout 300 python tools/reuse_probe.py --steps 11 --precision 32 > gpurun_out/reuse_probe_f32.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-other-build > gpurun_out/bench_c5_reuse.json 2> gpurun_out/bench_c5_reuse.err; echo "bench rc=$?"
for P in 4 6; do SPHB_REUSE_PERIOD=$P timeout 300 python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu --no-other-build > gpurun_out/bench_c5_p$P.json 2> gpurun_out/bench_c5_p$P.err; done
SPHB_REUSE_PERIOD=4 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_knn_(tile|reuse)' -s 1 -c 2 -o gpurun_out/r02_knn_reuse \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other-build > gpurun_out/ncu_reuse.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
cat gpurun_out/reuse_probe.txt
for f in gpurun_out/bench_c5_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['ms_per_step'], d.get('knn_fallback_particles'), d.get('reuse'))"; done
