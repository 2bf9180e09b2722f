#!/bin/bash
# round 2, call 2: first run of the list-reuse path
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_reuse.py -q -x --tb=short > gpurun_out/pytest_reuse.log 2>&1; echo "reuse rc=$?"
tail -30 gpurun_out/pytest_reuse.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-other-build > gpurun_out/bench_c5_reuse.json 2> gpurun_out/bench_c5_reuse.err; echo "bench rc=$?"
SPHB_REUSE=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-other-build > gpurun_out/bench_c5_noreuse.json 2> gpurun_out/bench_c5_noreuse.err
for P in 4 6 8; do SPHB_REUSE_PERIOD=$P timeout 300 python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu --no-other-build > gpurun_out/bench_c5_p$P.json 2> gpurun_out/bench_c5_p$P.err; done
timeout 300 python tools/reuse_probe.py > gpurun_out/reuse_probe.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_c5_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['ms_per_step'], d.get('knn_fallback_particles'), d.get('reuse'), {k:round(v['ms'],3) for k,v in d['roofline']['phases'].items()})"; done
