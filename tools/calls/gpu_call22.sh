#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-legs --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print(d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['roofline']['phases'].items() if isinstance(v,dict) and 'ms' in v}, d['other_build']['ms_per_step'], d['knn_fallback_particles'])
PY
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
