#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_reuse.py -q --tb=short > gpurun_out/pytest_reuse.log 2>&1; echo "reuse rc=$?"; tail -5 gpurun_out/pytest_reuse.log
SPHB_REUSE_PERIOD=16 timeout 300 python tools/reuse_probe.py --steps 18 > gpurun_out/probe_local.txt 2>&1; grep -E "build|reuse|rror" gpurun_out/probe_local.txt
echo global; SPHB_REUSE_LOCAL=0 SPHB_REUSE_PERIOD=10 timeout 300 python tools/reuse_probe.py --steps 11 > gpurun_out/probe_global.txt 2>&1; grep -E "build|reuse|rror" gpurun_out/probe_global.txt | tail -5
cat > /tmp/legs.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
for prec in (64, 32):
    r = bench.run_leg("c5", prec, 42, 6, 0, flags=2)
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items() if k in ('leg', 'dtype', 'ms_per_step', 'fallback_fraction', 'reuse_steps')}, flush=True)
PY
python /tmp/legs.py
