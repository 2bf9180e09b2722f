#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
cat > /tmp/legs.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
for name in sys.argv[1:]:
    r = bench.run_leg(name, 64, 20, 6, 0, fresh=(name == "speed"))
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items() if k in ('leg', 'ms_per_step', 'fallback_fraction')}, flush=True)
PY
python /tmp/legs.py c3u c4dam speed c3p c4
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
