#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
cat > /tmp/legs.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
for name in sys.argv[1:]:
    bench.run_leg(name, 64, 20, 6, 0, fresh=(name == "speed"))
    r = bench.run_leg(name, 64, 20, 6, 0, fresh=(name == "speed"))
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items() if k in ('leg', 'ms_per_step', 'fallback_fraction', 'wall_ms_per_step')}, flush=True)
PY
echo default; python /tmp/legs.py speed c3p
echo norecord; SPHB_NO_RECORD=1 python /tmp/legs.py speed c3p
echo fixedlevel; SPHB_GUESS_MARGIN=0.02 python /tmp/legs.py speed c3p
