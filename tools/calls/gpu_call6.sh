#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
cat > /tmp/legs.py <<'PY'
import sys, json, os
sys.path.insert(0, os.getcwd())
import bench
from sphugo_b200 import build as B
B.build()
for name in sys.argv[1:]:
    for prec in (64,):
        r = bench.run_leg(name, prec, 10, 3, 0, fresh=(name == "speed"))
        print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items() if k in ('leg', 'dtype', 'ms_per_step', 'fallback_fraction', 'reuse_steps')}, flush=True)
PY
echo "== reuse off"; SPHB_REUSE=0 python /tmp/legs.py c3p c3u c4 c4dam speed speed
echo "== reuse on, period 4"; SPHB_REUSE_PERIOD=4 python /tmp/legs.py c3u c4 speed
echo "== probe c4 period 4"; SPHB_REUSE_PERIOD=4 timeout 300 python tools/reuse_probe.py --workload c4 --steps 9
