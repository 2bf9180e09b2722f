#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
cat > /tmp/legs.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
for name in sys.argv[1:]:
    r = bench.run_leg(name, 64, 10, 3, 0, fresh=(name == "speed"))
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items() if k in ('leg', 'ms_per_step', 'fallback_fraction')}, flush=True)
PY
for M in 0.02 0.04 0.06; do for C in 47 56; do echo "margin $M cap $C"; SPHB_GUESS_MARGIN=$M SPHB_KNN_CAP=$C python /tmp/legs.py c3u c4dam speed c3p; done; done
