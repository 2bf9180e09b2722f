#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python bench.py --steps 28 --warmup 3 --no-e2e --no-other-build --no-cpu > gpurun_out/bench_legs.json 2> gpurun_out/bench_legs.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_legs.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_legs.json'))
print(d['ms_per_step'], d['reuse'])
for l in d['legs']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in l.items() if k in ('leg','dtype','ms_per_step','step_roofline_frac','fallback_fraction','reuse_steps','error','ratio_vs_cpu_same_input')})
PY
timeout 200 python tools/reuse_probe.py --steps 16 > gpurun_out/probe.txt 2>&1; grep -E "build|reuse|rror" gpurun_out/probe.txt
timeout 900 python -m pytest tests/test_gpu_reuse.py tests/test_gpu_ring.py -q --tb=short > gpurun_out/pytest_reuse.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_reuse.log
