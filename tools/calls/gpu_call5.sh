#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
SPHB_REUSE_PERIOD=6 timeout 200 python tools/reuse_probe.py --steps 8 > gpurun_out/probe_p6.txt 2>&1; grep -E "build|reuse" gpurun_out/probe_p6.txt
SPHB_REUSE_PERIOD=6 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_knn_tile' -s 1 -c 1 -o gpurun_out/r02_knn_ext \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other-build --no-legs > gpurun_out/ncu_ext.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ring.py -q --tb=short > gpurun_out/pytest_ring.log 2>&1; echo "ring rc=$?"; tail -25 gpurun_out/pytest_ring.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-other-build > gpurun_out/bench_legs.json 2> gpurun_out/bench_legs.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_legs.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_legs.json'))
print(d['ms_per_step'], d['reuse'])
for l in d['legs']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in l.items() if k in ('leg','dtype','ms_per_step','step_roofline_frac','fallback_fraction','reuse_steps','error','ratio_vs_cpu_same_input')})
PY
