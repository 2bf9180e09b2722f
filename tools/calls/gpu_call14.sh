#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5.json'))
print(d['ms_per_step'], d['reuse'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step'], d['other_build'])
print(d['roofline']['phases']['by_kind_ms'])
for l in d['legs']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in l.items() if k in ('leg','dtype','ms_per_step','step_roofline_frac','fallback_fraction','reuse_steps','error','ratio_vs_cpu_same_input')})
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_workload.py > gpurun_out/sanitizer.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer.txt
