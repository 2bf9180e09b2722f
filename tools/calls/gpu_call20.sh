#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c5.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --precision 32 --steps 20 --warmup 3 --no-legs --no-cpu --no-other-build > gpurun_out/bench_c5_f32.json 2> gpurun_out/bench_c5_f32.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['serial']['value'], d['roofline']['frac'], d['roofline']['step']['frac'], d['other_build']['ms_per_step'], d['cpu_baseline']['value'])
for l in d['legs']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in l.items() if k in ('leg','dtype','ms_per_step','step_roofline_frac','fallback_fraction','reuse_steps','error','ratio_vs_cpu_same_input')})
r=json.load(open('gpurun_out/bench_reference.json')); print('reference', r['value'], r['ms_per_step'], r['wall_s'])
f=json.load(open('gpurun_out/bench_c5_f32.json')); print('f32', f['ms_per_step'], f['e2e']['value'], f['roofline']['step']['frac'])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -q --tb=short -k "middle_of_a_run or split_upload" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_new.log
