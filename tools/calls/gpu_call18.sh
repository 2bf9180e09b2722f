#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q --tb=short -k "split_upload or by_id" > gpurun_out/pytest_split.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_split.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-legs --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5.json'))
print(d['ms_per_step'], d['e2e']['value'], d['e2e']['serial']['value'], d['roofline']['frac'], d['roofline']['step']['frac'])
PY
