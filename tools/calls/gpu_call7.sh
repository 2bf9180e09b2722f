#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ring_nccl_check.py > gpurun_out/ring_nccl.log 2>&1; echo "nccl ring rc=$?"; tail -5 gpurun_out/ring_nccl.log
timeout 900 python -m pytest tests/test_gpu_ring.py tests/test_gpu_reuse.py tests/test_multi_gpu.py -q --tb=short > gpurun_out/pytest_ring.log 2>&1; echo "ring rc=$?"; tail -15 gpurun_out/pytest_ring.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 rc=$?"; tail -3 gpurun_out/bench_g2.err; cut -c1-400 gpurun_out/bench_g2.json
SPHB_REUSE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_g2_noreuse.json 2> gpurun_out/bench_g2_noreuse.err; cut -c1-400 gpurun_out/bench_g2_noreuse.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-other-build > gpurun_out/bench_legs.json 2> gpurun_out/bench_legs.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_legs.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_legs.json'))
print(d['ms_per_step'], d['reuse'])
for l in d['legs']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in l.items() if k in ('leg','dtype','ms_per_step','step_roofline_frac','fallback_fraction','reuse_steps','error','ratio_vs_cpu_same_input')})
for f in ('bench_g2','bench_g2_noreuse'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['reuse'], d['phases']['device_ms'])
PY
