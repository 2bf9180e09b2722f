#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ring_nccl_check.py > gpurun_out/ring_nccl.log 2>&1; echo "nccl ring rc=$?"; tail -2 gpurun_out/ring_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 28 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 rc=$?"; tail -2 gpurun_out/bench_g2.err
SPHB_REUSE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_g2_noreuse.json 2> gpurun_out/bench_g2_noreuse.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_g2_c4.json 2> gpurun_out/bench_g2_c4.err; tail -2 gpurun_out/bench_g2_c4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload c4dam --steps 10 --warmup 3 > gpurun_out/bench_g2_c4dam.json 2> gpurun_out/bench_g2_c4dam.err; tail -2 gpurun_out/bench_g2_c4dam.err
python - <<'PY'
import json
for f in ('bench_g2','bench_g2_noreuse','bench_g2_c4','bench_g2_c4dam'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['ms_per_step'],3), d['reuse'], {k:{a:round(b,3) for a,b in v.items()} for k,v in d['phases']['device_ms'].items()}, 'wall', round(d['config']['wall_ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 900 python -m pytest tests/test_gpu_reuse.py tests/test_gpu_ring.py tests/test_multi_gpu.py -q --tb=short > gpurun_out/pytest_reuse.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_reuse.log
