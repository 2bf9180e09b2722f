#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L | wc -l
RING_CHECK_NX=256 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 tools/ring_nccl_check.py > gpurun_out/ring_nccl4.log 2>&1; echo "nccl ring 4 rc=$?"; grep "ring over" gpurun_out/ring_nccl4.log
SPHB_REUSE_PERIOD=4 RING_CHECK_NX=256 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516 tools/ring_nccl_check.py > gpurun_out/ring_nccl4r.log 2>&1; echo "nccl ring 4 reuse rc=$?"; grep "ring over" gpurun_out/ring_nccl4r.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err; echo "bench g4 rc=$?"; tail -2 gpurun_out/bench_g4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_g4_c4.json 2> gpurun_out/bench_g4_c4.err; tail -2 gpurun_out/bench_g4_c4.err
python - <<'PY'
import json
for f in ('bench_g4','bench_g4_c4'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['ms_per_step'],3), d['value'], d['reuse'], 'wall', round(d['config']['wall_ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
PY
