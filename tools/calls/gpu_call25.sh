#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ring_nccl_check.py > gpurun_out/ring_nccl.log 2>&1; echo "nccl ring rc=$?"; grep "ring over" gpurun_out/ring_nccl.log
SPHB_REUSE_PERIOD=4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 tools/ring_nccl_check.py > gpurun_out/ring_ncclr.log 2>&1; echo "nccl ring reuse rc=$?"; grep "ring over" gpurun_out/ring_ncclr.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 rc=$?"; tail -2 gpurun_out/bench_g2.err

python - <<'PY'
import json
for f in ('bench_g2','bench_g2_c4'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['ms_per_step'],3), d['value'], {k:round(b,3) for k,b in d['phases']['device_ms']['rebuild'].items()}, 'wall', round(d['config']['wall_ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
PY
