#!/bin/bash
# final library on 4 GPUs (left and right neighbours are different peers): NCCL checks against the single handle, bench
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
RING_CHECK_NX=256 timeout 300 $T --master-port 29517 tools/ring_nccl_check.py > gpurun_out/ring_nccl4.log 2>&1; echo "nccl ring rc=$?"; grep "ring over" gpurun_out/ring_nccl4.log
RING_CHECK_NX=256 RING_CHECK_VX=0.1 RING_CHECK_STEPS=30 timeout 300 $T --master-port 29519 tools/ring_nccl_check.py > gpurun_out/ring_nccl4_slow.log 2>&1; echo "nccl ring slow rc=$?"; grep "ring over" gpurun_out/ring_nccl4_slow.log
timeout 600 $T --master-port 29518 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err; echo "bench g4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_g4.json')); print('bench_g4', round(d['ms_per_step'],3), d['value'], {k:round(b,3) for k,b in d['phases']['device_ms']['rebuild'].items()})
PY
