#!/bin/bash
set -u
mkdir -p gpurun_out
RING_CHECK_VX=0.1 RING_CHECK_STEPS=30 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ring_nccl_check.py > gpurun_out/ring_nccl_slow.log 2>&1; echo "nccl ring slow rc=$?"; grep "ring over" gpurun_out/ring_nccl_slow.log
RING_CHECK_VX=0.4 RING_CHECK_STEPS=24 RING_CHECK_NX=256 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/ring_nccl_check.py > gpurun_out/ring_nccl_mid.log 2>&1; echo "nccl ring mid rc=$?"; grep "ring over" gpurun_out/ring_nccl_mid.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --workload c4 > gpurun_out/bench_g2_c4.json 2> gpurun_out/bench_g2_c4.err; echo "bench g2 c4 rc=$?"; cut -c1-200 gpurun_out/bench_g2_c4.json
