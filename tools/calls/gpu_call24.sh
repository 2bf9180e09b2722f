#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "bench g8 rc=$?"; tail -2 gpurun_out/bench_g8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_g8_c4.json 2> gpurun_out/bench_g8_c4.err; tail -2 gpurun_out/bench_g8_c4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload c4dam --steps 10 --warmup 3 > gpurun_out/bench_g8_c4dam.json 2> gpurun_out/bench_g8_c4dam.err; tail -2 gpurun_out/bench_g8_c4dam.err
python - <<'PY'
import json
for f in ('bench_g8','bench_g8_c4','bench_g8_c4dam'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['ms_per_step'],3), d['value'], 'wall', round(d['config']['wall_ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
PY
