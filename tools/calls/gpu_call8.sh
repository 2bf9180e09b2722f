#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_reuse.py -q --tb=short -x > gpurun_out/pytest_reuse.log 2>&1; echo "reuse rc=$?"; tail -15 gpurun_out/pytest_reuse.log
SPHB_REUSE_PERIOD=7 timeout 200 python tools/reuse_probe.py --steps 9 > gpurun_out/probe_p7.txt 2>&1; grep -E "build|reuse|Error|error" gpurun_out/probe_p7.txt
SPHB_REUSE_PERIOD=7 SPHB_REUSE_NCW=320 timeout 200 python tools/reuse_probe.py --steps 9 > gpurun_out/probe_p7b.txt 2>&1; echo ncw320; grep -E "build|reuse|Error|error" gpurun_out/probe_p7b.txt
SPHB_REUSE_PERIOD=7 timeout 200 python tools/reuse_probe.py --steps 9 --precision 32 > gpurun_out/probe_p7_f32.txt 2>&1; echo f32; grep -E "build|reuse|Error|error" gpurun_out/probe_p7_f32.txt
timeout 600 python -m pytest tests/test_gpu_ring.py -q --tb=short -x > gpurun_out/pytest_ring.log 2>&1; echo "ring rc=$?"; tail -8 gpurun_out/pytest_ring.log
SPHB_REUSE_PERIOD=6 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_knn_(annulus|reuse|tile)' -s 1 -c 3 -o gpurun_out/r02_knn_staged \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other-build --no-legs > gpurun_out/ncu_staged.log 2>&1
