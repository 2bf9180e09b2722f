#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_reuse.py -q --tb=short -x > gpurun_out/pytest_reuse.log 2>&1; echo "reuse rc=$?"; tail -5 gpurun_out/pytest_reuse.log
for N in 384 448 512; do echo ncw $N; SPHB_REUSE_NCW=$N SPHB_REUSE_PERIOD=7 timeout 200 python tools/reuse_probe.py --steps 8 > gpurun_out/probe_$N.txt 2>&1; grep -E "build|reuse|rror" gpurun_out/probe_$N.txt | sed -n '1p;6,8p'; done
