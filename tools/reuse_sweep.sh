#!/bin/bash
# sweep of the list-reuse knobs on bench.py's workload (one GPU): per-step probe for each setting
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
run() { # name, env...
  name=$1; shift
  env "$@" SPHB_REUSE_PERIOD=7 timeout 200 python tools/reuse_probe.py --steps 9 > gpurun_out/sweep_$name.txt 2>&1
  echo "== $name: $*"; grep -E "build|reuse" gpurun_out/sweep_$name.txt | awk '{print $2, $3, $7, $8, $10}' | tr '\n' ';'; echo
}
run base
run ncw448 SPHB_REUSE_NCW=448
run ncw448_c127 SPHB_REUSE_NCW=448 SPHB_CELL_PER_H=1.27
run ncw320_c127 SPHB_REUSE_NCW=320 SPHB_CELL_PER_H=1.27
run ncw256_c127 SPHB_REUSE_NCW=256 SPHB_CELL_PER_H=1.27
run ncw320_c130_s20 SPHB_REUSE_NCW=320 SPHB_CELL_PER_H=1.23 SPHB_REUSE_SKIN=0.20
run ncw320_c135_s30 SPHB_REUSE_NCW=384 SPHB_CELL_PER_H=1.33 SPHB_REUSE_SKIN=0.30
run capb24_c127 SPHB_REUSE_NCW=320 SPHB_CELL_PER_H=1.27 SPHB_REUSE_CAPB=24
timeout 600 python -m pytest tests/test_gpu_reuse.py tests/test_gpu_parity.py -q --tb=short > gpurun_out/pytest_reuse.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_reuse.log
timeout 600 python -m pytest tests/test_gpu_ring.py -q --tb=short -x > gpurun_out/pytest_ring.log 2>&1; echo "ring rc=$?"; tail -25 gpurun_out/pytest_ring.log
