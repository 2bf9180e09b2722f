#!/bin/bash
# One gpurun call's worth of validation after a kernel change (run from the repo root on the GPU box):
#   gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'         (QUICK=1: without the ncu launch list)
# Writes everything under gpurun_out/: the GPU test log, smoke, the bench lines, the ncu launch list of the bench command
# and compute-sanitizer memcheck over tools/sanitize_workload.py.  Bench numbers come from runs that are NOT under a profiler.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
if [ -z "${QUICK:-}" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-legs --no-cpu --no-other-build > gpurun_out/ncu_bench.log 2>&1
fi
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_workload.py > gpurun_out/sanitizer.txt 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_c5.json; tail -3 gpurun_out/sanitizer.txt
