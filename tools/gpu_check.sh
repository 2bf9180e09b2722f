#!/bin/bash
# One gpurun call's worth of validation after a kernel change (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
# Writes everything under gpurun_out/: the GPU test log, smoke, both bench lines, the ncu launch list of the bench
# command, and compute-sanitizer memcheck over tools/sanitize_workload.py.  Bench numbers come from the runs that are
# NOT under a profiler.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
# Go toolchain probe (the cgo shim can only be compiled where `go version` works)
{ echo "PATH=$PATH"; which go gccgo tinygo 2>&1; go version 2>&1; ls -d /usr/local/go /usr/lib/go* 2>&1; } > gpurun_out/go_probe.txt 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"
if [ -z "${QUICK:-}" ]; then
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1
fi
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_workload.py > gpurun_out/sanitizer.txt 2>&1; echo "memcheck rc=$?"
timeout 300 python tools/speed_test.py --impl gpu > gpurun_out/speed_test_gpu.json 2> gpurun_out/speed_test_gpu.err
timeout 300 python tools/speed_test.py --impl cpu > gpurun_out/speed_test_cpu.json 2> gpurun_out/speed_test_cpu.err
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_c5.json
