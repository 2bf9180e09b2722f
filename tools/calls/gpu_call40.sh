#!/bin/bash
# final library: ncu --set full of the three big kernels of the default path (one launch each)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-legs --no-cpu --no-other-build"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_(knn_tile|force_st|reorder)' -s 9 -c 3 -o gpurun_out/r02_final_full_c5_f64 $B > gpurun_out/ncu_f1.log 2>&1
ls -la gpurun_out/*.ncu-rep
