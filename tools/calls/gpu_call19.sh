#!/bin/bash
# round 2: the ncu evidence of the final kernels (launch lists of the bench command, --set full of the top kernels)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-legs --no-cpu --no-other-build"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5.csv $B > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c5_reuse.csv $B --reuse --steps 8 > gpurun_out/ncu_l2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_(knn_tile|force_st|reorder)' -s 9 -c 3 -o gpurun_out/r02_full_c5_f64 $B > gpurun_out/ncu_f1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_(knn_tile|force_st32|reorder)' -s 9 -c 3 -o gpurun_out/r02_full_c5_f32 $B --precision 32 > gpurun_out/ncu_f2.log 2>&1
SPHB_REUSE_PERIOD=6 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_knn_(annulus|reuse)|k_predict' -s 0 -c 3 -o gpurun_out/r02_full_c5_reuse $B --steps 4 > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
