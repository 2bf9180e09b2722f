#!/bin/bash
# force kernel: blocks per SM x staged records (variant libraries built with -DFORCE_MINB)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-legs --no-cpu --no-other-build"
run() { $B > gpurun_out/bench_v.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/bench_v.json').read().strip().splitlines()[-1]);p=d['roofline']['phases']['by_kind_ms']['rebuild'];print('$1', round(d['ms_per_step'],3), 'force', round(p['force'],3), 'knn', round(p['knn'],3))"; }
cp sphugo_b200/libsphb.so /tmp/lib_default.so
for n in 576 672 736; do SPHB_FORCE_NREC=$n run "mb5 nrec=$n"; done
cp build/var/libsphb_mb6.so sphugo_b200/libsphb.so
for n in 480 528 560; do SPHB_FORCE_NREC=$n run "mb6 nrec=$n"; done
cp build/var/libsphb_mb4.so sphugo_b200/libsphb.so
for n in 672 840; do SPHB_FORCE_NREC=$n run "mb4 nrec=$n"; done
cp /tmp/lib_default.so sphugo_b200/libsphb.so
