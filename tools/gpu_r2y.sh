#!/bin/bash
# round-2 GPU call Y: full capture of k_align<2> on configs[4] (multi-slab engine with row-only slabs), 6000 reads
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'k_align<\(int\)2>' -s 0 -c 1 \
    -o gpurun_out/prof_align2_c5_r2y python bench.py --workload c5 --reads 6000 --steps 1 \
    --warmup 1 --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2y_ncu_c5.log 2>&1
ls -la gpurun_out/prof_align2_c5_r2y.ncu-rep
