#!/bin/bash
# round-2 GPU call G: full captures of k_theil_sen and k_align<1> (first-iteration launches, 20000 reads)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_theil_sen -s 0 -c 1 \
    -o gpurun_out/prof_theil_sen_r2g python bench.py --reads 20000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2g_ncu_ts.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 0 -c 1 \
    -o gpurun_out/prof_align1_r2g python bench.py --reads 20000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2g_ncu_al.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
