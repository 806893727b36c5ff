#!/bin/bash
# round-2 GPU call X: launch lists (kernel time shares) of the configs[2]-like mix and configs[4], final kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv \
    --log-file gpurun_out/launches_r2x_mixed.csv python bench.py --workload mixed --reads 10000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2x_ncu_mixed.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv \
    --log-file gpurun_out/launches_r2x_c5.csv python bench.py --workload c5 --reads 3000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2x_ncu_c5.log 2>&1
ls -la gpurun_out/launches_r2x_*.csv
