#!/bin/bash
# round-2 GPU call F: sort-and-sweep Theil-Sen, wavefront raw DP in k_resolve
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2f_tests.log
timeout 1200 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file gpurun_out/launches_r2f.csv python bench.py --reads 30000 --steps 2 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2f_ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv \
    --log-file gpurun_out/launches_r2f_c5.csv python bench.py --workload c5 --reads 2000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2f_ncu_c5.log 2>&1
tail -5 gpurun_out/r2f_tests.log
