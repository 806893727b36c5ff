#!/bin/bash
# round-2 GPU call D: wide-band (three chunks per lane) adaptive engine; Theil-Sen loop reverted
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2d_tests.log
timeout 1200 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 3 \
    -o gpurun_out/prof_align_c5_r2d python bench.py --workload c5 --reads 600 --steps 1 \
    --warmup 1 --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2d_ncu_c5.log 2>&1
tail -5 gpurun_out/r2d_tests.log
