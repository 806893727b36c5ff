#!/bin/bash
# round-2 GPU call H: Theil-Sen with the 4096-pair list (one listing sweep for n ~ 444)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2h_tests.log
timeout 1200 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_theil_sen -s 0 -c 1 \
    -o gpurun_out/prof_theil_sen_r2h python bench.py --reads 20000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2h_ncu_ts.log 2>&1
tail -4 gpurun_out/r2h_tests.log
