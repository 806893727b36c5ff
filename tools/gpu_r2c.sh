#!/bin/bash
# round-2 GPU call C: Theil-Sen exact pass without divergent pushes, k_resolve with parallel z
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2c_tests.log
timeout 1200 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file gpurun_out/launches_r2c.csv python bench.py --reads 30000 --steps 2 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2c_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_theil_sen -s 3 -c 1 \
    -o gpurun_out/prof_theil_sen_r2c python bench.py --reads 20000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2c_ncu_ts.log 2>&1
tail -5 gpurun_out/r2c_tests.log
