#!/bin/bash
# round-2 GPU call B: register engine for the adaptive band -- tests, bench, "after" ncu
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2b_tests.log
timeout 1200 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 2 -c 4 \
    -o gpurun_out/prof_align_mixed_r2after python bench.py --workload mixed --reads 6000 --steps 1 \
    --warmup 1 --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2b_ncu_mixed.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file gpurun_out/launches_r2b.csv python bench.py --reads 30000 --steps 2 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2b_ncu_list.log 2>&1
tail -5 gpurun_out/r2b_tests.log
