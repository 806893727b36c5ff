#!/bin/bash
# round-2 GPU call A: tests, bench (headline + extras), "before" ncu of the adaptive engine, sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2a_tests.log
timeout 1200 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 2 -c 4 \
    -o gpurun_out/prof_align_mixed_r2before python bench.py --workload mixed --reads 6000 --steps 1 \
    --warmup 1 --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2a_ncu_mixed.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_dp_gpu.py \
    tests/test_pipeline_gpu.py -x -q -k "not large_batch" > gpurun_out/r2a_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_dp_gpu.py \
    tests/test_pipeline_gpu.py -x -q -k "not large_batch" > gpurun_out/r2a_racecheck.log 2>&1
tail -5 gpurun_out/r2a_tests.log
