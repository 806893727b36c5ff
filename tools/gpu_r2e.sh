#!/bin/bash
# round-2 GPU call E: queue mode at N=1, c5 launch list, sanitizer on the new engines
mkdir -p gpurun_out
timeout 900 python bench.py --workload mixed --queue --reads 20000 --steps 2 --warmup 1 --gpus 1 \
    > gpurun_out/r2e_queue1.json 2> gpurun_out/r2e_queue1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv \
    --log-file gpurun_out/launches_r2e_c5.csv python bench.py --workload c5 --reads 2000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2e_ncu_c5.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_dp_gpu.py \
    tests/test_pipeline_gpu.py tests/test_region_stats_gpu.py -x -q -k "not large_batch" > gpurun_out/r2e_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_dp_gpu.py \
    tests/test_pipeline_gpu.py -x -q -k "not large_batch" > gpurun_out/r2e_racecheck.log 2>&1
tail -3 gpurun_out/r2e_memcheck.log; tail -3 gpurun_out/r2e_racecheck.log; tail -c 600 gpurun_out/r2e_queue1.err
