#!/bin/bash
# round-2 final evidence call R (1 GPU): tests, default bench, reference arm, launch list, full capture of
# k_align<1> (-> roofline.traffic), memcheck + racecheck
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2r_smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2r_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" > gpurun_out/r2r_smoke.log 2>&1; tail -n 1 gpurun_out/r2r_smoke.log
timeout 1500 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file gpurun_out/launches_r2r.csv python bench.py --reads 30000 --steps 2 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2r_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 0 -c 1 \
    -o gpurun_out/prof_align1_r2r python bench.py --reads 20000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2r_ncu_al.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_dp_gpu.py \
    tests/test_pipeline_gpu.py -x -q -k "not large_batch" > gpurun_out/r2r_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_dp_gpu.py \
    tests/test_pipeline_gpu.py -x -q -k "not large_batch" > gpurun_out/r2r_racecheck.log 2>&1
tail -3 gpurun_out/r2r_tests.log
tail -n 2 gpurun_out/r2r_memcheck.log; tail -n 2 gpurun_out/r2r_racecheck.log
