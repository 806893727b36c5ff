#!/bin/bash
# round-2 GPU call J: wavefront steady state with the shared-memory ring, 4-operation exact divide,
# predicate move bits: tests, bench, full capture of k_align<1>, racecheck of the DP tests
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2j_tests.log
timeout 1200 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 0 -c 1 \
    -o gpurun_out/prof_align1_r2j python bench.py --reads 20000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2j_ncu_al.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_dp_gpu.py \
    tests/test_pipeline_gpu.py -x -q -k "not large_batch" > gpurun_out/r2j_racecheck.log 2>&1
tail -4 gpurun_out/r2j_tests.log
python - <<'PY'
import json
for l in open('gpurun_out/r2j_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','parity')}, d['e2e']['value'], [(k,v.get('value')) for k,v in d.get('extra_configs',{}).items()])
PY
