#!/bin/bash
# round-2 GPU call T: launch list of configs[1] on integer-valued signal (which kernel is slower on DAC-like data?)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > gpurun_out/r2t_tests.log; tail -2 gpurun_out/r2t_tests.log
TB2_BENCH_ROUND_RAW=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file gpurun_out/launches_r2t_int.csv python bench.py --reads 30000 --steps 2 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2t_ncu_list.log 2>&1
TB2_BENCH_ROUND_RAW=1 timeout 600 python bench.py --no-cpu-baseline --extras "" --no-int16 \
    > gpurun_out/r2t_c1_int.json 2> gpurun_out/r2t_c1_int.err
python - <<'PY'
import json
for l in open('gpurun_out/r2t_c1_int.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], (d.get('parity') or {}).get('mismatches'))
PY
