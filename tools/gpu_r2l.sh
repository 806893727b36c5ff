#!/bin/bash
# round-2 GPU call L: register-engine trims (arg-max index, fix-up without the band test, 4-op divide):
# tests, mixed + c5 bench, queue mode with pinned result buffers, full capture of k_align<2>
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2l_tests.log
for w in mixed c5; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --extras "" --no-int16 \
      > gpurun_out/r2l_$w.json 2> gpurun_out/r2l_$w.err
done
timeout 900 python bench.py --workload mixed --queue --reads 40000 --bucket-samples 20000000 --steps 2 --warmup 1 \
    > gpurun_out/r2l_queue_n1.json 2> gpurun_out/r2l_queue_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 2 -c 1 \
    -o gpurun_out/prof_align2_r2l python bench.py --workload mixed --reads 6000 --steps 1 \
    --warmup 1 --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2l_ncu_mixed.log 2>&1
tail -4 gpurun_out/r2l_tests.log
for f in gpurun_out/r2l_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], (d.get('parity') or {}).get('mismatches'))
PY
done
