#!/bin/bash
# round-2 GPU call S: configs[4] with more reads per batch (the multi-slab engine's occupancy after the
# shared-memory halving), RNA / mixed with larger batches
mkdir -p gpurun_out
timeout 900 python bench.py --workload c5 --reads 6000 --steps 2 --warmup 1 --no-cpu-baseline --extras "" --no-int16 \
    > gpurun_out/r2s_c5_6000.json 2> gpurun_out/r2s_c5_6000.err
timeout 900 python bench.py --workload mixed --reads 40000 --steps 3 --warmup 2 --no-cpu-baseline --extras "" --no-int16 \
    > gpurun_out/r2s_mixed_40000.json 2> gpurun_out/r2s_mixed_40000.err
for f in gpurun_out/r2s_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], (d.get('parity') or {}).get('mismatches'), d['roofline'].get('cell_updates_per_s'))
PY
done
