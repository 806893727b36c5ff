#!/bin/bash
# round-2 GPU call Z (last of the round): chunk-transposed event means of the wide-band engine --
# tests first (stop on failure), configs[4] bench, capture of k_align<1> for roofline.traffic
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r2z_tests.log
tail -n 2 gpurun_out/r2z_tests.log
grep -q " passed" gpurun_out/r2z_tests.log && ! grep -q "failed" gpurun_out/r2z_tests.log || exit 1
timeout 600 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --extras "" --no-int16 \
    > gpurun_out/r2z_c5.json 2> gpurun_out/r2z_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 0 -c 1 \
    -o gpurun_out/prof_align1_r2r python bench.py --reads 20000 --steps 1 --warmup 1 \
    --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2z_ncu_al.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2z_c5.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('c5', round(d['value']), round(d['e2e']['value']), d['parity']['mismatches'], d['roofline'].get('cell_updates_per_s'))
PY
