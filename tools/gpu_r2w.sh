#!/bin/bash
# round-2 GPU call W: the default bench line once more, now that profiles/k_align_traffic.json matches the final kernel sources
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2w_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['e2e']['value']), d['roofline']['traffic'], d['parity']['mismatches'])
PY
