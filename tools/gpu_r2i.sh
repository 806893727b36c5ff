#!/bin/bash
# round-2 GPU call I: k_align<2> 4 vs 5 CTAs/SM on the adaptive-band workloads; queue mode N=1 reference point
mkdir -p gpurun_out
for w in mixed c5; do
  for c in 4 5; do
    TB2_ALIGN2_CTAS=$c timeout 600 python bench.py --workload $w --no-cpu-baseline --extras "" --no-int16 \
        > gpurun_out/r2i_${w}_ctas$c.json 2> gpurun_out/r2i_${w}_ctas$c.err
  done
done
timeout 600 python bench.py --workload mixed --queue --reads 40000 --bucket-samples 20000000 --no-cpu-baseline \
    --extras "" --no-int16 > gpurun_out/r2i_queue_n1.json 2> gpurun_out/r2i_queue_n1.err
for f in gpurun_out/r2i_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','e2e','parity')})
PY
done
