#!/bin/bash
# round-2 GPU call V (1 GPU): work-queue worker threads x bucket schedule on one GPU (80000 mixed reads)
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 900 python bench.py --workload mixed --queue --reads 80000 --bucket-samples 16000000 --steps 2 --warmup 1 "$@" \
      > gpurun_out/r2v_$tag.json 2> gpurun_out/r2v_$tag.err
  python - gpurun_out/r2v_$tag.json $tag <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], round(d['value']), d['ms_per_step'])
PY
}
run t2_plain --queue-threads 2 --bucket-tail 0
run t4_plain --queue-threads 4 --bucket-tail 0
run t3_plain --queue-threads 3 --bucket-tail 0
run t4_guided --queue-threads 4 --bucket-tail 0.25
