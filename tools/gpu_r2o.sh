#!/bin/bash
# round-2 GPU call O: pipelined host call variants on configs[1] (threads x SM share x chunk size)
mkdir -p gpurun_out
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --extras "" --no-parity --no-int16 --steps 4 --warmup 2 \
      > gpurun_out/r2o_$tag.json 2> gpurun_out/r2o_$tag.err
  python - gpurun_out/r2o_$tag.json $tag <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], round(d['value']), round(d['e2e']['value']))
PY
}
run t1_u8 TB2_PIPELINE_THREADS=1
run t2_share2 TB2_PIPELINE_THREADS=2 TB2_PIPELINE_SHARE=2
run t2_share1 TB2_PIPELINE_THREADS=2 TB2_PIPELINE_SHARE=1
run t1_u12 TB2_PIPELINE_THREADS=1 TB2_PIPELINE_CHUNK_U=12
run t1_u6 TB2_PIPELINE_THREADS=1 TB2_PIPELINE_CHUNK_U=6
run t2_share2_u6 TB2_PIPELINE_THREADS=2 TB2_PIPELINE_SHARE=2 TB2_PIPELINE_CHUNK_U=6
