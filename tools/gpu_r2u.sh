#!/bin/bash
# round-2 GPU call U (4 GPUs): shared-queue strong scaling with guided tail buckets (last 25 % of the samples in
# 4x smaller buckets), N=4 and N=1 on the same box
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 4 --workload mixed --queue --reads 80000 --bucket-samples 16000000 --steps 2 --warmup 1 \
    > gpurun_out/r2u_queue_n4.json 2> gpurun_out/r2u_queue_n4.err
timeout 900 python bench.py --workload mixed --queue --reads 80000 --bucket-samples 16000000 --steps 2 --warmup 1 \
    > gpurun_out/r2u_queue_n1.json 2> gpurun_out/r2u_queue_n1.err
for f in gpurun_out/r2u_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d['config'].get('rank_time_min_over_max'), d['config'].get('queue'))
PY
done
