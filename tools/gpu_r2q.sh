#!/bin/bash
# round-2 GPU call Q (4 GPUs): weak scaling of the default bench at N=2 and 4; strong scaling of the shared work
# queue (ONE mixed read set of 80000 reads, 16M-sample buckets) at N=1, 2, 4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2q_topo.txt 2>&1
P=29520
for n in ${NS:-2 4}; do
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_${n}gpu.json 2> gpurun_out/r2q_bench_${n}gpu.err
done
timeout 900 python bench.py --workload mixed --queue --reads 80000 --bucket-samples 16000000 --steps 2 --warmup 1 \
    > gpurun_out/r2q_queue_n1.json 2> gpurun_out/r2q_queue_n1.err
for n in ${NS:-2 4}; do
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $n --workload mixed --queue --reads 80000 --bucket-samples 16000000 --steps 2 --warmup 1 \
      > gpurun_out/r2q_queue_n$n.json 2> gpurun_out/r2q_queue_n$n.err
done
for f in gpurun_out/r2q_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['config'].get('rank_time_min_over_max'))
PY
done
for f in gpurun_out/r2q_*.err; do tail -n 2 $f; done
