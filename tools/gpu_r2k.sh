#!/bin/bash
# round-2 GPU call K (2 GPUs): weak scaling of the default bench at N=2, strong scaling of the shared
# work queue (one mixed read set) at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2k_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err
timeout 900 python bench.py --workload mixed --queue --reads 40000 --bucket-samples 20000000 --steps 2 --warmup 1 \
    > gpurun_out/r2k_queue_n1.json 2> gpurun_out/r2k_queue_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload mixed --queue --reads 40000 --bucket-samples 20000000 --steps 2 --warmup 1 \
    > gpurun_out/r2k_queue_n2.json 2> gpurun_out/r2k_queue_n2.err
timeout 600 python bench.py --workload mixed --queue --queue-threads 1 --reads 40000 --bucket-samples 20000000 --steps 2 --warmup 1 \
    > gpurun_out/r2k_queue_n1_t1.json 2> gpurun_out/r2k_queue_n1_t1.err
for f in gpurun_out/r2k_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'])
PY
done
tail -3 gpurun_out/r2k_*.err
