#!/bin/bash
# round-2 GPU call N: pipelined host call with one host thread per lane (A/B against TB2_PIPELINE_THREADS=1),
# full capture of k_align<2>
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2n_tests.log
for t in 2 1; do
  TB2_PIPELINE_THREADS=$t timeout 600 python bench.py --no-cpu-baseline --extras "" --no-parity \
      > gpurun_out/r2n_c1_threads$t.json 2> gpurun_out/r2n_c1_threads$t.err
  TB2_PIPELINE_THREADS=$t timeout 600 python bench.py --workload mixed --no-cpu-baseline --extras "" --no-parity --no-int16 \
      > gpurun_out/r2n_mixed_threads$t.json 2> gpurun_out/r2n_mixed_threads$t.err
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'k_align<\(int\)2>' -s 0 -c 1 \
    -o gpurun_out/prof_align2_r2n python bench.py --workload mixed --reads 6000 --steps 1 \
    --warmup 1 --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2n_ncu_mixed.log 2>&1
tail -4 gpurun_out/r2n_tests.log
ls -la gpurun_out/prof_align2_r2n.ncu-rep
for f in gpurun_out/r2n_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], (d.get('e2e_int16') or {}).get('value'))
PY
done
