#!/bin/bash
# round-2 GPU call M: chunk trace of the pipelined host call (f64 and int16 input), full capture of k_align<2>,
# launch list of the int16 path
mkdir -p gpurun_out
TB2_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras "" --no-parity \
    > gpurun_out/r2m_trace.json 2> gpurun_out/r2m_trace.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'k_align<2>' -s 0 -c 1 \
    -o gpurun_out/prof_align2_r2m python bench.py --workload mixed --reads 6000 --steps 1 \
    --warmup 1 --no-cpu-baseline --extras "" --no-parity --no-int16 > gpurun_out/r2m_ncu_mixed.log 2>&1
ls -la gpurun_out/prof_align2_r2m.ncu-rep
grep -c "chunk" gpurun_out/r2m_trace.err
