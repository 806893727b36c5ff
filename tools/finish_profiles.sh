#!/bin/bash
# turn the outputs of tools/gpu_r2r.sh (gpurun_out/r2r_*) into the tracked files under profiles/
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/r2r_bench.json profiles/r2_final_bench.json
cp gpurun_out/r2r_bench_reference.json profiles/r2_final_bench_reference.json
python - <<'PY'
import sys
sys.path.insert(0, 'profiles')
import summarize
summarize.launches('r2r')
PY
rm -f profiles/r2r_launches.csv
mv profiles/r2r_launch_shares.txt profiles/r2_final_launch_shares.txt
python profiles/ncu_table.py gpurun_out/prof_align1_r2r.ncu-rep profiles/k_align1_r2_final_ncu_full.txt \
    "k_align<1>, final kernels of round 2; first-iteration launch over 20000 configs[1] reads (static band W~748)"
python tools/make_traffic_json.py gpurun_out/prof_align1_r2r.ncu-rep 20000
{
  echo "# compute-sanitizer on the final round-2 kernels (GPU call R, B200): tests/test_dp_gpu.py + tests/test_pipeline_gpu.py -k 'not large_batch'"
  echo -n "memcheck : "; grep -E "passed|failed" gpurun_out/r2r_memcheck.log | tail -n 1; grep "ERROR SUMMARY" gpurun_out/r2r_memcheck.log | tail -n 1
  echo -n "racecheck: "; grep -E "passed|failed" gpurun_out/r2r_racecheck.log | tail -n 1; grep "RACECHECK SUMMARY" gpurun_out/r2r_racecheck.log | tail -n 1
} > profiles/r2_sanitizer_after.txt
cat profiles/r2_sanitizer_after.txt
head -12 profiles/r2_final_launch_shares.txt
cat profiles/k_align_traffic.json
