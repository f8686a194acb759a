#!/bin/bash
# One B200 round trip: GPU parity tests, the two FCI bench workloads, and (optionally) an
# `ncu --set full` capture of the kernels named in $1 on the Cr2 CAS(12,12) workload.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_check.sh "k_rows_product|k_spmv"'
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload hubbard_4x3 --steps 3 --warmup 3 --cpu-seconds ${CPU_SECONDS:-0} > gpurun_out/bench_hubbard.json 2> gpurun_out/bench_hubbard.err
timeout 300 python bench.py --workload cr2_cas12 --no-also --steps 3 --warmup 3 --cpu-seconds ${CPU_SECONDS:-0} > gpurun_out/bench_cr2.json 2> gpurun_out/bench_cr2.err
cat gpurun_out/bench_hubbard.json gpurun_out/bench_cr2.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: d.get(k) for k in ['value','hbuild_ms','hbuild_setup_ms','hbuild_count_ms','hbuild_fill_ms','hbuild_thresh_ms','sigma_iter_ms','gpu_launches']}, d['config']['workload'], 'fill frac', d['roofline']['frac'], 'sigma frac', d['roofline_sigma']['frac'], 'e2e', d['e2e']['value'], d.get('davidson'))
"
tail -n 3 gpurun_out/bench_hubbard.err gpurun_out/bench_cr2.err
if [ -n "$1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-2} -o gpurun_out/prof_cr2 -f \
    python bench.py --workload cr2_cas12 --no-also --steps 1 --warmup 3 --no-davidson --cpu-seconds 0 > gpurun_out/ncu_full.log 2>&1
  tail -n 2 gpurun_out/ncu_full.log | cut -c1-200
fi
