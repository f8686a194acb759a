#!/bin/bash
# final single-GPU evidence of the round: tests, default bench + reference arm, launch list, ncu of the headline
# kernels, general-build phase traces. Small files only, into gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_gpu_tests.txt
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err
python bench.py --impl reference > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err
python bench.py --workload hubbard_4x3 --no-asci --no-also > gpurun_out/final_bench_hubbard.json 2>> gpurun_out/final_bench.err
for w in n2_asci26 cr2_asci30; do
  B2CI_NO_INCREMENTAL=1 B2CI_HBUILD_TRACE=1 python scripts/asci_scale.py $w 1000000 max_refine_iter=0 > gpurun_out/final_asci_$w.json 2> gpurun_out/final_asci_$w.err
  grep -E "hbuild scan|\[hbuild\]" gpurun_out/final_asci_$w.err | tail -14 > gpurun_out/final_scan_trace_$w.txt
  python scripts/asci_scale.py $w 1000000 max_refine_iter=0 > gpurun_out/final_asci_incremental_$w.json 2>/dev/null
  rm -f gpurun_out/final_asci_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-asci --no-also --no-plugin-e2e --cpu-seconds 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_rows_dense|^k_spmv" -s 6 -c 2 -o gpurun_out/final_dense_spmv -f \
  python bench.py --steps 2 --warmup 3 --no-asci --no-also --no-plugin-e2e --no-davidson --cpu-seconds 1 > /dev/null 2>&1
ls -la gpurun_out | head -40
