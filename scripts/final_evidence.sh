#!/bin/bash
# Round-end evidence on one B200: default bench line (both arms), per-workload bench lines, the ncu
# launch list of the bench command and a --set full capture of the two dominant kernels.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python bench.py --workload cr2_cas12 --no-also --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_cr2.json 2> gpurun_out/bench_cr2.err
timeout 300 python bench.py --workload hubbard_4x3 --no-also --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_hubbard.json 2> gpurun_out/bench_hubbard.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cr2.csv \
  python bench.py --workload cr2_cas12 --no-also --steps 2 --warmup 1 --no-davidson --cpu-seconds 0 > gpurun_out/launches_cr2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_rows_product|k_spmv" -s 6 -c 2 -o gpurun_out/prof_cr2 -f \
  python bench.py --workload cr2_cas12 --no-also --steps 1 --warmup 3 --no-davidson --cpu-seconds 0 > gpurun_out/ncu_full.log 2>&1
for f in bench_default bench_reference bench_cr2 bench_hubbard; do echo "== $f"; tail -c 1500 gpurun_out/$f.json; echo; tail -n 2 gpurun_out/$f.err; done
tail -n 2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/prof_cr2.ncu-rep
