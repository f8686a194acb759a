#!/bin/bash
# compute-sanitizer over the kernels written in round 2 (small parity tests; each tool in its own process)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_r02.txt
: > $OUT
python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a $OUT
SEL='hit_list_fill_equals_rescan_fill and (hits or tile_all or wide) or balanced or water_cisd_csr or davidson_water'
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool  pytest tests/test_gpu_parity.py -k \"$SEL\"" >> $OUT
  timeout 1500 compute-sanitizer --tool $tool --print-limit 40 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$SEL" 2>&1 \
    > gpurun_out/san_$tool.log 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|=========.*at |b2ci::" gpurun_out/san_$tool.log | sort | uniq -c | sort -rn | head -14 >> $OUT
done
echo "== compute-sanitizer --tool memcheck  pytest tests/test_gpu_asci.py -k 'search or wide'" >> $OUT
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_asci.py -x -q -m gpu -k "search or wide" 2>&1 \
  | grep -E "passed|failed|ERROR SUMMARY|Invalid" | tail -6 >> $OUT
cat $OUT
