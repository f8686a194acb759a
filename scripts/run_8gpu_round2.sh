#!/bin/bash
# One 8-GPU box visit: configs[4] (Cr2 1e7) with the piecewise reference checks, the multi-GPU parity tests, bench.
# Everything lands in gpurun_out/ (small files only).
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
B2CI_LOG=info timeout 600 $TR --master-port 29521 scripts/asci_scale.py cr2_asci30 10000000 max_refine_iter=0 \
  save=/tmp/cr2_1e7.npz > gpurun_out/asci8.json 2> gpurun_out/asci8.err
grep -E "^\[(h_build|asci_search|asci_grow)\]" gpurun_out/asci8.err | tail -60 > gpurun_out/asci8_phase_log.txt
tail -c 400 gpurun_out/asci8.json; echo
( timeout 900 python scripts/verify_scale.py /tmp/cr2_1e7.npz ncore=10000 nkeep=100000 rows_per_block=64 nblocks=8 \
    > gpurun_out/verify8.json 2> gpurun_out/verify8.err; echo "verify rc=$?" ) &
VPID=$!
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/multi8.txt
wait $VPID
cat gpurun_out/verify8.json | cut -c1-1500; tail -3 gpurun_out/verify8.err
timeout 900 $TR --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8.json 2> gpurun_out/bench8.err
tail -c 300 gpurun_out/bench8.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'hbuild_ms', 'hbuild_setup_ms', 'hbuild_count_ms', 'hbuild_fill_ms', 'sigma_iter_ms')})
print(d.get('e2e', {}).get('value'), (d.get('davidson') or {}).get('other_ms_per_iter'), d.get('parity'))
PY
rm -f gpurun_out/asci8.err
ls -la gpurun_out | head -30
