#!/bin/bash
# second 8-GPU visit: connection-balanced row blocks for the Cr2 1e7 ASCI run, bench at 8 GPUs
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
B2CI_LOG=info timeout 600 $TR --master-port 29541 scripts/asci_scale.py cr2_asci30 10000000 max_refine_iter=0 \
  > gpurun_out/asci8b.json 2> gpurun_out/asci8b.err
grep -E "^\[(h_build|asci_search|asci_grow|ci_solver)\]" gpurun_out/asci8b.err | tail -60 > gpurun_out/asci8b_phase_log.txt
grep -E "^\[h_build\]" gpurun_out/asci8b.err | tail -8
tail -3 gpurun_out/asci8b.err | cut -c1-300
rm -f gpurun_out/asci8b.err
tail -c 700 gpurun_out/asci8b.json; echo
timeout 900 $TR --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8b.json 2> gpurun_out/bench8b.err
tail -c 300 gpurun_out/bench8b.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench8b.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'hbuild_ms', 'hbuild_setup_ms', 'hbuild_count_ms', 'hbuild_fill_ms', 'sigma_iter_ms')})
print(d.get('e2e', {}).get('value'), (d.get('davidson') or {}))
PY
