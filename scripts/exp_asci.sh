# A/B of environment switches on the N2-like ASCI(14e,26o) run, same box, warm GPU.
#   bash scripts/exp_asci.sh 1000000 "A=1" "B2CI_HBUILD_NO_HITLIST=1" ...
mkdir -p gpurun_out
N=$1; shift
timeout 300 python scripts/asci_scale.py n2_asci26 200000 max_refine_iter=1 > /dev/null 2>&1   # warm-up
for rep in 1 2; do
  for v in "$@"; do
    env $v timeout 300 python scripts/asci_scale.py n2_asci26 $N max_refine_iter=2 refine_energy_tol=1e-12 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'wall', round(d['wall_s'], 3), 'h_build', round(d['h_build_ms']), 'count', round(d['h_build_count_ms']), 'fill', round(d['h_build_fill_ms']), 'last', round(d['h_build_last_ms']), 'search', round(d['asci_search_ms']), 'dav', round(d['davidson_sigma_ms'] + d['davidson_other_ms']), 'patched', d.get('h_build_patched'))
"
  done
done
