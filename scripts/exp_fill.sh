mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "cr2 or cas or hubbard or product or path" 2>&1 | tail -4)
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --workload ${WL:-cr2_cas12} --no-also --no-davidson --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/exp_$name.json').read().strip().splitlines()[-1])
    print('$name', 'build', round(d['hbuild_ms'],3), 'fill', round(d['hbuild_fill_ms'],3), 'setup', round(d['hbuild_setup_ms'],3), 'count', round(d['hbuild_count_ms'],3), 'sigma', round(d['sigma_iter_ms'],3), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'][:40])
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/exp_$name.err').read()[-500:])
P
}
run base A=1
for v in "$@"; do run $v B2CI_LIB_PATH=$PWD/qdk_chemistry_b200/libb2ci_$v.so; done
WL=hubbard_4x3 run hub A=1
