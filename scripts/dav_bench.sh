for w in cr2_cas12 hubbard_4x3; do python bench.py --workload $w --steps 3 --warmup 3 --no-asci --no-also --no-plugin-e2e --cpu-seconds 1 2>/dev/null > /tmp/b_$w.json; python - <<PY
import json
d=json.loads(open('/tmp/b_$w.json').read().strip().splitlines()[-1]); v=d["davidson"]
print(d["config"]["workload"], v["niter"], "other %.3f rr %.2f res %.2f gs %.2f" % (v["other_ms_per_iter"], v["rr_ms_total"], v["res_ms_total"], v["gs_ms_total"]), d["parity"]["energy"]["ok"], d["parity"]["energy"]["abs_err"])
PY
done
