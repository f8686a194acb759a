#!/usr/bin/env python
"""Time the RDM build on a full-CI workload: python scripts/rdm_time.py cr2_cas12 [spin_dep]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qdk_chemistry_b200 import device, workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "cr2_cas12"
sd = len(sys.argv) > 2 and sys.argv[2] == "spin_dep"
sp = W.config(name)
ctx = device.Context(0)
ctx.upload_integrals(sp.norb, sp.T, sp.V)
dets = ctx.generate_fci(sp.norb, sp.nalpha, sp.nbeta)
rng = np.random.default_rng(0)
C = rng.normal(size=len(dets)); C /= np.linalg.norm(C)
out = {}
for rep in range(2):
    t0 = time.perf_counter()
    r = ctx.form_rdms(dets, C, spin_dep=sd)
    out[f"wall_s_{rep}"] = time.perf_counter() - t0
    out[f"pattern_ms_{rep}"] = ctx.timer_ms("rdm.pattern"); out[f"scatter_ms_{rep}"] = ctx.timer_ms("rdm.scatter")
out["trace"] = float(np.trace(r[0]) + (np.trace(r[1]) if sd else 0.0)); out["ndets"] = len(dets); out["workload"] = name; out["spin_dep"] = sd
print(json.dumps(out))
