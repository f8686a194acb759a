#!/usr/bin/env python
"""Run the b200_asci plugin on a named synthetic workload and print per-phase statistics.
    python scripts/asci_scale.py n2_asci26 100000 [key=value ...]
Under torchrun (one process per GPU) the rows of H and the key partitions of the search are
sharded over the ranks; rank 0 prints."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qdk_chemistry_b200 import algorithms as alg, data, workloads as W  # noqa: E402

name, nt = sys.argv[1], int(float(sys.argv[2]))
kw = {}
SAVE = None
for a in sys.argv[3:]:
    k, v = a.split("=")
    if k == "save":
        SAVE = v
        continue
    try:
        kw[k] = int(v)
    except ValueError:
        try:
            kw[k] = float(v)
        except ValueError:
            kw[k] = v
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    alg.init_distributed_from_torch(local)
# CUDA context + module load of this process (once per process, 0.3-1.5 s on a fresh box): timed on its own so that
# wall_s below is the ASCI run, like the reference's CPU timings
t_init = time.perf_counter()
if world == 1:
    from qdk_chemistry_b200 import device as _dev
    _c = _dev.Context(0)
    _c.synchronize()
    _c.close()
t_init = time.perf_counter() - t_init
sp = W.config(name)
ham = data.Hamiltonian(sp.T, sp.V, sp.core_energy)
c = alg.create("multi_configuration_calculator", "b200_asci", ntdets_max=nt, ci_residual_tolerance=1e-8, **kw)
t0 = time.perf_counter()
try:
    E, w = c.run(ham, sp.nalpha, sp.nbeta)
    out = {"E": E, "ndets": w.size()}
    if SAVE and rank == 0:  # determinant words (alpha, beta) and coefficients, for scripts/verify_scale.py
        import numpy as np
        np.savez(SAVE, words=w.determinant_words(), coeffs=np.asarray(w.get_coefficients()), E=E - sp.core_energy,
                 workload=name)
except Exception as e:  # report what failed and the statistics so far
    out = {"error": str(e)[:300]}
out["wall_s"] = time.perf_counter() - t0
out["cuda_init_s"] = t_init
out.update(alg.last_run_stats())
out["world"] = world
# candidates/s of the generator and the bandwidth of the radix sort (12 B per record and
# pass read + written, 2 * ceil(norb / 8) passes), from the device timers of all searches of the run
if out.get("asci_contributions") and out.get("asci_pair_ms"):
    npass = 2 * ((sp.norb + 7) // 8)
    out["asci_candidates_per_s"] = out["asci_contributions"] / (out["asci_pair_ms"] * 1e-3)
    out["asci_sort_accumulate_GBps"] = out["asci_contributions"] * 24.0 * npass / (out["asci_sort_acc_ms"] * 1e-3) / 1e9
if out.get("sigma_last_ms"):
    # sigma on the final matrix as a fraction of the measured HBM copy peak (MEASURED_PEAKS.json)
    peak = 6550.4
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except OSError:
        pass
    gbs = out["sigma_last_bytes"] / (out["sigma_last_ms"] * 1e-3) / 1e9
    out["sigma_last_GBps"] = gbs
    out["sigma_last_frac_of_hbm_peak"] = gbs / peak
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
