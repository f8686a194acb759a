#!/usr/bin/env python
"""bench.py -- headline benchmark of the CI hot path (H build + Davidson sigma).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A *step* is one pass of the hot path over one synthetic workload: the Hamiltonian build of
this rank's row block (b2ci_hbuild_csr) followed by one sigma = H c application
(b2ci_sigma_sharded: NCCL all-gather of the trial vector + SpMV). The default workload is
BASELINE.json configs[2] (Cr2-like CAS(12e,12o), dense synthetic integrals, 853,776 determinants,
1.55e9 non-zeros = 18.6 GB of CSR: the largest full-CI configuration that fits one GPU and the
one SURVEY.md section 8(d) sizes the sigma roofline on); configs[1] (2D extended Hubbard 4x3,
same dimension, 1.7e7 non-zeros after thresholding) is measured in the same run and reported
under "also". Other workloads are selectable with --workload.

Metric (BASELINE.json: "H-build nnz/s and Davidson sigma-iter time (ms)"): `value` is the
whole-job H-build throughput in nnz/s with the determinant list and integrals resident in
HBM; the sigma half of the metric is reported beside it (`sigma_iter_ms`, `roofline_sigma`).
`e2e` is the same H-build throughput measured through the C ABI with HOST buffers
(integrals + determinant words uploaded, row pointer read back, inside the timed region).
Rows are sharded over the N ranks (strong scaling: the problem is fixed); timing is CUDA
events on the library's stream, max over ranks. `--impl reference` times the reference's own
CPU implementation (oracle/_ref = MACIS compiled unmodified, else the plain-C oracle port) on
a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the CPU legs run OpenMP inside this process: idle worker threads must sleep, not spin, while the GPU legs
# (host-latency sensitive: ASCI growth through the plugin) run afterwards
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

EPS = float(np.finfo(np.float64).eps)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.first = 0
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def wait_first(self, timeout_s):
        t0 = time.time()
        while self.proc and not self.samples and time.time() - t0 < timeout_s:
            time.sleep(0.01)

    def mark(self):
        """samples from here on belong to the timed region"""
        self.first = len(self.samples)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples[max(0, self.first - 1):]:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def asci_leg(ntdets=100000, with_reference=True):
    """BASELINE configs[3] workload through the plugin API: N2-like ASCI(14e,26o) grown from the HF determinant
    to 1e5 determinants (search + H build + Davidson per iteration). Reported next to the headline, never part
    of it. The unmodified reference (oracle/_ref, asci_grow with the same settings) is timed on this box's host
    cores in the same run when it is available."""
    from qdk_chemistry_b200 import algorithms as alg, data
    from qdk_chemistry_b200 import workloads as W
    sp = W.config("n2_asci26")
    ham = data.Hamiltonian(sp.T, sp.V, sp.core_energy)
    calc = alg.create("multi_configuration_calculator", "b200_asci", ntdets_max=int(ntdets), max_refine_iter=0,
                      ci_residual_tolerance=1e-8)
    t0 = time.perf_counter()
    E, w = calc.run(ham, sp.nalpha, sp.nbeta)
    wall = time.perf_counter() - t0
    st = alg.last_run_stats()
    out = {"config": "BASELINE configs[3] workload: N2-like ASCI(14e,26o) grown from HF to 1e5 determinants through "
                     "the plugin API (wall time includes creating the plugin's CUDA context)",
           "ndets": int(w.size()), "E": E - sp.core_energy, "wall_s": wall,
           "asci_iterations": st.get("asci_iterations"), "nnz": st.get("nnz_local"),
           "h_build_ms": st.get("h_build_ms"), "asci_search_ms": st.get("asci_search_ms"),
           "davidson_ms": (st.get("davidson_sigma_ms") or 0.0) + (st.get("davidson_other_ms") or 0.0),
           "davidson_iterations": st.get("davidson_iterations")}
    if with_reference:
        out["reference"] = reference_asci(sp, ntdets)
        if "E" in out["reference"]:
            out["E_minus_reference"] = out["E"] - out["reference"]["E"]
            out["speedup_vs_reference_wall"] = out["reference"]["seconds"] / wall
    return out


def reference_asci(sp, ntdets):
    """The unmodified reference's asci_grow (oracle/_ref) on this box's host cores, same settings."""
    try:
        from oracle import ref
        if not ref.available():
            return {"unavailable": "oracle/_ref not built"}
        try:
            ref.set_num_threads(len(os.sched_getaffinity(0)))
        except (AttributeError, OSError):
            pass
        hg = ref.HamGen(sp.norb, sp.T, sp.V)
        t0 = time.perf_counter()
        E, dets, C = hg.asci_run(ref.AsciOpts(ntdets_max=int(ntdets), max_refine_iter=0), sp.nalpha, sp.nbeta,
                                 refine=False)
        return {"E": E, "seconds": time.perf_counter() - t0, "cores": ref.num_threads(), "ndets": int(C.size),
                "what": "macis::asci_grow, QDK defaults, measured in this run on this box"}
    except Exception as e:
        return {"error": repr(e)[:300]}


def fci_workload(name):
    from qdk_chemistry_b200 import workloads as W
    sp = W.config(name)
    return sp


def fci_config(workload, sp):
    """the `config` object of both arms (identical by construction): the workload and its size; nnz is the
    reference's own count for this workload where a full reference build is on record (tests/golden)"""
    g = golden_energy(workload)
    return {"workload": workload, "norb": sp.norb, "nalpha": sp.nalpha, "nbeta": sp.nbeta, "ndets": sp.fci_dimension,
            "nnz": (g or {}).get("nnz"), "h_thresh": EPS}


def split_rows(n, nranks):
    """contiguous equal row blocks, remainder spread over the first ranks"""
    base, rem = divmod(n, nranks)
    offs = [0]
    for r in range(nranks):
        offs.append(offs[-1] + base + (1 if r < rem else 0))
    return offs


# ---------------------------------------------------------------------------------------
# CPU legs (oracle/ is only ever the checker / the baseline, never the product path)
# ---------------------------------------------------------------------------------------
def cpu_sample(sp, budget_rows_frac=None, target_seconds=12.0):
    """Reference CPU implementation on a bounded sample of the workload: H build of the rows
    of a few alpha blocks spread over the list against ALL kets (make_csr_hamiltonian_block),
    then gespmbv on those rows. Returns dict with nnz/s, sigma ms (scaled to the full matrix)."""
    from oracle import port
    try:
        from oracle import ref
        use_ref = ref.available()
    except Exception:
        use_ref = False
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    n = len(a)
    nbeta_str = int(np.count_nonzero(a == a[0]))
    nruns = n // nbeta_str
    words = port.pack(a, b)
    out = {}
    if use_ref:
        hg = ref.HamGen(sp.norb, sp.T, sp.V)
        # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its
        # workers, which would leave the reference arm on one core at N > 1
        try:
            ref.set_num_threads(len(os.sched_getaffinity(0)))
        except (AttributeError, OSError):
            pass
        cores = ref.num_threads()
        kind = "reference"

        def build(rows_idx):
            H, sec = hg.hbuild(words[rows_idx], EPS, kets=words)
            return H, sec
    else:
        hp = port.Ham(sp.norb, sp.T, sp.V)
        cores = port.lib().op_num_threads()
        kind = "port"
    # calibrate with one alpha block, then size the sample for ~target_seconds
    def rows_of(runs):
        return np.concatenate([np.arange(r * nbeta_str, (r + 1) * nbeta_str) for r in runs])

    def timed_build(runs):
        idx = rows_of(runs)
        if use_ref:
            H, sec = build(idx)
            return H, sec, H.nnz, idx
        t0 = time.perf_counter()
        nnz = 0
        for r in runs:  # contiguous row blocks for the port
            rp, ci, nz = hp.hbuild(a, b, EPS, rows=(r * nbeta_str, (r + 1) * nbeta_str))
            nnz += int(rp[-1])
        return None, time.perf_counter() - t0, nnz, idx

    # the reference parallelises over bra alpha blocks (omp for, sorted_double_loop.hpp:145):
    # the sample must hold several blocks per thread or the cores idle
    ncal = int(min(nruns, max(16, 2 * cores)))
    cal_runs = sorted(set(np.linspace(0, nruns - 1, ncal).astype(int).tolist()))
    _, t1, nnz1, _ = timed_build(cal_runs)
    nsample = int(max(ncal, min(nruns, round(len(cal_runs) * target_seconds / max(t1, 1e-3)))))
    runs = sorted(set(np.linspace(0, nruns - 1, nsample).astype(int).tolist()))
    H, sec, nnz, idx = timed_build(runs)
    out["hbuild_nnz_per_s"] = nnz / sec
    out["hbuild_sample_seconds"] = sec
    out["sample_rows"] = int(len(idx))
    out["sample_nnz"] = int(nnz)
    out["cores"] = int(cores)
    out["kind"] = kind
    out["sample"] = (f"{len(runs)} of {nruns} alpha blocks ({len(idx)} of {n} rows, all {n} kets), "
                     f"make_csr_hamiltonian_block<int64>, H_thresh=eps; non-mirrored row block")
    out["rect_block_nnz_per_s"] = out["hbuild_nnz_per_s"]
    out["_H"], out["_runs"], out["_nbeta_str"] = H, runs, nbeta_str
    if use_ref:
        # the reference's single-process path is ONE symmetric make_csr_hamiltonian (upper triangle + mirror,
        # sorted_double_loop.hpp:96,154,183,206): timed on the sub-list of the first K alpha strings x all beta
        # strings, K sized for ~target_seconds. (The full 853,776-determinant symmetric build ran at 1.08e7 nnz/s on
        # 8 threads in the build container -- tests/golden/fullsize_meta.json -- i.e. slower per nnz than these
        # prefixes, so the prefix figure is the conservative baseline.)
        k0 = int(max(8, min(nruns, 2 * cores)))
        Hs, t0s = hg.hbuild(words[: k0 * nbeta_str], EPS)
        # cost grows ~ K^2: scale to the budget
        K = int(max(k0, min(nruns, round(k0 * (max(target_seconds, 1.0) / max(t0s, 1e-3)) ** 0.5))))
        del Hs
        Hs, ts = hg.hbuild(words[: K * nbeta_str], EPS)
        out["symmetric_prefix"] = {"alpha_strings": K, "rows": int(K * nbeta_str), "nnz": int(Hs.nnz), "seconds": ts,
                                   "nnz_per_s": Hs.nnz / ts}
        del Hs
        if out["symmetric_prefix"]["nnz_per_s"] > out["hbuild_nnz_per_s"]:
            out["hbuild_nnz_per_s"] = out["symmetric_prefix"]["nnz_per_s"]
            out["hbuild_sample_seconds"] = ts
            out["sample"] = (f"symmetric make_csr_hamiltonian<int64> (mirror shortcut) of the first {K} of {nruns} alpha "
                             f"strings x all {nbeta_str} beta strings ({K * nbeta_str} determinants), H_thresh=eps; "
                             f"beside it a rectangular block of {len(idx)} rows x all kets ran at "
                             f"{out['rect_block_nnz_per_s']:.3e} nnz/s")
    # sigma on the sampled rows (rectangular block), scaled by rows to the full matrix
    if use_ref and H is not None and H.nnz > 0:
        x = np.random.default_rng(0).normal(size=n)
        # the block is rows x n: gespmbv reads V of length n and writes len(rows)
        H.spmv(x, nrep=2)
        _, tsp = H.spmv(x, nrep=10)
        out["sigma_ms_sample"] = tsp * 1e3
        out["sigma_nnz_per_s"] = H.nnz / tsp
        out["sigma_ms_full_est"] = tsp * 1e3 * (n / len(idx))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sp = fci_workload(args.workload)
    vals, sig = [], []
    info = None
    if args.workload in ("n2_asci26", "cr2_asci30"):
        return run_reference_asci(args, sp)
    for it in range(args.warmup + args.steps):
        info = cpu_sample(sp, target_seconds=args.cpu_seconds)
        info.pop("_H", None)
        if it >= args.warmup:
            vals.append(info["hbuild_nnz_per_s"])
            sig.append(info.get("sigma_ms_full_est"))
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "hbuild_nnz_per_s", "value": v, "unit": "nnz/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(info["hbuild_sample_seconds"] * 1e3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": fci_config(args.workload, sp),
        "sigma_iter_ms": float(np.mean([s for s in sig if s is not None])) if any(sig) else None,
        "cpu_baseline": {"value": v, "unit": "nnz/s", "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"], "rect_block_nnz_per_s": info.get("rect_block_nnz_per_s"),
                         "symmetric_prefix": info.get("symmetric_prefix")},
        "e2e": {"value": v, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


ASCI_NDETS = {"n2_asci26": 100000, "cr2_asci30": 100000}


def run_reference_asci(args, sp):
    """--impl reference --workload n2_asci26|cr2_asci30: the unmodified asci_grow on the host cores; one step = one
    growth from the HF determinant to ASCI_NDETS determinants (bounded: ~10-40 s per step)."""
    nt = ASCI_NDETS[args.workload]
    secs, last = [], None
    for it in range(1 + max(1, min(args.steps, 2))):  # one warm-up, at most two timed growths
        last = reference_asci(sp, nt)
        if "seconds" not in last:
            print(json.dumps({"impl": "reference", "unavailable": json.dumps(last)[:200]}))
            return
        if it >= 1:
            secs.append(last["seconds"])
    v = float(np.mean(secs))
    print(json.dumps({
        "impl": "reference", "metric": "asci_grow_seconds", "value": v, "unit": "s", "n_gpus": args.gpus,
        "steps": len(secs), "warmup": 1, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "norb": sp.norb, "nalpha": sp.nalpha, "nbeta": sp.nbeta, "ntdets_max": nt,
                   "max_refine_iter": 0},
        "energy": last["E"], "cpu_baseline": {"value": v, "unit": "s", "cores": last["cores"], "kind": "reference",
                                              "sample": "the whole growth (asci_grow, QDK defaults)"},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_b200_asci(args):
    """--workload n2_asci26|cr2_asci30: the plugin's ASCI growth, one step = one run() from the HF determinant."""
    import torch
    from qdk_chemistry_b200 import algorithms as alg, data
    sp = fci_workload(args.workload)
    nt = ASCI_NDETS[args.workload]
    ham = data.Hamiltonian(sp.T, sp.V, sp.core_energy)
    secs, E, st = [], None, {}
    for it in range(1 + max(1, args.steps)):
        calc = alg.create("multi_configuration_calculator", "b200_asci", ntdets_max=nt, max_refine_iter=0,
                          ci_residual_tolerance=1e-8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        E, w = calc.run(ham, sp.nalpha, sp.nbeta)
        torch.cuda.synchronize()
        if it >= 1:
            secs.append(time.perf_counter() - t0)
        st = alg.last_run_stats()
    v = float(np.mean(secs))
    print(json.dumps({
        "metric": "asci_grow_seconds", "value": v, "unit": "s", "n_gpus": 1, "steps": len(secs), "warmup": 1,
        "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "norb": sp.norb, "nalpha": sp.nalpha, "nbeta": sp.nbeta, "ntdets_max": nt,
                   "max_refine_iter": 0},
        "energy": E - sp.core_energy, "stats": {k: st.get(k) for k in (
            "asci_iterations", "h_build_ms", "asci_search_ms", "davidson_sigma_ms", "davidson_other_ms", "nnz_local")},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": int(sp.T.nbytes + sp.V.nbytes),
                "d2h_bytes_per_step": int(nt * 16)},
        "gpu_launches": int(st.get("launches", 0) or 0)}))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def load_traffic(kernel_key):
    """DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant
    kernels, from the committed `ncu --set full` capture (profiles/r02_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(p) as fh:
            return json.load(fh).get(kernel_key)
    except OSError:
        return None


class _DevArr:
    """a raw device pointer as a __cuda_array_interface__ object (zero-copy view for torch.as_tensor)"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def shard_checksums(H, n, r0, r1, world, rank, dist, torch):
    """N-independent fingerprints of the row-sharded CSR (checked by comparing the N = 1 and N = 8 lines): sha256
    of the per-row counts of the whole matrix, exact wrap-around sums of the column indices and of the matrix
    elements' bit patterns (plain and position-weighted), and a seeded sigma. Outside the timed region; the sums
    are torch reductions over the library's device arrays."""
    import hashlib
    rp = H.download_rowptr()
    counts = np.diff(rp).astype(np.int64)
    nnz_local = int(rp[-1])
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, counts.tobytes())
        offs = [None] * world
        dist.all_gather_object(offs, nnz_local)
        base = int(sum(offs[:rank]))
        counts_all = b"".join(allc)
    else:
        base, counts_all = 0, counts.tobytes()
    _, pci, pnz = H.device_ptrs()
    sums = torch.zeros(4, dtype=torch.int64, device="cuda")
    if nnz_local:
        ci = torch.as_tensor(_DevArr(pci, nnz_local, "<i4"), device="cuda")
        nz = torch.as_tensor(_DevArr(pnz, nnz_local, "<i8"), device="cuda")  # bit patterns of the doubles
        step = 1 << 26
        for b in range(0, nnz_local, step):
            e = min(nnz_local, b + step)
            wgt = (torch.arange(base + b, base + e, dtype=torch.int64, device="cuda") % 65521) + 1
            c64 = ci[b:e].to(torch.int64)
            sums[0] += c64.sum()
            sums[1] += (c64 * wgt).sum()
            sums[2] += nz[b:e].sum()
            sums[3] += (nz[b:e] * wgt).sum()
    if world > 1:
        dist.all_reduce(sums)
    # seeded sigma: y = H x on every rank's rows, z . y all-reduced
    x = np.random.default_rng(7).normal(size=n)
    z = np.random.default_rng(8).normal(size=n)
    xl = torch.from_numpy(x[r0:r1].copy()).cuda()
    xf = torch.from_numpy(x).cuda()
    yl = torch.empty(r1 - r0, dtype=torch.float64, device="cuda")
    H.sigma_sharded(xl.data_ptr(), 0 if world > 1 else xf.data_ptr(), yl.data_ptr())
    torch.cuda.synchronize()
    dot = (yl * torch.from_numpy(z[r0:r1].copy()).cuda()).sum().reshape(1)
    if world > 1:
        dist.all_reduce(dot)
    u = lambda v: int(v) & 0xFFFFFFFFFFFFFFFF
    return {"row_counts_sha256": hashlib.sha256(counts_all).hexdigest(), "colind_sum": u(sums[0]),
            "colind_weighted_sum": u(sums[1]), "nzval_bits_sum": u(sums[2]), "nzval_bits_weighted_sum": u(sums[3]),
            "sigma_seeded_dot": float(dot.item()),
            "note": "identical for every N iff the sharded CSR equals the single-GPU CSR (sigma to rounding)"}


def parity_against_reference(ctx, dets, H_ref, runs, nbeta_str):
    """rows of the alpha blocks the CPU leg built with the unmodified reference, against the GPU's rows
    b2ci_hbuild_csr(row_begin, row_end) for the same blocks: rowptr / colind / nzval compared exactly"""
    n_rows = nnz = 0
    equal = True
    first_bad = None
    ref_row = 0
    for r in runs:
        r0, r1 = r * nbeta_str, (r + 1) * nbeta_str
        B = ctx.hbuild(dets, EPS, (r0, r1))
        rp, ci, nz = B.download()
        B.free()
        rrp, rci, rnz = H_ref.rows(ref_row, ref_row + nbeta_str)
        ref_row += nbeta_str
        ok = np.array_equal(rp, rrp) and np.array_equal(ci, rci) and np.array_equal(nz, rnz)
        if not ok and first_bad is None:
            first_bad = int(r)
        equal = equal and ok
        n_rows += nbeta_str
        nnz += int(rrp[-1])
    return {"rows": int(n_rows), "nnz": int(nnz), "equal": bool(equal), "first_unequal_alpha_block": first_bad,
            "what": "rowptr, colind (int64) and nzval (bitwise) of the reference's make_csr_hamiltonian_block rows vs "
                    "b2ci_hbuild_csr row blocks, same alpha blocks"}


def golden_energy(name):
    p = os.path.join(ROOT, "tests", "golden", "fullsize_meta.json")
    try:
        with open(p) as fh:
            g = json.load(fh).get(name)
        return g
    except OSError:
        return None


def plugin_e2e(sp, args, world, torch):
    """The call a user of the reference makes: create("multi_configuration_calculator", "macis_cas").run(ham, na, nb)
    with HOST integrals in and (E, wavefunction) out -- upload, determinant generation, H build of this rank's rows,
    Davidson to ci_residual_tolerance, coefficients back. Wall clock around run()."""
    from qdk_chemistry_b200 import algorithms as alg, data
    ham = data.Hamiltonian(sp.T, sp.V, sp.core_energy)
    walls, E, st, nd = [], None, {}, 0
    for it in range(3):
        calc = alg.create("multi_configuration_calculator", "macis_cas", ci_residual_tolerance=1e-8,
                          max_solver_iterations=int(args.max_m))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        E, w = calc.run(ham, sp.nalpha, sp.nbeta)
        walls.append(time.perf_counter() - t0)
        st = alg.last_run_stats()
        nd = int(w.size())
    wall = float(min(walls[1:]))
    nnz = float(st.get("nnz_local") or 0.0)
    return {"call": 'create("multi_configuration_calculator","macis_cas").run(hamiltonian, n_alpha, n_beta)',
            "wall_s": wall, "wall_s_first_call": float(walls[0]), "E_total": E, "ndets": nd,
            "h_build_ms": st.get("h_build_ms"), "davidson_iterations": st.get("davidson_iterations"),
            "davidson_sigma_ms": st.get("davidson_sigma_ms"), "davidson_other_ms": st.get("davidson_other_ms"),
            "launches": st.get("launches"), "nnz_local": nnz,
            "h2d_bytes": int(sp.T.nbytes + sp.V.nbytes), "d2h_bytes": int(nd * 8 + nd * 16)}


def measure(ctx, sp, name, args, world, rank, local_rank, dist, torch, steps, warmup, full):
    """One workload: `warmup` untimed + `steps` timed passes of (H build of this rank's rows,
    one sigma), CUDA events on the library's stream, max over ranks. `full` adds the e2e leg
    (host buffers through the C ABI), Davidson to 1e-8 Eh and the clock samples."""
    from qdk_chemistry_b200 import device

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    mx = lambda x: reduce(x, dist.ReduceOp.MAX) if world > 1 else float(x)
    sm = lambda x: reduce(x, dist.ReduceOp.SUM) if world > 1 else float(x)

    n = sp.fci_dimension
    offs = split_rows(n, world)
    r0, r1 = offs[rank], offs[rank + 1]
    hbm_peak, peak_src = measured_peaks()

    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.generate_fci(sp.norb, sp.nalpha, sp.nbeta)
    x_local = torch.randn(r1 - r0, dtype=torch.float64, device="cuda")
    x_full = torch.empty(n, dtype=torch.float64, device="cuda")
    y_local = torch.empty(r1 - r0, dtype=torch.float64, device="cuda")
    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > L2 (126 MB)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    T = {k: [] for k in ("build", "fill", "count", "setup", "thresh", "sigma", "fill_kernel")}
    H = None
    sampler = ClockSampler(local_rank)
    launches0 = wall0 = 0
    for it in range(warmup + steps):
        timed = it >= warmup
        if it == 0 and rank == 0 and full:
            # nvidia-smi is started during the warm-up: its start-up (NVML initialisation takes the
            # driver lock for tens of ms and stalls the host-synchronising steps of a build) must
            # not land in the timed region; the 100 ms samples keep coming throughout it
            sampler.start()
            sampler.wait_first(3.0)
        if it == warmup:
            barrier()
            launches0 = ctx.launch_count
            sampler.mark()
            wall0 = time.perf_counter()
        if H is not None:
            H.free()
        flush.zero_()  # evict L2 between iterations (outside the timed events)
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        H = ctx.hbuild(dets, EPS, (r0, r1))
        if world > 1:
            H.set_row_partition(offs)
        e1.record()
        flush.zero_()
        e2.record()
        # x_full = NULL: the library's own exchange buffers are used (peer-to-peer stores over NVLink)
        H.sigma_sharded(x_local.data_ptr(), 0 if world > 1 else x_full.data_ptr(), y_local.data_ptr())
        e3.record()
        torch.cuda.synchronize()
        if timed:
            T["build"].append(e0.elapsed_time(e1))
            T["sigma"].append(e2.elapsed_time(e3))
            for k in ("count", "fill", "setup", "thresh", "fill_kernel"):
                T[k].append(ctx.timer_ms("h_build." + k))
    barrier()
    wall1 = time.perf_counter()
    if os.environ.get("B2CI_BENCH_VERBOSE") and rank == 0:
        print(name, {k: [round(x, 3) for x in v] for k, v in T.items()}, file=sys.stderr)
    clocks = sampler.stop() if (rank == 0 and full) else None
    launches = ctx.launch_count - launches0
    nnz_local = H.nnz
    group = ctx.timer_ms("h_build.group_width")
    slices = ctx.timer_ms("h_build.smem_slices")
    dense_mode = int(ctx.timer_ms("h_build.dense_fill"))

    res = {"workload": name}
    # ---- e2e leg: same H build through the C ABI with HOST buffers (pinned determinant words,
    # integrals), row pointer (= per-row nnz, the step's result) read back, per step
    if full:
        words_host = dets.download(1)
        words_pinned = torch.from_numpy(words_host.view(np.int64)).pin_memory()
        t_e2e = []
        ne_w, ne_s = max(1, warmup // 2), max(2, steps)
        for it in range(ne_w + ne_s):
            H.free()
            H = None
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            ctx.upload_integrals(sp.norb, sp.T, sp.V)
            d2 = ctx.upload_dets(words_pinned.numpy().view(np.uint64), 1)
            H = ctx.hbuild(d2, EPS, (r0, r1))
            H.download_rowptr()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            d2.free()
            if it >= ne_w:
                t_e2e.append((t1 - t0) * 1e3)
        res["e2e_ms"] = mx(float(np.mean(t_e2e)))
        res["h2d"] = int(sp.T.nbytes + sp.V.nbytes + words_host.nbytes)
        res["d2h"] = int((r1 - r0 + 1) * 8)

    nrows = r1 - r0
    res.update(
        n=n, nnz_total=sm(nnz_local), nnz_local=int(nnz_local),
        build_ms=mx(float(np.mean(T["build"]))), sigma_ms=mx(float(np.mean(T["sigma"]))),
        fill_ms=mx(float(np.mean(T["fill"]))), count_ms=mx(float(np.mean(T["count"]))),
        setup_ms=mx(float(np.mean(T["setup"]))), thresh_ms=mx(float(np.mean(T["thresh"]))),
        launches=int(launches), clocks=clocks, p2p=ctx.timer_ms("comm.p2p"), wall=wall1 - wall0, group=int(group), slices=bool(slices), dense=dense_mode,
        # algorithmic bytes per launch on THIS rank
        B_sigma=int(nnz_local * 12 + (nrows + 1) * 8 + n * 8 + nrows * 8),
        B_fill=int(n * 16 + nnz_local * 12 + (nrows + 1) * 8),
        fill_ms_local=float(np.mean(T["fill"])), sigma_ms_local=float(np.mean(T["sigma"])),
        fill_kernel_ms_local=float(np.mean(T["fill_kernel"])),
        hbm_peak=hbm_peak, peak_src=peak_src)

    if full and args.davidson:
        try:
            E, X, niter, _ = H.davidson(args.max_m, 1e-8)
            ncalls = max(1.0, ctx.timer_ms("davidson.OP_CALLS"))
            other = sum(ctx.timer_ms("davidson." + k) for k in ("RR_DUR", "RES_DUR", "GS_DUR"))
            res["davidson"] = {"niter": int(niter), "E0_electronic": E, "E0_total": E + sp.core_energy,
                               "sigma_ms_mean": ctx.timer_ms("davidson.OP_DUR") / ncalls,
                               "other_ms_per_iter": other / max(1, niter),
                               "rr_ms_total": ctx.timer_ms("davidson.RR_DUR"),
                               "res_ms_total": ctx.timer_ms("davidson.RES_DUR"),
                               "gs_ms_total": ctx.timer_ms("davidson.GS_DUR")}
        except device.B2ciError as e:
            res["davidson"] = {"error": str(e)}
    if full:
        try:
            res["shard_checksums"] = shard_checksums(H, n, r0, r1, world, rank, dist, torch)
        except Exception as e:  # reported, never required
            res["shard_checksums"] = {"error": repr(e)[:300]}
    H.free()
    H = None
    if full and world == 1 and rank == 0 and args.cpu_seconds > 0:
        # CPU leg here (the determinant list is still resident): reference rows vs GPU rows
        try:
            cpu = cpu_sample(sp, target_seconds=args.cpu_seconds)
            Href, runs, nbs = cpu.pop("_H", None), cpu.pop("_runs", None), cpu.pop("_nbeta_str", None)
            if Href is not None:
                res["parity_rows"] = parity_against_reference(ctx, dets, Href, runs, nbs)
            del Href
            res["cpu"] = cpu
        except Exception as e:
            res["cpu"] = {"error": repr(e)[:300]}
    dets.free()
    del flush, x_full
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    from qdk_chemistry_b200 import device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream().cuda_stream
    ctx = device.Context(local_rank, stream)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(device.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    sp = fci_workload(args.workload)
    m = measure(ctx, sp, args.workload, args, world, rank, local_rank, dist, torch, args.steps, args.warmup, True)
    ctx.trim()  # the plugin leg below builds the same matrix in its own context
    plug = None
    if args.plugin_e2e:
        try:
            from qdk_chemistry_b200 import algorithms as alg
            if world > 1:
                alg.init_distributed_from_torch(local_rank)
            else:
                alg.set_device(local_rank)
            plug = plugin_e2e(sp, args, world, torch)
        except Exception as e:
            plug = {"error": repr(e)[:300]}
    also = None
    if args.also and args.workload != "hubbard_4x3":
        sp2 = fci_workload("hubbard_4x3")
        a = measure(ctx, sp2, "hubbard_4x3", args, world, rank, local_rank, dist, torch, max(3, args.steps),
                    args.warmup, True)
        also = {"hubbard_4x3": {
            "config": "BASELINE configs[1], 853,776 dets; structurally connected entries are evaluated "
                      "and |h| <= eps ones dropped (nnz is the surviving count)",
            "hbuild_nnz_per_s": a["nnz_total"] / (a["build_ms"] * 1e-3), "nnz": int(a["nnz_total"]),
            "hbuild_ms": a["build_ms"], "hbuild_fill_ms": a["fill_ms"], "sigma_iter_ms": a["sigma_ms"],
            "sigma_frac_of_hbm": a["B_sigma"] / (a["sigma_ms_local"] * 1e-3) / 1e9 / a["hbm_peak"],
            "e2e_nnz_per_s": a["nnz_total"] / (a["e2e_ms"] * 1e-3), "e2e_ms": a["e2e_ms"],
            "davidson": a.get("davidson")}}

    # configs[3] beside the headline (single GPU only; reported, never required)
    if args.also and world == 1 and not args.no_asci:
        try:
            asci = asci_leg()
        except Exception as e:
            asci = {"error": repr(e)[:300]}
        also = dict(also or {})
        also["n2_asci26_1e5"] = asci

    cpu = m.get("cpu")

    if rank == 0:
        # the dominant kernel's own launch duration (events around that launch alone); the compacting fill is
        # a single launch, so there the phase and the kernel coincide
        kern_ms = m["fill_kernel_ms_local"] if m.get("dense") == 2 and m["fill_kernel_ms_local"] > 0 else m["fill_ms_local"]
        fill_gbs = m["B_fill"] / (kern_ms * 1e-3) / 1e9
        phase_gbs = m["B_fill"] / (m["fill_ms_local"] * 1e-3) / 1e9
        sig_gbs = m["B_sigma"] / (m["sigma_ms_local"] * 1e-3) / 1e9
        if m.get("dense") == 2:
            rect = ("k_rows_dense<EVAL> (H-build fill of a uniform list with dense integrals: position-ordered, one "
                    "contiguous aligned store range per warp instruction); 'phase' adds its four pre-pass launches "
                    "(k_dense_tables, k_dense_rec, k_dense_sab x2)")
            rect_key = "k_rows_dense"
        else:
            rect = "k_rows_product<EVAL,G=%d,SLICES=%d,DENSE=%d> (H-build fill: evaluates and writes the CSR)" % (
                m["group"], int(m["slices"]), int(m.get("dense") or 0))
            rect_key = "k_rows_product"
        line = {
            "metric": "hbuild_nnz_per_s", "value": m["nnz_total"] / (m["build_ms"] * 1e-3), "unit": "nnz/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["build_ms"] + m["sigma_ms"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": fci_config(args.workload, sp),
            "config_detail": {"nnz_built": int(m["nnz_total"]), "row_sharding": f"{world} contiguous row blocks",
                              "l2": "192 MiB buffer rewritten between timed kernels"},
            "hbuild_ms": m["build_ms"], "hbuild_setup_ms": m["setup_ms"], "hbuild_count_ms": m["count_ms"],
            "hbuild_fill_ms": m["fill_ms"], "hbuild_thresh_ms": m["thresh_ms"],
            "sigma_iter_ms": m["sigma_ms"], "sigma_nnz_per_s": m["nnz_total"] / (m["sigma_ms"] * 1e-3),
            "roofline": {"kernel": rect, "bound": "hbm", "achieved": fill_gbs, "peak": m["hbm_peak"],
                         "unit": "GB/s", "frac": fill_gbs / m["hbm_peak"],
                         "traffic": load_traffic(rect_key) if args.workload == "cr2_cas12" else None,
                         "bytes_per_launch": m["B_fill"], "peak_source": m["peak_src"], "kernel_ms": kern_ms,
                         "phase": {"ms": m["fill_ms_local"], "achieved": phase_gbs, "frac": phase_gbs / m["hbm_peak"]},
                         "note": "algorithmic bytes = determinants read + CSR (12 B/nnz) + rowptr written by "
                                 "rank 0's launch; duration = CUDA events around the launch on the library stream"},
            "roofline_sigma": {"kernel": "k_spmv", "bound": "hbm", "achieved": sig_gbs, "peak": m["hbm_peak"],
                               "unit": "GB/s", "frac": sig_gbs / m["hbm_peak"],
                               "traffic": load_traffic("k_spmv") if args.workload == "cr2_cas12" else None,
                               "bytes_per_launch": m["B_sigma"]},
            "e2e": {"value": m["nnz_total"] / (m["e2e_ms"] * 1e-3), "unit": "nnz/s", "ms": m["e2e_ms"],
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "what": "b2ci_integrals_upload + b2ci_dets_upload + b2ci_hbuild_csr + row pointer back, host buffers",
                    "plugin_run": plug},
            "gpu_launches": m["launches"], "clocks": m["clocks"],
            "sigma_exchange": ("single GPU" if world == 1 else
                               ("peer-to-peer stores over NVLink (k_push + flag wait)" if m["p2p"] == 1.0
                                else "ncclAllGather")), "davidson": m.get("davidson"),
            "energy_total": (m.get("davidson") or {}).get("E0_total"),
            "wall_s_timed_region": m["wall"],
        }
        # what keeps the step from scaling with N (max over ranks of every phase; the string-table setup is replicated
        # on every rank, fill / count / sigma shard with the rows)
        other_ms = max(0.0, m["build_ms"] - m["setup_ms"] - m["count_ms"] - m["fill_ms"] - m["thresh_ms"])
        parts = {"replicated_setup_ms": m["setup_ms"], "count_ms": m["count_ms"], "fill_ms": m["fill_ms"],
                 "thresh_ms": m["thresh_ms"], "host_gaps_ms": other_ms}
        line["scaling_limiter"] = {
            "hbuild": dict(parts, largest_non_sharded=max(("replicated_setup_ms", "host_gaps_ms"), key=lambda q: parts[q]),
                           non_sharded_share=(m["setup_ms"] + other_ms) / m["build_ms"]),
            "sigma": {"ms": m["sigma_ms"], "exchange": line["sigma_exchange"],
                      "note": "at N > 1 the product reads the gathered vector; exchange and product run back to back "
                              "(the overlapped variant measured slower, DESIGN.md section 5)"}}
        g = golden_energy(args.workload)
        dav = m.get("davidson") or {}
        line["parity"] = {
            "rows_vs_reference": m.get("parity_rows"),
            "energy": None if not (g and "E0_electronic" in dav) else {
                "E0_electronic": dav["E0_electronic"], "reference_E0_electronic": g["E0_electronic"],
                "abs_err": abs(dav["E0_electronic"] - g["E0_electronic"]), "tolerance": 1e-8,
                "ok": bool(abs(dav["E0_electronic"] - g["E0_electronic"]) < 1e-8),
                "iterations": dav.get("niter"), "reference_iterations": g["davidson_iterations"],
                "source": "tests/golden/fullsize_meta.json: the unmodified reference's full symmetric build + davidson "
                          "on this workload (make_golden_fullsize.py)"},
            "nnz_equals_reference": None if not g else bool(int(m["nnz_total"]) == g["nnz"]),
            "shard_checksums": m.get("shard_checksums")}
        if also is not None:
            line["also"] = also
        if cpu is not None:
            if "error" in cpu:
                line["cpu_baseline"] = cpu
            else:
                line["cpu_baseline"] = {"value": cpu["hbuild_nnz_per_s"], "unit": "nnz/s",
                                        "cores": cpu["cores"], "kind": cpu["kind"],
                                        "sample": cpu["sample"],
                                        "sigma_ms_full_est": cpu.get("sigma_ms_full_est"),
                                        "sigma_nnz_per_s": cpu.get("sigma_nnz_per_s")}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cr2_cas12",
                    choices=["hubbard_4x3", "cr2_cas12", "n2_cas10", "small_cas8", "hubbard_4x2", "n2_asci26",
                             "cr2_asci30"])
    ap.add_argument("--max-m", type=int, default=100, dest="max_m")
    ap.add_argument("--no-davidson", action="store_false", dest="davidson")
    ap.add_argument("--no-asci", action="store_true", dest="no_asci",
                    help="skip the ASCI (BASELINE configs[3]) leg reported under 'also'")
    ap.add_argument("--no-also", action="store_false", dest="also",
                    help="skip the secondary hubbard_4x3 measurement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, dest="cpu_seconds",
                    help="CPU work budget of the cpu_baseline sample (0 disables)")
    ap.add_argument("--no-plugin-e2e", action="store_false", dest="plugin_e2e",
                    help="skip the end-to-end leg through the plugin API")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ASCI_NDETS:
        run_b200_asci(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
