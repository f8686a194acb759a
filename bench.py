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


def asci_leg(ntdets=100000):
    """BASELINE configs[3] workload through the plugin API: N2-like ASCI(14e,26o) grown from the HF determinant
    to 1e5 determinants (search + H build + Davidson per iteration). Reported next to the headline, never part
    of it. The unmodified reference ends this growth at E = -21.321911887616803 after 19.2 s on 8 CPU threads
    (profiles/r01_reference_cpu_asci_n2_14e26o.json)."""
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
    ref_E = -21.321911887616803
    return {"config": "BASELINE configs[3] workload: N2-like ASCI(14e,26o) grown from HF to 1e5 determinants through "
                      "the plugin API (wall time includes creating the plugin's CUDA context)",
            "ndets": int(w.size()), "E": E - sp.core_energy, "E_minus_reference": (E - sp.core_energy) - ref_E,
            "reference_cpu_seconds_8_threads": 19.2, "wall_s": wall,
            "asci_iterations": st.get("asci_iterations"), "nnz": st.get("nnz_local"),
            "h_build_ms": st.get("h_build_ms"), "asci_search_ms": st.get("asci_search_ms"),
            "davidson_ms": (st.get("davidson_sigma_ms") or 0.0) + (st.get("davidson_other_ms") or 0.0),
            "davidson_iterations": st.get("davidson_iterations")}


def fci_workload(name):
    from qdk_chemistry_b200 import workloads as W
    sp = W.config(name)
    return sp


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
    for it in range(args.warmup + args.steps):
        info = cpu_sample(sp, target_seconds=args.cpu_seconds)
        if it >= args.warmup:
            vals.append(info["hbuild_nnz_per_s"])
            sig.append(info.get("sigma_ms_full_est"))
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "hbuild_nnz_per_s", "value": v, "unit": "nnz/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(info["hbuild_sample_seconds"] * 1e3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "norb": sp.norb, "nalpha": sp.nalpha, "nbeta": sp.nbeta,
                   "ndets": sp.fci_dimension, "h_thresh": EPS},
        "sigma_iter_ms": float(np.mean([s for s in sig if s is not None])) if any(sig) else None,
        "cpu_baseline": {"value": v, "unit": "nnz/s", "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": v, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def load_traffic(kernel_key):
    """DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant
    kernels, from the committed `ncu --set full` capture (profiles/r01_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(p) as fh:
            return json.load(fh).get(kernel_key)
    except OSError:
        return None


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
    T = {k: [] for k in ("build", "fill", "count", "setup", "thresh", "sigma")}
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
            for k in ("count", "fill", "setup", "thresh"):
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
        launches=int(launches), clocks=clocks, p2p=ctx.timer_ms("comm.p2p"), wall=wall1 - wall0, group=int(group), slices=bool(slices),
        # algorithmic bytes per launch on THIS rank
        B_sigma=int(nnz_local * 12 + (nrows + 1) * 8 + n * 8 + nrows * 8),
        B_fill=int(n * 16 + nnz_local * 12 + (nrows + 1) * 8),
        fill_ms_local=float(np.mean(T["fill"])), sigma_ms_local=float(np.mean(T["sigma"])),
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
    H.free()
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

    cpu = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        try:
            cpu = cpu_sample(sp, target_seconds=args.cpu_seconds)
        except Exception as e:  # the baseline is reported, never required
            cpu = {"error": repr(e)}

    if rank == 0:
        fill_gbs = m["B_fill"] / (m["fill_ms_local"] * 1e-3) / 1e9
        sig_gbs = m["B_sigma"] / (m["sigma_ms_local"] * 1e-3) / 1e9
        rect = "k_rows_product<EVAL,G=%d,SLICES=%d> (H-build fill: evaluates and writes the CSR)" % (
            m["group"], int(m["slices"]))
        line = {
            "metric": "hbuild_nnz_per_s", "value": m["nnz_total"] / (m["build_ms"] * 1e-3), "unit": "nnz/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["build_ms"] + m["sigma_ms"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "norb": sp.norb, "nalpha": sp.nalpha,
                       "nbeta": sp.nbeta, "ndets": m["n"], "nnz": int(m["nnz_total"]), "h_thresh": EPS,
                       "row_sharding": f"{world} contiguous row blocks",
                       "l2": "192 MiB buffer rewritten between timed kernels"},
            "hbuild_ms": m["build_ms"], "hbuild_setup_ms": m["setup_ms"], "hbuild_count_ms": m["count_ms"],
            "hbuild_fill_ms": m["fill_ms"], "hbuild_thresh_ms": m["thresh_ms"],
            "sigma_iter_ms": m["sigma_ms"], "sigma_nnz_per_s": m["nnz_total"] / (m["sigma_ms"] * 1e-3),
            "roofline": {"kernel": rect, "bound": "hbm", "achieved": fill_gbs, "peak": m["hbm_peak"],
                         "unit": "GB/s", "frac": fill_gbs / m["hbm_peak"],
                         "traffic": load_traffic("k_rows_product") if args.workload == "cr2_cas12" else None,
                         "bytes_per_launch": m["B_fill"], "peak_source": m["peak_src"],
                         "note": "algorithmic bytes = determinants read + CSR (12 B/nnz) + rowptr written by "
                                 "rank 0's launch; duration = CUDA events around the launch on the library stream"},
            "roofline_sigma": {"kernel": "k_spmv", "bound": "hbm", "achieved": sig_gbs, "peak": m["hbm_peak"],
                               "unit": "GB/s", "frac": sig_gbs / m["hbm_peak"],
                               "traffic": load_traffic("k_spmv") if args.workload == "cr2_cas12" else None,
                               "bytes_per_launch": m["B_sigma"]},
            "e2e": {"value": m["nnz_total"] / (m["e2e_ms"] * 1e-3), "unit": "nnz/s", "ms": m["e2e_ms"],
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
            "gpu_launches": m["launches"], "clocks": m["clocks"],
            "sigma_exchange": ("single GPU" if world == 1 else
                               ("peer-to-peer stores over NVLink (k_push + flag wait)" if m["p2p"] == 1.0
                                else "ncclAllGather")), "davidson": m.get("davidson"),
            "energy_total": (m.get("davidson") or {}).get("E0_total"),
            "wall_s_timed_region": m["wall"],
        }
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
                    choices=["hubbard_4x3", "cr2_cas12", "n2_cas10", "small_cas8", "hubbard_4x2"])
    ap.add_argument("--max-m", type=int, default=100, dest="max_m")
    ap.add_argument("--no-davidson", action="store_false", dest="davidson")
    ap.add_argument("--no-asci", action="store_true", dest="no_asci",
                    help="skip the ASCI (BASELINE configs[3]) leg reported under 'also'")
    ap.add_argument("--no-also", action="store_false", dest="also",
                    help="skip the secondary hubbard_4x3 measurement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, dest="cpu_seconds",
                    help="CPU work budget of the cpu_baseline sample (0 disables)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
