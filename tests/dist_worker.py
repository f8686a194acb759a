"""Worker of the multi-GPU tests (one process per GPU, launched by torchrun): row-sharded H
build, sigma (NCCL all-gather + SpMV) and Davidson through the C ABI and through the plugin
layer, checked against the single-GPU result computed by every rank. Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qdk_chemistry_b200 import algorithms as alg  # noqa: E402
from qdk_chemistry_b200 import data, device  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

EPS = float(np.finfo(np.float64).eps)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world}
    sp = W.config("small_cas8")
    # ---- C ABI level
    ctx = device.Context(local, torch.cuda.current_stream().cuda_stream)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(device.Context.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.generate_fci(sp.norb, sp.nalpha, sp.nbeta)
    n = len(dets)
    # deliberately unequal blocks: exercises the grouped-broadcast all-gather
    cuts = [0] + [int(n * (r + 1) / world) - (7 if r % 2 == 0 and r + 1 < world else 0) for r in range(world)]
    cuts[-1] = n
    r0, r1 = cuts[rank], cuts[rank + 1]
    Hloc = ctx.hbuild(dets, EPS, (r0, r1))
    # single-GPU reference on a second, communicator-free context
    ctx1 = device.Context(local, torch.cuda.current_stream().cuda_stream)
    ctx1.upload_integrals(sp.norb, sp.T, sp.V)
    d1 = ctx1.generate_fci(sp.norb, sp.nalpha, sp.nbeta)
    Hfull = ctx1.hbuild(d1, EPS)
    rp, ci, nz = Hfull.download()
    rpl, cil, nzl = Hloc.download()
    out["block_bit_exact"] = bool(np.array_equal(rpl, rp[r0:r1 + 1] - rp[r0]) and
                                  np.array_equal(cil, ci[rp[r0]:rp[r1]]) and
                                  np.array_equal(nzl, nz[rp[r0]:rp[r1]]))
    x = np.random.default_rng(0).normal(size=n)
    xl = torch.from_numpy(x[r0:r1].copy()).cuda()
    xf = torch.empty(n, dtype=torch.float64, device="cuda")
    yl = torch.empty(r1 - r0, dtype=torch.float64, device="cuda")
    os.environ["B2CI_SIGMA_OVERLAP"] = "1"   # exercise the overlapped exchange whatever the rank count
    Hloc.sigma_sharded(xl.data_ptr(), xf.data_ptr(), yl.data_ptr())
    torch.cuda.synchronize()
    yref = Hfull.spmv(x)
    out["gather_exact"] = bool(np.array_equal(xf.cpu().numpy(), x))
    # overlapped exchange: own columns first, the others after the wait -- a different (fixed) association
    # order of the FP64 sums than the single-GPU kernel, so the bar is the sigma bound of the parity tests:
    # |dy| <= 1e-13 * sum |h||x| per row
    bound = 1e-13 * ctx1.upload_csr(rp, ci, np.abs(nz)).spmv(np.abs(x))[r0:r1] + 1e-300
    out["sigma_within_bound"] = bool(np.all(np.abs(yl.cpu().numpy() - yref[r0:r1]) <= bound))
    # the plain path (exchange, then one SpMV over all columns) is bit-identical to the single-GPU product
    os.environ["B2CI_NO_SIGMA_OVERLAP"] = "1"
    Hloc.sigma_sharded(xl.data_ptr(), xf.data_ptr(), yl.data_ptr())
    torch.cuda.synchronize()
    del os.environ["B2CI_NO_SIGMA_OVERLAP"]
    out["sigma_bit_exact"] = bool(np.array_equal(yl.cpu().numpy(), yref[r0:r1]))
    out["p2p"] = ctx.timer_ms("comm.p2p")
    # repeated sigmas with changing vectors (exchange-buffer parity, epochs) and x_full = NULL
    ok = True
    for rep in range(5):
        xr = np.random.default_rng(100 + rep).normal(size=n)
        xl2 = torch.from_numpy(xr[r0:r1].copy()).cuda()
        Hloc.sigma_sharded(xl2.data_ptr(), 0, yl.data_ptr())
        torch.cuda.synchronize()
        ok = ok and np.allclose(yl.cpu().numpy(), Hfull.spmv(xr)[r0:r1], rtol=0, atol=1e-11)
    out["sigma_repeat_exact"] = bool(ok)
    # the same with the split supplied by the caller (no exchange of block sizes)
    Hloc2 = ctx.hbuild(dets, EPS, (r0, r1))
    Hloc2.set_row_partition(cuts)
    xf.zero_(); yl.zero_()
    Hloc2.sigma_sharded(xl.data_ptr(), xf.data_ptr(), yl.data_ptr())
    torch.cuda.synchronize()
    out["sigma_within_bound"] = bool(out["sigma_within_bound"] and np.all(np.abs(yl.cpu().numpy() - yref[r0:r1]) <= bound))
    try:
        Hloc2.set_row_partition([0] + [c + 1 for c in cuts[1:-1]] + [n])
        out["bad_partition_rejected"] = world == 1
    except device.B2ciError:
        out["bad_partition_rejected"] = True
    E1, X1, it1, _ = Hfull.davidson(200, 1e-8)
    Ed, Xd, itd, _ = Hloc.davidson(200, 1e-8)
    out.update(E_single=E1, E_sharded=Ed, niter_single=it1, niter_sharded=itd,
               overlap=float(abs(X1 @ Xd)), norm=float(Xd @ Xd))
    ctx.close()
    ctx1.close()
    # ---- plugin level
    alg.init_distributed_from_torch(local)
    ham = data.Hamiltonian(sp.T, sp.V, sp.core_energy)
    Ep, wp = alg.create("multi_configuration_calculator", "macis_cas", ci_residual_tolerance=1e-8).run(
        ham, sp.nalpha, sp.nbeta)
    out.update(E_plugin=Ep - sp.core_energy, plugin_norm=wp.norm(), plugin_ndets=wp.size())
    # (connection-balanced row blocks for the ASCI lists, which are far below the size where that starts by default)
    os.environ["B2CI_BALANCE_MIN"] = "16"
    Ea, wa = alg.create("multi_configuration_calculator", "macis_asci", ntdets_max=600,
                        ci_residual_tolerance=1e-8).run(ham, sp.nalpha, sp.nbeta)
    out["asci_row_partition_max_over_mean"] = alg.last_run_stats().get("row_partition_max_over_mean")
    alg.clear_communicator()
    Ea1, wa1 = alg.create("multi_configuration_calculator", "macis_asci", ntdets_max=600,
                          ci_residual_tolerance=1e-8).run(ham, sp.nalpha, sp.nbeta)
    out.update(E_asci_sharded=Ea, E_asci_single=Ea1, asci_ndets=wa.size(),
               asci_same_dets=bool(np.array_equal(wa.determinant_words(), wa1.determinant_words())))
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("DIST_RESULT " + json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
