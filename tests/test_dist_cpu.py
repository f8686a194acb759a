"""N > 1 host logic on CPU: two processes over gloo. The CUDA kernels cannot run here; what is
covered is everything around them on the sharded path -- the row partition every rank derives
on its own, the unique-id hand-over, and the identity the sharded sigma rests on (row blocks
built independently + all-gathered trial vector == the full product), with the oracle standing
in for the device kernels."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import port as op
        from qdk_chemistry_b200 import algorithms as alg
        from qdk_chemistry_b200 import workloads as W
        res = {}
        # 1. unique-id hand-over: rank 0's 128 bytes arrive everywhere
        uid = bytes(range(128)) if rank == 0 else None
        res["uid_ok"] = alg.broadcast_unique_id(uid) == bytes(range(128))
        # 2. row partition: blocks tile [0, n) in rank order, sizes differ by at most one
        tiles = {}
        for n in (0, 1, 7, 3920, 853776):
            blocks = [None] * world
            dist.all_gather_object(blocks, tuple(alg.row_block(n, rank, world)))
            tiles[n] = blocks
        res["tiles"] = tiles
        # 3. sharded sigma identity with oracle kernels
        sp = W.config("tiny_cas6")
        a, b = op.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
        n = len(a)
        h = op.Ham(sp.norb, sp.T, sp.V)
        r0, r1 = alg.row_block(n, rank, world)
        eps = float(np.finfo(np.float64).eps)
        rp, ci, nz = h.hbuild(a, b, eps, rows=(r0, r1))
        x = np.random.default_rng(1).normal(size=n)
        blocks = [None] * world
        dist.all_gather_object(blocks, x[r0:r1].copy())
        xf = np.concatenate(blocks)
        # rectangular block: rows r0..r1 against all columns
        y_loc = np.array([nz[rp[i]:rp[i + 1]] @ xf[ci[rp[i]:rp[i + 1]]] for i in range(r1 - r0)])
        ys = [None] * world
        dist.all_gather_object(ys, y_loc)
        rpf, cif, nzf = h.hbuild(a, b, eps)
        res["gather_ok"] = bool(np.array_equal(xf, x))
        res["sigma_err"] = float(np.max(np.abs(np.concatenate(ys) - op.spmv(rpf, cif, nzf, x))))
        res["nnz_sum_ok"] = None
        nnzs = [None] * world
        dist.all_gather_object(nnzs, int(rp[-1]))
        res["nnz_sum_ok"] = sum(nnzs) == int(rpf[-1])
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        r = out[rank]
        assert r["uid_ok"] and r["gather_ok"] and r["nnz_sum_ok"]
        assert r["sigma_err"] < 1e-12
        for n, blocks in r["tiles"].items():
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_row_block_argument_checks():
    from qdk_chemistry_b200 import algorithms as alg
    assert alg.row_block(10, 0, 3) == (0, 4) and alg.row_block(10, 2, 3) == (7, 10)
    with pytest.raises(ValueError):
        alg.row_block(10, 3, 3)
