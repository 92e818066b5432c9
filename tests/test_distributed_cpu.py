"""CPU: the multi-GPU path's host logic under a real 2-process ``gloo`` group -- row-block
ownership, the NCCL-id bootstrap broadcast, and a numpy emulation of the distributed matvec
(local row-block product -> all-gather of padded slices -> un-permute) that uses the same index
formulas as the CUDA kernels (mmr_assemble's target mapping, gmres.cu's unpermute_kernel)."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from magpylib_material_response_b200 import distributed as mdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_cells, tgt_block, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert mdist.is_distributed()
        # --- bootstrap: 128-byte id travels intact from rank 0
        payload = bytes(range(128)) if rank == 0 else b""
        got = mdist.broadcast_bytes(payload, src=0)
        assert got == bytes(range(128))

        # --- ownership
        mine = mdist.owned_cells(n_cells, tgt_block, world, rank)
        t = np.arange(len(mine))
        assert np.array_equal(mdist.local_to_global_cell(t, tgt_block, world, rank), mine)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.tolist())
        allc = np.sort(np.concatenate([np.asarray(g, dtype=np.int64) for g in gathered]))
        assert np.array_equal(allc, np.arange(n_cells))  # every cell exactly once

        # --- distributed matvec emulation
        N = 3 * n_cells
        rng = np.random.default_rng(0)  # same matrix on every rank
        Q = rng.normal(size=(N, N))
        x = rng.normal(size=N)
        rows = (3 * mine[:, None] + np.arange(3)[None, :]).reshape(-1)
        y_local = Q[rows] @ x
        nblk = -(-n_cells // tgt_block)
        maxrows = 3 * tgt_block * (-(-nblk // world))
        pad = torch.zeros(maxrows, dtype=torch.float64)
        pad[: len(y_local)] = torch.from_numpy(y_local)
        bufs = [torch.zeros(maxrows, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(bufs, pad)
        flat = torch.cat(bufs).numpy()
        owner, lrow = mdist.global_row_owner(n_cells, tgt_block, world)
        y = flat[owner * maxrows + lrow]
        np.testing.assert_allclose(y, Q @ x, rtol=1e-13, atol=1e-13)
        results[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("n_cells", "tgt_block"), [(100, 8), (37, 40), (250, 40), (64, 32), (500, 80), (173, 85)])
def test_two_rank_gloo_ownership_bootstrap_and_matvec(n_cells, tgt_block):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_cells, tgt_block, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world))


def test_ownership_maps_single_process():
    for n, tb, R in [(8000, 40, 8), (64000, 40, 8), (1, 40, 2), (41, 40, 3)]:
        seen = np.concatenate([mdist.owned_cells(n, tb, R, r) for r in range(R)])
        assert np.array_equal(np.sort(seen), np.arange(n))
        counts = [len(mdist.owned_cells(n, tb, R, r)) for r in range(R)]
        assert max(counts) - min(counts) <= tb  # balanced to within one block
        owner, lrow = mdist.global_row_owner(n, tb, R)
        for r in range(R):
            mine = mdist.owned_cells(n, tb, R, r)
            rows = (3 * mine[:, None] + np.arange(3)[None, :]).reshape(-1)
            assert np.all(owner[rows] == r)
            assert np.array_equal(lrow[rows], np.arange(len(rows)))
    assert not mdist.is_distributed()


# ---------------------------------------------------------------------------------------------------------
# The distributed triangular sweeps on inverse-format factors (lu.cu: lu_sweeps_inv / lu_sweep_kernel) under a
# real 2-process gloo group: same ownership, the same far-column ranges per step, the same order of operations;
# the finished block travels with a broadcast (the GPU path: peer-memory stores + a flag, or ncclBroadcast).
# ---------------------------------------------------------------------------------------------------------
def _sweep_worker(rank, world, port, N, obw, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(1)
        A = rng.normal(size=(N, N)) * 0.1 + np.eye(N) * 4.0  # column-major view A = Q^T; no pivoting needed
        b = rng.normal(size=N)
        LU = A.copy()
        for j in range(N):
            LU[j + 1:, j] /= LU[j, j]
            LU[j + 1:, j + 1:] -= np.outer(LU[j + 1:, j], LU[j, j + 1:])
        # inverse format: every diagonal panel block (<= 128 wide, two per ownership block) holds L11^-1 \ U11^-1
        S = -(-N // obw)
        F = LU.copy()
        for s in range(S):
            k0 = s * obw
            ob = min(obw, N - k0)
            for lo, hi in ((k0, k0 + min(128, ob)), (k0 + min(128, ob), k0 + ob)):
                if hi > lo:
                    D = LU[lo:hi, lo:hi]
                    Li = np.linalg.inv(np.tril(D, -1) + np.eye(hi - lo))
                    Ui = np.linalg.inv(np.triu(D))
                    F[lo:hi, lo:hi] = np.tril(Li, -1) + Ui
        my_blocks = list(range(rank, S, world))
        cols = np.concatenate([np.zeros(0, dtype=np.int64)] + [np.arange(t * obw, min(N, (t + 1) * obw)) for t in my_blocks])
        Al = F[:, cols]  # my local columns, global rows
        lcol0 = {t: j * obw for j, t in enumerate(my_blocks)}
        x = b.copy()  # replicated; I only maintain the entries of my own columns until a block is broadcast

        def finish(fwd, s_done, s_next):
            d0, nd = s_next * obw, min(obw, N - s_next * obw)
            na = min(128, nd)
            blk = Al[:, lcol0[s_next]:lcol0[s_next] + nd]
            y = x[d0:d0 + nd].copy()
            if s_done >= 0:
                u0, nu = s_done * obw, min(obw, N - s_done * obw)
                y -= blk[u0:u0 + nu].T @ x[u0:u0 + nu]
            D = blk[d0:d0 + nd]
            z = np.empty(nd)
            if fwd:
                z[:na] = np.triu(D[:na, :na]).T @ y[:na]
                y[na:] -= D[:na, na:].T @ z[:na]
                z[na:] = np.triu(D[na:, na:]).T @ y[na:]
            else:
                z[na:] = y[na:] + np.tril(D[na:, na:], -1).T @ y[na:]
                y[:na] -= D[na:, :na].T @ z[na:]
                z[:na] = y[:na] + np.tril(D[:na, :na], -1).T @ y[:na]
            x[d0:d0 + nd] = z

        def step(fwd, s_done, s_next):
            own_next = s_next >= 0 and s_next % world == rank
            if s_done >= 0:
                u0, nu = s_done * obw, min(obw, N - s_done * obw)
                if fwd:
                    jlo = (s_done - rank) // world + 1 if s_done >= rank else 0
                    jlo += 1 if own_next else 0
                    far = range(jlo * obw, Al.shape[1])
                else:
                    nleft = (s_done - 1 - rank) // world + 1 if s_done > rank else 0
                    nleft -= 1 if own_next else 0
                    far = range(0, nleft * obw)
                for lc in far:
                    g = ((lc // obw) * world + rank) * obw + lc % obw
                    if g < N:
                        x[g] -= Al[u0:u0 + nu, lc] @ x[u0:u0 + nu]
            if own_next:
                finish(fwd, s_done, s_next)

        def bcast(s):
            d0, nd = s * obw, min(obw, N - s * obw)
            t = torch.from_numpy(x[d0:d0 + nd].copy())
            dist.broadcast(t, src=s % world)
            x[d0:d0 + nd] = t.numpy()

        step(True, -1, 0)
        bcast(0)
        for s in range(S - 1):
            step(True, s, s + 1)
            bcast(s + 1)
        step(False, -1, S - 1)
        bcast(S - 1)
        for s in range(S - 1, 0, -1):
            step(False, s, s - 1)
            bcast(s - 1)
        ref = np.linalg.solve(A.T, b)  # Q x = b with Q = A^T = U^T L^T
        results[rank] = float(np.abs(x - ref).max() / np.abs(ref).max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("N", "obw"), [(700, 240), (513, 120), (256, 256)])
def test_distributed_block_sweeps_gloo(N, obw):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_sweep_worker, args=(world, _free_port(), N, obw, results), nprocs=world, join=True)
    assert len(results) == world
    assert all(err < 1e-11 for err in results.values()), dict(results)
