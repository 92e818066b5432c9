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
