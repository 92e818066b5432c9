"""Multi-GPU plumbing for the demag path: one process per GPU, ``torch.distributed`` for
bootstrap only, NCCL inside libmmr_b200 for the data path.

The demag matrix is sharded by TARGET-ROW blocks: block ``s`` (the 3*tgt_block rows of
``tgt_block`` consecutive cells) belongs to rank ``s % nranks`` (block-cyclic, so the LU's
shrinking trailing matrix stays balanced).  Assembly needs no exchange at all (every rank holds
the O(n) geometry); the solves exchange the factored panels (LU) or the matvec slices (GMRES).
"""

from __future__ import annotations

import numpy as np

DEFAULT_TGT_BLOCK = 80  # 3*80 = 240-row ownership blocks = two LU panels (128 + 112) and one rank-240 DMMA update; the distributed LU needs 3*tgt_block <= 256


def owned_blocks(n_cells, tgt_block, nranks, rank):
    """Global block indices owned by ``rank``."""
    nblk = -(-n_cells // tgt_block)
    return list(range(rank, nblk, nranks))


def owned_cells(n_cells, tgt_block, nranks, rank):
    """Global cell indices owned by ``rank`` in LOCAL order (the order of its matrix rows)."""
    parts = [
        np.arange(b * tgt_block, min(n_cells, (b + 1) * tgt_block))
        for b in owned_blocks(n_cells, tgt_block, nranks, rank)
    ]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)


def local_to_global_cell(t, tgt_block, nranks, rank):
    """The mapping the assembly kernel applies (include/mmr_b200.h, mmr_assemble)."""
    t = np.asarray(t)
    return ((t // tgt_block) * nranks + rank) * tgt_block + t % tgt_block


def global_row_owner(n_cells, tgt_block, nranks):
    """(owner rank, local row) of every global row 3*cell+comp -- what the GMRES un-permute
    kernel computes on the device."""
    cell = np.arange(n_cells)
    blk = cell // tgt_block
    owner = blk % nranks
    lcell = (blk // nranks) * tgt_block + cell % tgt_block
    rows_owner = np.repeat(owner, 3)
    rows_local = (3 * lcell[:, None] + np.arange(3)[None, :]).reshape(-1)
    return rows_owner, rows_local


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from ``src`` to every rank of the default process group (any
    backend; used to ship the 128-byte NCCL unique id)."""
    import torch.distributed as dist

    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def is_distributed():
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def ensure_comm(ctx):
    """Create the context's NCCL communicator once per process group."""
    if ctx.nranks == 1 and is_distributed():
        ctx.comm_init_from_torch()
    return ctx.nranks
