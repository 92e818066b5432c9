"""Multi-GPU parity check, launched by torchrun (one process per GPU), NOT collected by pytest:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_parity.py

Every rank solves the same reduced config-2 problem through the public array API with the
distributed LU and the distributed GMRES and compares with the CPU oracle."""
import os
import pathlib
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from bench import make_workload  # noqa: E402
from magpylib_material_response_b200.demag import apply_demag_arrays  # noqa: E402
from oracle import demag_ref as oref  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    w = make_workload("two_magnets_432")
    ref = oref.apply_demag_ref(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], H_ext=w["H_ext"])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    for solver in ("lu", "gmres"):
        got, info = apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"],
                                       solver=solver, return_info=True)
        err = float(np.max(np.abs(got - ref) / scale))
        assert info.get("n_gpus") == world, info
        assert err <= 1e-8, (solver, err)
        if rank == 0:
            print(f"multi-GPU parity ok: {world} ranks, solver={solver}, n={len(ref)}, max rel err {err:.2e}", flush=True)
    # the keyword features of the public API on the sharded path: source tiling, the
    # per-rank unique-pair table, the max_dist mask and the stored sharded factors
    for name, kw in (("two_magnets_432", {"split": 8}), ("lattice_2x5", {"pairs_matching": True}),
                     ("array_3x4", {"max_dist": 12.0, "split": 4})):
        wv = make_workload(name)
        geo = (wv["pos"], wv["quat"], wv["dim"], wv["pol"], wv["chi"])
        refv = oref.apply_demag_ref(*geo, H_ext=wv["H_ext"], max_dist=kw.get("max_dist", 0))
        for solver in ("lu", "gmres"):
            got, info = apply_demag_arrays(*geo, wv["H_ext"], solver=solver, return_info=True, **kw)
            err = float(np.abs(got - refv).max() / np.abs(refv).max())
            assert err <= 1e-8, (name, kw, solver, err)
            if kw.get("pairs_matching"):
                assert info["unique_pairs_this_rank"] is not None and info["unique_pairs_this_rank"] < len(refv) ** 2 / world
        if rank == 0:
            print(f"multi-GPU parity ok: {world} ranks, {name} {kw}, max rel err {err:.2e}", flush=True)
    store = {}
    for solver in ("lu", "gmres"):
        first, i1 = apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"], solver=solver,
                                       return_info=True, demag_store=store)
        pol2 = w["pol"] * 0.5 + 0.1
        again, i2 = apply_demag_arrays(w["pos"], w["quat"], w["dim"], pol2, w["chi"], w["H_ext"], solver=solver,
                                       return_info=True, demag_store=store)
        ref2 = oref.apply_demag_ref(w["pos"], w["quat"], w["dim"], pol2, w["chi"], H_ext=w["H_ext"])
        assert (i1["demag_store"], i2["demag_store"]) == ("miss", "hit"), (i1, i2)
        err = float(np.abs(again - ref2).max() / np.abs(ref2).max())
        assert err <= 1e-8, (solver, err)
        if rank == 0:
            print(f"multi-GPU demag_store ok: {world} ranks, solver={solver}, second rhs max rel err {err:.2e}", flush=True)

    # general dense systems through the C ABI: a diagonally dominant head (speculative panels) followed by
    # a general tail (pivoting strip kernels), row blocks dealt block-cyclically
    from magpylib_material_response_b200 import _native
    from magpylib_material_response_b200 import distributed as mdist

    ctx = _native.default_context(local_rank)
    for n, tgt_block, ndom in ((500, 80, 1500), (700, 40, 0), (900, 85, 1200)):
        N = 3 * n
        rng = np.random.default_rng(n)
        Q = rng.normal(size=(N, N))
        Q[np.arange(ndom), np.arange(ndom)] += 8.0 * np.sqrt(N)
        b = rng.normal(size=N)
        cells = mdist.owned_cells(n, tgt_block, world, rank)
        rows = (3 * cells[:, None] + np.arange(3)[None, :]).reshape(-1)
        ld = (N + 15) // 16 * 16
        dQ = torch.zeros(len(rows), ld, dtype=torch.float64, device=ctx.device)
        dQ[:, :N] = torch.from_numpy(Q[rows]).to(ctx.device)
        dx = torch.from_numpy(b).to(ctx.device)
        info = ctx.lu_solve_dist(N, len(rows), tgt_block, dQ, ld, dx)
        assert info == 0
        x = dx.cpu().numpy()
        resid = np.linalg.norm(Q @ x - b) / (np.linalg.norm(Q) * np.linalg.norm(x))
        assert resid < 1e-14, (n, tgt_block, resid)
        if rank == 0:
            print(f"multi-GPU dense LU ok: {world} ranks, N={N}, block={3 * tgt_block}, scaled residual {resid:.1e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
