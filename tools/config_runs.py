"""BASELINE configs[2] and configs[4] at full size through the public array API
(`apply_demag_arrays`, host numpy in/out), with the size-independent checks the parity tests use at
reduced size: true residual of the returned solution against a freshly assembled system, symmetry of
the solution under the lattice's mirror planes, and (config 3) dedupe table == dense assembly on a
row sample.  Prints one JSON line.  Single GPU:  python tools/config_runs.py lattice_27000 lu
Multi GPU (config 5 does not fit one GPU):  torchrun --nproc-per-node 2 tools/config_runs.py array_51200 gmres
"""
import json, os, pathlib, sys, time
import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
_OUT = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
from bench import make_workload  # noqa: E402
from magpylib_material_response_b200 import _native  # noqa: E402
from magpylib_material_response_b200.demag import apply_demag_arrays  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "lattice_27000"
solver = sys.argv[2] if len(sys.argv) > 2 else "lu"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
w = make_workload(name)
n = len(w["pos"]); N = 3 * n
kw = dict(pairs_matching=w.get("pairs_matching", False), max_dist=w.get("max_dist", 0), split=w.get("split", 1))
t0 = time.perf_counter()
x, info = apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"], solver=solver, tol=1e-12,
                             return_info=True, **kw)
torch.cuda.synchronize()
t_first = time.perf_counter() - t0
t0 = time.perf_counter()
x, info = apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"], solver=solver, tol=1e-12,
                             return_info=True, **kw)
torch.cuda.synchronize()
t_second = time.perf_counter() - t0
out = {"workload": name, "desc": w["desc"], "n_cells": n, "matrix_order": N, "matrix_GB": 8.0 * N * N / 1e9,
       "solver": solver, "n_gpus": world, "seconds_first_call": t_first, "seconds_per_apply_demag": t_second,
       "info": {k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in info.items()}, **{k: v for k, v in kw.items()}}
# mirror symmetry of the geometry (x <-> -x about the lattice centre): J_z even, J_x odd
ctr = w["pos"].mean(axis=0)
key = np.round((w["pos"] - ctr) * 1e9).astype(np.int64)
mir = key.copy(); mir[:, 0] *= -1
order = np.lexsort(key.T[::-1]); order_m = np.lexsort(mir.T[::-1])
perm = np.empty(n, dtype=np.int64); perm[order] = order_m  # cell i <-> mirrored cell perm[i]
xm = x[perm]
out["mirror_sym_err"] = float(max(np.abs(x[:, 2] - xm[:, 2]).max(), np.abs(x[:, 0] + xm[:, 0]).max()) / np.abs(x).max())
if world == 1:
    # true residual on a sample of rows of a densely assembled Q (independent of the dedupe table / LU)
    ctx = _native.default_context(local_rank)
    rng = np.random.default_rng(0)
    rows_cells = np.sort(rng.choice(n, size=min(n, 64), replace=False))
    ld = (N + 15) // 16 * 16
    d_pos, d_quat, d_dim = ctx.to_device(w["pos"]), ctx.to_device(w["quat"]), ctx.to_device(w["dim"])
    d_chi = ctx.to_device(w["chi"].reshape(N))
    from scipy.spatial.transform import Rotation as R
    rot = R.from_quat(w["quat"])
    xg = rot.apply(x).reshape(N)
    rhs = (rot.apply(w["pol"]) + w["chi"] * w["H_ext"]).reshape(N)
    worst = 0.0
    for c in rows_cells:
        Qrow = ctx.empty(3, ld)
        ctx.assemble(d_pos, d_quat, d_dim, Qrow, ld, chi=d_chi, max_dist=kw["max_dist"], tgt_first=int(c), n_tgt_local=1)
        r = Qrow[:, :N].cpu().numpy() @ xg - rhs[3 * c:3 * c + 3]
        worst = max(worst, float(np.abs(r).max()))
    out["sampled_row_residual_rel"] = worst / float(np.abs(rhs).max())
if rank == 0:
    _OUT.write(json.dumps(out) + "\n"); _OUT.flush()
if world > 1:
    dist.destroy_process_group()
