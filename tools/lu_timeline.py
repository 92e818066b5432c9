"""Timeline of ONE distributed-LU step (the panel-by-panel code path, any number of ranks incl. 1) from the
library's event records: per class the summed time on the main / side stream, the per-block period, and the
gaps of the main stream (time it waits for the panel chain).  Prints one JSON line on rank 0.

    python tools/lu_timeline.py [workload] [blocks_to_list]
    torchrun --nproc-per-node 2 tools/lu_timeline.py two_magnets_8000
"""
import json, os, pathlib, sys
import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
_OUT = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
from bench import make_workload  # noqa: E402
from magpylib_material_response_b200 import _native  # noqa: E402
from magpylib_material_response_b200.demag import _padded_ld  # noqa: E402
from magpylib_material_response_b200.distributed import DEFAULT_TGT_BLOCK, owned_cells  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "two_magnets_8000"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
ctx = _native.default_context(local_rank)
if world > 1:
    ctx.comm_init_from_torch()
w = make_workload(name)
n = len(w["pos"]); N = 3 * n; ld = _padded_ld(N)
tgt_block = DEFAULT_TGT_BLOCK
n_local = len(owned_cells(n, tgt_block, world, rank))
from scipy.spatial.transform import Rotation as R
rhs = (R.from_quat(w["quat"]).apply(w["pol"]) + w["chi"] * w["H_ext"]).reshape(N)
d_pos, d_quat, d_dim = ctx.to_device(w["pos"]), ctx.to_device(w["quat"]), ctx.to_device(w["dim"])
d_chi, d_rhs = ctx.to_device(w["chi"].reshape(N)), ctx.to_device(rhs)
Q = ctx.empty(3 * n_local, ld); d_x = ctx.empty(N); ipiv = ctx.empty(N, dtype=torch.int32)


def step(profile):
    ctx.assemble(d_pos, d_quat, d_dim, Q, ld, chi=d_chi, n_tgt_local=n_local, tgt_block=tgt_block, tgt_nparts=world, tgt_part=rank)
    d_x.copy_(d_rhs)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if profile:
        ctx.set_profiling(True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    assert ctx.lu_factor_dist(N, 3 * n_local, tgt_block, Q, ld, ipiv) == 0
    e[1].record()
    ctx.lu_trisolve_dist(N, 3 * n_local, tgt_block, Q, ld, ipiv, d_x)
    e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])


for _ in range(2):
    step(False)
plain = step(False)
timed = step(True)
tl = ctx.profile_timeline()
ctx.set_profiling(False)
out = {"workload": name, "n_gpus": world, "rank": rank, "N": N, "factor_ms": plain[0], "trisolve_ms": plain[1],
       "factor_ms_profiled": timed[0], "trisolve_ms_profiled": timed[1]}
per = {}
for kind, t0, dt in tl:
    a = per.setdefault(kind, [0.0, 0])
    a[0] += dt; a[1] += 1
out["per_kind_ms"] = {k: [round(v[0], 3), v[1]] for k, v in sorted(per.items())}
# main-stream idle time: gaps between consecutive main-stream records inside the factorization
main = sorted((t0, dt) for kind, t0, dt in tl if "@side" not in kind and kind != "trisolve")
gaps = [b[0] - (a[0] + a[1]) for a, b in zip(main[:-1], main[1:])]
out["main_stream_gap_ms_total"] = round(sum(g for g in gaps if g > 0.002), 3)
out["main_stream_busy_ms"] = round(sum(dt for _, dt in main), 3)
side = sorted((t0, dt, kind) for kind, t0, dt in tl if "@side" in kind)
out["side_stream_busy_ms"] = round(sum(dt for _, dt, _ in side), 3)
# period of the panel chain: start-to-start distance of consecutive first-panel records on the side stream
pan = [t0 for t0, dt, kind in side if kind.startswith("panel@")]
if len(pan) > 4:
    d = np.diff(pan[::2])
    out["block_period_ms"] = {"first10": round(float(np.mean(d[:10])), 4), "mid": round(float(np.mean(d[len(d) // 2 - 5:len(d) // 2 + 5])), 4),
                              "last10": round(float(np.mean(d[-10:])), 4)}
nlist = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if nlist:
    t_first = pan[2 * 20] if len(pan) > 44 else 0.0
    out["events_from_block_20"] = [[k, round(t0 - t_first, 4), round(dt, 4)] for k, t0, dt in tl if t0 >= t_first][:nlist]
if rank == 0:
    _OUT.write(json.dumps(out) + "\n"); _OUT.flush()
if world > 1:
    dist.destroy_process_group()
