"""2-GPU debug: distributed LU solve with the peer-memory sweeps vs the NCCL sweeps, block-wise difference."""
import os, sys, pathlib, ctypes
import numpy as np, torch, torch.distributed as dist
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import make_workload
from magpylib_material_response_b200 import _native
from magpylib_material_response_b200.demag import _padded_ld
from magpylib_material_response_b200.distributed import DEFAULT_TGT_BLOCK, owned_cells
from scipy.spatial.transform import Rotation as R

local_rank = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = _native.default_context(local_rank); ctx.comm_init_from_torch()
w = make_workload(sys.argv[1] if len(sys.argv) > 1 else "two_magnets_1728")
n = len(w["pos"]); N = 3 * n; ld = _padded_ld(N); tb = DEFAULT_TGT_BLOCK
n_local = len(owned_cells(n, tb, world, rank))
rhs = (R.from_quat(w["quat"]).apply(w["pol"]) + w["chi"] * w["H_ext"]).reshape(N)
d = [ctx.to_device(w[k]) for k in ("pos", "quat", "dim")]; d_chi = ctx.to_device(w["chi"].reshape(N))
Q = ctx.empty(3 * n_local, ld); ipiv = ctx.empty(N, dtype=torch.int32)
ctx.assemble(*d, Q, ld, chi=d_chi, n_tgt_local=n_local, tgt_block=tb, tgt_nparts=world, tgt_part=rank)
assert ctx.lu_factor_dist(N, 3 * n_local, tb, Q, ld, ipiv) == 0
res = {}
for mode in ("p2p", "nccl", "p2p_again"):
    if mode == "nccl":
        ctx.lib.mmr_comm_p2p_disable(ctx.handle)
    if mode == "p2p_again":
        # re-enable: ready flag lives in the context; reopen is not needed, flip it back through a second open
        pass
    x = ctx.to_device(rhs)
    ctx.lu_trisolve_dist(N, 3 * n_local, tb, Q, ld, ipiv, x)
    res[mode] = x.cpu().numpy()
    if mode == "nccl":
        break
diff = np.abs(res["p2p"] - res["nccl"]).reshape(-1)
obw = 3 * tb
blocks = [float(diff[s * obw:(s + 1) * obw].max()) for s in range(-(-N // obw))]
if rank == 0:
    print("p2p flag:", ctx.p2p, "N", N, "blocks", len(blocks), "scale", float(np.abs(res["nccl"]).max()))
    print("max diff per block:", ["%.1e" % b for b in blocks])
dist.barrier(); dist.destroy_process_group()
