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
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
