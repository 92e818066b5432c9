"""Times the assembly kernel alone (mmr_assemble, fused Q) on n cells, aligned or rotated."""
import json, sys, pathlib, numpy as np, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from magpylib_material_response_b200 import _native
from bench import make_workload
name = sys.argv[1] if len(sys.argv) > 1 else "soft_cube_8000"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w = make_workload(name)
ctx = _native.default_context(0)
n = len(w["pos"]); N = 3 * n; ld = (N + 15) // 16 * 16
d = [ctx.to_device(w[k]) for k in ("pos", "quat", "dim")]
chi = ctx.to_device(w["chi"].reshape(-1))
Q = ctx.empty(N, ld)
ctx.assemble(*d, Q, ld, chi=chi); ctx.sync()
best = 1e9
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.assemble(*d, Q, ld, chi=chi); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"workload": name, "n": n, "ms": best, "entries_per_s": 9.0 * n * n / (best * 1e-3), "pairs_per_s": n * n / (best * 1e-3)}))
