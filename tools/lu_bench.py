"""LU factor + solve on a random dense N x N system through the C ABI; prints stage timings."""
import json, sys, pathlib, numpy as np, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from magpylib_material_response_b200 import _native
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = _native.default_context(0)
ld = (N + 15) // 16 * 16
g = torch.Generator(device="cuda"); g.manual_seed(0)
Q0 = torch.randn(N, ld, dtype=torch.float64, device="cuda", generator=g)
Q0[:, :N] += torch.eye(N, dtype=torch.float64, device="cuda") * 4.0
b = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
ipiv = torch.zeros(N, dtype=torch.int32, device="cuda")
best = None
for r in range(reps):
    Q = Q0.clone(); x = b.clone()
    ctx.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); info = ctx.lu_factor(Q, ld, ipiv); e1.record(); ctx.lu_solve(Q, ld, ipiv, x); torch.cuda.synchronize()
    prof = ctx.profile(); recs = ctx.profile_records('panel'); ctx.set_profiling(False)
    t = e0.elapsed_time(e1)
    if best is None or t < best[0]: best = (t, prof)
y = torch.empty_like(b); ctx.matvec(Q0, ld, x, y)
res = float(torch.linalg.norm(y - b) / torch.linalg.norm(b))
import numpy as _np; _r=_np.array(recs); print("panel records: n", len(_r), "sum", _r.sum(), "max", _r.max(), "argmax", int(_r.argmax()), "n>1ms", int((_r>1).sum()), "idx>1ms", _np.nonzero(_r>1)[0][:20].tolist(), file=sys.stderr)
print(json.dumps({"N": N, "factor_ms": best[0], "tflops": (2 / 3) * N**3 / (best[0] * 1e-3) / 1e12, "info": info, "relres": res,
                  "stages_ms": {k: round(v[0], 2) for k, v in best[1].items() if v[2]}}))
