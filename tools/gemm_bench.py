"""Times the DMMA DGEMM (mmr_dgemm_nn) on the LU trailing-update shape: C(m x m) -= A(m x k) B(k x m)."""
import json, sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from magpylib_material_response_b200 import _native
ctx = _native.default_context(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
k = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ld = m + 16
A = torch.randn(k, ld, dtype=torch.float64, device="cuda")
B = torch.randn(m, k + 16, dtype=torch.float64, device="cuda")
C = torch.randn(m, ld, dtype=torch.float64, device="cuda")
ref = (C[:, :m].T - A[:, :m].T @ B[:, :k].T).T.contiguous()
ctx.dgemm_nn(m, m, k, -1.0, A, ld, B, k + 16, 1.0, C, ld); ctx.sync()
err = float((C[:, :m] - ref).abs().max())
best = 1e9
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.dgemm_nn(m, m, k, -1.0, A, ld, B, k + 16, 1.0, C, ld); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"m": m, "k": k, "ms": best, "tflops": 2.0 * m * m * k / (best * 1e-3) / 1e12, "max_abs_err": err}))
