"""Times the speculative panel kernels (diag-block LU + triangular inverses) in isolation: LU of a
diagonally dominant N x N matrix, N small, stage 'panel' = lu_diag_nopiv_kernel + lu_tri_inv_kernel."""
import json, sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from magpylib_material_response_b200 import _native
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = _native.default_context(0)
ld = (N + 15) // 16 * 16
g = torch.Generator(device="cuda"); g.manual_seed(0)
Q0 = torch.randn(N, ld, dtype=torch.float64, device="cuda", generator=g)
Q0[:, :N] += torch.eye(N, dtype=torch.float64, device="cuda") * 8.0 * N ** 0.5
ipiv = torch.zeros(N, dtype=torch.int32, device="cuda")
Q = Q0.clone(); ctx.lu_factor(Q, ld, ipiv)
ctx.set_profiling(True)
for _ in range(reps):
    Q.copy_(Q0); ctx.lu_factor(Q, ld, ipiv)
torch.cuda.synchronize()
prof = ctx.profile(); ctx.set_profiling(False)
print(json.dumps({"N": N, "us_per_call": {k: round(1e3 * v[0] / v[2], 2) for k, v in prof.items() if v[2]}}))
