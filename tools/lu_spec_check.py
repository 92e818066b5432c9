"""LU factor + solve of a diagonally dominant system (speculative panels, deferred TRSM: the demag code path);
prints the relative residual.  Used to compare GEMM variants through the env switches of lu.cu."""
import json, os, sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from magpylib_material_response_b200 import _native
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5184
ctx = _native.default_context(0)
ld = (N + 15) // 16 * 16
g = torch.Generator(device="cuda"); g.manual_seed(0)
Q0 = torch.randn(N, ld, dtype=torch.float64, device="cuda", generator=g) * (0.5 / N)
Q0[:, :N] += torch.eye(N, dtype=torch.float64, device="cuda")
b = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
ipiv = torch.zeros(N, dtype=torch.int32, device="cuda")
Q = Q0.clone(); x = b.clone()
info = ctx.lu_factor(Q, ld, ipiv); ctx.lu_solve(Q, ld, ipiv, x); torch.cuda.synchronize()
y = torch.empty_like(b); ctx.matvec(Q0, ld, x, y)
res = float(torch.linalg.norm(y - b) / torch.linalg.norm(b))
print(json.dumps({"N": N, "info": info, "relres": res, "env": {k: v for k, v in os.environ.items() if k.startswith("MMR_")}}))

