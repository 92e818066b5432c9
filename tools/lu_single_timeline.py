"""Where the single-GPU factorization's main stream is NOT running a trailing GEMM: timeline of one lu_factor of a
diagonally dominant N x N matrix (speculative panels, W path) from the library's event records."""
import json, sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from magpylib_material_response_b200 import _native
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24000
ctx = _native.default_context(0)
ld = (N + 15) // 16 * 16
g = torch.Generator(device="cuda"); g.manual_seed(0)
Q0 = torch.randn(N, ld, dtype=torch.float64, device="cuda", generator=g) * (0.5 / N)
Q0[:, :N] += torch.eye(N, dtype=torch.float64, device="cuda")
ipiv = torch.zeros(N, dtype=torch.int32, device="cuda")
for rep in range(2):
    Q = Q0.clone()
    ctx.set_profiling(rep == 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); info = ctx.lu_factor(Q, ld, ipiv); e1.record(); torch.cuda.synchronize()
tl = ctx.profile_timeline()
ctx.set_profiling(False)
main = [(t, d) for k, t, d in tl if k == "gemm"]
main.sort()
t_first, t_last = main[0][0], main[-1][0] + main[-1][1]
gaps = [(main[i + 1][0] - (main[i][0] + main[i][1]), i) for i in range(len(main) - 1)]
third = len(main) // 3
out = {"N": N, "factor_ms_profiled": e0.elapsed_time(e1), "main_gemms": len(main), "first_gemm_starts_ms": t_first,
       "gemm_busy_ms": sum(d for _, d in main), "span_ms": t_last - t_first,
       "gap_ms_total": sum(g for g, _ in gaps), "gap_ms_by_third": [sum(g for g, i in gaps if lo <= i < hi) for lo, hi in
                                                                      ((0, third), (third, 2 * third), (2 * third, len(main)))],
       "largest_gaps": sorted(gaps, reverse=True)[:8],
       "gemm_ms_first_mid_last": [main[0][1], main[len(main) // 2][1], main[-1][1]]}
by = {}
for k, t, d in tl:
    by.setdefault(k, [0, 0.0]); by[k][0] += 1; by[k][1] += d
out["per_kind"] = {k: (v[0], round(v[1], 2)) for k, v in sorted(by.items())}
# what runs during the 3 largest gaps
for gsz, i in sorted(gaps, reverse=True)[:3]:
    a, b = main[i][0] + main[i][1], main[i + 1][0]
    out.setdefault("during_gap", []).append({"gap_ms": round(gsz, 3), "after_gemm": i, "others": [(k, round(t - a, 3), round(d, 3)) for k, t, d in tl if k != "gemm" and t < b and t + d > a][:14]})
print(json.dumps(out))
