"""Diagonal-block panel step (mmr_lu_diag_block) alone and while a long DMMA GEMM occupies every SM on a
lower-priority stream -- the situation of the LU look-ahead.  MMR_LU_PANEL_V1=1 selects the round-1 pair."""
import json, os, sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from magpylib_material_response_b200 import _native

nb = 128
lo, hi = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
with torch.cuda.stream(lo):
    ctx_lo = _native.default_context(0)
with torch.cuda.stream(hi):
    ctx_hi = _native.default_context(0)
g = torch.Generator(device="cuda"); g.manual_seed(0)
A = torch.randn(nb, nb + 8, dtype=torch.float64, device="cuda", generator=g) * 0.2
A[:, :nb] += torch.eye(nb, dtype=torch.float64, device="cuda") * 30.0
S = torch.zeros_like(A); L = torch.zeros(nb, nb, dtype=torch.float64, device="cuda"); U = torch.zeros_like(L)
m, k = 12288, 256
Ga = torch.randn(k, m + 16, dtype=torch.float64, device="cuda"); Gb = torch.randn(m, k + 16, dtype=torch.float64, device="cuda")
Gc = torch.randn(m, m + 16, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()


def time_panel(n):
    ts = []
    with torch.cuda.stream(hi):
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(hi); rej = ctx_hi.lu_diag_block(A, nb + 8, nb, S, nb + 8, L, U); e1.record(hi)
            e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
            assert not rej
    ts.sort()
    return {"median_us": round(ts[len(ts) // 2], 1), "min_us": round(ts[0], 1), "p90_us": round(ts[int(0.9 * len(ts))], 1)}


time_panel(5)
alone = time_panel(100)
clk = ctx_hi.lu_diag_block_clocks() if os.environ.get("MMR_LU_PANEL_V1") != "1" else []
phases = [clk[i + 1] - clk[i] for i in range(len(clk) - 1) if clk[i + 1] > 0 and clk[i] > 0]
with torch.cuda.stream(lo):
    for _ in range(40):
        ctx_lo.dgemm_nn(m, m, k, -1.0, Ga, m + 16, Gb, k + 16, 1.0, Gc, m + 16)
busy = time_panel(100)
torch.cuda.synchronize()
print(json.dumps({"variant": "v1 (one-CTA LU + 8-CTA inverses)" if os.environ.get("MMR_LU_PANEL_V1") == "1" else "cluster kernel",
                  "alone": alone, "under_gemm": busy, "phase_cycles": phases,
                  "note": "event pair around memset + kernel(s) + 4-byte D2H; 'under_gemm': 40 x (12288^2, k=256) DMMA GEMMs queued on a priority-0 stream"}))
