"""cuBLAS DGEMM / cuSOLVER getrf throughput on this GPU (peak PROXIES for the LU roofline; not
part of the product path).  Prints one JSON line."""
import json, time, torch
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best
out = {}
for n in (8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    out[f"cublas_dgemm_{n}_tflops"] = 2 * n**3 / t(lambda: a @ b) / 1e12
# rank-128 update shape used by the LU trailing update
n, k = 16384, 128
a = torch.randn(n, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, n, dtype=torch.float64, device="cuda"); c = torch.randn(n, n, dtype=torch.float64, device="cuda")
out["cublas_rank128_update_16384_tflops"] = 2 * n * n * k / t(lambda: torch.addmm(c, a, b, alpha=-1.0, out=c)) / 1e12
for n in (8192, 24000):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda") + n ** 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    out[f"cusolver_getrf_{n}_tflops"] = (2 / 3) * n**3 / t(lambda: torch.linalg.lu_factor(a), reps=2) / 1e12
print(json.dumps(out))
