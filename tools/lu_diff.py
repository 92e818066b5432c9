"""Factor the same diagonally dominant matrix in this process and save it; with a second argument compare with the
saved factors of another run and print where they differ (256 x 256 block map)."""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from magpylib_material_response_b200 import _native
N = int(sys.argv[1]); out = sys.argv[2]; ref = sys.argv[3] if len(sys.argv) > 3 else None
ctx = _native.default_context(0)
ld = (N + 15) // 16 * 16
g = torch.Generator(device="cuda"); g.manual_seed(0)
Q = torch.randn(N, ld, dtype=torch.float64, device="cuda", generator=g) * (0.5 / N)
Q[:, :N] += torch.eye(N, dtype=torch.float64, device="cuda")
ipiv = torch.zeros(N, dtype=torch.int32, device="cuda")
info = ctx.lu_factor(Q, ld, ipiv); torch.cuda.synchronize()
torch.save(Q.cpu(), out)
if ref:
    R = torch.load(ref).cuda()
    D = (Q - R).abs()[:, :N]          # D[col, row] (column-major storage: Q[j, i] = A(i, j))
    print("max abs diff", float(D.max()), "entries > 1e-12:", int((D > 1e-12).sum()))
    nb = (N + 255) // 256
    bad = torch.zeros(nb, nb, dtype=torch.int64)
    for cj in range(nb):
        for ri in range(nb):
            bad[ri, cj] = int((D[cj*256:(cj+1)*256, ri*256:(ri+1)*256] > 1e-12).sum())
    print("rows = row blocks, cols = column blocks (count of differing entries, 256 x 256 blocks)")
    for ri in range(nb):
        print(" ".join("%5d" % int(bad[ri, cj]) for cj in range(nb)))
if ref:
    bad_rc = (D > 1e-12).nonzero()   # (col, row)
    cols = bad_rc[:, 0]; rows = bad_rc[:, 1]
    print("first bad row", int(rows.min()), "cols with any bad entry:", int(cols.unique().numel()))
    r0 = int(rows.min()) // 128 * 128
    for rb in range(r0, min(N, r0 + 1024), 128):
        sel = (rows >= rb) & (rows < rb + 128)
        cu = cols[sel].unique()
        if cu.numel():
            print("rows [%d,%d): %d bad entries in %d columns: %s" % (rb, rb + 128, int(sel.sum()), int(cu.numel()), cu[:80].tolist()))
