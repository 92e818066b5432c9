"""CPU checks of the shared-memory layout arithmetic of the TMA-fed trailing-update kernel
(magpylib_material_response_b200/csrc/dgemm_tma.cuh, dgemm_sub_tma2_kernel): the formulas below restate the
kernel's address computation and the TMA 128-byte swizzle; the tests assert that

  * the four DMMA groups of a k-tile use every k of the tile exactly once (any assignment of the summation index
    to DMMA slots is valid as long as A and B use the same one),
  * a thread's fragment addresses address exactly the element the DMMA fragment layout asks for, in the swizzled
    tile the TMA writes,
  * every half-warp's 16 eight-byte loads fall into 16 different bank pairs (conflict-free LDS.64).

No GPU involved: this pins the design so that a change of the constants cannot silently introduce
wrong operands or bank conflicts."""
import itertools

BM, BN, BK = 128, 64, 16
A_BYTES = BM * BK * 8  # 8 boxes of 16 (m) x 16 (k) doubles, 2048 B each
XS = (0, 1, 4, 5)      # k of lane slot kq in DMMA group gi: kp(kq) ^ XS[gi]


def kp(kq):
    return (kq & 1) * 3 + (kq >> 1) * 12  # {0, 3, 12, 15}


def swizzle128(byte_addr):
    """TMA SWIZZLE_128B on a 1024-byte aligned tile: the 16-byte chunk index (bits 4..6) is XORed with the
    128-byte row index mod 8 (bits 7..9)."""
    return byte_addr ^ (((byte_addr >> 7) & 7) << 4)


def a_tile_addr(m, k):
    """A k-tile as the TMA writes it: box j = m // 16 at j * 2048, inside a box row k (128 B) holds m % 16."""
    return swizzle128((m // 16) * 2048 + k * 128 + (m % 16) * 8)


def b_tile_addr(n, k):
    """B k-tile: one box of 16 (k, inner) x 64 (n): row n (128 B) holds k."""
    return A_BYTES + swizzle128(n * 128 + k * 8)


def thread_addrs(warp, lane):
    """The kernel's per-thread bases and the 4 x (4 + 4) fragment addresses of one k-tile (byte offsets in a stage)."""
    wm, wn = warp & 3, warp >> 2
    rr, kq = lane >> 2, lane & 3
    k0 = kp(kq)
    ta0 = (wm * 2) * 2048 + k0 * 128 + ((((rr >> 1) ^ k0) & 7) << 4) + (rr & 1) * 8
    tb0 = A_BYTES + (wn * 32 + rr) * 128 + ((((k0 >> 1) ^ rr) & 7) << 4) + (k0 & 1) * 8
    out = []
    for gi in range(4):
        xa, xb = XS[gi] * 144, XS[gi] * 8
        a = [(ta0 ^ xa ^ ((f & 1) * 64)) + (f >> 1) * 2048 for f in range(4)]
        b = [(tb0 ^ xb) + g * 1024 for g in range(4)]
        out.append((a, b))
    return out


def test_every_k_of_a_tile_is_used_exactly_once():
    ks = sorted(kp(kq) ^ x for x in XS for kq in range(4))
    assert ks == list(range(BK))


def test_fragment_addresses_hit_the_elements_the_dmma_layout_asks_for():
    # m8n8k4 fragments: A lane holds A[row = lane >> 2][k slot = lane & 3], B lane holds B[k slot = lane & 3][col = lane >> 2]
    for warp, lane in itertools.product(range(8), range(32)):
        wm, wn = warp & 3, warp >> 2
        rr, kq = lane >> 2, lane & 3
        for gi, (a, b) in enumerate(thread_addrs(warp, lane)):
            k = kp(kq) ^ XS[gi]
            for f in range(4):
                assert a[f] == a_tile_addr(wm * 32 + 8 * f + rr, k), (warp, lane, gi, f)
            for g in range(4):
                assert b[g] == b_tile_addr(wn * 32 + 8 * g + rr, k), (warp, lane, gi, g)


def test_fragment_loads_are_bank_conflict_free():
    # an LDS.64 of a warp is served as two half-warps; 16 lanes x 8 B must cover 16 distinct 8-byte bank pairs
    for warp in range(8):
        per_lane = [thread_addrs(warp, lane) for lane in range(32)]
        for gi in range(4):
            for half in (range(0, 16), range(16, 32)):
                for f in range(4):
                    banks = {(per_lane[lane][gi][0][f] % 128) // 8 for lane in half}
                    assert len(banks) == 16, ("A", warp, gi, f)
                for g in range(4):
                    banks = {(per_lane[lane][gi][1][g] % 128) // 8 for lane in half}
                    assert len(banks) == 16, ("B", warp, gi, g)


def test_c_staging_buffer_is_conflict_free_and_fits():
    # C tile staged column by column, LDC_S = BM + 2 doubles per column; accumulator-layout access of a half-warp:
    # rows (lane >> 2) = 0..3, columns 2 * (lane & 3) + e
    ldc_s = BM + 2
    for e in (0, 1):
        banks = {(((2 * kq + e) * ldc_s + r) * 8 % 128) // 8 for kq in range(4) for r in range(4)}
        assert len(banks) == 16
    assert BN * ldc_s * 8 <= 4 * (A_BYTES + BN * BK * 8)       # fits into the 4 stages of the tensor-map kernel
    assert BN * ldc_s * 8 <= 3 * (BK * (BM + 4) + BN * (BK + 4)) * 8  # and into the 3 padded cp.async stages
