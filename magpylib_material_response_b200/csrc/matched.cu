// pairs_matching assembly (replaces match_pairs + the gather of demag.py:280-321, :198).
//
// The reference builds an n^2 x 14 key matrix on the host and runs np.unique(axis=0) over it
// (81.6 GB at n = 27000).  Here the unique pairs are found with a lock-free open-addressing table
// on the device, keyed like the reference:  rint((p_j - p_i + 1e-9) * 1e8) per axis (= its
// "(x + 1e-9).round(8)") plus the source cell's class id (rounded orientation + dimension, computed
// on the host over n rows).  A slot IS the key: 128 bits (three 32-bit
// displacements in units of 1e-8 m = +-21 m, 31-bit class id, valid bit) claimed with one 128-bit
// compare-and-swap (ATOMG.CAS.128), so a probe is a single 16-byte load and compare.  A parallel
// array keeps, via atomicMin, the smallest flat index i*n+j of the class -- the representative
// np.unique(return_index=True) picks.  Then
//   1. insert all n^2 pairs, remembering each pair's slot (4 B per pair, coalesced)
//   2. evaluate one 3x3 block per class (the only transcendental work)
//   3. scatter blocks to all pairs      (72 B written + 4 B read per pair: HBM-write bound)
// (The first version stored a member pair per slot and re-derived its key at every probe -- a 64-bit
// division and six scattered loads -- and looked every pair up twice: 85 ms at n = 27000, slower
// than evaluating all pairs densely.)
#include "cuboid.cuh"

namespace {

constexpr int SRC_TILE = 128;
constexpr int TGT_TILE = 32;
typedef unsigned long long u64;

struct MatchParams {
  int64_t n;
  const double* pos;
  const double* quat;
  const double* dim;
  const int32_t* cls;
  const double* chi;
  double* out;
  int64_t ld;
  u64* tab;        // 2 x u64 per slot: packed key, (0, 0) = empty
  u64* minrep;     // smallest pair index of the class
  double* blocks;  // 8 doubles (64 B) per slot: xx, xy, xz, yy, yz, zz of the symmetric block + padding
  unsigned* slot_of_pair;  // [j*n + i] slot of pair (source i, observer j), or NULL (look up again)
  u64 mask;        // capacity - 1 (capacity is a power of two)
  int* status;     // [0] = number of classes, [1] = overflow flag, [2] = key out of range
  int rotated;
  // target set of this rank (block-cyclic over cells, as mmr_assemble): local target t is global cell
  // ((t / tgt_block) * tgt_nparts + tgt_part) * tgt_block + t % tgt_block; rows of `out` are local targets
  int64_t n_tgt_local, tgt_block, tgt_nparts, tgt_part;
};

__device__ __forceinline__ int64_t global_target(const MatchParams& p, int64_t t) {
  return ((t / p.tgt_block) * p.tgt_nparts + p.tgt_part) * p.tgt_block + t % p.tgt_block;
}

struct PairKey {
  u64 lo, hi;  // lo = kx | ky << 32, hi = kz | cls << 32 | valid bit 63
  bool ok;     // every component fits its field
};

__device__ __forceinline__ PairKey make_key(const MatchParams& p, int64_t i, int64_t j) {
  // (p_j - p_i + 1e-9).round(8): numpy rounds via rint(x * 1e8) / 1e8
  const long long kx = __double2ll_rn(__dmul_rn(__dadd_rn(__dsub_rn(p.pos[3 * j + 0], p.pos[3 * i + 0]), 1e-9), 1e8));
  const long long ky = __double2ll_rn(__dmul_rn(__dadd_rn(__dsub_rn(p.pos[3 * j + 1], p.pos[3 * i + 1]), 1e-9), 1e8));
  const long long kz = __double2ll_rn(__dmul_rn(__dadd_rn(__dsub_rn(p.pos[3 * j + 2], p.pos[3 * i + 2]), 1e-9), 1e8));
  const int cls = p.cls[i];
  PairKey k;
  k.ok = (kx == (long long)(int)kx) && (ky == (long long)(int)ky) && (kz == (long long)(int)kz) && cls >= 0;
  k.lo = (u64)(unsigned)(int)kx | ((u64)(unsigned)(int)ky << 32);
  k.hi = (u64)(unsigned)(int)kz | ((u64)(unsigned)cls << 32) | (1ull << 63);
  return k;
}

__device__ __forceinline__ u64 hash_key(const PairKey& k) {
  u64 h = k.lo * 0x9E3779B97F4A7C15ull;
  h ^= h >> 32;
  h += k.hi * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

__device__ __forceinline__ void cas128(u64* addr, u64 new_lo, u64 new_hi, u64& old_lo, u64& old_hi) {
  // compare with (0, 0) = empty slot
  asm volatile(
      "{\n\t.reg .b128 e, n, o;\n\tmov.b128 e, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
      "atom.global.relaxed.gpu.cas.b128 o, [%6], e, n;\n\tmov.b128 {%0, %1}, o;\n\t}"
      : "=l"(old_lo), "=l"(old_hi)
      : "l"(0ull), "l"(0ull), "l"(new_lo), "l"(new_hi), "l"(addr)
      : "memory");
}

// slot of the class of pair (i,j); inserts if INSERT.  Returns ~0 on overflow / missing.
template <bool INSERT>
__device__ __forceinline__ u64 find_slot(const MatchParams& p, const PairKey& key, u64 me) {
  u64 h = hash_key(key) & p.mask;
  for (u64 probes = 0; probes <= p.mask; ++probes) {
    const ulonglong2 cur = __ldcg(reinterpret_cast<const ulonglong2*>(p.tab + 2 * h));  // L2: where the CAS lands
    u64 lo = cur.x, hi = cur.y;
    if (hi == 0) {  // empty (a valid key has bit 63 of hi set)
      if (!INSERT) return ~0ull;
      cas128(p.tab + 2 * h, key.lo, key.hi, lo, hi);
      if (hi == 0) {
        atomicAdd(&p.status[0], 1);
        atomicMin(&p.minrep[h], me);
        return h;
      }
    }
    if (hi == key.hi && lo == 0) {
      // The 16-byte vector load is not guaranteed single-copy atomic against the 128-bit CAS that publishes
      // a slot: a reader can see the new `hi` next to a stale (zero) `lo`.  Such a slot may hold this key, a
      // genuine key whose low word IS zero (pairs displaced along z only), or another key with the same
      // high word.  Read it once more ATOMICALLY before comparing (a 128-bit CAS of 0 -> 0 returns the slot
      // and cannot modify a published one).
      cas128(p.tab + 2 * h, 0ull, 0ull, lo, hi);
    }
    if (lo == key.lo && hi == key.hi) {
      if (INSERT && me < p.minrep[h]) atomicMin(&p.minrep[h], me);
      return h;
    }
    h = (h + 1) & p.mask;
  }
  return ~0ull;
}

// thread = source i (fast index, so the slot record [j*n + i] is written coalesced), loop over observers
__global__ void __launch_bounds__(256) match_insert_kernel(const MatchParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  for (int64_t t = blockIdx.y; t < p.n_tgt_local; t += gridDim.y) {
    if (p.status[1] | p.status[2]) return;
    const int64_t j = global_target(p, t);
    const PairKey key = make_key(p, i, j);
    if (!key.ok) {
      atomicExch(&p.status[2], 1);
      return;
    }
    const u64 h = find_slot<true>(p, key, (u64)(i * p.n + j));
    if (h == ~0ull) atomicExch(&p.status[1], 1);
    else if (p.slot_of_pair) p.slot_of_pair[t * p.n + i] = (unsigned)h;
    // load factor guard: stop early once the table is more than half full
    if (threadIdx.x == 0 && (u64)p.status[0] * 2 > p.mask + 1) atomicExch(&p.status[1], 1);
  }
}

__global__ void __launch_bounds__(128) match_eval_kernel(const MatchParams p) {
  const u64 h = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (h > p.mask) return;
  if (p.tab[2 * h + 1] == 0) return;
  const u64 rep = p.minrep[h];
  const int64_t i = (int64_t)(rep / (u64)p.n), j = (int64_t)(rep % (u64)p.n);
  const double dx = p.pos[3 * j + 0] - p.pos[3 * i + 0];
  const double dy = p.pos[3 * j + 1] - p.pos[3 * i + 1];
  const double dz = p.pos[3 * j + 2] - p.pos[3 * i + 2];
  const double ha = 0.5 * p.dim[3 * i + 0], hb = 0.5 * p.dim[3 * i + 1], hc = 0.5 * p.dim[3 * i + 2];
  double blk[9];
  SymBlock N;
  if (p.rotated) {
    Rot3 R;
    quat_to_rot(p.quat + 4 * i, R);
    const double lx = R.m[0] * dx + R.m[3] * dy + R.m[6] * dz;
    const double ly = R.m[1] * dx + R.m[4] * dy + R.m[7] * dz;
    const double lz = R.m[2] * dx + R.m[5] * dy + R.m[8] * dz;
    cuboid_block_auto<true>(lx, ly, lz, ha, hb, hc, N);
    rotate_block(R, N, blk);
  } else {
    cuboid_block_auto<true>(dx, dy, dz, ha, hb, hc, N);
    blk[0] = N.xx; blk[1] = N.xy; blk[2] = N.xz;
    blk[3] = N.xy; blk[4] = N.yy; blk[5] = N.yz;
    blk[6] = N.xz; blk[7] = N.yz; blk[8] = N.zz;
  }
  // R N R^T is symmetric like N: 6 values, one 64-byte line per class (two sectors per gather
  // instead of nine scattered 8-byte loads)
  double2* dst = reinterpret_cast<double2*>(p.blocks + h * 8);
  dst[0] = make_double2(blk[0], blk[1]);
  dst[1] = make_double2(blk[2], blk[4]);
  dst[2] = make_double2(blk[5], blk[8]);
}

template <bool FUSE_Q>
__global__ void __launch_bounds__(SRC_TILE) match_scatter_kernel(const MatchParams p) {
  __shared__ double s_tchi[TGT_TILE][3];
  __shared__ double s_stage[SRC_TILE / 32][3][96];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t t0 = (int64_t)blockIdx.y * TGT_TILE;
  const int ntgt = (int)min((int64_t)TGT_TILE, p.n_tgt_local - t0);
  if (FUSE_Q && tid < ntgt) {
    const int64_t jg = global_target(p, t0 + tid);
    s_tchi[tid][0] = p.chi[3 * jg + 0];
    s_tchi[tid][1] = p.chi[3 * jg + 1];
    s_tchi[tid][2] = p.chi[3 * jg + 2];
  }
  const int64_t i_warp0 = (int64_t)blockIdx.x * SRC_TILE + warp * 32;
  const int64_t i = i_warp0 + lane;
  const int64_t ic = i < p.n ? i : p.n - 1;
  __syncthreads();
  const int64_t ncols_warp = min((int64_t)96, 3 * (p.n - i_warp0));
  for (int tt = 0; tt < ntgt; ++tt) {
    const int64_t j = global_target(p, t0 + tt);
    const u64 h = p.slot_of_pair ? (u64)p.slot_of_pair[(t0 + tt) * p.n + ic]
                                 : find_slot<false>(p, make_key(p, ic, j), 0ull);
    double blk[9];
    {
      double2 a = make_double2(0.0, 0.0), b = a, c = a;
      if (h != ~0ull) {
        const double2* src = reinterpret_cast<const double2*>(p.blocks + h * 8);
        a = src[0];
        b = src[1];
        c = src[2];
      }
      blk[0] = a.x; blk[1] = a.y; blk[2] = b.x;
      blk[3] = a.y; blk[4] = b.y; blk[5] = c.x;
      blk[6] = b.x; blk[7] = c.x; blk[8] = c.y;
    }
    if (FUSE_Q) {
      const bool diag = (j == i);
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double chi = s_tchi[tt][l];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double v = -(chi * blk[3 * l + k]);
          blk[3 * l + k] = (diag && l == k) ? (1.0 + v) : v;
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int k = 0; k < 3; ++k) s_stage[warp][l][3 * lane + k] = blk[3 * l + k];
    __syncwarp();
    double* row_base = p.out + (3 * (t0 + tt)) * p.ld + 3 * i_warp0;
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int c = q * 32 + lane;
        if (c < ncols_warp) row_base[l * p.ld + c] = s_stage[warp][l][c];
      }
  }
}

}  // namespace

int mmr_detect_rotated(mmr_ctx* ctx, int64_t n, const double* d_quat, int* rotated_out);

extern "C" int mmr_assemble_matched(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
                                    const double* d_dim, const int32_t* d_cls, const double* d_chi,
                                    double* d_out, int64_t ld, int64_t table_cap_slots,
                                    int64_t* n_unique_out) {
  return mmr_assemble_matched_rows(ctx, n, d_pos, d_quat, d_dim, d_cls, d_chi, d_out, ld, table_cap_slots, n, n > 0 ? n : 1, 1,
                                   0, n_unique_out);
}

extern "C" int mmr_assemble_matched_rows(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
                                         const double* d_dim, const int32_t* d_cls, const double* d_chi,
                                         double* d_out, int64_t ld, int64_t table_cap_slots, int64_t n_tgt_local,
                                         int64_t tgt_block, int64_t tgt_nparts, int64_t tgt_part,
                                         int64_t* n_unique_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(n >= 0, "negative n");
  if (n_unique_out) *n_unique_out = 0;
  if (n == 0 || n_tgt_local == 0) return 0;
  MMR_ARG(d_pos && d_quat && d_dim && d_cls && d_out, "NULL pointer");
  MMR_ARG(ld >= 3 * n, "ld must be >= 3*n");
  MMR_ARG(table_cap_slots >= 2, "table_cap_slots too small");
  MMR_ARG(n_tgt_local > 0 && tgt_block >= 1 && tgt_nparts >= 1 && tgt_part >= 0 && tgt_part < tgt_nparts, "bad target set");
  {
    const int64_t last = n_tgt_local - 1;
    const int64_t jlast = ((last / tgt_block) * tgt_nparts + tgt_part) * tgt_block + last % tgt_block;
    MMR_ARG(jlast < n, "target set exceeds n");
  }
  DeviceGuard g(ctx->device);
  int rc = mmr_ws_reserve(ctx, 256);
  if (rc) return rc;
  rc = mmr_hpin_reserve(ctx, 256);
  if (rc) return rc;
  int rotated = 0;
  rc = mmr_detect_rotated(ctx, n, d_quat, &rotated);
  if (rc) return rc;

  // capacity: power of two, grown on overflow up to table_cap_slots (88 B per slot)
  const u64 npairs = (u64)n * (u64)n_tgt_local;
  u64 cap = 1ull << 16;
  while (cap < (1ull << 21) && cap < 4 * npairs) cap <<= 1;
  // one slot id per pair, so the scatter pass needs no second lookup (skipped beyond 12 GB)
  const size_t slot_bytes = (sizeof(unsigned) * npairs <= (size_t)12e9) ? ((sizeof(unsigned) * npairs + 255) / 256) * 256 : 0;
  for (;;) {
    while (cap > (u64)table_cap_slots) cap >>= 1;
    const size_t tab_bytes = 256 + cap * (sizeof(u64) * 3 + sizeof(double) * 8);
    rc = mmr_ws2_reserve(ctx, tab_bytes + slot_bytes);
    if (rc) return rc;
    char* base = (char*)ctx->ws2;
    MatchParams p;
    p.n = n; p.pos = d_pos; p.quat = d_quat; p.dim = d_dim; p.cls = d_cls; p.chi = d_chi;
    p.out = d_out; p.ld = ld;
    p.status = (int*)base;
    p.tab = (u64*)(base + 256);
    p.minrep = p.tab + 2 * cap;
    p.blocks = (double*)(p.minrep + cap);
    p.slot_of_pair = slot_bytes ? (unsigned*)(base + tab_bytes) : nullptr;
    p.mask = cap - 1;
    p.rotated = rotated;
    p.n_tgt_local = n_tgt_local; p.tgt_block = tgt_block; p.tgt_nparts = tgt_nparts; p.tgt_part = tgt_part;
    MMR_CUDA(cudaMemsetAsync(p.status, 0, 256, ctx->stream));
    MMR_CUDA(cudaMemsetAsync(p.tab, 0, sizeof(u64) * 2 * cap, ctx->stream));
    MMR_CUDA(cudaMemsetAsync(p.minrep, 0xff, sizeof(u64) * cap, ctx->stream));
    dim3 gi((unsigned)ceil_div64(n, 256), (unsigned)(n_tgt_local < 2048 ? n_tgt_local : 2048));
    match_insert_kernel<<<gi, 256, 0, ctx->stream>>>(p);
    MMR_CHECK_LAUNCH(ctx);
    MMR_CUDA(cudaMemcpyAsync(ctx->hpin, p.status, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MMR_CUDA(cudaStreamSynchronize(ctx->stream));
    const int nclasses = ((int*)ctx->hpin)[0];
    const int overflow = ((int*)ctx->hpin)[1];
    if (((int*)ctx->hpin)[2]) {
      mmr_set_error("pairs_matching: a pair displacement exceeds the +-21 m key range (1e-8 m units in 32 bits)");
      return -3;
    }
    if (overflow) {
      if (cap * 2 > (u64)table_cap_slots || cap >= 2 * npairs) {
        mmr_set_error("pairs_matching: more than %llu unique pairs (table cap %lld slots)",
                      (unsigned long long)(cap / 2), (long long)table_cap_slots);
        return -3;
      }
      cap <<= 2;
      continue;
    }
    match_eval_kernel<<<(unsigned)ceil_div64((int64_t)cap, 128), 128, 0, ctx->stream>>>(p);
    MMR_CHECK_LAUNCH(ctx);
    dim3 gs((unsigned)ceil_div64(n, SRC_TILE), (unsigned)ceil_div64(n_tgt_local, TGT_TILE));
    const int pidx = mmr_prof_begin(ctx, MMR_PROF_ASSEMBLE, ctx->stream);
    if (d_chi) match_scatter_kernel<true><<<gs, SRC_TILE, 0, ctx->stream>>>(p);
    else match_scatter_kernel<false><<<gs, SRC_TILE, 0, ctx->stream>>>(p);
    mmr_prof_end(ctx, pidx, (double)n * (double)n_tgt_local, ctx->stream);
    MMR_CHECK_LAUNCH(ctx);
    if (n_unique_out) *n_unique_out = nclasses;
    // the per-pair slot record (4 B x n x n_tgt_local, up to 12 GB) is dead now: do not keep it for the
    // lifetime of the process-wide context
    if (slot_bytes > ((size_t)1 << 30)) {
      MMR_CUDA(cudaStreamSynchronize(ctx->stream));
      MMR_CUDA(cudaFree(ctx->ws2));
      ctx->ws2 = nullptr;
      ctx->ws2_bytes = 0;
    }
    return 0;
  }
}
