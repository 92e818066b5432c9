// Persistent form of the LU trailing update  C (m x n) -= A (m x k) * B (k x n)  (column-major, FP64 DMMA).
//
// dgemm_nn_fast_kernel (dgemm.cuh) keeps the DMMA pipe 89 % busy at k = 256: every 128 x 64 tile is its own CTA whose
// 32 strided C loads and 32 strided C stores per thread go through the LSU in 8-row x 4-column pieces.  Probes
// (tools/gemm_lab.cu, profiles/gemm_lab_r2.md) put the cost of that C traffic at 5 % of the kernel at k = 256 -- the
// other CTA of the SM slows down while it lasts -- and of the cp.async pipeline restart at ~1 %.  This kernel:
//
//   * one CTA per SM, 512 threads = two independent 8-warp groups, each with the fast kernel's 128 x 64 tile,
//     3-stage cp.async pipeline and warp layout, synchronised by its own named barrier (bar.sync 1+g, 256);
//   * a group walks its tiles (gid, gid + G, ...) with ONE flattened (tile, k-tile) pipeline: the A/B stages of the
//     next tile are already in flight while the last k-tiles of the current one are multiplied;
//   * C never touches the LSU's global path: the C tile of a group's NEXT tile is staged in shared memory by the
//     bulk-copy engine (TMA, cp.async.bulk: one 1 KB copy per column, completion counted on an mbarrier) while the
//     group is still multiplying; at the tile change every thread SWAPS its 32 accumulators against the staged
//     values (LDS.64 / STS.64, conflict-free), and the finished tile leaves by bulk stores from the same buffer.
//     One 65 KB buffer serves both groups alternately.
//
// Measured on B200, m = n = 12288 (tools/gemm_lab): 33.5 TF at k = 256 (per-tile kernel 32.9), 34.2 at k = 512 (34.0).
// Tried and rejected here: A/B k-tiles by 1-D bulk copies issued by lane 0 of every warp (80 copies of 128 B .. 1 KB
// per k-tile and group: 31.2 TF), handing the buffer over one k-tile later instead of waiting for the stores' reads
// inside the tile change (33.2 TF: the extra state in the main loop costs more than the wait).
//
// Shared memory: 2 groups x 3 stages x 27 136 B + 66 560 B (C tile, columns padded to 130 doubles so that the
// accumulator-layout accesses are bank-conflict free) + 2 mbarriers = 229 392 B of the 232 448 B a CTA may use.
#pragma once

#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include "dgemm.cuh"

namespace mmr_gemm {

constexpr int P_THREADS = 512;
constexpr int LDC_S = BM + 2;       // doubles per staged C column: 2 columns apart = 2080 B = 32 B mod 128
constexpr int C_STAGE = BN * LDC_S;
constexpr int GROUP_STAGES = STAGES * (A_STAGE + B_STAGE);
constexpr size_t P_SMEM_BYTES = (size_t)(2 * GROUP_STAGES + C_STAGE) * sizeof(double) + 16;

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 256;\n" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// shared -> global bulk copy (SASS UBLKCP.G.S), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// Returns 0 through a shared-memory load (a leftover of the hunt for the stage-release bug, profiles/gemm_lab_r2.md §3:
// one hypothesis was that a bulk-copy issue right behind a CTA barrier is not held back by BAR.SYNC.DEFER_BLOCKING the
// way a dependent memory instruction is; it was not the cause).  Kept because the kernels were validated with it in
// place -- their main loops are sensitive to any change of the surrounding code -- and it costs nothing.
__device__ __forceinline__ unsigned after_barrier_zero(const void* smem_word) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];\n\tand.b32 %0, %0, 0;\n" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(smem_word)) : "memory");
  return v;
}
// global -> shared bulk copy (SASS UBLKCP), bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// requires: A, B, C 16-byte aligned; lda, ldb, ldc, m even; k % BK == 0, k > 0.
// grid = (#SMs), block = 512, dynamic shared memory = P_SMEM_BYTES.  tiles_m = ceil(m / BM), ntiles = tiles_m * ceil(n / BN).
//
// The staging buffer is handed back and forth in "uses": use j of group g first receives the C tile of the group's
// tile j (bulk copies, has_in), is then swapped against the accumulators of tile j - 1 (has_out; STORE_BULK) and
// written back by bulk stores, and is finally handed to the other group.  Phase j of bars[g] (8 arrivals: lane 0 of
// every warp of the group that hands over, each with the bytes of its own 8 columns) says "use j of group g may
// start".  Warp w owns columns 8w .. 8w+7 of the buffer on the async side: its lane 0 stores them, waits until
// the stores have read them and refills them, so the hand-over needs no second barrier.
// PROBE (timing experiments only, wrong results): 1 = no output, 2 = no C input, 4 = no barrier per k-tile,
// 8 = no cp.async in the main loop.
template <bool STORE_BULK, int PROBE>
__global__ void __launch_bounds__(P_THREADS, 1)
    dgemm_sub_persist_kernel(int m, int n, int k, const double* __restrict__ A, int64_t lda, const double* __restrict__ B,
                             int64_t ldb, double* __restrict__ C, int64_t ldc, int tiles_m, int ntiles,
                             const int* __restrict__ fail, int dep) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, g = tid >> 8, t = tid & 255, lane = t & 31, warp = t >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  double* As = smem + g * GROUP_STAGES;
  double* Bs = As + STAGES * A_STAGE;
  double* Cs = smem + 2 * GROUP_STAGES;
  uint64_t* bars = (uint64_t*)(Cs + C_STAGE);
  const bool poisoned = lu_poisoned(lu_failword(fail), dep);
  const int G = gridDim.x * 2, gid = blockIdx.x * 2 + g;
  const int KT = k / BK;

  if (tid == 0) {
    mbar_init(&bars[0], 8);
    mbar_init(&bars[1], 8);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (gid >= ntiles) return;  // (group 1 of the last CTAs; the other group then works alone)

  const int ntl = (ntiles - gid + G - 1) / G;                                 // my tiles
  const int pgid = gid ^ 1;                                                   // the other group of this CTA
  const int pntl = pgid < ntiles ? (ntiles - pgid + G - 1) / G : 0;           // its tiles
  constexpr int XU = STORE_BULK ? 1 : 0;                                      // one more (store-only) use at the end

  // lane 0 of every warp: open use `uj` of the group with id `ugid` / barrier `ug` (refill my 8 columns or just arrive)
  auto open_use = [&](int ug, int ugid, int untl, int uj) {
    if (uj < untl && !(PROBE & 2)) {
      const int tile = ugid + uj * G;
      const int tm = tile % tiles_m, tn = tile / tiles_m;
      const int rows = min(BM, m - tm * BM), cols = min(BN, n - tn * BN);
      const int c0 = warp * (BN / 8);
      const int mine = max(0, min(BN / 8, cols - c0));
      mbar_arrive_expect_tx(&bars[ug], (unsigned)(rows * mine) * 8u);
      const double* src = C + (int64_t)(tn * BN + c0) * ldc + tm * BM;
#pragma unroll
      for (int c = 0; c < BN / 8; ++c)
        if (c < mine) bulk_g2s(Cs + (c0 + c) * LDC_S, src + (int64_t)c * ldc, (unsigned)rows * 8u, &bars[ug]);
    } else {
      mbar_arrive(&bars[ug]);
    }
  };
  // after my use j: whose use comes next?
  auto hand_over = [&](int j) {
    if (pntl > 0) {
      const int pj = j + g;
      if (pj < pntl + XU) open_use(1 - g, pgid, pntl, pj);
    } else if (j + 1 < ntl + XU) {
      open_use(g, gid, ntl, j + 1);
    }
  };

  // producer (cp.async) state: tile ptile, k-tile pkt, `pleft` k-tiles still to be requested
  const int a_mc = (t & 63) * 2, a_kr = t >> 6;
  const int b_kc = (t & 7) * 2, b_nr = t >> 3;
  const int a_dst = a_kr * LDA_S + a_mc, b_dst = b_nr * LDB_S + b_kc;
  const int64_t a_step4 = 4 * lda;
  const double *a_src, *b_src0, *b_src1;
  int a_bytes, b_bytes0, b_bytes1;
  auto set_prod = [&](int tile) {
    const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
    const int a_gm = m0 + a_mc;
    a_bytes = (a_gm < m) ? 16 : 0;
    a_src = A + (int64_t)a_kr * lda + (a_bytes ? a_gm : 0);
    b_bytes0 = (n0 + b_nr < n) ? 16 : 0;
    b_bytes1 = (n0 + b_nr + 32 < n) ? 16 : 0;
    b_src0 = B + (b_bytes0 ? (int64_t)(n0 + b_nr) * ldb : 0) + b_kc;
    b_src1 = B + (b_bytes1 ? (int64_t)(n0 + b_nr + 32) * ldb : 0) + b_kc;
  };
  auto copy_a = [&](int stage, int c0) {  // chunks c0, c0+1 of the producer's current k-tile
    double* as = As + stage * A_STAGE + a_dst;
    cp_async16(as + (c0 * 4) * LDA_S, a_src + c0 * a_step4, a_bytes);
    cp_async16(as + (c0 * 4 + 4) * LDA_S, a_src + (c0 + 1) * a_step4, a_bytes);
  };
  auto copy_b = [&](int stage) {
    double* bs = Bs + stage * B_STAGE + b_dst;
    cp_async16(bs, b_src0, b_bytes0);
    cp_async16(bs + 32 * LDB_S, b_src1, b_bytes1);
  };
  int ptile = gid, pkt = 0;
  auto advance = [&]() {  // the producer moves on by one k-tile
    if (++pkt == KT) {
      pkt = 0;
      ptile += G;
      if (ptile < ntiles) set_prod(ptile);
    } else {
      a_src += (int64_t)BK * lda;
      b_src0 += BK;
      b_src1 += BK;
    }
  };

  const int q_total = ntl * KT;
  int pleft = q_total;
  set_prod(gid);
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (pleft > 0) {
      copy_a(s, 0);
      copy_a(s, 2);
      copy_b(s);
      --pleft;
      advance();
    }
    cp_async_commit();
  }

  double acc[4][4][2];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) acc[f][gq][0] = acc[f][gq][1] = 0.0;

  // Tile change = use j of my group: accumulators of tile j-1 out (tile `out_tile`), -C of tile j in.
  auto change_tile = [&](int j, int out_tile) {
    const bool has_in = j < ntl && !(PROBE & 2);
    const bool has_out = j > 0 && !poisoned && (!(PROBE & 1) || m < 0);  // (m < 0 never holds: keeps the products alive)
    if (!STORE_BULK && has_out) {  // direct stores from the accumulator layout
      const int m0 = (out_tile % tiles_m) * BM, n0 = (out_tile / tiles_m) * BN;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int gm = m0 + wm * 32 + 8 * f + (lane >> 2);
        if (gm >= m) continue;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int gn = n0 + wn * 32 + 8 * gq + 2 * (lane & 3) + e;
            if (gn < n) C[(int64_t)gn * ldc + gm] = flip_sign(acc[f][gq][e]);
          }
      }
    }
    if (!STORE_BULK && !has_in && !(j < ntl)) return;  // (no store-only use in this mode)
    mbar_wait(&bars[g], j & 1);
    double* cb = Cs + (wn * 32 + 2 * (lane & 3)) * LDC_S + wm * 32 + (lane >> 2);
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        double* c0 = cb + (8 * gq) * LDC_S + 8 * f;
        double t0 = 0.0, t1 = 0.0;
        if (has_in) {
          t0 = c0[0];
          t1 = c0[LDC_S];
        }
        if (STORE_BULK && has_out) {
          c0[0] = flip_sign(acc[f][gq][0]);
          c0[LDC_S] = flip_sign(acc[f][gq][1]);
        }
        if (has_in) {
          acc[f][gq][0] = flip_sign(t0);
          acc[f][gq][1] = flip_sign(t1);
        }
      }
    if (STORE_BULK && has_out) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    group_bar(g);  // the buffer is read (and rewritten) by all 8 warps
    if (lane == 0) {
      if (STORE_BULK && has_out) {
        const int tm = out_tile % tiles_m, tn = out_tile / tiles_m;
        const int rows = min(BM, m - tm * BM), cols = min(BN, n - tn * BN);
        const int c0 = warp * (BN / 8);
        double* dst = C + (int64_t)(tn * BN + c0) * ldc + tm * BM;
        const unsigned zs = after_barrier_zero(Cs);
#pragma unroll
        for (int c = 0; c < BN / 8; ++c)
          if (c0 + c < cols) bulk_s2g(dst + (int64_t)c * ldc, Cs + (c0 + c) * LDC_S, (unsigned)rows * 8u + zs);
        bulk_commit();
        bulk_wait_read0();
      } else {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      }
      hand_over(j);
    }
    __syncwarp();
  };

  int ctile = gid, ckt = 0, j = 0;
  if (lane == 0 && (g == 0 || pntl == 0)) open_use(g, gid, ntl, 0);  // buffer use 0 (group 1 receives its first use from group 0)
  __syncwarp();
  change_tile(0, 0);

  const int a_row = wm * 32 + (lane >> 2);
  const int b_row = wn * 32 + (lane >> 2);
  const int kq = lane & 3;
  int cs = 0, ls = STAGES - 1;

  for (int q = 0; q < q_total; ++q) {
    cp_async_wait<STAGES - 2>();
    if (!(PROBE & 4)) group_bar(g);  // stage cs has landed for everybody, everybody is done with stage ls
    const bool do_load = pleft > 0 && !(PROBE & 8);
    const double* as = As + cs * A_STAGE;
    const double* bs = Bs + cs * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) af[f] = as[(kk + kq) * LDA_S + a_row + 8 * f];
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) bf[gq] = bs[(b_row + 8 * gq) * LDB_S + kk + kq];
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) dmma884(acc[f][gq][0], acc[f][gq][1], af[f], bf[gq]);
      if (do_load) {
        if (kk == 0) copy_a(ls, 0);
        if (kk == 4) copy_a(ls, 2);
        if (kk == 8) copy_b(ls);
      }
    }
    cp_async_commit();
    if (do_load) {
      --pleft;
      advance();
    }
    cs = (cs + 1 == STAGES) ? 0 : cs + 1;
    ls = (ls + 1 == STAGES) ? 0 : ls + 1;

    if (++ckt == KT) {  // tile finished
      ckt = 0;
      ++j;
      change_tile(j, ctile);
      ctile += G;
    }
  }
  cp_async_wait<0>();
  if (STORE_BULK && lane == 0) bulk_wait0();
}

// ---------------------------------------------------------------------------------------
// Per-tile form with the same C path: one CTA per 128 x 64 tile, 2 CTAs per SM like dgemm_nn_fast_kernel, so CTAs
// of other streams (the LU's look-ahead panel chain) still find SM slots between tiles -- a persistent grid that
// fills every SM would serialise them behind the whole update.  The product P = A B is accumulated from zero while
// the CTA's C tile is pulled into L2 (prefetch at the start); after the main loop the stage buffers are free and
// take the C tile (bulk loads, mbarrier), every thread forms C - P in place (LDS.64 / STS.64) and the tile leaves
// by bulk stores.  No global load or store of C goes through the LSU.
// requires: A, B, C 16-byte aligned; lda, ldb, ldc, m even; k % BK == 0, k > 0.  dynamic shared memory = T_SMEM_BYTES.
// ---------------------------------------------------------------------------------------
constexpr size_t T_SMEM_BYTES = SMEM_BYTES + 16;
static_assert((size_t)C_STAGE * sizeof(double) <= SMEM_BYTES, "the C tile must fit into the stage buffers");

template <int PROBE>
__global__ void __launch_bounds__(THREADS, 2)
    dgemm_sub_tma_kernel(int m, int n, int k, const double* __restrict__ A, int64_t lda, const double* __restrict__ B,
                         int64_t ldb, double* __restrict__ C, int64_t ldc, const int* __restrict__ fail, int dep) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;
  uint64_t* bar = (uint64_t*)(smem + STAGES * (A_STAGE + B_STAGE));
  const int failword = lu_failword(fail);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int KT = k / BK;
  const int rows = min(BM, m - m0), cols = min(BN, n - n0);
  if (tid == 0) {
    mbar_init(bar, 8);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  double acc[4][4][2];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[f][g][0] = acc[f][g][1] = 0.0;

  const int a_mc = (tid & 63) * 2;
  const int a_kr = tid >> 6;
  const int a_gm = m0 + a_mc;
  const int a_bytes = (a_gm < m) ? 16 : 0;
  const double* a_src = A + (int64_t)a_kr * lda + (a_bytes ? a_gm : 0);
  const int64_t a_step4 = 4 * lda;
  const int64_t a_stepk = (int64_t)BK * lda;
  const int a_dst = a_kr * LDA_S + a_mc;
  const int b_kc = (tid & 7) * 2;
  const int b_nr = tid >> 3;
  const int b_bytes0 = (n0 + b_nr < n) ? 16 : 0;
  const int b_bytes1 = (n0 + b_nr + 32 < n) ? 16 : 0;
  const double* b_src0 = B + (b_bytes0 ? (int64_t)(n0 + b_nr) * ldb : 0) + b_kc;
  const double* b_src1 = B + (b_bytes1 ? (int64_t)(n0 + b_nr + 32) * ldb : 0) + b_kc;
  const int b_dst = b_nr * LDB_S + b_kc;

  auto copy_a = [&](int stage, int kt, int c0) {
    double* as = As + stage * A_STAGE + a_dst;
    const double* src = a_src + (int64_t)kt * a_stepk;
    cp_async16(as + (c0 * 4) * LDA_S, src + c0 * a_step4, a_bytes);
    cp_async16(as + (c0 * 4 + 4) * LDA_S, src + (c0 + 1) * a_step4, a_bytes);
  };
  auto copy_b = [&](int stage, int kt) {
    double* bs = Bs + stage * B_STAGE + b_dst;
    cp_async16(bs, b_src0 + kt * BK, b_bytes0);
    cp_async16(bs + 32 * LDB_S, b_src1 + kt * BK, b_bytes1);
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) {
      copy_a(s, s, 0);
      copy_a(s, s, 2);
      copy_b(s, s);
    }
    cp_async_commit();
  }
  {  // my own C tile -> L2 (64 columns x 8 lines of 128 B): the bulk loads after the main loop then hit L2
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int line = tid + q * THREADS;
      const int col = line >> 3, row = (line & 7) * 16;
      if (col < cols && row < rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(C + (int64_t)(n0 + col) * ldc + m0 + row));
    }
  }

  const int a_row = wm * 32 + (lane >> 2);
  const int b_row = wn * 32 + (lane >> 2);
  const int kq = lane & 3;
  int cs = 0, ls = STAGES - 1;

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    const bool do_load = nxt < KT;
    const double* as = As + cs * A_STAGE;
    const double* bs = Bs + cs * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) af[f] = as[(kk + kq) * LDA_S + a_row + 8 * f];
#pragma unroll
      for (int g = 0; g < 4; ++g) bf[g] = bs[(b_row + 8 * g) * LDB_S + kk + kq];
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) dmma884(acc[f][g][0], acc[f][g][1], af[f], bf[g]);
      if (do_load) {
        if (kk == 0) copy_a(ls, nxt, 0);
        if (kk == 4) copy_a(ls, nxt, 2);
        if (kk == 8) copy_b(ls, nxt);
      }
    }
    cp_async_commit();
    cs = (cs + 1 == STAGES) ? 0 : cs + 1;
    ls = (ls + 1 == STAGES) ? 0 : ls + 1;
  }
  cp_async_wait<0>();
  if (lu_poisoned(failword, dep)) return;
  __syncthreads();  // every warp is done with the stage buffers: they become the C tile (columns of LDC_S doubles)
  double* Cs = smem;
  const int c0 = warp * (BN / 8);
  const int mine = max(0, min(BN / 8, cols - c0));
  if (lane == 0 && !(PROBE & 2)) {
    // generic-proxy reads of the stages (all warps, ordered by the barrier) before the async-proxy writes below
    const unsigned z = after_barrier_zero(bar);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    mbar_arrive_expect_tx(bar, (unsigned)(rows * mine) * 8u + z);
    const double* src = C + (int64_t)(n0 + c0) * ldc + m0;
#pragma unroll
    for (int c = 0; c < BN / 8; ++c)
      if (c < mine) bulk_g2s(Cs + (c0 + c) * LDC_S, src + (int64_t)c * ldc, (unsigned)rows * 8u + z, bar);
  }
  __syncwarp();
  if (!(PROBE & 2)) mbar_wait(bar, 0);
  double* cb = Cs + (wn * 32 + 2 * (lane & 3)) * LDC_S + wm * 32 + (lane >> 2);
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      double* cp = cb + (8 * g) * LDC_S + 8 * f;
      const double t0 = cp[0], t1 = cp[LDC_S];
      cp[0] = t0 - acc[f][g][0];
      cp[LDC_S] = t1 - acc[f][g][1];
    }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  if (lane == 0 && (!(PROBE & 1) || m < 0)) {
    const unsigned zs = after_barrier_zero(Cs);  // every warp's C - P is in the buffer before any column leaves
    double* dst = C + (int64_t)(n0 + c0) * ldc + m0;
#pragma unroll
    for (int c = 0; c < BN / 8; ++c)
      if (c < mine) bulk_s2g(dst + (int64_t)c * ldc, Cs + (c0 + c) * LDC_S, (unsigned)rows * 8u + zs);
    bulk_commit();
    bulk_wait_read0();
  }
}

// ---------------------------------------------------------------------------------------
// Per-tile kernel with A and B fed by the tensor-map TMA as well.  No thread computes a load address, there is no
// LDGSTS and no CTA barrier in the main loop:
//   * 4 stages of 24 KB: A k-tile = 8 boxes of 16 (m) x 16 (k) doubles, B k-tile = one box of 16 (k) x 64 (n), both with
//     the 128-byte swizzle (rows of 128 B; the 16-byte chunk index is XORed with the row index mod 8), landing on a
//     "full" mbarrier per stage (expect_tx = 24 576 B).  9 UTMALDG per k-tile, issued by lane 0 of warp (kt mod 8),
//     two k-tiles ahead, after an "empty" mbarrier (8 arrivals: every warp, after its last LDS of the stage's
//     previous use) -- with 4 stages and a prefetch distance of 2 that wait is over before it starts;
//   * bank-conflict-free fragment loads from the dense swizzled tiles need the four k of one DMMA to differ in
//     ((k & 7) >> 1) [A: 4 k-rows x 4 consecutive m per half-warp] and in (k & 1, k >> 3) [B: 4 k x 4 n-rows]:
//     the k of a k-tile are dealt to the four DMMA groups as {0,3,12,15} ^ x, x = 0, 1, 4, 5 (any assignment of
//     the summation index to DMMA slots is valid as long as A and B use the same one).  A thread's 8 + 4 fragment
//     addresses are then TA0 ^ (144 x) ^ (64 (f & 1)) + 2048 (f >> 1) and TB0 ^ (8 x) + 1024 g: two registers
//     and compile-time constants;
//   * C as in dgemm_sub_tma_kernel (bulk copies into the freed stages, C - P in place, bulk stores).
// A stage is released (arrive on its "empty" mbarrier) at the TOP of the following iteration, behind the wait loop
// of the next k-tile, never at the end of its own iteration: there ptxas schedules the arrive in front of the last
// 24 DMMAs, where the fragment loads that feed them may still be queued, and the producer's next TMA box can then
// land under them.  Alone on the GPU that never showed; inside the LU, with the look-ahead kernels competing for
// the LSU, it corrupted a few half-tiles per factorization by one k-tile's worth of products (4 runs of 5).  Behind
// the wait loop every DMMA of the released k-tile has been issued, so its operands have arrived (in-order issue).
// PROBE: 1 = no output, 2 = no C input, 8 = no loads (timing probes, wrong results); 16 = CTA barrier per k-tile
// instead of the empty barriers (33.4 TF), 32 = direct C - P epilogue, 64 = C tile by coalesced generic loads.
// ---------------------------------------------------------------------------------------
constexpr int T2_STAGES = 4;
constexpr int T2_A_BYTES = BM * BK * 8, T2_B_BYTES = BN * BK * 8, T2_STAGE_BYTES = T2_A_BYTES + T2_B_BYTES;
constexpr size_t T2_SMEM_BYTES = 1024 + (size_t)T2_STAGES * T2_STAGE_BYTES + 128;
static_assert((size_t)C_STAGE * sizeof(double) <= (size_t)T2_STAGES * T2_STAGE_BYTES, "the C tile must fit into the stages");

__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

template <int PROBE>
__global__ void __launch_bounds__(THREADS, 2)
    dgemm_sub_tma2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int m, int n, int k,
                          double* __restrict__ C, int64_t ldc, const int* __restrict__ fail, int dep) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // 1024: the CTA's window itself must sit on a swizzle-pattern boundary
  const unsigned raw = (unsigned)__cvta_generic_to_shared(smem_raw);
  const unsigned sbase = (raw + 1023u) & ~1023u;  // the swizzle pattern is a function of the absolute address
  unsigned char* sgen = smem_raw + (sbase - raw);
  uint64_t* bars = (uint64_t*)(sgen + T2_STAGES * T2_STAGE_BYTES);  // full[4], empty[4], cbar
  const unsigned bar0 = sbase + T2_STAGES * T2_STAGE_BYTES;
  const bool poisoned = lu_poisoned(lu_failword(fail), dep);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int KT = k / BK;
  const int rows = min(BM, m - m0), cols = min(BN, n - n0);

  auto produce = [&](int kt) {  // one thread: the 9 boxes of k-tile kt
    const int st = kt & (T2_STAGES - 1);
    const unsigned fb = bar0 + 8u * st;
    const unsigned dst = sbase + st * T2_STAGE_BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"((unsigned)T2_STAGE_BYTES) : "memory");
#pragma unroll
    for (int j = 0; j < BM / 16; ++j) tma_load_2d(dst + j * 2048, &mapA, m0 + 16 * j, kt * BK, fb);
    tma_load_2d(dst + T2_A_BYTES, &mapB, kt * BK, n0, fb);
  };

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < T2_STAGES; ++i) {
      mbar_init(&bars[i], 1);
      mbar_init(&bars[T2_STAGES + i], 8);
    }
    mbar_init(&bars[2 * T2_STAGES], 8);
    mbar_init(&bars[2 * T2_STAGES + 1], 8);
    mbar_init(&bars[2 * T2_STAGES + 2], 8);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (!(PROBE & 8)) {
      produce(0);
      if (KT > 1) produce(1);
      if ((PROBE & 16) && KT > 2) produce(2);
    }
  }
  __syncthreads();

  double acc[4][4][2];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[f][g][0] = acc[f][g][1] = 0.0;

  {  // my own C tile -> L2 (64 columns x 8 lines of 128 B): the loads after the main loop then hit L2
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int line = tid + q * THREADS;
      const int col = line >> 3, row = (line & 7) * 16;
      if (col < cols && row < rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(C + (int64_t)(n0 + col) * ldc + m0 + row));
    }
  }

  const int rr = lane >> 2, kq = lane & 3;
  const int kp = (kq & 1) * 3 + (kq >> 1) * 12;  // this lane's k in DMMA group 0: {0, 3, 12, 15}
  const unsigned ta0 = (unsigned)(wm * 2) * 2048u + kp * 128u + ((((rr >> 1) ^ kp) & 7) << 4) + (rr & 1) * 8u;
  const unsigned tb0 = T2_A_BYTES + (unsigned)(wn * 32 + rr) * 128u + ((((kp >> 1) ^ rr) & 7) << 4) + (kp & 1) * 8u;

  for (int kt = 0; kt < KT; ++kt) {
    const int cs = kt & (T2_STAGES - 1);
    if (!(PROBE & 8)) mbar_wait(&bars[cs], (kt / T2_STAGES) & 1);
    // Release the stage of k-tile kt - 1 only now, behind the wait loop: every DMMA of that k-tile has been issued
    // (in-order issue), so every LDS that fed them has returned.  At the end of iteration kt - 1 ptxas moves the
    // arrive up in front of the last 24 DMMAs, i.e. to a point where the last fragment loads may still be in flight.
    if (!(PROBE & 16) && kt > 0) {
      if (lane == 0) mbar_arrive(&bars[T2_STAGES + ((kt - 1) & (T2_STAGES - 1))]);
      __syncwarp();
    }
    if (PROBE & 16) {
      // CTA barrier per k-tile: everybody is done with k-tile kt - 1, whose stage takes k-tile kt + 3
      __syncthreads();
      if (warp == (kt & 7) && lane == 0 && kt + 3 < KT && !(PROBE & 8)) produce(kt + 3);
    } else if (warp == (kt & 7) && lane == 0 && kt + 2 < KT && !(PROBE & 8)) {
      const int nx = kt + 2;
      if (nx >= T2_STAGES) mbar_wait(&bars[T2_STAGES + (nx & (T2_STAGES - 1))], (nx / T2_STAGES - 1) & 1);
      produce(nx);
    }
    __syncwarp();
    const unsigned char* st = sgen + cs * T2_STAGE_BYTES;
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      constexpr int xs[4] = {0, 1, 4, 5};
      const unsigned xa = xs[gi] * 144u, xb = xs[gi] * 8u;
      double af[4], bf[4];
#pragma unroll
      for (int f = 0; f < 4; ++f)
        af[f] = *(const double*)(st + ((ta0 ^ xa ^ ((f & 1) * 64u)) + (f >> 1) * 2048u));
#pragma unroll
      for (int g = 0; g < 4; ++g) bf[g] = *(const double*)(st + ((tb0 ^ xb) + g * 1024u));
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) dmma884(acc[f][g][0], acc[f][g][1], af[f], bf[g]);
    }
  }
  if (poisoned) return;
  if (PROBE & 32) {  // direct epilogue (debugging): C - P through the LSU
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int gm = m0 + wm * 32 + 8 * f + (lane >> 2);
      if (gm >= m) continue;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gn = n0 + wn * 32 + 8 * g + 2 * (lane & 3) + e;
          if (gn < n) C[(int64_t)gn * ldc + gm] -= acc[f][g][e];
        }
    }
    return;
  }
  // every warp is done with the stages: they become the C tile (columns of LDC_S doubles)
  __syncwarp();
  if (lane == 0) mbar_arrive(&bars[2 * T2_STAGES + 1]);
  mbar_wait(&bars[2 * T2_STAGES + 1], 0);
  double* Cs = (double*)sgen;
  uint64_t* cbar = &bars[2 * T2_STAGES];
  const int c0 = warp * (BN / 8);
  const int mine = max(0, min(BN / 8, cols - c0));
  if (PROBE & 64) {  // C tile by coalesced 16-byte loads (thread t: rows 2 (t % 64), 2 (t % 64) + 1 of columns t / 64 + 4 i)
    const int r = 2 * (tid & 63), cq = tid >> 6;
    double2 v[BN / 4];
#pragma unroll
    for (int i2 = 0; i2 < BN / 4; ++i2) {
      const int c = cq + 4 * i2;
      v[i2] = (r < rows && c < cols) ? __ldcs((const double2*)(C + (int64_t)(n0 + c) * ldc + m0 + r)) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int i2 = 0; i2 < BN / 4; ++i2) *(double2*)(Cs + (cq + 4 * i2) * LDC_S + r) = v[i2];
    __syncthreads();
  }
  if (lane == 0 && !(PROBE & 2) && !(PROBE & 64)) {
    const unsigned z = after_barrier_zero(cbar);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // (see dgemm_sub_tma_kernel)
    mbar_arrive_expect_tx(cbar, (unsigned)(rows * mine) * 8u + z);
    const double* src = C + (int64_t)(n0 + c0) * ldc + m0;
#pragma unroll
    for (int c = 0; c < BN / 8; ++c)
      if (c < mine) bulk_g2s(Cs + (c0 + c) * LDC_S, src + (int64_t)c * ldc, (unsigned)rows * 8u + z, cbar);
  }
  __syncwarp();
  if (!(PROBE & 2) && !(PROBE & 64)) mbar_wait(cbar, 0);
  double* cb = Cs + (wn * 32 + 2 * (lane & 3)) * LDC_S + wm * 32 + (lane >> 2);
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      double* cp = cb + (8 * g) * LDC_S + 8 * f;
      const double t0 = cp[0], t1 = cp[LDC_S];
      cp[0] = t0 - acc[f][g][0];
      cp[LDC_S] = t1 - acc[f][g][1];
    }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive(&bars[2 * T2_STAGES + 2]);
  mbar_wait(&bars[2 * T2_STAGES + 2], 0);
  if (lane == 0 && (!(PROBE & 1) || m < 0)) {
    const unsigned zs = after_barrier_zero(Cs);  // every warp's C - P is in the buffer before any column leaves
    double* dst = C + (int64_t)(n0 + c0) * ldc + m0;
#pragma unroll
    for (int c = 0; c < BN / 8; ++c)
      if (c < mine) bulk_s2g(dst + (int64_t)c * ldc, Cs + (c0 + c) * LDC_S, (unsigned)rows * 8u + zs);
    bulk_commit();
    bulk_wait_read0();
  }
}

// Host side: tensor maps of the operands of one product.  A: m x k column-major (lda), boxes of 16 x 16; B: k x n
// column-major (ldb), boxes of 16 x 64; both FP64, 128-byte swizzle, out-of-bounds elements read as zero.
typedef CUresult (*mmr_tmap_encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline mmr_tmap_encode_t mmr_tmap_encoder() {
  static mmr_tmap_encode_t fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (mmr_tmap_encode_t)p;
    cudaGetLastError();
  }
  return fn;
}
inline bool mmr_tmap_2d(CUtensorMap* map, const double* base, uint64_t inner, uint64_t outer, int64_t ld, uint32_t box_inner,
                        uint32_t box_outer) {
  mmr_tmap_encode_t enc = mmr_tmap_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t es[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mmr_gemm
