/*
 * mmr_b200.h -- C ABI of the B200-native demagnetization hot path.
 *
 * The reference (magpylib-material-response, pure Python) has no FFI layer; the seam this
 * library fills is what sits underneath its two hot functions:
 *
 *   reference call (file:line, relative to /root/reference)            replaced by
 *   ------------------------------------------------------------------  -----------------------
 *   3 x magpy.getH(...)   src/magpylib_material_response/demag.py:190,  mmr_assemble
 *                         :216, :220  (demag_tensor :124-224)
 *   filter_distance       demag.py:227-277 (max_dist mask)              mmr_assemble (max_dist)
 *   match_pairs           demag.py:280-321 (pairs_matching dedupe)      mmr_assemble_matched
 *   T *= mu_0; T.swapaxes(2,3).reshape(3n,3n).T   demag.py:451-452      mmr_assemble (layout)
 *   Q = np.eye(3n) - np.matmul(S, T)              demag.py:432,:466     mmr_assemble (chi != NULL)
 *   np.linalg.solve(Q, rhs)                       demag.py:469-470      mmr_lu_factor + mmr_lu_solve
 *                                                                       or mmr_gmres
 *   collection.getB(observers) of the solved cells                      mmr_field
 *       (tests/test_isotropic_anisotropic.py:23, magpylib call)
 *   magpy.getB(currents_list, pos, sumup=True)    demag.py:456-463      mmr_field_currents
 *
 * Conventions
 *   - All floating point is FP64.  All pointers named d_* are DEVICE pointers owned by the
 *     caller (torch tensors); the library allocates only workspace inside the context.
 *   - Geometry is array-of-structs, row-major: pos[n][3] global cell barycenters (metres),
 *     quat[n][4] scalar-LAST (x,y,z,w) exactly as scipy Rotation.as_quat() (demag.py:178),
 *     dim[n][3] full edge lengths.
 *   - Cell-major ("interleaved") vector/matrix ordering: index 3*cell + component.  The
 *     reference's component-major ordering (comp*n + cell, demag.py:426-428 order="F") is a
 *     symmetric permutation of it; the Python host layer converts at the boundary.
 *   - Matrices are ROW-major with leading dimension ld (in doubles): element (r,c) at
 *     d_A[r*ld + c]; row r = 3*target + field component, column c = 3*source + polarization
 *     component.  Rows are what multi-GPU sharding splits ("target-row blocks").
 *   - Every function returns an int status: 0 ok, <0 invalid argument, >0 a CUDA (1000+code)
 *     or NCCL (2000+code) error.  Nothing throws across the boundary.  mmr_last_error()
 *     returns a thread-local message for the last non-zero status.
 *   - One context per thread/stream; contexts are independent.  All work is enqueued on the
 *     context's stream; functions that return host scalars synchronize that stream.
 */
#ifndef MMR_B200_H
#define MMR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMR_ABI_VERSION 1

typedef struct mmr_ctx mmr_ctx;

/* layouts for mmr_assemble */
#define MMR_LAYOUT_CELL_MAJOR 0 /* out[(3*(j-j0)+l)*ld + 3*i+k], dimensionless mu0*H        */
#define MMR_LAYOUT_REFERENCE 1  /* out[((k*n+i)*n+j)*3+l] = demag.py:224 (3,n,n,3), A/m per T */

/* field selector for mmr_field */
#define MMR_FIELD_B 0
#define MMR_FIELD_H 1

int mmr_abi_version(void);
const char* mmr_last_error(void);

/* ---- context -------------------------------------------------------------------------- */
/* stream: a cudaStream_t (as void*) or NULL for the legacy default stream. */
int mmr_ctx_create(int device, void* stream, mmr_ctx** out);
int mmr_ctx_destroy(mmr_ctx* ctx);
int mmr_ctx_sync(mmr_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t mmr_ctx_launch_count(mmr_ctx* ctx);

/* Optional per-kernel-class timing with CUDA events on the context's stream (bench.py's live
 * roofline).  set_profiling(on) clears the records; profile() synchronizes the stream and
 * returns the summed duration (ms), summed algorithmic units and launch count of one class. */
#define MMR_PROF_ASSEMBLE 0   /* units: cell pairs evaluated                    */
#define MMR_PROF_GEMM 1       /* units: flops of the LU trailing DMMA updates   */
#define MMR_PROF_PANEL 2      /* units: panel columns factored (strip kernels)  */
#define MMR_PROF_TRSM 3       /* units: flops                                   */
#define MMR_PROF_LASWP 4      /* units: bytes moved                             */
#define MMR_PROF_MATVEC 5     /* units: matrix bytes streamed                   */
#define MMR_PROF_PANEL_GEMM 6 /* units: flops of the in-panel rank-W updates    */
#define MMR_PROF_TRISOLVE 7   /* units: matrix bytes streamed                   */
#define MMR_PROF_COMM 8       /* units: bytes broadcast / gathered (NCCL calls)   */
#define MMR_PROF_NKINDS 9
/* on = 0: off; 1: every class; otherwise a mask with bit (k + 1) set for every class k to record (event pairs
 * cost host time per launch: a timed region records only the classes it reports). */
int mmr_ctx_set_profiling(mmr_ctx* ctx, int on);
int mmr_ctx_profile(mmr_ctx* ctx, int kind, double* ms_out, double* units_out,
                    int64_t* launches_out);
/* Individual record durations (ms) of one class, in launch order; returns the count written. */
int64_t mmr_ctx_profile_records(mmr_ctx* ctx, int kind, double* ms_out, int64_t cap);
/* All records in launch order as a timeline: kind (| 0x100: side stream, | 0x200: communication stream), start (ms after mmr_ctx_set_profiling(on)) and
 * duration (ms) -- records of both streams of the context share one device clock.  Returns the count. */
int64_t mmr_ctx_profile_timeline(mmr_ctx* ctx, int* kind_out, double* start_ms_out, double* dur_ms_out,
                                 int64_t cap);

/* FP64 roofline denominators of this device, measured now (about 0.1 s): the FMA-pipe peak that bounds
 * the assembly kernel and the tensor-pipe (DMMA.8x8x4) peak that bounds the LU trailing update, in
 * TFLOP/s (MEASURED_PEAKS.json has no FP64 entry).  Host pointers, may be NULL. */
int mmr_fp64_peaks(mmr_ctx* ctx, double* fma_tflops_out, double* dmma_tflops_out);

/* ---- kernel 1: demag tensor assembly (demag.py:124-224 + :451-452 + :466) ------------- */
/*
 * Evaluates the analytic cuboid H-field block of every source cell i in [src_begin,src_end)
 * at the barycenter of every target cell j of the target set, for the three global unit
 * polarizations at once (the 9 entries of a pair share their 6 transcendental terms).
 *
 * Target set: block-cyclic over cells -- local target t (0 <= t < n_tgt_local) is global cell
 *     j = ((t / tgt_block) * tgt_nparts + tgt_part) * tgt_block + t % tgt_block.
 * Single GPU / contiguous range: tgt_block = n_tgt_local, tgt_nparts = 1, tgt_part = 0 gives
 * j = t (use tgt_first to offset: j = tgt_first + ...).
 *
 * d_chi == NULL : out = N   (N[l,j;k,i] = mu0 * H_l at j for unit J_k of i; inside -> includes -1)
 * d_chi != NULL : out = I - diag(chi) * N  with chi[3*j+l] (cell-major), i.e. the system matrix Q
 *                 of demag.py:466 fused into the store.
 * max_dist != 0 : pairs with |p_j - p_i| / max(dim_i U dim_j) >= max_dist contribute N = 0
 *                 (evident intent of demag.py:227-277, which raises as written; Appendix A Q2).
 * layout        : MMR_LAYOUT_CELL_MAJOR (ld >= 3*n doubles, rows are local targets) or
 *                 MMR_LAYOUT_REFERENCE (d_chi must be NULL, full target range, ld ignored;
 *                 values divided by mu0 to match demag.py:224 units).
 */
int mmr_assemble(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
                 const double* d_dim, int64_t src_begin, int64_t src_end, int64_t tgt_first,
                 int64_t n_tgt_local, int64_t tgt_block, int64_t tgt_nparts, int64_t tgt_part,
                 const double* d_chi, double max_dist, int layout, double* d_out, int64_t ld);

/*
 * pairs_matching variant (demag.py:280-321): pairs that are equivalent -- same rounded
 * displacement p_j - p_i (8 decimals after +1e-9, as demag.py:302-304) and same source class
 * (d_cls[i], host-computed id of the cell's rounded orientation+dimension) -- are evaluated
 * once (representative = smallest flat index i*n+j, like np.unique's return_index) and copied
 * to all duplicates.  Output and arguments as mmr_assemble (cell-major only).  n_unique_out
 * (host pointer, may be NULL) receives the number of unique pairs evaluated.
 * Returns status -3 if the unique-pair table would exceed table_cap_slots (caller falls back
 * to mmr_assemble, which produces the same matrix).
 */
int mmr_assemble_matched(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
                         const double* d_dim, const int32_t* d_cls, const double* d_chi,
                         double* d_out, int64_t ld, int64_t table_cap_slots,
                         int64_t* n_unique_out);

/* The same for a row shard (multi-GPU): only the n_tgt_local targets of the block-cyclic target set
 * (tgt_block, tgt_nparts, tgt_part), as in mmr_assemble; rows of d_out are the local targets.  The unique-pair
 * table covers this shard's pairs, so a class representative is the smallest flat index i*n+j among them. */
int mmr_assemble_matched_rows(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
                              const double* d_dim, const int32_t* d_cls, const double* d_chi,
                              double* d_out, int64_t ld, int64_t table_cap_slots, int64_t n_tgt_local,
                              int64_t tgt_block, int64_t tgt_nparts, int64_t tgt_part,
                              int64_t* n_unique_out);

/* ---- kernel 2a: dense LU with partial pivoting (np.linalg.solve / dgesv, demag.py:470) -- */
/*
 * Factorizes the row-major N x N matrix Q in place.  Pivoting is over COLUMNS of Q (it is
 * partial pivoting of the column-major view Q^T = P L U), which keeps every pivot search
 * inside one row block -- the property the multi-GPU row sharding relies on.
 * d_ipiv: N int32 (device).  info_out (host, may be NULL): 0, or k>0 if pivot k-1 is exactly 0.
 *
 * Factor format.  The factors are consumed by mmr_lu_solve only.  When no panel of the factorization needed an
 * interchange (always the case for the demag matrices Q = I - chi N), every diagonal panel block (128 columns,
 * at multiples of 128) is stored as L11^-1 (strictly lower part) and U11^-1 (upper part) instead of L11 and U11,
 * and bit 30 of d_ipiv[0] is set (MMR_LU_FORMAT_INVDIAG); the pivot index itself is d_ipiv[0] & (2^30 - 1).
 * The solves on such factors are block products without substitution.  Otherwise plain LAPACK-style L\U.
 */
#define MMR_LU_FORMAT_INVDIAG (1 << 30)
int mmr_lu_factor(mmr_ctx* ctx, int64_t N, double* d_Q, int64_t ld, int32_t* d_ipiv,
                  int* info_out);
/* Solves Q X = B for nrhs right-hand sides; d_B is row-major nrhs x N (each RHS contiguous),
 * overwritten by the solution. */
int mmr_lu_solve(mmr_ctx* ctx, int64_t N, const double* d_LU, int64_t ld, const int32_t* d_ipiv,
                 double* d_B, int64_t nrhs);

/* The diagonal-block step of a speculative LU panel on its own (exported for tests and micro-benchmarks):
 * LU WITHOUT interchanges of the nb x nb (nb <= 128) column-major block d_A into d_S (packed L and U, ld lds)
 * plus the dense inverses of both factors (nb x nb, ld nb).  *rejected_out = 1 when partial pivoting would
 * have interchanged rows (a multiplier above 1 in magnitude) or a pivot is zero; outputs are then undefined. */
int mmr_lu_diag_block(mmr_ctx* ctx, const double* d_A, int64_t ld, int nb, double* d_S, int64_t lds,
                      double* d_Linv, double* d_Uinv, int* rejected_out);

/* SM clock stamps (30 x int64) at the phase boundaries of the last mmr_lu_diag_block call (micro-benchmarks). */
int mmr_lu_diag_block_clocks(mmr_ctx* ctx, long long* out30);

/* C = alpha * op(A) op(B) + beta * C on the FP64 tensor-core (DMMA) path used by the LU
 * trailing update; exported for tests and the roofline measurement.  COLUMN-major (BLAS
 * convention, "N","N" only): C is m x n, A is m x k, B is k x n. */
int mmr_dgemm_nn(mmr_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha, const double* d_A,
                 int64_t lda, const double* d_B, int64_t ldb, double beta, double* d_C,
                 int64_t ldc);

/* B (nb0 + nb1 rows x ncols, COLUMN-major, ld ldb, in place) <- L_blk^-1 B for the unit-lower two-panel block
 * L_blk = [L00 0; L10 L11], given L00^-1 (nb0 x nb0, ld nb0), L11^-1 (nb1 x nb1, ld nb1) and L10 (nb1 x nb0, ld ldl):
 * the U12 step of the blocked LU (the triangular solve inside LAPACK dgetrf behind np.linalg.solve, demag.py:470)
 * as ONE launch.  nb0 = 128, 0 < nb1 <= 128 with nb1 % 16 == 0, 16-byte aligned operands, even leading dimensions;
 * exported for tests. */
int mmr_block_trsm(mmr_ctx* ctx, int nb0, int nb1, const double* d_Linv0, const double* d_Linv1, const double* d_L10,
                   int64_t ldl, double* d_B, int64_t ldb, int ncols);

/* ---- kernel 2b: restarted GMRES on the same Q with iterative-refinement check ----------- */
/*
 * Solves Q x = b (row-major Q, ld).  x holds the initial guess on entry.  Stops when the TRUE
 * residual ||b - Q x|| / ||b|| (recomputed at every restart = the iterative-refinement
 * check) <= tol, or after maxit matvecs.  iters_out / relres_out are host pointers.
 * Returns 0 on convergence, -4 if maxit was reached or the true residual stagnated above tol for two
 * consecutive restart cycles (relres_out then holds the residual that was reached).
 */
int mmr_gmres(mmr_ctx* ctx, int64_t N, const double* d_Q, int64_t ld, const double* d_b,
              double* d_x, double tol, int restart, int maxit, int* iters_out,
              double* relres_out);

/* GMRES with a block-Jacobi right preconditioner: the diagonal blocks Q[r0:r0+rows, r0:r0+rows] of the `nblocks`
 * listed row ranges (host arrays; one magnet of a multi-magnet collection each, typically only the
 * high-susceptibility ones) are copied, LU-factorized with the DMMA LU and applied as M^-1 in Q M^-1 u = b,
 * x = M^-1 u.  The stopping test is still the true residual ||b - Q x|| / ||b||.  nblocks = 0: plain GMRES. */
int mmr_gmres_bj(mmr_ctx* ctx, int64_t N, const double* d_Q, int64_t ld, const double* d_b, double* d_x,
                 double tol, int restart, int maxit, int nblocks, const int64_t* block_row0,
                 const int64_t* block_rows, int* iters_out, double* relres_out);

/* y = Q x for a row block: y[r] = sum_c Q[r*ld + c] x[c], r in [0, nrows). */
int mmr_matvec(mmr_ctx* ctx, int64_t nrows, int64_t ncols, const double* d_Q, int64_t ld,
               const double* d_x, double* d_y);

/* ---- field of the (solved) cells at observers (magpylib getB/getH) ---- */
/* d_pol: [n][3] LOCAL-frame polarizations (tesla); d_obs: [m][3]; d_out: [m][3] global-frame
 * B (tesla) or H (A/m), summed over all n cells. */
int mmr_field(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
              const double* d_dim, const double* d_pol, int64_t m, const double* d_obs,
              int field, double* d_out);

/* ---- field of current sources at observers (magpy.getB(currents_list, pos, sumup=True),
 *      demag.py:456-463: the current-source term of the right-hand side) ---- */
/* d_seg : [nseg][7]  straight segments of current.Polyline objects in the GLOBAL frame:
 *                    (p1x, p1y, p1z, p2x, p2y, p2z, current in A)
 * d_loop: [nloop][9] current.Circle objects: (centre xyz, orientation quaternion x,y,z,w
 *                    scalar-last, diameter in m, current in A); the loop lies in its local xy plane
 * d_out : [m][3] global-frame B (tesla) or H (A/m), summed over all elements.  Observers on a
 *         conductor get no contribution from it (magpylib convention). */
int mmr_field_currents(mmr_ctx* ctx, int64_t m, const double* d_obs, int64_t nseg,
                       const double* d_seg, int64_t nloop, const double* d_loop, int field,
                       double* d_out);

/* ---- multi-GPU (one process per GPU; NCCL communicator owned by the context) ------------ */
/* rank 0 calls mmr_nccl_unique_id, the 128-byte id is shipped to the other ranks by the host
 * (torch.distributed broadcast), then every rank calls mmr_comm_init. */
int mmr_nccl_unique_id(void* id128_out);
int mmr_comm_init(mmr_ctx* ctx, const void* id128, int nranks, int rank);
int mmr_comm_destroy(mmr_ctx* ctx);
/* Optional: peer-memory buffers for the flag-synchronised triangular sweeps of the distributed LU (CUDA IPC, one
 * NVSwitch box).  Every rank exports the 64-byte handle of its buffer, the host gathers the handles of ranks
 * 0..nranks-1 into one table (64 bytes each) and every rank opens it.  Without them the sweeps use one small
 * ncclBroadcast per block. */
int mmr_comm_p2p_export(mmr_ctx* ctx, void* handle64_out);
int mmr_comm_p2p_open(mmr_ctx* ctx, const void* handles, int nranks, int rank);
int mmr_comm_p2p_disable(mmr_ctx* ctx); /* back to the NCCL broadcasts (all ranks must agree) */

/*
 * Distributed GMRES: this rank owns nrows_local rows of Q (row-major, ld >= N), the rows of the
 * target set given by (tgt_block, nranks, rank) as in mmr_assemble (row = 3*cell + comp).
 * d_b and d_x are full-length (N) and replicated; every rank returns the same x.
 * Per iteration: local row-block matvec, ncclAllGather of the y slices, replicated
 * orthogonalization (no further collective).
 */
int mmr_gmres_dist(mmr_ctx* ctx, int64_t N, int64_t nrows_local, int64_t tgt_block,
                   const double* d_Qlocal, int64_t ld, const double* d_b, double* d_x,
                   double tol, int restart, int maxit, int* iters_out, double* relres_out);

/*
 * Distributed LU + solve: 1-D block-cyclic over row blocks of Q (= column blocks of Q^T) of
 * width 3*tgt_block <= 256 (one block = up to two 128-column LU panels = one rank-3*tgt_block
 * DMMA update; 3*tgt_block a multiple of 16, e.g. tgt_block = 80, keeps the fast GEMM kernel).
 * The block factorization is local to its owner; the packed block, both inverted diagonal blocks
 * and the pivots travel as ONE ncclBroadcast per block on the context's side stream, overlapped
 * with the update by the previous block; every rank swaps and updates its own trailing blocks.
 * d_b (N, replicated) is overwritten by the solution on every rank.  *info_out > 0: zero pivot
 * seen by THIS rank (reduce over ranks with max).
 */
int mmr_lu_solve_dist(mmr_ctx* ctx, int64_t N, int64_t nrows_local, int64_t tgt_block,
                      double* d_Qlocal, int64_t ld, double* d_b, int* info_out);
/* The two phases of mmr_lu_solve_dist on their own, so that a factorized (sharded) system can be kept
 * and solved again for further right-hand sides (apply_demag's demag_store on several GPUs):
 * d_ipiv (N int32, device) receives / provides the replicated pivot sequence. */
int mmr_lu_factor_dist(mmr_ctx* ctx, int64_t N, int64_t nrows_local, int64_t tgt_block,
                       double* d_Qlocal, int64_t ld, int32_t* d_ipiv_out, int* info_out);
int mmr_lu_trisolve_dist(mmr_ctx* ctx, int64_t N, int64_t nrows_local, int64_t tgt_block,
                         const double* d_Qlocal, int64_t ld, const int32_t* d_ipiv, double* d_b);

#ifdef __cplusplus
}
#endif
#endif /* MMR_B200_H */
