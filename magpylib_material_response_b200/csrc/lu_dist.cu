// Distributed LU (1-D block-cyclic over row blocks of Q = column panels of Q^T).
#include "comm.h"
#include "common.cuh"

extern "C" int mmr_lu_solve_dist(mmr_ctx* ctx, int64_t N, int64_t nrows_local, int64_t tgt_block,
                                 double* d_Qlocal, int64_t ld, double* d_b, int* info_out) {
  (void)ctx; (void)N; (void)nrows_local; (void)tgt_block; (void)d_Qlocal; (void)ld; (void)d_b; (void)info_out;
  mmr_set_error("mmr_lu_solve_dist: not implemented yet");
  return -7;
}
