// pairs_matching assembly (demag.py:280-321) -- unique-pair table on device.
#include "cuboid.cuh"

extern "C" int mmr_assemble_matched(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
                                    const double* d_dim, const int32_t* d_cls, const double* d_chi,
                                    double* d_out, int64_t ld, int64_t table_cap_slots,
                                    int64_t* n_unique_out) {
  (void)ctx; (void)n; (void)d_pos; (void)d_quat; (void)d_dim; (void)d_cls; (void)d_chi; (void)d_out; (void)ld;
  (void)table_cap_slots; (void)n_unique_out;
  mmr_set_error("mmr_assemble_matched: not implemented yet");
  return -7;
}
