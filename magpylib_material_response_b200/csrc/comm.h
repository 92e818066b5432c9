// Internal helpers implemented in comm.cu (NCCL via dlopen).
#pragma once
#include "common.cuh"

int mmr_comm_allgather_f64(mmr_ctx* ctx, const double* send, double* recv, size_t count, cudaStream_t s);
int mmr_comm_bcast_bytes(mmr_ctx* ctx, void* buf, size_t bytes, int root, cudaStream_t s);
