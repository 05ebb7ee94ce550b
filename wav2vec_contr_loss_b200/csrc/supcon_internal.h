// Internal launch interfaces between the C-ABI dispatcher and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "supcon_b200.h"

namespace supcon {

struct FfmaArgs {
  const void* z;             // [n_total][d]
  const int32_t* labels;     // [n_total]
  float* row_stats;          // fwd out / topk-idx in: [n_rows][STRIDE]
  const float* stats_all;    // bwd in: [n_total][STRIDE]
  double* partials;          // fwd: out [8]; bwd: in (global sums)
  float* loss_out;           // fwd: optional
  double* block_partials;    // workspace [blocks][8]
  unsigned* ticket;          // workspace counter (zero between launches)
  const float* grad_out;     // bwd: optional device scalar
  void* dz_out;              // bwd: [n_rows][d]
  int n_total, row_offset, n_rows, d;
  int z_dtype, similarity, topk, kcap, mine, vec_ok;
  float tau, alpha, lambda_uni, uni_t;
};

int ffma_kcap();
size_t ffma_workspace_bytes(int n_rows);
cudaError_t ffma_forward(const FfmaArgs& a, cudaStream_t stream);
cudaError_t ffma_backward(const FfmaArgs& a, int dz_dtype, cudaStream_t stream);
cudaError_t ffma_topk_indices(const FfmaArgs& a, int32_t* idx_out, cudaStream_t stream);

// tcgen05 building-block diagnostic (supcon_tc_debug.cu)
int tc_debug_tile(const void* z_bf16, int n, int d, int row_i, int row_j, float* s_out, float* o_out,
                  cudaStream_t stream, const char** err);

}  // namespace supcon
