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
  // column splits (ffma_plan): partial records / candidate lists / partial dz in the workspace
  int splits, col_tiles, rows_pad;
  float* part;      // [splits][rows_pad][8]
  float* lv_part;   // [splits][rows_pad][kcap]
  int* li_part;
  float* dz_part;   // [splits][rows_pad][d]
};

struct FfmaPlan {
  int fwd_splits, bwd_splits, col_tiles, rows_pad, kcap, merge_blocks;
  size_t off_block_partials, off_part, off_lv, off_li, off_dz, total_bytes;
};

int ffma_kcap();
FfmaPlan ffma_plan(const supcon_problem_t* p);
void ffma_bind_plan(FfmaArgs& a, const FfmaPlan& pl, void* workspace, bool backward);
size_t ffma_workspace_bytes(const supcon_problem_t* p);
cudaError_t ffma_forward(const FfmaArgs& a, cudaStream_t stream);
cudaError_t ffma_backward(const FfmaArgs& a, int dz_dtype, cudaStream_t stream);
cudaError_t ffma_topk_indices(const FfmaArgs& a, int32_t* idx_out, cudaStream_t stream);

// ---- bf16 tensor-core path (supcon_tc.cu) ----
struct TcSched {   // flattened (row block, column tile) work list cut into P contiguous CTA ranges
  int T;           // column tiles per row block
  int P;           // CTAs
  long long U;     // row_blocks * T
  // Column PANELS (backward at large N): the list is ordered (panel, row block, tile within the panel) instead of
  // (row block, tile), so that the CTAs running at the same time work on a few panels of z instead of all of it and
  // the Z_J tiles they stream stay in L2.  NP == 1: one panel = the plain row-major order.
  int RB;          // row blocks
  int NP;          // panels
  int Tp;          // tiles per panel (the last panel holds the remainder)
};
struct TcPlan {
  TcSched fwd_sched, bwd_sched;
  TcSched fwd_sched_local, fwd_sched_remote;   // two-phase forward (own columns first, then the others)
  TcSched bwd_sched_local, bwd_sched_remote;   // two-phase backward, likewise (64-column tiles)
  bool two_phase;
  int local_ct0, local_cts, slots_local;       // forward: own-column window in 128-column tiles
  int bwd_local_ct0, bwd_local_cts, bwd_slots_local;   // backward: the same window in 64-column tiles
  int n_pad, rows_pad, row_blocks, fwd_row_blocks, fwd_col_tiles, bwd_col_tiles, fwd_slots, bwd_slots, merge_blocks;
  int bwd_spp;                                 // backward: partial-record slots per column panel
  size_t off_block_partials, off_lab, off_nrm, off_colA, off_colAm, off_colB, off_colThr, off_colThrIdx,
      off_scalars, off_hkeys, off_hcounts, off_topk_v, off_topk_i, off_part, total_bytes;
  size_t off_hids, off_csum_part, off_csum;   // positives by linearity: class ids, per-block and total class sums
  size_t off_cls;                             // dense class id of every (padded) column, for the backward's reduce
  int csum_blocks;                            // blocks of tc_class_sum_kernel (256 columns each)
  uint32_t hash_size;
};
constexpr int TC_CMAX = 32;         // most classes the class-sum route handles (more: per-pair sums in the sweep)
constexpr int TC_MAX_PASSES = 4;    // passes of a multi-pass forward (own columns + up to three groups of peers)
constexpr int TC_MAX_BLOCKS = 16;   // rank blocks one pass may list
struct TcBlockList {                // column blocks a launch sweeps, in sweep order: block b covers `len` units from start[b]
  int n, len;
  int start[TC_MAX_BLOCKS];
};
struct TcFwdArgs {
  const int32_t* lab_pad;
  const float* nrm_pad;
  const unsigned long long* hkeys;   // label -> class size table (tc_prep_fwd_kernel)
  const int* hcounts;
  uint32_t hmask;
  const unsigned* nrm2_max;          // bits of max_j |z_j|^2 over the columns swept so far (workspace header)
  // positives by linearity (cosine, no mining, whole forward): sum_{j in pos(i)} z_i.z_j = z_i . C[class_i] - |z_i|^2
  // with C[c] = sum of the rows of class c.  n_classes == nullptr: off.  plin_twin: this launch is the per-pair
  // kernel that only runs when there are more than TC_CMAX classes (its twin returns at once in that case).
  const int* n_classes;              // distinct labels seen by the label table (workspace header)
  const int* hids;                   // dense class id of every table slot
  const float* csum;                 // [TC_CMAX][256] class sums
  const void* z_rows;                // z (bf16), read by the merge for the rows' dot products with the class sums
  int plin_twin;
  float* part;       // [slots][rows_pad][8]
  float* topk_v;     // [splits][rows_pad][kcap]  per-split hard-negative candidates (mining)
  int32_t* topk_i;
  TcSched sched;
  int ct_base, ex_lo, ex_len, slot_base;
  TcBlockList blocks;   // n > 0: the logical -> physical column-tile map of this launch (instead of ct_base / ex_*)
  // merge only: every pass whose partial records are summed (sched + first slot)
  int npass;
  TcSched msched[TC_MAX_PASSES];
  int mslot[TC_MAX_PASSES];
  int n_total, n_pad, row_offset, n_rows, rows_pad, topk, mine, kcap;
  float inv_tau, c1, c0, ut2;
  float m_limit;     // tau / 0.025: largest fixed maximum that cannot underflow a row's dominant terms
};
struct TcBwdPrepArgs {
  const float* stats;        // statistics of rows [stats_row0, ...): [..][STRIDE]
  int stats_row0;
  int j_lo, j_cnt;           // columns whose coefficients this launch computes
  int use_label_counts;      // |A_f|, |A_m| from the label-derived global counts (own partials, before any exchange)
  const double* partials;
  const int32_t* labels;
  float *colA, *colAm, *colB, *colThr;
  int32_t *colThrIdx, *lab_pad;
  // positives by linearity: dense class id of every column (hkeys == nullptr: not wanted)
  const unsigned long long* hkeys;
  const int* hids;
  uint32_t hmask;
  int32_t* cls_pad;
  float* scalars;
  int n_total, n_pad, topk;
  float tau, alpha, lambda_uni, uni_t;
};
struct TcBwdArgs {
  const int32_t* lab_pad;
  const float* nrm_pad;
  const float *colA, *colB;
  const float *colAm, *colThr;   // mining
  const int32_t* colThrIdx;
  const float* scalars;  // [0] = uniformity coefficient cu, [1] = exponent offset -M/tau log2(e)
  float* dz_part;        // [slots][rows_pad][256]
  TcSched sched;
  TcSched sched_b;       // reduce only: second pass of the two-phase backward (P == 0: none)
  int ct_base, ex_lo, ex_len, slot_base, slot_base_b;
  int spp;               // partial-record slots per panel of `sched`
  int n_total, n_pad, row_offset, n_rows, rows_pad;
  float c1, c0, ut2;
  // positives by linearity (cosine, no mining, single-phase backward): B is constant within a class, so
  // -sum_{j in pos(i)} (B_i + B_j) z_j = -2 B_i (C[class_i] - z_i) is added in fp32 by the reduce kernel and the
  // sweep forms H without the per-pair label compare.  n_classes == nullptr: off; plin_twin as in TcFwdArgs.
  const int* n_classes;
  const unsigned long long* hkeys;
  const int* hids;
  uint32_t hmask;
  const float* csum;     // [TC_CMAX][256]
  const int32_t* cls_pad;  // dense class id per column (tc_prep_bwd_kernel)
  int plin_twin;
};
TcPlan tc_plan(const supcon_problem_t* p);
bool tc_supported(const supcon_problem_t* p);
bool tc_two_phase(const supcon_problem_t* p);
bool tc_bwd_two_phase(const supcon_problem_t* p);
int tc_debug_plan(const supcon_problem_t* p, int32_t* out, int n_out);
int tc_debug_sched(int T, int P, long long U, int cta, int row_block, long long* range_begin, long long* range_end,
                   int* first_cta, int* last_cta);
int tc_forward(const supcon_problem_t* p, const void* z_all, const int32_t* labels_all, float* row_stats,
               double* partials, float* loss_out, void* workspace, cudaStream_t stream, const char** err,
               int phase = 0);
// multi-pass forward: pass `pass_index` sweeps the rank blocks blocks[first .. first + pass_sizes[pass_index]);
// pass 0 clears the workspace, the last pass merges all of them
int tc_forward_pass(const supcon_problem_t* p, const void* z_all, const int32_t* labels_all, const int32_t* blocks,
                    const int32_t* pass_sizes, int n_passes, int pass_index, int skip_norms, float* row_stats,
                    double* partials, void* workspace, cudaStream_t stream, const char** err);
// phase: 0 = whole backward; 1 = only the columns this rank owns (`stats` = the rank's OWN statistics
// [n_rows][STRIDE], `partials` = its own forward partials; partial dz records, nothing else is written);
// 2 = all other columns (`stats` = everyone's, `partials` = the global sums) + reduce over both phases.
int tc_backward(const supcon_problem_t* p, const void* z_all, const int32_t* labels_all, const float* stats,
                const double* partials, const float* grad_out, void* dz_out, int dz_dtype, void* workspace,
                cudaStream_t stream, const char** err, int phase = 0);

// ---- exchange through peer memory (supcon_peer.cu) ----
int peer_check(const supcon_peer_t* pe, const char** err);
cudaError_t peer_push(const supcon_peer_t& pe, const void* src0, size_t bytes0, uint64_t off0, const void* src1,
                      size_t bytes1, uint64_t off1, int flag_id, int wait_flag_id, int include_self,
                      cudaStream_t stream);
cudaError_t peer_push_ordered(const supcon_peer_t& pe, const void* src0, size_t bytes0, uint64_t off0, const void* src1,
                              size_t bytes1, uint64_t off1, int flag_id, int wait_flag_id, cudaStream_t stream);
cudaError_t peer_wait(const supcon_peer_t& pe, int flag_id, uint64_t mask, cudaStream_t stream);
cudaError_t peer_end_step(const supcon_peer_t& pe, int flag_id, cudaStream_t stream);

// ---- single-launch small-batch path (supcon_small.cu) ----
struct SmallArgs {
  const void* z;
  const int32_t* labels;
  float* row_stats;       // optional [n][STRIDE]
  double* partials;       // optional [8]
  float* loss_out;        // optional
  void* dz_out;           // optional [n][d]
  const float* grad_out;  // optional device scalar
  int n, d, z_dtype, dz_dtype, similarity, topk, mine;
  float tau, alpha, lambda_uni, uni_t;
};
bool small_supported(const supcon_problem_t* p, const void* z);
cudaError_t small_launch(const SmallArgs& a, cudaStream_t stream);

// ---- single-launch mid-size path, 160 < N <= 512 (supcon_mid.cu); *taken = false when the device cannot
//      co-schedule the cluster (the caller then uses the tiled kernels) ----
bool mid_supported(const supcon_problem_t* p, const void* z);
cudaError_t mid_launch(const SmallArgs& a, cudaStream_t stream, bool* taken);

// ---- producer of z: compression head up to the Linear layer, fused with the time mean (supcon_head.cu) ----
struct HeadPoolArgs {
  const float* hs;                      // [B][K][F][T]
  float* pooled;                        // fwd out [B][F]
  const unsigned long long* rng_state;  // device {seed, offset}; NULL = no dropout
  int B, K, F, T;
  float dropout_p, negative_slope;
  int rows_per_block;                   // filled by the launcher
};
int head_pool_rows_per_block(int F, int T);
cudaError_t head_pool_forward(HeadPoolArgs a, cudaStream_t stream);
cudaError_t head_pool_backward(HeadPoolArgs a, const float* dpooled, float* dhs, cudaStream_t stream);

// tcgen05 building-block diagnostic (supcon_tc_debug.cu)
int tc_debug_tile(const void* z_bf16, int n, int d, int row_i, int row_j, float* s_out, float* o_out,
                  cudaStream_t stream, const char** err);

}  // namespace supcon
