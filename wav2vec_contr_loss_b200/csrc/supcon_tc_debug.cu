// Diagnostic kernel for the tcgen05 building blocks (used by tests/test_gpu_tc.py):
// one CTA computes, for a 128-row block I and a 128-row block J of a bf16 matrix z,
//   S  = Z_I . Z_J^T                       (K-major A and B from TMA-staged smem)
//   O  = bf16(S) . Z_J                     (A written to swizzled smem by threads,
//                                           B = the same Z_J tile used MN-major)
// which exercises every descriptor form the fused kernels rely on.
#include "supcon_common.cuh"
#include "supcon_internal.h"
#include "tc_ptx.cuh"
#include "tc_tmap.cuh"

namespace supcon {
namespace {

constexpr int TD = 256;  // d
constexpr int TR = 128;  // rows per tile
constexpr uint32_t BOX_BYTES = TR * 128;         // one 128-row x 64-col bf16 box
constexpr uint32_t TILE_BYTES = 4 * BOX_BYTES;   // 128 x 256 bf16

__global__ void __launch_bounds__(128) tc_debug_kernel(const __grid_constant__ CUtensorMap tmap, int row_i, int row_j,
                                                       float* __restrict__ s_out, float* __restrict__ o_out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // SW128 atoms: 1024-B aligned
  unsigned char* sA = smem;                  // Z_I
  unsigned char* sB = smem + TILE_BYTES;     // Z_J
  unsigned char* sH = smem + 2 * TILE_BYTES; // bf16(S): 2 K-blocks of 128 rows x 128 B
  __shared__ __align__(8) uint64_t bar_tma, bar_mma1, bar_mma2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    ptx::mbar_init(&bar_tma, 1);
    ptx::mbar_init(&bar_mma1, 1);
    ptx::mbar_init(&bar_mma2, 1);
    ptx::fence_mbar_init();
    ptx::tma_prefetch_desc(&tmap);
  }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_base_s);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    ptx::mbar_expect_tx(&bar_tma, 2 * TILE_BYTES);
    for (int b = 0; b < 4; ++b) {
      ptx::tma_load_2d(sA + b * BOX_BYTES, &tmap, &bar_tma, 64 * b, row_i);
      ptx::tma_load_2d(sB + b * BOX_BYTES, &tmap, &bar_tma, 64 * b, row_j);
    }
  }
  ptx::mbar_wait(&bar_tma, 0);

  if (tid == 0) {
    ptx::tc_fence_after_sync();
    constexpr uint32_t idesc = ptx::idesc_bf16(128, 128, false, false);
    const uint32_t a0 = ptx::smem_u32(sA), b0 = ptx::smem_u32(sB);
#pragma unroll
    for (int ks = 0; ks < TD / 16; ++ks) {
      uint32_t off = (ks >> 2) * BOX_BYTES + (ks & 3) * 32;
      ptx::mma_ss(tmem, ptx::smem_desc_sw128(a0 + off, 16, 1024), ptx::smem_desc_sw128(b0 + off, 16, 1024), idesc,
                  ks > 0);
    }
    ptx::mma_commit(&bar_mma1);
  }
  ptx::mbar_wait(&bar_mma1, 0);
  ptx::tc_fence_after_sync();

  const int row = 32 * warp + lane;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 32 * c, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) s_out[row * 128 + 32 * c + e] = __uint_as_float(r[e]);
    // bf16 copy into the K-major swizzled A tile: 16-byte chunk q (8 columns) of row `row`
    const int kb = (32 * c) / 64;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 p = __floats2bfloat162_rn(__uint_as_float(r[8 * q + 2 * e]), __uint_as_float(r[8 * q + 2 * e + 1]));
        w[e] = *reinterpret_cast<uint32_t*>(&p);
      }
      int chunk = ((32 * c) % 64) / 8 + q;
      uint32_t addr = ptx::smem_u32(sH) + kb * BOX_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                   : "memory");
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();

  if (tid == 0) {
    ptx::tc_fence_after_sync();
    constexpr uint32_t idesc2 = ptx::idesc_bf16(128, 256, false, true);
    const uint32_t h0 = ptx::smem_u32(sH), b0 = ptx::smem_u32(sB);
#pragma unroll
    for (int kk = 0; kk < 128 / 16; ++kk) {
      uint64_t da = ptx::smem_desc_sw128(h0 + (kk >> 2) * BOX_BYTES + (kk & 3) * 32, 16, 1024);
      uint64_t db = ptx::smem_desc_sw128(b0 + kk * 16 * 128, BOX_BYTES, 1024);  // MN-major: LBO = next 64 columns
      ptx::mma_ss(tmem + 128, da, db, idesc2, kk > 0);
    }
    ptx::mma_commit(&bar_mma2);
  }
  ptx::mbar_wait(&bar_mma2, 0);
  ptx::tc_fence_after_sync();
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 128 + 32 * c, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) o_out[row * 256 + 32 * c + e] = __uint_as_float(r[e]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}

// Same two products with the A operands in tensor memory (tcgen05.mma [d], [a], b-desc):
//   Z_I rows are written to TMEM columns [256,384) from registers (packed bf16 pairs),
//   bf16(S) to columns [384,448); S accumulates in [0,128), O in [128,... wait: O needs 256
//   columns, so the layout is  O [0,256)  |  A = Z_I [256,384)  |  S [384,512) with bf16(S)
//   aliased onto the first 64 columns of S after it has been read out.
__global__ void __launch_bounds__(128) tc_debug_ts_kernel(const __grid_constant__ CUtensorMap tmap,
                                                          const __nv_bfloat16* __restrict__ z, int n, int row_i,
                                                          int row_j, float* __restrict__ s_out,
                                                          float* __restrict__ o_out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sB = smem;  // Z_J
  __shared__ __align__(8) uint64_t bar_tma, bar_mma1, bar_mma2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TM_O = 0, TM_A = 256, TM_S = 384;

  if (tid == 0) {
    ptx::mbar_init(&bar_tma, 1);
    ptx::mbar_init(&bar_mma1, 1);
    ptx::mbar_init(&bar_mma2, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_base_s);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = (uint32_t)(32 * warp) << 16;
  const int row = 32 * warp + lane;

  if (tid == 0) {
    ptx::mbar_expect_tx(&bar_tma, TILE_BYTES);
    for (int b = 0; b < 4; ++b) ptx::tma_load_2d(sB + b * BOX_BYTES, &tmap, &bar_tma, 64 * b, row_j);
  }
  // Z_I row -> TMEM: 256 bf16 = 128 packed words, 4 stores of 32 columns
  {
    const int gi = row_i + row;
    const uint4* src = reinterpret_cast<const uint4*>(z + (int64_t)min(gi, n - 1) * TD);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t w[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 v = (gi < n) ? __ldg(src + 8 * c + q) : make_uint4(0, 0, 0, 0);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      ptx::tmem_st32(tmem + lane_addr + TM_A + 32 * c, w);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::mbar_wait(&bar_tma, 0);

  if (tid == 0) {
    ptx::tc_fence_after_sync();
    constexpr uint32_t idesc = ptx::idesc_bf16(128, 128, false, false);
    const uint32_t b0 = ptx::smem_u32(sB);
#pragma unroll
    for (int ks = 0; ks < TD / 16; ++ks) {
      uint32_t off = (ks >> 2) * BOX_BYTES + (ks & 3) * 32;
      ptx::mma_ts(tmem + TM_S, tmem + TM_A + 8 * ks, ptx::smem_desc_sw128(b0 + off, 16, 1024), idesc, ks > 0);
    }
    ptx::mma_commit(&bar_mma1);
  }
  ptx::mbar_wait(&bar_mma1, 0);
  ptx::tc_fence_after_sync();

  uint32_t hw[64];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + lane_addr + TM_S + 32 * c, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) s_out[row * 128 + 32 * c + e] = __uint_as_float(r[e]);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      __nv_bfloat162 p = __floats2bfloat162_rn(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
      hw[16 * c + e] = *reinterpret_cast<uint32_t*>(&p);
    }
  }
  // bf16(S) back into TMEM over the first 64 columns of S
  {
    uint32_t w0[32], w1[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) { w0[e] = hw[e]; w1[e] = hw[32 + e]; }
    ptx::tmem_st32(tmem + lane_addr + TM_S, w0);
    ptx::tmem_st32(tmem + lane_addr + TM_S + 32, w1);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();

  if (tid == 0) {
    ptx::tc_fence_after_sync();
    constexpr uint32_t idesc2 = ptx::idesc_bf16(128, 256, false, true);
    const uint32_t b0 = ptx::smem_u32(sB);
#pragma unroll
    for (int kk = 0; kk < 128 / 16; ++kk) {
      uint64_t db = ptx::smem_desc_sw128(b0 + kk * 16 * 128, BOX_BYTES, 1024);
      ptx::mma_ts(tmem + TM_O, tmem + TM_S + 8 * kk, db, idesc2, kk > 0);
    }
    ptx::mma_commit(&bar_mma2);
  }
  ptx::mbar_wait(&bar_mma2, 0);
  ptx::tc_fence_after_sync();
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + lane_addr + TM_O + 32 * c, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) o_out[row * 256 + 32 * c + e] = __uint_as_float(r[e]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}

}  // namespace

int tc_debug_tile(const void* z_bf16, int n, int d, int row_i, int row_j, float* s_out, float* o_out,
                  cudaStream_t stream, const char** err) {
  const bool ts_mode = row_i >= (1 << 24);   // test hook: high bit selects the A-in-TMEM variant
  if (ts_mode) row_i -= (1 << 24);
  if (d != TD) { *err = "tc_debug_tile needs d == 256"; return SUPCON_E_UNSUPPORTED; }
  if (row_i < 0 || row_j < 0 || row_i >= n || row_j >= n) { *err = "row out of range"; return SUPCON_E_INVALID; }
  CUtensorMap tmap;
  if (make_bf16_rowmajor_tmap(&tmap, z_bf16, (uint64_t)n, (uint64_t)d, TR) != 0) {
    *err = "cuTensorMapEncodeTiled failed";
    return SUPCON_E_INVALID;
  }
  const size_t smem = 2 * TILE_BYTES + 2 * BOX_BYTES + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *err = cudaGetErrorString(e); return (int)e; }
  if (ts_mode) {
    e = cudaFuncSetAttribute(tc_debug_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return (int)e; }
    tc_debug_ts_kernel<<<1, 128, smem, stream>>>(tmap, reinterpret_cast<const __nv_bfloat16*>(z_bf16), n, row_i, row_j,
                                                 s_out, o_out);
  } else {
    tc_debug_kernel<<<1, 128, smem, stream>>>(tmap, row_i, row_j, s_out, o_out);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { *err = cudaGetErrorString(e); return (int)e; }
  return 0;
}

}  // namespace supcon
