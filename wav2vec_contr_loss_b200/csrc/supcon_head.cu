// Producer of the loss's input: the compression head up to the Linear layer, fused with the time mean.
//
// Reference (compression_module.py:48-65, stage1_utils.py:122-123):
//     x   = LeakyReLU(Dropout(mean_k hs[b,k,f,t]))            (B, F, T)
//     seq = Linear(x^T)^T                                      (B, D, T)
//     z   = normalize(mean_t seq)                              (B, D)
// The Linear layer and the time mean are both linear, so  mean_t seq = W * (mean_t x) + bias : the per-frame
// GEMM over (B*T, F) collapses to one over (B, F), and everything left of it is ONE pass over hs:
//     pooled[b,f] = (1/T) sum_t LeakyReLU(Dropout((1/K) sum_k hs[b,k,f,t]))
// This file is that pass (forward) and its transpose (backward, only needed when the encoder is fine-tuned).
// HBM-bound: the algorithmic traffic is hs read once (B*K*F*T*4 bytes; 1.3 GB at B = 64, K = 25, F = 1024, T = 199),
// the output is B*F*4 bytes.
//
// Layout: hs is (B, K, F, T) contiguous, T innermost.  For one b the (f, t) plane is contiguous, so a block takes
// `rows` consecutive feature rows = one contiguous run of rows*T floats per layer k (stride F*T between layers),
// reads it with 128-bit loads when the run is 16-byte aligned, keeps the activated values in shared memory and
// reduces each row over t in a fixed order (deterministic, no atomics).
//
// Dropout: counter-based Philox4x32-10 keyed by a (seed, offset) pair read from DEVICE memory, so a captured CUDA
// graph draws a new mask every replay (the caller bumps the offset in-stream).  Element (b,f,t) uses word
// (i & 3) of the block i >> 2, i = (b*F + f)*T + t; the backward regenerates the same mask.
#include "supcon_common.cuh"
#include "supcon_internal.h"

namespace supcon {
namespace {

constexpr int HP_THREADS = 256;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// multiplier applied by Dropout to element i: 0 or 1/(1-p)
__device__ __forceinline__ float dropout_scale(unsigned long long seed, unsigned long long offset, long long i,
                                               float p, float keep_scale) {
  const unsigned long long blk = (unsigned long long)i >> 2;
  const uint4 r = philox4x32_10(make_uint4((unsigned)blk, (unsigned)(blk >> 32), (unsigned)offset,
                                           (unsigned)(offset >> 32)),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const int w = (int)(i & 3);
  const unsigned word = w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
  const float u = (float)(word >> 8) * (1.0f / 16777216.0f);   // [0, 1)
  return u >= p ? keep_scale : 0.0f;
}

struct RunView {            // the contiguous run of one block
  const float* base;        // hs + ((b*K)*F + f0)*T
  long long plane;          // F*T
  long long first_elem;     // (b*F + f0)*T : index of the run's first element in the (B,F,T) mask space
  int rows, len, f0, b;
};

__device__ __forceinline__ RunView run_of_block(const HeadPoolArgs& a) {
  RunView v;
  v.b = blockIdx.y;
  v.f0 = blockIdx.x * a.rows_per_block;
  v.rows = min(a.rows_per_block, a.F - v.f0);
  v.len = v.rows * a.T;
  v.plane = (long long)a.F * a.T;
  v.base = a.hs + ((long long)v.b * a.K) * v.plane + (long long)v.f0 * a.T;
  v.first_elem = ((long long)v.b * a.F + v.f0) * a.T;
  return v;
}

template <int VEC>
struct Pack;
template <>
struct Pack<4> { typedef float4 type; };
template <>
struct Pack<1> { typedef float type; };

template <int VEC>
__device__ __forceinline__ void add_pack(float (&acc)[VEC], const typename Pack<VEC>::type& x) {
  if constexpr (VEC == 4) { acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w; }
  else acc[0] += x;
}

// sum over the K layers of VEC consecutive elements starting at run offset e, layers added in order k = 0..K-1;
// loads are issued eight layers at a time so that each thread keeps 8 x 16 bytes in flight
template <int VEC>
__device__ __forceinline__ void layer_sum(const RunView& v, int K, int e, float (&acc)[VEC]) {
  typedef typename Pack<VEC>::type pack_t;
  constexpr int BATCH = 8;
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  const float* p = v.base + e;
  int k = 0;
  for (; k + BATCH <= K; k += BATCH) {
    pack_t x[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j) x[j] = __ldg(reinterpret_cast<const pack_t*>(p + (long long)(k + j) * v.plane));
#pragma unroll
    for (int j = 0; j < BATCH; ++j) add_pack<VEC>(acc, x[j]);
  }
  for (; k < K; ++k) add_pack<VEC>(acc, __ldg(reinterpret_cast<const pack_t*>(p + (long long)k * v.plane)));
}

template <int VEC>
__global__ void __launch_bounds__(HP_THREADS, 4) head_pool_fwd_kernel(HeadPoolArgs a) {
  extern __shared__ float act[];                      // [rows*T] activated values of this block's run
  const RunView v = run_of_block(a);
  const bool drop = a.dropout_p > 0.f && a.rng_state != nullptr;
  unsigned long long seed = 0, offset = 0;
  if (drop) { seed = a.rng_state[0]; offset = a.rng_state[1]; }
  const float inv_k = 1.0f / (float)a.K;
  const float keep_scale = drop ? 1.0f / (1.0f - a.dropout_p) : 1.0f;

  for (int e = threadIdx.x * VEC; e < v.len; e += HP_THREADS * VEC) {
    float acc[VEC];
    layer_sum<VEC>(v, a.K, e, acc);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float x = acc[i] * inv_k;
      if (drop) x *= dropout_scale(seed, offset, v.first_elem + e + i, a.dropout_p, keep_scale);
      act[e + i] = x > 0.f ? x : x * a.negative_slope;
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_t = 1.0f / (float)a.T;
  for (int r = warp; r < v.rows; r += HP_THREADS / 32) {
    float s = 0.f;
    for (int t = lane; t < a.T; t += 32) s += act[r * a.T + t];
    s = warp_sum(s);
    if (lane == 0) a.pooled[(long long)v.b * a.F + v.f0 + r] = s * inv_t;
  }
}

// dhs[b,k,f,t] = dpooled[b,f] * (1/T) * LeakyReLU'(x) * dropout multiplier * (1/K), the same for every k
template <int VEC>
__global__ void __launch_bounds__(HP_THREADS, 4) head_pool_bwd_kernel(HeadPoolArgs a, const float* __restrict__ dpooled,
                                                                   float* __restrict__ dhs) {
  const RunView v = run_of_block(a);
  const bool drop = a.dropout_p > 0.f && a.rng_state != nullptr;
  unsigned long long seed = 0, offset = 0;
  if (drop) { seed = a.rng_state[0]; offset = a.rng_state[1]; }
  const float inv_k = 1.0f / (float)a.K;
  const float keep_scale = drop ? 1.0f / (1.0f - a.dropout_p) : 1.0f;
  const float inv_kt = inv_k / (float)a.T;
  float* out = dhs + (v.base - a.hs);

  for (int e = threadIdx.x * VEC; e < v.len; e += HP_THREADS * VEC) {
    float acc[VEC], g[VEC];
    layer_sum<VEC>(v, a.K, e, acc);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float m = drop ? dropout_scale(seed, offset, v.first_elem + e + i, a.dropout_p, keep_scale) : 1.0f;
      const float x = acc[i] * inv_k * m;
      const float up = dpooled[(long long)v.b * a.F + v.f0 + (e + i) / a.T];
      g[i] = up * inv_kt * m * (x > 0.f ? 1.0f : a.negative_slope);
    }
    float* q = out + e;
    for (int k = 0; k < a.K; ++k) {
      if constexpr (VEC == 4)
        *reinterpret_cast<float4*>(q + (long long)k * v.plane) = make_float4(g[0], g[1], g[2], g[3]);
      else
        q[(long long)k * v.plane] = g[0];
    }
  }
}

// 128-bit accesses need every run (and every layer plane) to start on a 16-byte boundary and to hold a multiple of
// four floats - including the shorter run of the last block of a row of blocks.
bool vector_ok(const HeadPoolArgs& a, const void* other) {
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const long long run = (long long)a.rows_per_block * a.T;
  const long long tail = (long long)(a.F % a.rows_per_block) * a.T;
  return aligned(a.hs) && (other == nullptr || aligned(other)) && ((long long)a.F * a.T) % 4 == 0 && run % 4 == 0 &&
         tail % 4 == 0;
}

}  // namespace

int head_pool_rows_per_block(int F, int T) {
  // 16 rows per block when they fit 48 KB of shared memory (T <= 768), fewer for longer clips; 0 = unsupported
  if (T < 1 || F < 1) return 0;
  int rows = 12288 / T;
  if (rows < 1) return 0;
  if (rows > 16) rows = 16;
  if (rows > 4) rows &= ~3;           // keep runs a multiple of four floats for any T
  return rows < F ? rows : F;
}

cudaError_t head_pool_forward(HeadPoolArgs a, cudaStream_t stream) {
  a.rows_per_block = head_pool_rows_per_block(a.F, a.T);
  if (a.rows_per_block < 1) return cudaErrorInvalidValue;
  dim3 grid((a.F + a.rows_per_block - 1) / a.rows_per_block, a.B);
  const size_t smem = (size_t)a.rows_per_block * a.T * sizeof(float);
  if (vector_ok(a, nullptr))
    head_pool_fwd_kernel<4><<<grid, HP_THREADS, smem, stream>>>(a);
  else
    head_pool_fwd_kernel<1><<<grid, HP_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t head_pool_backward(HeadPoolArgs a, const float* dpooled, float* dhs, cudaStream_t stream) {
  a.rows_per_block = head_pool_rows_per_block(a.F, a.T);
  if (a.rows_per_block < 1) return cudaErrorInvalidValue;
  dim3 grid((a.F + a.rows_per_block - 1) / a.rows_per_block, a.B);
  if (vector_ok(a, dhs))
    head_pool_bwd_kernel<4><<<grid, HP_THREADS, 0, stream>>>(a, dpooled, dhs);
  else
    head_pool_bwd_kernel<1><<<grid, HP_THREADS, 0, stream>>>(a, dpooled, dhs);
  return cudaGetLastError();
}

}  // namespace supcon
