// Single-launch SupCon forward + backward for small batches (the repo's native
// batch of 64; N <= 160).  At these sizes the path is latency-bound: HBM time
// for z is ~30 ns and the math ~4 MFLOP, so the cost is launches and dependent
// phases.  One thread-block cluster of 8 CTAs does everything in one launch:
//
//   A  every CTA stages all of z (fp32) in shared memory; CTA c owns rows
//      [c R, (c+1) R), R = ceil(N/8), and computes their dot products (FFMA)
//   B  one warp per owned row: similarity, online-free exact softmax stats,
//      hard-negative selection by K rounds of warp arg-max (value desc, index asc)
//   C  cluster barrier; row statistics and partial sums of all CTAs are read
//      through distributed shared memory (no global round trip, no second launch)
//   D  H rows = G + G^T from the gathered statistics, dz rows = H z (FFMA)
//
// Results follow the same formulas as supcon_ffma.cu and
// are written in the same row-statistics format.
#include <cooperative_groups.h>

#include "supcon_common.cuh"
#include "supcon_internal.h"

namespace cg = cooperative_groups;

namespace supcon {
namespace {

constexpr int SC = 8;      // CTAs per cluster
constexpr int SNT = 256;   // threads per CTA
constexpr int SMAXR = 20;  // owned rows per CTA (N <= 160)
constexpr int SMAXCOLS = 5;  // columns per lane in phase B (N <= 160)

struct SmallLayout {
  int ld;        // padded row stride of z in smem (floats)
  int np;        // N rounded up to a multiple of 32
  int groups;    // k-split groups in phase A
  size_t off_z, off_c, off_h, off_ps, off_nrm, off_lab, off_stats_local, off_stats_all, off_part, total;
};

__host__ __device__ inline SmallLayout small_layout(int n, int d, int rows) {
  SmallLayout L;
  L.ld = d + 4;
  L.np = (n + 31) / 32 * 32;
  L.groups = SNT / L.np;
  if (L.groups < 1) L.groups = 1;
  if (L.groups > 4) L.groups = 4;
  size_t o = 0;
  L.off_z = o; o += (size_t)n * L.ld * 4;
  L.off_c = o; o += (size_t)rows * L.np * 4;
  L.off_h = o; o += (size_t)rows * L.np * 4;
  L.off_ps = o; o += (size_t)L.groups * rows * L.np * 4;
  L.off_nrm = o; o += (size_t)L.np * 4;
  L.off_lab = o; o += (size_t)L.np * 4;
  L.off_stats_local = o; o += (size_t)rows * SUPCON_STATS_STRIDE * 4;
  L.off_stats_all = o; o += (size_t)L.np * SUPCON_STATS_STRIDE * 4;
  o = (o + 7) & ~(size_t)7;
  L.off_part = o; o += SUPCON_N_PARTIALS * sizeof(double);
  L.total = o;
  return L;
}

template <typename T, typename TO>
__global__ void __cluster_dims__(SC, 1, 1) __launch_bounds__(SNT, 1) small_kernel(SmallArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int n = a.n, d = a.d;
  const int R = (n + SC - 1) / SC;
  const SmallLayout L = small_layout(n, d, R);
  float* zs = reinterpret_cast<float*>(smem + L.off_z);
  float* cs = reinterpret_cast<float*>(smem + L.off_c);      // dot products of owned rows  [R][np]
  float* hs = reinterpret_cast<float*>(smem + L.off_h);      // H rows                       [R][np]
  float* ps = reinterpret_cast<float*>(smem + L.off_ps);     // k-split partial dots         [G][R][np]
  float* nrm = reinterpret_cast<float*>(smem + L.off_nrm);
  int* lab = reinterpret_cast<int*>(smem + L.off_lab);
  float* st_local = reinterpret_cast<float*>(smem + L.off_stats_local);
  float* st_all = reinterpret_cast<float*>(smem + L.off_stats_all);
  double* part = reinterpret_cast<double*>(smem + L.off_part);

  const T* z = reinterpret_cast<const T*>(a.z);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = rank * R;                       // first owned row
  const int nrows = max(0, min(R, n - r0));      // owned rows in range
  const bool geo = a.similarity == SUPCON_GEODESIC;
  const bool uni = a.lambda_uni > 0.f;
  const bool mine = a.mine != 0;

  // ---- A0: stage z (fp32) and labels ----
  for (int idx = tid; idx < n * (d / 4); idx += SNT) {
    const int r = idx / (d / 4), q = idx % (d / 4);
    float4 v = ld_row4<T>(z, r, true, 4 * q, d, true);
    *reinterpret_cast<float4*>(&zs[r * L.ld + 4 * q]) = v;
  }
  for (int j = tid; j < L.np; j += SNT) lab[j] = j < n ? a.labels[j] : 0;
  __syncthreads();

  // ---- A1: squared norms of all rows (uniformity) ----
  if (uni) {
    for (int j = tid; j < n; j += SNT) {
      float s = 0.f;
      for (int k = 0; k < d; ++k) s = fmaf(zs[j * L.ld + k], zs[j * L.ld + k], s);
      nrm[j] = s;
    }
  }

  // ---- A2: dot products c[r][j] of the owned rows; thread = (column j, k-slice g) ----
  {
    const int j = tid % L.np, g = tid / L.np;
    if (g < L.groups) {
      const int klen = (d / 4 + L.groups - 1) / L.groups * 4;  // multiple of 4
      const int k_begin = g * klen, k_end = min(d, k_begin + klen);
      float acc[SMAXR];
#pragma unroll
      for (int r = 0; r < SMAXR; ++r) acc[r] = 0.f;
      if (j < n) {
        for (int k = k_begin; k < k_end; k += 4) {
          const float4 b = *reinterpret_cast<const float4*>(&zs[j * L.ld + k]);
#pragma unroll
          for (int r = 0; r < SMAXR; ++r) {
            if (r < nrows) {
              const float4 av = *reinterpret_cast<const float4*>(&zs[(r0 + r) * L.ld + k]);
              float c = acc[r];
              c = fmaf(av.x, b.x, c); c = fmaf(av.y, b.y, c); c = fmaf(av.z, b.z, c); c = fmaf(av.w, b.w, c);
              acc[r] = c;
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < SMAXR; ++r)
        if (r < R) ps[(g * R + r) * L.np + j] = acc[r];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < R * L.np; idx += SNT) {  // fixed-order combine: symmetric in (i, j)
    float c = ps[idx];
    for (int g = 1; g < L.groups; ++g) c += ps[g * R * L.np + idx];
    cs[idx] = c;
  }
  __syncthreads();

  // ---- B: row statistics, one warp per owned row ----
  double acc_full = 0.0, acc_cf = 0.0, acc_mined = 0.0, acc_cm = 0.0, acc_w = 0.0;  // lane 0 of each warp
  for (int r = warp; r < nrows; r += SNT / 32) {
    const int gi = r0 + r;
    const int lab_r = lab[gi];
    float sv[SMAXCOLS];
    float m = -INFINITY;
#pragma unroll
    for (int q = 0; q < SMAXCOLS; ++q) {
      const int j = lane + 32 * q;
      sv[q] = -INFINITY;
      if (j < n && j != gi) {
        const float c = cs[r * L.np + j];
        sv[q] = geo ? geodesic_sim(c) : c;
        m = fmaxf(m, __fdiv_rn(sv[q], a.tau));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum_all = 0.f, sum_pos_e = 0.f, sum_pos_s = 0.f, wsum = 0.f;
    int npos = 0, nneg = 0;
    unsigned negmask = 0;  // my columns that are negatives
#pragma unroll
    for (int q = 0; q < SMAXCOLS; ++q) {
      const int j = lane + 32 * q;
      if (j < n && j != gi) {
        const float e = expf(__fdiv_rn(sv[q], a.tau) - m);
        sum_all += e;
        if (lab[j] == lab_r) { npos++; sum_pos_e += e; sum_pos_s += sv[q]; }
        else { nneg++; negmask |= 1u << q; }
        if (uni) {
          const float d2 = fmaxf(nrm[gi] + nrm[j] - 2.f * cs[r * L.np + j], 0.f);
          wsum += expf(-a.uni_t * d2);
        }
      }
    }
    sum_all = warp_sum(sum_all); sum_pos_e = warp_sum(sum_pos_e); sum_pos_s = warp_sum(sum_pos_s);
    wsum = warp_sum(wsum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      npos += __shfl_xor_sync(0xffffffffu, npos, o);
      nneg += __shfl_xor_sync(0xffffffffu, nneg, o);
    }
    const float lse = m + logf(sum_all);
    float lse_m = lse, thr_val = -INFINITY;
    int thr_idx = SUPCON_INT_MAX;
    if (mine && nneg > a.topk) {
      // K rounds of warp arg-max over the not-yet-selected negatives, order (value desc, index asc)
      float sum_top = 0.f;
      unsigned avail = negmask;
      for (int round = 0; round < a.topk; ++round) {
        float bv = -INFINITY;
        int bj = SUPCON_INT_MAX;
#pragma unroll
        for (int q = 0; q < SMAXCOLS; ++q)
          if ((avail >> q) & 1u) {
            const int j = lane + 32 * q;
            if (sv[q] > bv || (sv[q] == bv && j < bj)) { bv = sv[q]; bj = j; }
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
          if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
        }
        if ((bj & 31) == lane) avail &= ~(1u << (bj >> 5));
        sum_top += expf(__fdiv_rn(bv, a.tau) - m);
        thr_val = bv; thr_idx = bj;
      }
      lse_m = m + logf(sum_pos_e + sum_top);
    } else if (a.topk < 1) {
      thr_val = INFINITY; thr_idx = -1;
    }
    if (lane == 0) {
      const float pos_mean = npos > 0 ? __fdiv_rn(__fdiv_rn(sum_pos_s, a.tau), (float)npos) : 0.f;
      float* so = st_local + r * SUPCON_STATS_STRIDE;
      so[SUPCON_ST_LSE] = lse; so[SUPCON_ST_LSE_M] = lse_m;
      reinterpret_cast<int*>(so)[SUPCON_ST_NPOS] = npos; reinterpret_cast<int*>(so)[SUPCON_ST_NNEG] = nneg;
      so[SUPCON_ST_THR_VAL] = thr_val; reinterpret_cast<int*>(so)[SUPCON_ST_THR_IDX] = thr_idx;
      so[SUPCON_ST_WSUM] = wsum; so[SUPCON_ST_POS_MEAN] = pos_mean;
      if (a.row_stats) {
        float* go = a.row_stats + (int64_t)gi * SUPCON_STATS_STRIDE;
#pragma unroll
        for (int w = 0; w < SUPCON_STATS_STRIDE; ++w) go[w] = so[w];
      }
      if (npos > 0) {
        acc_full += (double)(lse - pos_mean); acc_cf += 1.0;
        if (nneg > 0 && a.topk >= 1) { acc_mined += (double)(lse_m - pos_mean); acc_cm += 1.0; }
      }
      acc_w += (double)wsum;
    }
  }
  // per-CTA partial sums, fixed order over warps
  {
    __shared__ double wred[SNT / 32][5];
    if (lane == 0) {
      wred[warp][0] = acc_full; wred[warp][1] = acc_cf; wred[warp][2] = acc_mined; wred[warp][3] = acc_cm;
      wred[warp][4] = acc_w;
    }
    __syncthreads();
    if (tid < SUPCON_N_PARTIALS) {
      double s = 0.0;
      if (tid < 5)
        for (int w = 0; w < SNT / 32; ++w) s += wred[w][tid];
      part[tid] = s;
    }
  }
  cluster.sync();  // C: statistics and partials of every CTA are complete

  // ---- gather through distributed shared memory ----
  for (int idx = tid; idx < n * SUPCON_STATS_STRIDE; idx += SNT) {
    const int i = idx / SUPCON_STATS_STRIDE, w = idx % SUPCON_STATS_STRIDE;
    const float* remote = cluster.map_shared_rank(st_local, i / R);
    st_all[idx] = remote[(i % R) * SUPCON_STATS_STRIDE + w];
  }
  __shared__ double gpart[SUPCON_N_PARTIALS];
  if (tid < SUPCON_N_PARTIALS) {
    double s = 0.0;
    for (int c = 0; c < SC; ++c) s += cluster.map_shared_rank(part, c)[tid];
    gpart[tid] = s;
  }
  __syncthreads();
  cluster.sync();  // nobody may leave (or reuse st_local/part) while peers still read them
  const GlobalCoef g = global_coef(gpart, n, a.tau, a.alpha, a.lambda_uni, a.uni_t);
  if (rank == 0 && tid == 0) {
    if (a.loss_out) *a.loss_out = g.loss;
    if (a.partials)
      for (int w = 0; w < SUPCON_N_PARTIALS; ++w) a.partials[w] = gpart[w];
  }
  if (!a.dz_out) return;

  // ---- D1: H rows ----
  const bool mining = g.a_mined != 0.f;
  for (int idx = tid; idx < nrows * L.np; idx += SNT) {
    const int r = idx / L.np, j = idx % L.np;
    const int gi = r0 + r;
    float h = 0.f;
    if (j < n && j != gi) {
      const float* si = st_all + gi * SUPCON_STATS_STRIDE;
      const float* sj = st_all + j * SUPCON_STATS_STRIDE;
      const int npos_i = reinterpret_cast<const int*>(si)[SUPCON_ST_NPOS], nneg_i = reinterpret_cast<const int*>(si)[SUPCON_ST_NNEG];
      const int npos_j = reinterpret_cast<const int*>(sj)[SUPCON_ST_NPOS], nneg_j = reinterpret_cast<const int*>(sj)[SUPCON_ST_NNEG];
      const float af_i = npos_i > 0 ? g.a_full : 0.f, af_j = npos_j > 0 ? g.a_full : 0.f;
      const float am_i = (npos_i > 0 && nneg_i > 0 && a.topk >= 1) ? g.a_mined : 0.f;
      const float am_j = (npos_j > 0 && nneg_j > 0 && a.topk >= 1) ? g.a_mined : 0.f;
      const float cv = cs[r * L.np + j];
      const float s = geo ? geodesic_sim(cv) : cv;
      const float lg = __fdiv_rn(s, a.tau);
      const float e_r = expf(lg - si[SUPCON_ST_LSE]), e_c = expf(lg - sj[SUPCON_ST_LSE]);
      const bool pos = lab[gi] == lab[j];
      h = af_i * e_r + af_j * e_c;
      if (mining) {
        const float thr_i = si[SUPCON_ST_THR_VAL], thr_j = sj[SUPCON_ST_THR_VAL];
        const int ti = reinterpret_cast<const int*>(si)[SUPCON_ST_THR_IDX], tj = reinterpret_cast<const int*>(sj)[SUPCON_ST_THR_IDX];
        const bool mem_r = pos || s > thr_i || (s == thr_i && j <= ti);
        const bool mem_c = pos || s > thr_j || (s == thr_j && gi <= tj);
        if (mem_r && am_i != 0.f) h = fmaf(am_i * expf(si[SUPCON_ST_LSE] - si[SUPCON_ST_LSE_M]), e_r, h);
        if (mem_c && am_j != 0.f) h = fmaf(am_j * expf(sj[SUPCON_ST_LSE] - sj[SUPCON_ST_LSE_M]), e_c, h);
      }
      if (pos) {
        const float bp_i = npos_i > 0 ? __fdiv_rn(af_i + am_i, (float)npos_i) : 0.f;
        const float bp_j = npos_j > 0 ? __fdiv_rn(af_j + am_j, (float)npos_j) : 0.f;
        h -= bp_i + bp_j;
      }
      if (geo) h *= geodesic_slope_exact(cv);
      if (g.cu != 0.f) {
        const float d2 = fmaxf(nrm[gi] + nrm[j] - 2.f * cv, 0.f);
        h = fmaf(-g.cu, expf(-a.uni_t * d2), h);
      }
    }
    hs[idx] = h;
  }
  __syncthreads();

  // ---- D2: dz rows = H z (+ uniformity diagonal), thread = output column ----
  const float gscale = a.grad_out ? *a.grad_out : 1.0f;
  TO* out = reinterpret_cast<TO*>(a.dz_out);
  for (int dd = tid; dd < d; dd += SNT) {
    float acc[SMAXR];
#pragma unroll
    for (int r = 0; r < SMAXR; ++r) acc[r] = 0.f;
    for (int j = 0; j < n; j += 4) {
      float zv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) zv[u] = (j + u < n) ? zs[(j + u) * L.ld + dd] : 0.f;
#pragma unroll
      for (int r = 0; r < SMAXR; ++r) {
        if (r < nrows) {
          const float4 h4 = *reinterpret_cast<const float4*>(&hs[r * L.np + j]);
          float v = acc[r];
          v = fmaf(h4.x, zv[0], v); v = fmaf(h4.y, zv[1], v); v = fmaf(h4.z, zv[2], v); v = fmaf(h4.w, zv[3], v);
          acc[r] = v;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < SMAXR; ++r) {
      if (r < nrows) {
        const int gi = r0 + r;
        float v = acc[r];
        if (g.cu != 0.f) v = fmaf(g.cu * st_all[gi * SUPCON_STATS_STRIDE + SUPCON_ST_WSUM], zs[gi * L.ld + dd], v);
        v *= gscale;
        if constexpr (sizeof(TO) == 4) out[(int64_t)gi * d + dd] = v;
        else out[(int64_t)gi * d + dd] = __float2bfloat16(v);
      }
    }
  }
}

}  // namespace

bool small_supported(const supcon_problem_t* p, const void* z) {
  if (p->row_offset != 0 || p->n_rows != p->n_total) return false;
  if (p->n_total > SC * SMAXR || p->n_total < 2) return false;
  if (p->d % 4 != 0 || p->d > 1024) return false;
  if ((reinterpret_cast<uintptr_t>(z) % 16) != 0) return false;
  const bool mine = p->alpha != 0.f && p->topk >= 1;
  if (mine && p->topk > 32 && p->topk < p->n_total - 1) return false;   // K rounds of arg-max: keep it short
  const int rows = (p->n_total + SC - 1) / SC;
  return small_layout(p->n_total, p->d, rows).total <= 224 * 1024;
}

cudaError_t small_launch(const SmallArgs& a, cudaStream_t stream) {
  const int rows = (a.n + SC - 1) / SC;
  const size_t smem = small_layout(a.n, a.d, rows).total;
  cudaError_t e;
#define SUPCON_SMALL(TI, TO)                                                                                 \
  do {                                                                                                       \
    e = cudaFuncSetAttribute(small_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e != cudaSuccess) return e;                                                                          \
    small_kernel<TI, TO><<<SC, SNT, smem, stream>>>(a);                                                      \
  } while (0)
  if (a.z_dtype == SUPCON_BF16) {
    if (a.dz_dtype == SUPCON_BF16) SUPCON_SMALL(__nv_bfloat16, __nv_bfloat16);
    else SUPCON_SMALL(__nv_bfloat16, float);
  } else {
    if (a.dz_dtype == SUPCON_BF16) SUPCON_SMALL(float, __nv_bfloat16);
    else SUPCON_SMALL(float, float);
  }
#undef SUPCON_SMALL
  return cudaGetLastError();
}

}  // namespace supcon
