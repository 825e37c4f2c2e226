// Stage 4 (multi-scale decomposition, also the ground-truth path) and stage 5 (log-space
// weighted combination + recombination), as stand-alone ops mirroring the reference functions
// one to one, plus rdm_fuse_tail which does stages 4+5 for one batch in a single launch.
//
// All of these are streaming kernels bounded by HBM bandwidth (O(1) flop per byte); the pyramids
// are tiny (<= 64x64 f64 = 32 KB per decoder) and live in shared memory, so each input is read
// once and each output written once.
//   quick_gm / gm_normalize : CP:244-255, RN:117, network/module.py:145-149
//   decompose               : CP:368-392 (+ CP:308-311 resize, CP:357-360 upsample)
//   log_stack               : CP:464-484 make_matrix
//   make_pred (+bwd)        : CP:512-528
//   recombination (+bwd)    : CP:394-421 (+ CP:362-366 multi_upsample)
//   fuse_tail               : RN:117-133 + network/module.py:132
#include "rdm_common.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace rdm {

constexpr int kMaxPtrs = 16;
constexpr int kPyrDoubles = 5461;   // levels 0..6: (4^7 - 1) / 3

__host__ __device__ __forceinline__ int off_level(int k) { return ((1 << (2 * k)) - 1) / 3; }   // levels 0..k-1
__host__ __device__ __forceinline__ int off_fine(int k) { return ((1 << (2 * k)) - 4) / 3; }    // F_1..F_{k-1}

struct PtrList {
  const void* p[kMaxPtrs];
  int32_t side[kMaxPtrs];
};
struct MutPtrList {
  void* p[kMaxPtrs];
  int32_t side[kMaxPtrs];
};

template <typename T>
__device__ __forceinline__ T block_prod(T v, T* scratch /* >= 32 */) {
  v = warp_prod(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = scratch[0];
  for (int i = 1; i < nw; ++i) r *= scratch[i];
  return r;
}

// torch.pow(t, 1/rc^2) keeps t's dtype (int64 -> default f32); evaluated through f64 and rounded
// so the f32 result is the correctly rounded power.
template <typename TAcc>
__device__ __forceinline__ TAcc pow_as(TAcc v, double e) {
  return (TAcc)pow((double)v, e);
}
// f64 rows (the ground truth, 16 384 values): prod_i x_i^e = exp(e * sum_i log x_i).  One log per
// element instead of a pow (log + exp), and closer to the exact product than multiplying 16 384
// individually rounded powers (2e-16 against ~1e-14 relative); it differs from the reference's
// own product by that ~1e-14.  Accumulated as (sum of logs) in the "product" slot.
template <typename TAcc>
struct GmAcc {
  static constexpr bool kLogSum = false;
};
template <>
struct GmAcc<double> {
  static constexpr bool kLogSum = true;
};
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch /* >= 32 */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = scratch[0];
  for (int i = 1; i < nw; ++i) r += scratch[i];
  return r;
}

// one CTA per batch row: gm[b] = prod_i pow(t[b,i], e);  optional normalised copy x / gm
template <typename TIn, typename TAcc>
__global__ void __launch_bounds__(256) gm_kernel(const TIn* __restrict__ t, int64_t n, double e, TAcc* __restrict__ gm_out,
                                                 TAcc* __restrict__ norm_out) {
  __shared__ TAcc scratch[32];
  const int64_t b = blockIdx.x;
  const TIn* row = t + b * n;
  TAcc gm;
  if constexpr (GmAcc<TAcc>::kLogSum) {
    double a0 = 0.0, a1 = 0.0;   // two chains: independent log() calls in flight
    for (int64_t i = threadIdx.x; i < n; i += 2 * blockDim.x) {
      a0 += log((double)row[i]);
      if (i + blockDim.x < n) a1 += log((double)row[i + blockDim.x]);
    }
    gm = (TAcc)exp(e * block_sum<double>(a0 + a1, reinterpret_cast<double*>(scratch)));
  } else {
    TAcc prod = (TAcc)1;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) prod *= pow_as<TAcc>((TAcc)row[i], e);
    gm = block_prod<TAcc>(prod, scratch);
  }
  if (gm_out && threadIdx.x == 0) gm_out[b] = gm;
  if (norm_out)
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) norm_out[b * n + i] = (TAcc)row[i] / gm;
}

// Large rows (the 128x128 ground truth, MOD:145-149: 16 384 f64 pow() per image) are split over a
// thread-block CLUSTER of 8 CTAs: each CTA multiplies its eighth, the partial products are
// exchanged through distributed shared memory (no workspace, no second launch), then every CTA
// normalises its own eighth with 128-bit accesses.
constexpr int kGmCluster = 8;
template <typename TIn, typename TAcc>
__global__ void __launch_bounds__(256) gm_cluster_kernel(const TIn* __restrict__ t, int64_t n, double e, TAcc* __restrict__ gm_out,
                                                         TAcc* __restrict__ norm_out) {
  __shared__ TAcc scratch[32];
  __shared__ TAcc part;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int64_t b = blockIdx.x / kGmCluster;
  const int64_t chunk = n / kGmCluster;
  const TIn* row = t + b * n + rank * chunk;
  TAcc mine;
  if constexpr (GmAcc<TAcc>::kLogSum) {
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = threadIdx.x; i < chunk; i += 2 * blockDim.x) {
      a0 += log((double)row[i]);
      if (i + blockDim.x < chunk) a1 += log((double)row[i + blockDim.x]);
    }
    mine = (TAcc)block_sum<double>(a0 + a1, reinterpret_cast<double*>(scratch));   // sum of logs of this eighth
  } else {
    TAcc prod = (TAcc)1;
    for (int64_t i = threadIdx.x; i < chunk; i += blockDim.x) prod *= pow_as<TAcc>((TAcc)row[i], e);
    mine = block_prod<TAcc>(prod, scratch);
  }
  if (threadIdx.x == 0) part = mine;
  cluster.sync();
  TAcc gm = GmAcc<TAcc>::kLogSum ? (TAcc)0 : (TAcc)1;
  for (int r = 0; r < kGmCluster; ++r) {   // rank order: deterministic
    const TAcc other = *cluster.map_shared_rank(&part, r);
    gm = GmAcc<TAcc>::kLogSum ? gm + other : gm * other;
  }
  if constexpr (GmAcc<TAcc>::kLogSum) gm = (TAcc)exp(e * (double)gm);
  cluster.sync();   // nobody leaves while its `part` may still be read
  if (gm_out && rank == 0 && threadIdx.x == 0) gm_out[b] = gm;
  if (norm_out) {
    TAcc* dst = norm_out + b * n + rank * chunk;
    for (int64_t i = threadIdx.x; i < chunk; i += blockDim.x) dst[i] = (TAcc)row[i] / gm;
  }
}

// ---------------------------------------------------------------------------------------------
// Pyramid of one map held in shared memory D[] (level k at off_level(k)), level `top` filled by
// the caller.  emit(k, idx, F) is called for every fine-detail value F_k[idx], k = top..1.
template <typename Emit>
__device__ __forceinline__ void pyramid_down(double* D, int top, Emit emit) {
  for (int k = top; k >= 1; --k) {
    const int side = 1 << k, half = side >> 1;
    const double* cur = D + off_level(k);
    double* nxt = D + off_level(k - 1);
    for (int idx = threadIdx.x; idx < half * half; idx += blockDim.x) {
      int y = idx / half, x = idx - y * half;
      nxt[idx] = bicubic_half_at([&](int r, int c) { return cur[r * side + c]; }, y, x, side);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < side * side; idx += blockDim.x) {
      int y = idx >> k, x = idx & (side - 1);
      emit(k, idx, cur[idx] / nxt[(y >> 1) * half + (x >> 1)]);   // CP:389 torch.div(dn, upsample(dn_1))
    }
  }
}

// Output is LEVEL-MAJOR: [D_0 of all images (unless relative)] [F_1 of all images] ... so every
// component is a dense (N,1,2^k,2^k) tensor for the caller.
// side <= 32: one CTA per image, the whole pyramid in shared memory (larger maps: decompose_cluster_kernel).
template <typename TIn>
__global__ void __launch_bounds__(256) decompose_kernel(const TIn* __restrict__ in, int side, int n, int relative, double* __restrict__ out, int64_t n_images) {
  extern __shared__ __align__(16) double D[];
  const int64_t img = blockIdx.x;
  const TIn* src = in + img * (int64_t)side * side;
  const int base = relative ? 0 : 1;
  auto fine = [&](int k) { return out + n_images * (base + off_fine(k)) + img * ((int64_t)1 << (2 * k)); };
  double* dn = D + off_level(n);
  for (int idx = threadIdx.x; idx < side * side; idx += blockDim.x) dn[idx] = (double)src[idx];
  __syncthreads();
  pyramid_down(D, n, [&](int k, int idx, double f) { fine(k)[idx] = f; });
  __syncthreads();
  if (!relative && threadIdx.x == 0) out[img] = D[0];
}

// ---------------------------------------------------------------------------------------------
// Backward of decompose.  With D_{k-1} = H(D_k) (H = the linear stride-2 bicubic filter) and
// F_k = D_k / U(D_{k-1}), the total gradient G_j = dL/dD_j obeys
//   G_j = [j >= 1] gF_j / U(D_{j-1})  -  [j < n] pool2x2(gF_{j+1} * D_{j+1}) / D_j^2
//         + [j >= 1] H^T G_{j-1}  + [j == 0] gD_0
// evaluated bottom-up; G_n is the input gradient.  H^T is applied as a deterministic gather.
__device__ __forceinline__ int adj_taps(int r, int side, int (&ys)[3], double (&ws)[3]) {
  const double w[4] = {RDM_W0, RDM_W1, RDM_W1, RDM_W0};
  const int half = side >> 1;
  int cnt = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int t = r + 1 - a;
    if (t >= 0 && !(t & 1) && (t >> 1) < half) {
      ys[cnt] = t >> 1;
      ws[cnt] = w[a];
      ++cnt;
    }
  }
  if (r == 0) {            // tap index -1 clamps to 0
    ys[cnt] = 0;
    ws[cnt] = w[0];
    ++cnt;
  }
  if (r == side - 1) {     // tap index `side` clamps to side-1
    ys[cnt] = half - 1;
    ws[cnt] = w[3];
    ++cnt;
  }
  return cnt;
}

__device__ __forceinline__ double adj_bicubic_at(const double* g, int r, int c, int side) {
  const int half = side >> 1;
  int ys[3], xs[3];
  double wy[3], wx[3];
  const int ny = adj_taps(r, side, ys, wy), nx = adj_taps(c, side, xs, wx);
  double acc = 0.0;
  for (int i = 0; i < ny; ++i) {
    double inner = 0.0;
    for (int j = 0; j < nx; ++j) inner = fma(g[ys[i] * half + xs[j]], wx[j], inner);
    acc = fma(inner, wy[i], acc);
  }
  return acc;
}

template <typename TIn>
__global__ void __launch_bounds__(256) decompose_bwd_kernel(const TIn* __restrict__ in, int side, int n, int relative,
                                                            const double* __restrict__ gpyr, int64_t n_images, TIn* __restrict__ gin) {
  extern __shared__ __align__(16) double sm[];
  double* D = sm;
  double* G = sm + kPyrDoubles;
  const int64_t img = blockIdx.x;
  const TIn* src = in + img * (int64_t)side * side;
  const int base = relative ? 0 : 1;
  auto gfine = [&](int k) { return gpyr + n_images * (base + off_fine(k)) + img * ((int64_t)1 << (2 * k)); };
  // ---- forward levels
  int top = n;
  if (n == 7) {
    double* d6 = D + off_level(6);
    for (int idx = threadIdx.x; idx < 4096; idx += blockDim.x)
      d6[idx] = bicubic_half_at([&](int r, int c) { return (double)src[r * 128 + c]; }, idx >> 6, idx & 63, 128);
    top = 6;
  } else {
    double* dn = D + off_level(n);
    for (int idx = threadIdx.x; idx < side * side; idx += blockDim.x) dn[idx] = (double)src[idx];
  }
  __syncthreads();
  for (int k = top; k >= 1; --k) {
    const int sd = 1 << k, half = sd >> 1;
    const double* cur = D + off_level(k);
    double* nxt = D + off_level(k - 1);
    for (int idx = threadIdx.x; idx < half * half; idx += blockDim.x)
      nxt[idx] = bicubic_half_at([&](int r, int c) { return cur[r * sd + c]; }, idx / half, idx % half, sd);
    __syncthreads();
  }
  // ---- gradients, bottom-up
  for (int j = 0; j <= n; ++j) {
    const int sd = 1 << j;
    const bool in_smem = j <= 6;
    for (int idx = threadIdx.x; idx < sd * sd; idx += blockDim.x) {
      const int y = idx >> j, x = idx & (sd - 1);
      const double dj = in_smem ? D[off_level(j) + idx] : (double)src[idx];
      double g = 0.0;
      if (j == 0 && !relative) g = gpyr[img];
      if (j >= 1) {
        const int hs = sd >> 1;
        g += gfine(j)[idx] / D[off_level(j - 1) + (y >> 1) * hs + (x >> 1)];
        g += adj_bicubic_at(G + off_level(j - 1), y, x, sd);
      }
      if (j < n) {
        const int us = sd << 1;
        const double* gf = gfine(j + 1);
        double pool = 0.0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int u = (2 * y + dy) * us + 2 * x + dx;
            const double du = (j + 1 <= 6) ? D[off_level(j + 1) + u] : (double)src[u];
            pool = fma(gf[u], du, pool);
          }
        g -= pool / (dj * dj);
      }
      if (j == n)
        gin[img * (int64_t)side * side + idx] = (TIn)g;
      else
        G[off_level(j) + idx] = g;
    }
    __syncthreads();
  }
}

// backward of log_stack: grad_cand_k[b,m] = g[b,k,m] / cand_k[b,m]
__global__ void __launch_bounds__(256) log_stack_bwd_kernel(PtrList cands, MutPtrList gc, int K, int64_t batch, int64_t M,
                                                            const double* __restrict__ g) {
  const int64_t total = batch * K * M;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = o % M;
    int64_t bk = o / M;
    int k = (int)(bk % K);
    int64_t b = bk / K;
    reinterpret_cast<double*>(gc.p[k])[b * M + m] = g[o] / reinterpret_cast<const double*>(cands.p[k])[b * M + m];
  }
}

// backward of gm_normalize / quick_gm for one batch row per CTA (x f32 or f64):
//   y_i = x_i / gm, gm = prod_j x_j^e  ->  dL/dx_j = gy_j / gm - (e / x_j) sum_i gy_i y_i  (+ ggm * e * gm / x_j)
template <typename T>
__global__ void __launch_bounds__(256) gm_bwd_kernel(const T* __restrict__ x, int64_t n, double e, const T* __restrict__ g_gm,
                                                     const T* __restrict__ g_norm, T* __restrict__ gx) {
  __shared__ double scratch[32];
  __shared__ double part[8];
  const int64_t b = blockIdx.x;
  const T* row = x + b * n;
  double prod = 1.0, dot = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    prod *= pow((double)row[i], e);
    if (g_norm) dot = fma((double)g_norm[b * n + i], (double)row[i], dot);
  }
  const double gm = block_prod<double>(prod, scratch);
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = dot;
  __syncthreads();
  double S = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) S += part[i];
  S /= gm;   // sum_i gy_i y_i
  const double gg = g_gm ? (double)g_gm[b] : 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double v = (gg * gm - S) * e / (double)row[i];
    if (g_norm) v += (double)g_norm[b * n + i] / gm;
    gx[b * n + i] = (T)v;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) log_stack_kernel(PtrList cands, int K, int64_t batch, int64_t M, double* __restrict__ out) {
  const int64_t total = batch * K * M;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = o % M;
    int64_t bk = o / M;
    int k = (int)(bk % K);
    int64_t b = bk / K;
    out[o] = log(reinterpret_cast<const double*>(cands.p[k])[b * M + m]);
  }
}

__global__ void __launch_bounds__(256) make_pred_kernel(const double* __restrict__ A, const float* __restrict__ w, int64_t batch,
                                                        int K, int64_t M, float* __restrict__ out) {
  const int64_t total = batch * M;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = o / M, m = o - b * M;
    const double* a = A + b * K * M + m;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf((float)a[k * M], w[k], acc);   // A[b].T.float() @ w.float()
    out[o] = acc;
  }
}

__global__ void __launch_bounds__(256) make_pred_bwd_a_kernel(const float* __restrict__ w, const float* __restrict__ g, int64_t batch,
                                                              int K, int64_t M, double* __restrict__ grad_A) {
  const int64_t total = batch * K * M;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = o % M;
    int64_t bk = o / M;
    int k = (int)(bk % K);
    int64_t b = bk / K;
    grad_A[o] = (double)(w[k] * g[b * M + m]);
  }
}

// one CTA per weight: grad_w[k] = sum_b sum_m f32(A[b,k,m]) * g[b,m]
__global__ void __launch_bounds__(512) make_pred_bwd_w_kernel(const double* __restrict__ A, const float* __restrict__ g, int64_t batch,
                                                              int K, int64_t M, float* __restrict__ grad_w) {
  __shared__ double part[16];
  const int k = blockIdx.x;
  double acc = 0.0;
  for (int64_t o = threadIdx.x; o < batch * M; o += blockDim.x) {
    int64_t b = o / M, m = o - b * M;
    acc = fma((double)(float)A[(b * K + k) * M + m], (double)g[o], acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    grad_w[k] = (float)t;
  }
}

// ---------------------------------------------------------------------------------------------
// recombination: comps (in list order) d0? then sides 2,4,...; out side S = 2^n.
// Four horizontally adjacent output pixels per thread (two 128-bit streaming stores; a warp writes
// 1 KB contiguous): every component coarser than the last two levels contributes ONE gathered value
// to all four, so the gather count per output pixel drops ~4x against a pixel-per-thread kernel.
template <typename TC>
__global__ void __launch_bounds__(256) recombination_kernel(PtrList comps, int n_comps, int has_d0, int n, int64_t batch,
                                                            double* __restrict__ out) {
  const int S = 1 << n, Q = S >> 2;          // Q quads per row (S >= 4)
  const int64_t per = (int64_t)S * Q;
  const int64_t total = batch * per;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = o / per;
    const int rem = (int)(o - b * per);
    const int y = rem / Q, x = (rem - y * Q) << 2;
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = has_d0; j < n_comps; ++j) {
      const int cs = comps.side[j];
      const int sh = n - (31 - __clz(cs));
      const TC* c = reinterpret_cast<const TC*>(comps.p[j]) + b * (int64_t)cs * cs + (y >> sh) * cs + (x >> sh);
      double v[4];
      if (sh >= 2) {
        v[0] = v[1] = v[2] = v[3] = (double)c[0];
      } else if (sh == 1) {
        v[0] = v[1] = (double)c[0];
        v[2] = v[3] = (double)c[1];
      } else {
        v[0] = (double)c[0];
        v[1] = (double)c[1];
        v[2] = (double)c[2];
        v[3] = (double)c[3];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = (j == has_d0) ? v[i] : a[i] + v[i];
    }
    if (has_d0) {
      const double d0 = (double)reinterpret_cast<const TC*>(comps.p[0])[b];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = d0 + a[i];
    }
    double* dst = out + b * (int64_t)S * S + (int64_t)y * S + x;
    stg_stream_f64x2(dst, a[0], a[1]);
    stg_stream_f64x2(dst + 2, a[2], a[3]);
  }
}

// Small outputs (side 2): two pixels per thread.
template <typename TC>
__global__ void __launch_bounds__(256) recombination_small_kernel(PtrList comps, int n_comps, int has_d0, int n, int64_t batch,
                                                                  double* __restrict__ out) {
  const int S = 1 << n;
  const int64_t per = (int64_t)S * S / 2;
  const int64_t total = batch * per;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = o / per;
    const int rem = (int)(o - b * per);
    const int y = rem / (S / 2), x = (rem - y * (S / 2)) * 2;
    double a0 = 0.0, a1 = 0.0;
    for (int j = has_d0; j < n_comps; ++j) {
      const int cs = comps.side[j];
      const int sh = n - (31 - __clz(cs));
      const TC* c = reinterpret_cast<const TC*>(comps.p[j]) + b * (int64_t)cs * cs + (y >> sh) * cs;
      const double v0 = (double)c[x >> sh], v1 = (double)c[(x + 1) >> sh];
      a0 = (j == has_d0) ? v0 : a0 + v0;
      a1 = (j == has_d0) ? v1 : a1 + v1;
    }
    if (has_d0) {
      const double d0 = (double)reinterpret_cast<const TC*>(comps.p[0])[b];
      a0 = d0 + a0;
      a1 = d0 + a1;
    }
    stg_stream_f64x2(out + b * (int64_t)S * S + (int64_t)y * S + x, a0, a1);
  }
}

// backward: successive 2x2 sum pooling of grad_out (exactly what autograd does through the
// chain of nearest x2 upsamples); component of side 2^k receives pooled level k.
template <typename TC>
__global__ void __launch_bounds__(256) recombination_bwd_kernel(const double* __restrict__ grad_out, MutPtrList gc, int n_comps, int n,
                                                                int64_t batch) {
  extern __shared__ __align__(16) double D[];
  const int64_t b = blockIdx.x;
  const int S = 1 << n;
  const double* g = grad_out + b * (int64_t)S * S;
  auto emit = [&](int k, const double* lvl) {
    for (int j = 0; j < n_comps; ++j)
      if (gc.side[j] == (1 << k) && gc.p[j]) {
        TC* dst = reinterpret_cast<TC*>(gc.p[j]) + b * (int64_t)(1 << (2 * k));
        for (int idx = threadIdx.x; idx < (1 << (2 * k)); idx += blockDim.x) dst[idx] = (TC)lvl[idx];
      }
  };
  // level n straight from global
  for (int j = 0; j < n_comps; ++j)
    if (gc.side[j] == S && gc.p[j]) {
      TC* dst = reinterpret_cast<TC*>(gc.p[j]) + b * (int64_t)S * S;
      for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) dst[idx] = (TC)g[idx];
    }
  if (n == 0) return;
  {
    const int half = S >> 1;
    double* nxt = D + off_level(n - 1);
    for (int idx = threadIdx.x; idx < half * half; idx += blockDim.x) {
      int y = idx / half, x = idx - y * half;
      const double* r0 = g + (2 * y) * S + 2 * x;
      nxt[idx] = (r0[0] + r0[1]) + (r0[S] + r0[S + 1]);
    }
    __syncthreads();
    emit(n - 1, nxt);
  }
  for (int k = n - 1; k >= 1; --k) {
    const int side = 1 << k, half = side >> 1;
    const double* cur = D + off_level(k);
    double* nxt = D + off_level(k - 1);
    for (int idx = threadIdx.x; idx < half * half; idx += blockDim.x) {
      int y = idx / half, x = idx - y * half;
      const double* r0 = cur + (2 * y) * side + 2 * x;
      nxt[idx] = (r0[0] + r0[1]) + (r0[side] + r0[side + 1]);
    }
    __syncthreads();
    emit(k - 1, nxt);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused tail.  Grid = n_images * bands, one thread-block CLUSTER of `bands` CTAs per image; every CTA
// holds the (tiny) pyramids of its image in shared memory and writes one horizontal band of the
// 128x128 f64 log-depth map, so the 128 KB per image output - the only significant HBM traffic of
// stages 4+5 - is spread over the chip.  The cheap bicubic levels are redone by every CTA; the f64
// logs are split over the cluster and exchanged through distributed shared memory.
// All decoders descend their pyramids TOGETHER down to 1x1 first (one barrier per large level), then
// the expensive f64 div + log of EVERY level is taken in one stage spread over all threads of the
// cluster, then the per-slot weighted sums are formed in candidate order (CP:521 sums decoder 1 first).
constexpr int kMaxRel = 6;
constexpr int kMaxDec = kMaxRel + 1;
struct TailParams {
  const int64_t* x_d1;
  const float* rel[kMaxRel];
  const float* w;
  float* yhat_out;
  double* depth_out;        // (N,128,128) f64 or NULL
  double* depth_compact;    // (N,2^kmax,2^kmax) f64 or NULL: one value per constant block of the map
  double* A_out[8];
  int32_t side[kMaxRel];
  int32_t n_rel;
  int32_t w_off[8];         // first weight of slot k
  int32_t K[8];             // candidates in slot k
  int32_t cand[kMaxDec][8]; // candidate index of decoder d in slot k
  int32_t doff[kMaxDec];    // start (doubles) of decoder d's pyramid in shared memory
  int32_t nlev[kMaxDec];    // levels n_d (side 2^n_d)
  int32_t act[8][kMaxDec];  // decoders that have level k, in candidate order
  int32_t nact[8];
  int32_t dtotal;           // doubles of pyramid storage
  int32_t lofs[8];          // first float of level k in the log buffer (levels kmax .. 1, in that order)
  int32_t ltotal;           // floats of the log buffer = sum_k nact[k] 4^k
  int32_t kmax;
  int32_t bands;
};

#ifndef RDM_TAIL_THREADS
#define RDM_TAIL_THREADS 256    // measured: 1024 threads do not shorten the CTA (f64 div+log latency ~1700 cycles per stage dominates) and hurt co-residency with the ALS kernel
#endif
__global__ void __launch_bounds__(RDM_TAIL_THREADS) fuse_tail_kernel(const __grid_constant__ TailParams P) {
  extern __shared__ __align__(16) double D[];                // P.dtotal doubles
  float* yh = reinterpret_cast<float*>(D + P.dtotal);         // slot k at off_level(k)
  const int ylen = off_level(P.kmax + 1);
  float* L = yh + ylen;                                       // P.ltotal floats: f32(log F) of every level
  cg::cluster_group cluster = cg::this_cluster();             // the P.bands CTAs of one image
  __shared__ float scratch[32];
  __shared__ float wsm[64];                                   // the (<= 4 + 6*6) weights, read from HBM once
  __shared__ float wl[8][kMaxDec];                            // wl[k][a]: weight of the a-th active decoder of slot k
  const int64_t img = blockIdx.x / P.bands;
  const int band = blockIdx.x - (int)(img * P.bands);
  const bool lead = band == 0;
  const int tid = threadIdx.x;
  if (tid < P.w_off[7] + P.K[7]) wsm[tid] = P.w[tid];
  if (tid < 8 * kMaxDec) {
    const int k = tid / kMaxDec, a = tid - k * kMaxDec;
    if (k >= 1 && k <= P.kmax && a < P.nact[k]) wl[k][a] = P.w[P.w_off[k] + P.cand[P.act[k][a]][k]];
  }
#ifdef RDM_TIMING
  long long tt[12];
  int ti = 0;
  tt[ti++] = clock64();
#endif

  // ---- top levels: decoder 1 = x / gm(x) in f32 (RN:117); relative decoders as they come
  {
    // the DORN counts are requested first and consumed last, so that their latency overlaps the map loads
    long long xi = 1;
    if (tid < 64) xi = P.x_d1[img * 64 + tid];
    for (int r = 0; r < P.n_rel; ++r) {
      const int side = P.side[r];
      const float* src = P.rel[r] + img * (int64_t)side * side;
      double* dn = D + P.doff[r + 1] + off_level(P.nlev[r + 1]);
      for (int idx = tid; idx < side * side; idx += blockDim.x) dn[idx] = (double)src[idx];
    }
    const float v = (float)xi;
    const float pw = (tid < 64) ? powf(v, 0.015625f) : 1.f;   // torch.pow(int64 -> f32, 1/64) is an f32 pow as well (CP:248-253)
    const float gm = block_prod<float>(pw, scratch);
    if (tid < 64) D[P.doff[0] + off_level(3) + tid] = (double)(v / gm);
  }
  __syncthreads();
#ifdef RDM_TIMING
  tt[ti++] = clock64();
#endif
  // ---- all pyramids down to 1x1 first (cheap, redone by every CTA of the cluster): large levels with all
  // threads, levels <= 3 (at most 84 values per decoder) in warp 0 back to back
  auto bicubic_level = [&](int k, int first, int stride) {   // level k-1 of every decoder that has level k
    const int side = 1 << k, half = side >> 1;
    for (int item = first; item < P.nact[k] * half * half; item += stride) {
      const int a = item >> (2 * (k - 1)), idx = item & (half * half - 1);
      const double* cur = D + P.doff[P.act[k][a]] + off_level(k);
      D[P.doff[P.act[k][a]] + off_level(k - 1) + idx] =
          bicubic_half_at([&](int r, int c) { return cur[r * side + c]; }, idx >> (k - 1), idx & (half - 1), side);
    }
  };
  for (int k = P.kmax; k >= 4; --k) {
    bicubic_level(k, tid, blockDim.x);
    __syncthreads();
  }
  const int ksmall = P.kmax < 3 ? P.kmax : 3;
  if (tid < 32)
    for (int k = ksmall; k >= 1; --k) {
      bicubic_level(k, tid, 32);
      __syncwarp();
    }
  cluster.sync();   // pyramids complete; every CTA of the cluster is running: its shared memory may be written now
#ifdef RDM_TIMING
  if (ti < 11) tt[ti++] = clock64();
#endif
  // ---- ONE log stage for all levels.  F_k = D_k / up2(D_{k-1}) (CP:389) and its log (CP:478-480); the f64
  // div + log is the expensive part of the tail (~110 instructions and ~1700 cycles of dependent latency per
  // value), so the items of all levels are SPLIT over the CTAs of the cluster (two independent chains per
  // thread at scales 8/16/32) and every result is stored into the log buffer of all of them through
  // distributed shared memory, instead of every band redoing all of it level after level.
  const int nb = P.bands, gthreads = blockDim.x * nb, gtid = band * blockDim.x + tid;
  for (int base = gtid; base < P.ltotal; base += 4 * gthreads) {
    double lg[4];
    int lev[4], itm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = base + j * gthreads;
      lev[j] = 0;
      if (t < P.ltotal) {
        int k = P.kmax;
        while (k > 1 && t >= P.lofs[k - 1]) --k;   // levels are laid out kmax first
        const int item = t - P.lofs[k];
        const int side = 1 << k, half = side >> 1;
        const int a = item >> (2 * k), idx = item & (side * side - 1);
        const double* bp = D + P.doff[P.act[k][a]];
        const int y = idx >> k, x = idx & (side - 1);
        lg[j] = log(bp[off_level(k) + idx] / bp[off_level(k - 1) + (y >> 1) * half + (x >> 1)]);
        lev[j] = k;
        itm[j] = item;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = lev[j];
      if (k) {
        const int side2 = 1 << (2 * k);
        const int a = itm[j] >> (2 * k), idx = itm[j] & (side2 - 1);
        const int d = P.act[k][a];
        if (P.A_out[k]) P.A_out[k][(img * P.K[k] + P.cand[d][k]) * (int64_t)side2 + idx] = lg[j];
        const float v = (float)lg[j];
        float* dst = L + base + j * gthreads;
        for (int b = 0; b < nb; ++b) *cluster.map_shared_rank(dst, b) = v;
      }
    }
  }
  cluster.sync();   // last remote access: a CTA may leave the cluster afterwards
#ifdef RDM_TIMING
  if (ti < 11) tt[ti++] = clock64();
#endif
  // slot k of y_hat: sum over candidates in decoder order (CP:521 / CP:526)
  for (int k = P.kmax; k >= 1; --k) {
    const int n = 1 << (2 * k), na = P.nact[k], lofs = P.lofs[k];
    for (int idx = tid; idx < n; idx += blockDim.x) {
      float y = 0.f;
      for (int a = 0; a < na; ++a) y = fmaf(L[lofs + a * n + idx], wl[k][a], y);
      yh[off_level(k) + idx] = y;
    }
  }
#ifdef RDM_TIMING
  if (ti < 11) tt[ti++] = clock64();
#endif
  if (tid == 0) {   // slot 0: D_0 of decoder 1
    const double lg = log(D[P.doff[0]]);
    if (lead && P.A_out[0]) P.A_out[0][img] = lg;
    yh[0] = fmaf((float)lg, wsm[P.w_off[0]], 0.f);
  }
  __syncthreads();
  if (lead && P.yhat_out)
    for (int i = tid; i < ylen; i += blockDim.x) P.yhat_out[img * ylen + i] = yh[i];
  // ---- recombination of this CTA's band (CP:394-421): d0 + ((f1 + f2) + ... + f_kmax)
  const int rows = 128 / P.bands;
  double* out = P.depth_out ? P.depth_out + img * 16384 + (int64_t)band * rows * 128 : nullptr;
  const double d0 = (double)yh[0];
  // Levels above kmax do not exist, so the sum is constant on blocks of bs x bs pixels (bs = 2^(7-kmax)):
  // one gather-sum per block, then bs rows of 128-bit stores.
  const int bsh = 7 - P.kmax, bs = 1 << bsh;
  if (bs >= 2) {
    const int bpr = 128 >> bsh;                                  // blocks per row
    for (int blk = tid; blk < (rows >> bsh) * bpr; blk += blockDim.x) {
      const int by = blk / bpr, bx = blk - by * bpr;
      const int y = band * rows + (by << bsh), x = bx << bsh;
      double a = 0.0;
      for (int k = 1; k <= P.kmax; ++k) {
        const int sh = 7 - k;
        const double v = (double)yh[off_level(k) + (y >> sh) * (1 << k) + (x >> sh)];
        a = (k == 1) ? v : a + v;
      }
      const double v = d0 + a;
      if (P.depth_compact) P.depth_compact[img * (int64_t)(bpr * bpr) + (y >> bsh) * bpr + bx] = v;
      if (out)
        for (int r = 0; r < bs; ++r)
          for (int c = 0; c < bs; c += 2) stg_stream_f64x2(out + ((by << bsh) + r) * 128 + x + c, v, v);
    }
  } else if (out) {
    for (int o = tid; o < rows * 64; o += blockDim.x) {
      const int y = band * rows + (o >> 6), x = (o & 63) * 2;
      double a0 = 0.0, a1 = 0.0;
      for (int k = 1; k <= P.kmax; ++k) {
        const int sh = 7 - k, cs = 1 << k;
        const float* c = yh + off_level(k) + (y >> sh) * cs;
        const double v0 = (double)c[x >> sh], v1 = (double)c[(x + 1) >> sh];
        a0 = (k == 1) ? v0 : a0 + v0;
        a1 = (k == 1) ? v1 : a1 + v1;
      }
      stg_stream_f64x2(out + (o >> 6) * 128 + x, d0 + a0, d0 + a1);
    }
  }
#ifdef RDM_TIMING
  tt[ti++] = clock64();
  if (tid == 0 && blockIdx.x == 5) {
    printf("fuse_tail block 5 (%d thr): load+gm %lld", (int)blockDim.x, tt[1] - tt[0]);
    for (int i = 2; i < ti - 1; ++i) printf(", +%lld", tt[i] - tt[i - 1]);
    printf(", yhat+band %lld, total %lld cycles\n", tt[ti - 1] - tt[ti - 2], tt[ti - 1] - tt[0]);
  }
#endif
}

// ---------------------------------------------------------------------------------------------
// Large maps (side 64 / 128): one thread-block CLUSTER of 8 CTAs per image, CTA r owns a band of side/8 rows.
// Shared by the stand-alone decomposition (rdm_decompose, side >= 64) and the fused ground-truth preparation.
constexpr int kBandCluster = 8;
constexpr int kBandMaxSide = 128;
constexpr int kBandMaxRows = kBandMaxSide / kBandCluster;      // 16

struct BandSmem {
  double Y[(kBandMaxRows + 2) * kBandMaxSide];   // staged rows of the band + one halo row either side (row stride = side)
  double Db[(kBandMaxRows / 2) * (kBandMaxSide / 2)];   // the band's rows of D_{n-1}
  double pyr[kPyrDoubles];                       // CTA 0: levels 0..6
  double scratch[32];
  double part;                                   // ground truth: sum of logs of the own rows
  int labels[64];                                // ground truth, CTA 0: SID labels of the 8x8 map
  float fscratch[32];
};

// From the staged rows sm.Y (global rows band*rank - 1 .. band*rank + band, border rows repeated) of a side x side map:
// the band's rows of D_{n-1} (stride-2 bicubic, CP:308-311) -> the band's rows of F_n = D_n / up2(D_{n-1}) (CP:389) with
// 128-bit stores; D_{n-1} goes to CTA 0's pyramid buffer through distributed shared memory, and CTA 0 finishes levels
// n-1 .. 1 from shared memory.  Returns with `true` in CTA 0 only (its sm.pyr holds levels 0 .. n-1).
__device__ __forceinline__ bool band_decompose(BandSmem& sm, cg::cluster_group& cluster, int rank, int side, int n, int base,
                                               double* __restrict__ out, int64_t n_images, int64_t img) {
  const int tid = threadIdx.x, half = side >> 1, band = side / kBandCluster;
  double* pyr0 = cluster.map_shared_rank(&sm.pyr[0], 0);
  for (int i = tid; i < (band / 2) * half; i += blockDim.x) {
    const int yy = i / half, x = i - yy * half;
    const double w[4] = {RDM_W0, RDM_W1, RDM_W1, RDM_W0};
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {            // source rows 2y-1 .. 2y+2, y = (band/2) rank + yy  ->  staged rows 2 yy + a
      const double* rowp = sm.Y + (2 * yy + a) * side;
      double inner = 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c = min(max(2 * x - 1 + b, 0), side - 1);
        inner = (b == 0) ? __dmul_rn(rowp[c], w[0]) : fma(rowp[c], w[b], inner);
      }
      acc = (a == 0) ? __dmul_rn(inner, w[0]) : fma(inner, w[a], acc);
    }
    sm.Db[i] = acc;
    pyr0[off_level(n - 1) + ((band / 2) * rank + yy) * half + x] = acc;
  }
  __syncthreads();
  {
    double* fn = out + n_images * (base + off_fine(n)) + img * (int64_t)side * side;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(fn) & 15u) == 0;
    for (int i = tid; i < band * half; i += blockDim.x) {   // two output pixels per thread
      const int rr = i / half, x2 = i - rr * half;
      const double d = sm.Db[(rr >> 1) * half + x2];
      const double* sp = sm.Y + (1 + rr) * side + 2 * x2;
      double* dst = fn + (band * rank + rr) * side + 2 * x2;
      if (vec_ok) {
        stg_stream_f64x2(dst, sp[0] / d, sp[1] / d);
      } else {   // level-major offsets are N * odd doubles when D_0 is present: only 8-byte aligned for odd N
        dst[0] = sp[0] / d;
        dst[1] = sp[1] / d;
      }
    }
  }
  cluster.sync();   // D_{n-1} is complete in CTA 0; nobody's shared memory is touched remotely after this
  if (rank != 0) return false;
  pyramid_down(sm.pyr, n - 1, [&](int k, int idx, double f) { out[n_images * (base + off_fine(k)) + img * ((int64_t)1 << (2 * k)) + idx] = f; });
  __syncthreads();
  return true;
}

// CP:368-392 for side 64 / 128 in one launch (no D_{n-1} parked in the output, no second kernel reading it back).
template <typename TIn>
__global__ void __launch_bounds__(256) decompose_cluster_kernel(const TIn* __restrict__ in, int side, int n, int relative,
                                                                double* __restrict__ out, int64_t n_images) {
  extern __shared__ __align__(16) unsigned char band_raw[];
  BandSmem& sm = *reinterpret_cast<BandSmem*>(band_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int64_t img = blockIdx.x / kBandCluster;
  const int band = side / kBandCluster;
  const TIn* src = in + img * (int64_t)side * side;
  const int row_lo = band * rank - 1;
  for (int i = threadIdx.x; i < (band + 2) * side; i += blockDim.x) {
    const int rr = i / side, c = i - rr * side;
    sm.Y[i] = (double)src[min(max(row_lo + rr, 0), side - 1) * side + c];
  }
  cluster.sync();   // the CTA barrier for Y, and: every CTA of the cluster is running before band_decompose writes into CTA 0's shared memory
  if (band_decompose(sm, cluster, rank, side, n, relative ? 0 : 1, out, n_images, img) && !relative && threadIdx.x == 0) out[img] = sm.pyr[0];
}

// ---------------------------------------------------------------------------------------------
// Ground-truth preparation of the training step in ONE launch (network/module.py:68, 74-78, 119-127, 134-149 +
// utils.py:195-211): bicubic resize to 128x128 (CP:308-311), mask (+1e-4 everywhere, invalid -> 1.0001), geometric-
// mean normalisation, decomposition n = 7, and the ordinal target: SID labels of the 8x8 resize of the masked map,
// whose normalised decomposition supplies D_0 of the component targets.
// One cluster of 8 CTAs per image; CTA r owns rows 16r..16r+15 of the 128x128 map:
//   1. resize + mask rows 16r-1 .. 16r+16 (one halo row either side, recomputed rather than exchanged) into shared
//      memory; the own rows go out as y;
//   2. row r of the 8x8 resize (taps 16r+6..16r+9: inside the band) -> SID labels -> ord_target and CTA 0;
//   3. sum of logs of the own rows, exchanged through distributed shared memory -> gm; the staged rows are divided by
//      it (the reference normalises BEFORE it decomposes);
//   4. band_decompose: F_7 rows of the band, D_6 to CTA 0, CTA 0 finishes levels 6..1;
//   5. CTA 0: D_0 of the ordinal pyramid.
constexpr int kGtSide = 128;
constexpr int kGtBand = kGtSide / kBandCluster;        // 16 rows
constexpr int kGtStageRows = kGtBand + 2;

#ifndef RDM_GT_THREADS
#define RDM_GT_THREADS 256   // every phase is a latency chain (f64 div ~500, f64 log ~1200 cycles): 256 threads 44.5k cycles per CTA, 512: 41.5k, 1024: 35.3k - but the training step, where this kernel runs beside the fusion path, is faster with 256 (0.191 vs 0.197 ms)
#endif
template <typename TIn>
__global__ void __launch_bounds__(RDM_GT_THREADS) gt_prepare_kernel(const TIn* __restrict__ y_raw, int ih, int iw, int64_t n_images, double sid_K,
                                                         double sid_alpha, double sid_log_ratio, double* __restrict__ y_out,
                                                         double* __restrict__ pyr_out, int32_t* __restrict__ ord_out) {
  extern __shared__ __align__(16) unsigned char band_raw[];
  BandSmem& sm = *reinterpret_cast<BandSmem*>(band_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int64_t img = blockIdx.x / kBandCluster;
  const int tid = threadIdx.x;
  const TIn* src = y_raw + img * (int64_t)ih * iw;
  const int row_lo = kGtBand * rank - 1;                       // global row of staged row 0
#ifdef RDM_GT_TIMING
  long long tt[10]; int ti = 0; tt[ti++] = clock64();
#define GT_MARK() do { __syncthreads(); tt[ti++] = clock64(); } while (0)
#else
#define GT_MARK() do {} while (0)
#endif
  // ---- 1. resize (torch bicubic, align_corners=False, A=-0.75, clamped taps, horizontal first) + mask
  {
    const double sy = (double)ih / (double)kGtSide, sx = (double)iw / (double)kGtSide;
    const float m_inv = 1.0f + 1e-4f, m_val = 1e-4f;           // `(y <= 0) + 1e-4` is an f32 tensor (MOD:77)
    for (int i = tid; i < kGtStageRows * kGtSide; i += blockDim.x) {
      const int rr = i / kGtSide, ox = i - rr * kGtSide;
      const int oy = min(max(row_lo + rr, 0), kGtSide - 1);    // halo rows beyond the map repeat the border row (clamped taps of the stride-2 filter)
      const double fy = sy * ((double)oy + 0.5) - 0.5, fx = sx * ((double)ox + 0.5) - 0.5;
      const double fly = floor(fy), flx = floor(fx);
      double wy[4], wx[4];
      cubic_coeffs(fy - fly, wy);
      cubic_coeffs(fx - flx, wx);
      const int iy = (int)fly, ix = (int)flx;
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int r = min(max(iy - 1 + a, 0), ih - 1);
        double inner = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int c = min(max(ix - 1 + b, 0), iw - 1);
          const double v = (double)src[r * iw + c];
          inner = (b == 0) ? __dmul_rn(v, wx[0]) : fma(v, wx[b], inner);
        }
        acc = (a == 0) ? __dmul_rn(inner, wy[0]) : fma(inner, wy[a], acc);
      }
      // MOD:74-78: y = gt * (gt > 0) + ((gt <= 0) + 1e-4)
      const double y = __dadd_rn(__dmul_rn(acc, acc > 0.0 ? 1.0 : 0.0), (double)(acc <= 0.0 ? m_inv : m_val));
      sm.Y[i] = y;
      if (rr >= 1 && rr <= kGtBand) y_out[img * (kGtSide * kGtSide) + (row_lo + rr) * kGtSide + ox] = y;
    }
  }
  __syncthreads();
  GT_MARK();
  // ---- 2. row `rank` of cp.resize(y, 8) (scale 16: taps 16 rank + 6 .. + 9), utils.depth2label_sid
  int my_label = 0;
  if (tid < 8) {
    const double s8 = (double)kGtSide / 8.0;
    const double fy = s8 * ((double)rank + 0.5) - 0.5, fx = s8 * ((double)tid + 0.5) - 0.5;
    const double fly = floor(fy), flx = floor(fx);
    double wy[4], wx[4];
    cubic_coeffs(fy - fly, wy);
    cubic_coeffs(fx - flx, wx);
    const int iy = (int)fly, ix = (int)flx;
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = min(max(iy - 1 + a, 0), kGtSide - 1) - row_lo;   // staged row
      double inner = 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c = min(max(ix - 1 + b, 0), kGtSide - 1);
        const double v = sm.Y[r * kGtSide + c];
        inner = (b == 0) ? __dmul_rn(v, wx[0]) : fma(v, wx[b], inner);
      }
      acc = (a == 0) ? __dmul_rn(inner, wy[0]) : fma(inner, wy[a], acc);
    }
    // label = K * log(depth / alpha) / log(beta / alpha), max(label, 0), .int()  (f32 scalars, f64 tensor: utils.py:195-211)
    const double label = __ddiv_rn(__dmul_rn(sid_K, log(__ddiv_rn(acc, sid_alpha))), sid_log_ratio);
    const int lab = __double2int_rz(fmax(label, 0.0));         // a NaN label (negative depth) becomes 0 like torch's CUDA cast
    ord_out[img * 64 + rank * 8 + tid] = lab;
    my_label = lab;
  }
  GT_MARK();
  // ---- 3. geometric mean over the whole map (MOD:145-149, rc = 128: the true geometric mean) as exp(mean log)
  {
    double a0 = 0.0;
    for (int i = tid; i < kGtBand * kGtSide; i += blockDim.x) a0 += log(sm.Y[kGtSide + i]);
    const double mine = block_sum<double>(a0, sm.scratch);
    if (tid == 0) sm.part = mine;
  }
  GT_MARK();
  cluster.sync();
  GT_MARK();
  // every CTA of the cluster is running now: remote shared memory may be written (CTA 0 reads the labels after
  // band_decompose's cluster barrier)
  if (tid < 8) *cluster.map_shared_rank(&sm.labels[rank * 8 + tid], 0) = my_label;
  double gm = 0.0;
  for (int r = 0; r < kBandCluster; ++r) gm += *cluster.map_shared_rank(&sm.part, r);   // rank order: deterministic
  gm = exp(gm * (1.0 / ((double)kGtSide * (double)kGtSide)));
  for (int i = tid; i < kGtStageRows * kGtSide; i += blockDim.x) sm.Y[i] = sm.Y[i] / gm;
  __syncthreads();
  GT_MARK();
  // ---- 4. F_7 band, D_6 -> CTA 0, levels 6..1 (D_0 of the ground truth itself is not a target: MOD:127 replaces it)
  if (!band_decompose(sm, cluster, rank, kGtSide, 7, 1, pyr_out, n_images, img)) return;
  GT_MARK();
  // ---- 5. ordinal D_0: normalize(labels) in f32 like torch (int -> pow(., 1/64) f32 product, int / gm in f32), then the
  // 8x8 map decomposed in f64 (MOD:126)
  {
    const float pw = tid < 64 ? pow_as<float>((float)sm.labels[tid], 1.0 / 64.0) : 1.0f;
    const float gmf = block_prod<float>(pw, sm.fscratch);
    __syncthreads();
    if (tid < 64) sm.pyr[off_level(3) + tid] = (double)((float)sm.labels[tid] / gmf);
    __syncthreads();
    pyramid_down(sm.pyr, 3, [&](int, int, double) {});
    __syncthreads();
    if (tid == 0) pyr_out[img] = sm.pyr[0];
  }
#ifdef RDM_GT_TIMING
  GT_MARK();
  if (tid == 0 && blockIdx.x == 0) {
    printf("gt_prepare CTA 0:");
    for (int i = 1; i < ti; ++i) printf(" +%lld", tt[i] - tt[i - 1]);
    printf(" = %lld cycles (resize+mask | labels | logs+blocksum | cluster.sync | gm+divide | band_decompose incl. levels 6..1 | ordinal D0)\n", tt[ti - 1] - tt[0]);
  }
#endif
}

static int grid_cap(int64_t items, int per_block) {
  int64_t blocks = (items + per_block - 1) / per_block;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;   // 8 resident CTAs of 256 threads per SM, two waves
  return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace rdm

using namespace rdm;

static size_t smem_set_decompose_bwd_kernel_double_[64];
static size_t smem_set_decompose_bwd_kernel_float_[64];
static size_t smem_set_decompose_kernel_double_[64];
static size_t smem_set_decompose_kernel_float_[64];
static size_t smem_set_decompose_cluster_double_[64];
static size_t smem_set_decompose_cluster_float_[64];
static size_t smem_set_fuse_tail_kernel[64];
static size_t smem_set_recombination_bwd_kernel_double_[64];
static size_t smem_set_recombination_bwd_kernel_float_[64];

extern "C" int64_t rdm_pyramid_len(int32_t side, int32_t relative_map) {
  if (!is_pow2(side)) return -1;
  int n = ilog2(side);
  return (relative_map ? 0 : 1) + (((int64_t)1 << (2 * (n + 1))) - 4) / 3;
}

template <typename TIn, typename TAcc>
static int gm_launch(const void* t, int64_t batch, int64_t n, double e, void* gm_out, void* norm_out, rdm_stream_t stream) {
  if (n >= 4096 && n % kGmCluster == 0 && batch * kGmCluster < (1ll << 31)) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(batch * kGmCluster));
    cfg.blockDim = dim3(256);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kGmCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t err = cudaLaunchKernelEx(&cfg, gm_cluster_kernel<TIn, TAcc>, (const TIn*)t, n, e, (TAcc*)gm_out, (TAcc*)norm_out);
    if (err != cudaSuccess) {
      set_error("gm_cluster_kernel: %s", cudaGetErrorString(err));
      return (int)err;
    }
    return launch_status("gm_cluster_kernel");
  }
  gm_kernel<TIn, TAcc><<<(unsigned)batch, 256, 0, (cudaStream_t)stream>>>((const TIn*)t, n, e, (TAcc*)gm_out, (TAcc*)norm_out);
  return launch_status("gm_kernel");
}

static int gm_dispatch(const char* name, const void* t, int32_t dtype, int64_t batch, int64_t n, int32_t rc, void* gm_out,
                       void* norm_out, rdm_stream_t stream) {
  RDM_REQUIRE(t && (gm_out || norm_out), "%s: null pointer", name);
  RDM_REQUIRE(dtype >= 0 && dtype <= 2, "%s: dtype must be 0 (f32), 1 (f64) or 2 (i64)", name);
  RDM_REQUIRE(batch >= 0 && n >= 1 && rc >= 1, "%s: bad shape", name);
  RDM_REQUIRE(batch < (1ll << 31), "%s: batch too large", name);
  if (batch == 0) return 0;
  const double e = 1.0 / ((double)rc * (double)rc);
  if (dtype == 0) return gm_launch<float, float>(t, batch, n, e, gm_out, norm_out, stream);
  if (dtype == 1) return gm_launch<double, double>(t, batch, n, e, gm_out, norm_out, stream);
  return gm_launch<int64_t, float>(t, batch, n, e, gm_out, norm_out, stream);
}

extern "C" int rdm_quick_gm(const void* t, int32_t dtype, int64_t batch, int64_t n, int32_t rc, void* out, rdm_stream_t stream) {
  return gm_dispatch("rdm_quick_gm", t, dtype, batch, n, rc, out, nullptr, stream);
}

extern "C" int rdm_gm_normalize(const void* x, int32_t dtype, int64_t n_images, int32_t side, void* out, rdm_stream_t stream) {
  RDM_REQUIRE(side >= 1, "rdm_gm_normalize: bad side");
  return gm_dispatch("rdm_gm_normalize", x, dtype, n_images, (int64_t)side * side, side, nullptr, out, stream);
}

extern "C" int rdm_decompose(const void* in, int32_t in_is_f64, int64_t n_images, int32_t side, int32_t relative_map,
                             double* pyramid_out, rdm_stream_t stream) {
  RDM_REQUIRE(in && pyramid_out, "rdm_decompose: null pointer");
  RDM_REQUIRE(is_pow2(side) && side <= 128, "rdm_decompose: side must be a power of two <= 128 (got %d)", side);
  RDM_REQUIRE(n_images >= 0 && n_images < (1ll << 31), "rdm_decompose: bad n_images");
  if (n_images == 0) return 0;
  const int n = ilog2(side);
  const int64_t len = rdm_pyramid_len(side, relative_map);
  if (len == 0) return 0;
  const size_t smem = kPyrDoubles * sizeof(double);
  cudaError_t e;
  if (side >= 64) {
    // one cluster of 8 CTAs per image: banded top level, D_{n-1} gathered in CTA 0 through distributed shared memory
    RDM_REQUIRE(n_images * kBandCluster < (1ll << 31), "rdm_decompose: too many images");
    const size_t bsm = sizeof(BandSmem);
    e = in_is_f64 ? ensure_dyn_smem(decompose_cluster_kernel<double>, bsm, smem_set_decompose_cluster_double_)
                  : ensure_dyn_smem(decompose_cluster_kernel<float>, bsm, smem_set_decompose_cluster_float_);
    if (e != cudaSuccess) {
      set_error("rdm_decompose: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_images * kBandCluster));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = bsm;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kBandCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (in_is_f64)
      e = cudaLaunchKernelEx(&cfg, decompose_cluster_kernel<double>, (const double*)in, (int)side, n, (int)relative_map, pyramid_out, n_images);
    else
      e = cudaLaunchKernelEx(&cfg, decompose_cluster_kernel<float>, (const float*)in, (int)side, n, (int)relative_map, pyramid_out, n_images);
    if (e != cudaSuccess) {
      set_error("rdm_decompose: launch: %s", cudaGetErrorString(e));
      return (int)e;
    }
    return launch_status("decompose_cluster_kernel");
  }
  if (in_is_f64) {
    e = ensure_dyn_smem(decompose_kernel<double>, smem, smem_set_decompose_kernel_double_);
    if (e == cudaSuccess)
      decompose_kernel<double><<<(unsigned)n_images, 256, smem, (cudaStream_t)stream>>>((const double*)in, side, n, relative_map, pyramid_out, n_images);
  } else {
    e = ensure_dyn_smem(decompose_kernel<float>, smem, smem_set_decompose_kernel_float_);
    if (e == cudaSuccess)
      decompose_kernel<float><<<(unsigned)n_images, 256, smem, (cudaStream_t)stream>>>((const float*)in, side, n, relative_map, pyramid_out, n_images);
  }
  if (e != cudaSuccess) {
    set_error("rdm_decompose: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return launch_status("decompose_kernel");
}

extern "C" int rdm_decompose_bwd(const void* in, int32_t in_is_f64, int64_t n_images, int32_t side, int32_t relative_map,
                                 const double* grad_pyramid, void* grad_in, rdm_stream_t stream) {
  RDM_REQUIRE(in && grad_pyramid && grad_in, "rdm_decompose_bwd: null pointer");
  RDM_REQUIRE(is_pow2(side) && side >= 2 && side <= 128, "rdm_decompose_bwd: side must be a power of two in 2..128 (got %d)", side);
  RDM_REQUIRE(n_images >= 0 && n_images < (1ll << 31), "rdm_decompose_bwd: bad n_images");
  if (n_images == 0) return 0;
  const int n = ilog2(side);
  const size_t smem = 2 * kPyrDoubles * sizeof(double);
  cudaError_t e;
  if (in_is_f64) {
    e = ensure_dyn_smem(decompose_bwd_kernel<double>, smem, smem_set_decompose_bwd_kernel_double_);
    if (e == cudaSuccess)
      decompose_bwd_kernel<double><<<(unsigned)n_images, 256, smem, (cudaStream_t)stream>>>((const double*)in, side, n, relative_map, grad_pyramid, n_images, (double*)grad_in);
  } else {
    e = ensure_dyn_smem(decompose_bwd_kernel<float>, smem, smem_set_decompose_bwd_kernel_float_);
    if (e == cudaSuccess)
      decompose_bwd_kernel<float><<<(unsigned)n_images, 256, smem, (cudaStream_t)stream>>>((const float*)in, side, n, relative_map, grad_pyramid, n_images, (float*)grad_in);
  }
  if (e != cudaSuccess) {
    set_error("rdm_decompose_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return launch_status("decompose_bwd_kernel");
}

extern "C" int rdm_gm_bwd(const void* x, int32_t is_f64, int64_t batch, int64_t n, int32_t rc, const void* grad_gm,
                          const void* grad_norm, void* grad_x, rdm_stream_t stream) {
  RDM_REQUIRE(x && grad_x && (grad_gm || grad_norm), "rdm_gm_bwd: null pointer");
  RDM_REQUIRE(batch >= 0 && batch < (1ll << 31) && n >= 1 && rc >= 1, "rdm_gm_bwd: bad shape");
  if (batch == 0) return 0;
  const double e = 1.0 / ((double)rc * (double)rc);
  if (is_f64)
    gm_bwd_kernel<double><<<(unsigned)batch, 256, 0, (cudaStream_t)stream>>>((const double*)x, n, e, (const double*)grad_gm, (const double*)grad_norm, (double*)grad_x);
  else
    gm_bwd_kernel<float><<<(unsigned)batch, 256, 0, (cudaStream_t)stream>>>((const float*)x, n, e, (const float*)grad_gm, (const float*)grad_norm, (float*)grad_x);
  return launch_status("gm_bwd_kernel");
}

extern "C" int rdm_log_stack_bwd(const double* const* cands, int32_t K, int64_t batch, int64_t M, const double* grad_out,
                                 double* const* grad_cands, rdm_stream_t stream) {
  RDM_REQUIRE(cands && grad_out && grad_cands, "rdm_log_stack_bwd: null pointer");
  RDM_REQUIRE(K >= 1 && K <= kMaxPtrs, "rdm_log_stack_bwd: K must be 1..%d (got %d)", kMaxPtrs, K);
  RDM_REQUIRE(batch >= 0 && M >= 1, "rdm_log_stack_bwd: bad shape");
  if (batch == 0) return 0;
  PtrList pl{};
  MutPtrList gl{};
  for (int k = 0; k < K; ++k) {
    RDM_REQUIRE(cands[k] && grad_cands[k], "rdm_log_stack_bwd: candidate %d is null", k);
    pl.p[k] = cands[k];
    gl.p[k] = grad_cands[k];
  }
  log_stack_bwd_kernel<<<grid_cap(batch * K * M, 256), 256, 0, (cudaStream_t)stream>>>(pl, gl, K, batch, M, grad_out);
  return launch_status("log_stack_bwd_kernel");
}

extern "C" int rdm_log_stack_f64(const double* const* cands, int32_t K, int64_t batch, int64_t M, double* out, rdm_stream_t stream) {
  RDM_REQUIRE(cands && out, "rdm_log_stack_f64: null pointer");
  RDM_REQUIRE(K >= 1 && K <= kMaxPtrs, "rdm_log_stack_f64: K must be 1..%d (got %d)", kMaxPtrs, K);
  RDM_REQUIRE(batch >= 0 && M >= 1, "rdm_log_stack_f64: bad shape");
  if (batch == 0) return 0;
  PtrList pl{};
  for (int k = 0; k < K; ++k) {
    RDM_REQUIRE(cands[k], "rdm_log_stack_f64: candidate %d is null", k);
    pl.p[k] = cands[k];
  }
  log_stack_kernel<<<grid_cap(batch * K * M, 256), 256, 0, (cudaStream_t)stream>>>(pl, K, batch, M, out);
  return launch_status("log_stack_kernel");
}

extern "C" int rdm_make_pred_f32(const double* A, const float* w, int64_t batch, int32_t K, int64_t M, float* out, rdm_stream_t stream) {
  RDM_REQUIRE(A && w && out, "rdm_make_pred_f32: null pointer");
  RDM_REQUIRE(batch >= 0 && K >= 1 && M >= 1, "rdm_make_pred_f32: bad shape");
  if (batch == 0) return 0;
  make_pred_kernel<<<grid_cap(batch * M, 256), 256, 0, (cudaStream_t)stream>>>(A, w, batch, K, M, out);
  return launch_status("make_pred_kernel");
}

extern "C" int rdm_make_pred_bwd(const double* A, const float* w, const float* grad_out, int64_t batch, int32_t K, int64_t M,
                                 double* grad_A, float* grad_w, rdm_stream_t stream) {
  RDM_REQUIRE(grad_out && (grad_A || grad_w), "rdm_make_pred_bwd: null pointer");
  RDM_REQUIRE((!grad_A || w) && (!grad_w || A), "rdm_make_pred_bwd: grad_A needs w, grad_w needs A");
  RDM_REQUIRE(batch >= 0 && K >= 1 && M >= 1, "rdm_make_pred_bwd: bad shape");
  if (batch == 0) {
    if (grad_w) cudaMemsetAsync(grad_w, 0, sizeof(float) * K, (cudaStream_t)stream);
    return launch_status("rdm_make_pred_bwd memset");
  }
  if (grad_A) {
    make_pred_bwd_a_kernel<<<grid_cap(batch * K * M, 256), 256, 0, (cudaStream_t)stream>>>(w, grad_out, batch, K, M, grad_A);
    int rc = launch_status("make_pred_bwd_a_kernel");
    if (rc) return rc;
  }
  if (grad_w) {
    make_pred_bwd_w_kernel<<<(unsigned)K, 512, 0, (cudaStream_t)stream>>>(A, grad_out, batch, K, M, grad_w);
    return launch_status("make_pred_bwd_w_kernel");
  }
  return 0;
}

static int check_comp_sides(const char* name, const int32_t* sides, int32_t n_comps, int32_t n, int* has_d0) {
  RDM_REQUIRE(n_comps >= 1 && n_comps <= kMaxPtrs, "%s: n_comps must be 1..%d (got %d)", name, kMaxPtrs, n_comps);
  RDM_REQUIRE(n >= 1 && n <= 12, "%s: n must be 1..12 (got %d)", name, n);
  *has_d0 = sides[0] == 1;
  RDM_REQUIRE(n_comps > *has_d0, "%s: need at least one component besides d_0", name);
  // CP:405-418 upsamples the j-th remaining component n-1-j times: its side must be 2^(j+1)
  for (int j = *has_d0; j < n_comps; ++j) {
    int want = 2 << (j - *has_d0);
    RDM_REQUIRE(sides[j] == want && want <= (1 << n), "%s: component %d has side %d, expected %d (<= %d)", name, j, sides[j], want, 1 << n);
  }
  return 0;
}

extern "C" int rdm_recombination_f64(const void* const* comps, const int32_t* sides, int32_t n_comps, int32_t comps_are_f64,
                                     int64_t batch, int32_t n, double* out, rdm_stream_t stream) {
  RDM_REQUIRE(comps && sides && out, "rdm_recombination_f64: null pointer");
  int has_d0 = 0;
  if (check_comp_sides("rdm_recombination_f64", sides, n_comps, n, &has_d0)) return -1;
  RDM_REQUIRE(aligned16(out), "rdm_recombination_f64: out must be 16-byte aligned");
  RDM_REQUIRE(batch >= 0, "rdm_recombination_f64: bad batch");
  if (batch == 0) return 0;
  PtrList pl{};
  for (int j = 0; j < n_comps; ++j) {
    RDM_REQUIRE(comps[j], "rdm_recombination_f64: component %d is null", j);
    pl.p[j] = comps[j];
    pl.side[j] = sides[j];
  }
  if (n >= 2) {
    const int64_t items = batch * ((int64_t)1 << (2 * n)) / 4;
    const int grid = grid_cap(items, 256);
    if (comps_are_f64)
      recombination_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(pl, n_comps, has_d0, n, batch, out);
    else
      recombination_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(pl, n_comps, has_d0, n, batch, out);
  } else {
    const int64_t items = batch * ((int64_t)1 << (2 * n)) / 2;
    if (comps_are_f64)
      recombination_small_kernel<double><<<grid_cap(items, 256), 256, 0, (cudaStream_t)stream>>>(pl, n_comps, has_d0, n, batch, out);
    else
      recombination_small_kernel<float><<<grid_cap(items, 256), 256, 0, (cudaStream_t)stream>>>(pl, n_comps, has_d0, n, batch, out);
  }
  return launch_status("recombination_kernel");
}

extern "C" int rdm_recombination_bwd(const double* grad_out, void* const* grad_comps, const int32_t* sides, int32_t n_comps,
                                     int32_t comps_are_f64, int64_t batch, int32_t n, rdm_stream_t stream) {
  RDM_REQUIRE(grad_out && grad_comps && sides, "rdm_recombination_bwd: null pointer");
  int has_d0 = 0;
  if (check_comp_sides("rdm_recombination_bwd", sides, n_comps, n, &has_d0)) return -1;
  RDM_REQUIRE(n <= 7, "rdm_recombination_bwd: n must be <= 7 (got %d)", n);
  RDM_REQUIRE(batch >= 0 && batch < (1ll << 31), "rdm_recombination_bwd: bad batch");
  if (batch == 0) return 0;
  MutPtrList pl{};
  for (int j = 0; j < n_comps; ++j) {
    pl.p[j] = grad_comps[j];
    pl.side[j] = sides[j];
  }
  const size_t smem = kPyrDoubles * sizeof(double);
  cudaError_t e;
  if (comps_are_f64) {
    e = ensure_dyn_smem(recombination_bwd_kernel<double>, smem, smem_set_recombination_bwd_kernel_double_);
    if (e == cudaSuccess)
      recombination_bwd_kernel<double><<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(grad_out, pl, n_comps, n, batch);
  } else {
    e = ensure_dyn_smem(recombination_bwd_kernel<float>, smem, smem_set_recombination_bwd_kernel_float_);
    if (e == cudaSuccess)
      recombination_bwd_kernel<float><<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(grad_out, pl, n_comps, n, batch);
  }
  if (e != cudaSuccess) {
    set_error("rdm_recombination_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return launch_status("recombination_bwd_kernel");
}

extern "C" int64_t rdm_fuse_tail_weight_count(const int32_t* sides, int32_t n_rel) {
  if (n_rel < 0 || n_rel > kMaxRel) return -1;
  int64_t total = 4;   // decoder 1: d0, f1, f2, f3
  for (int r = 0; r < n_rel; ++r) {
    if (!is_pow2(sides[r]) || sides[r] < 2 || sides[r] > 64) return -1;
    total += ilog2(sides[r]);
  }
  return total;
}

static int fuse_tail_impl(const int64_t* x_d1, const float* const* rel, const int32_t* sides, int32_t n_rel, const float* weights,
                          int64_t n_images, float* yhat_out, double* depth_out, double* depth_compact_out, double* const* A_out,
                          int32_t bands_req, rdm_stream_t stream) {
  RDM_REQUIRE(x_d1 && weights && (depth_out || depth_compact_out), "rdm_fuse_tail: null pointer");
  RDM_REQUIRE(bands_req == 0 || bands_req == 1 || bands_req == 2 || bands_req == 4 || bands_req == 8,
              "rdm_fuse_tail_bands: bands must be 0 (chosen from the batch), 1, 2, 4 or 8 (got %d)", bands_req);
  RDM_REQUIRE(n_rel >= 0 && n_rel <= kMaxRel, "rdm_fuse_tail: n_rel must be 0..%d (got %d)", kMaxRel, n_rel);
  RDM_REQUIRE(n_rel == 0 || (rel && sides), "rdm_fuse_tail: rel/sides required");
  RDM_REQUIRE(!depth_out || aligned16(depth_out), "rdm_fuse_tail: depth_out must be 16-byte aligned");
  RDM_REQUIRE(n_images >= 0, "rdm_fuse_tail: bad n_images");
  if (n_images == 0) return 0;
  TailParams P{};
  P.x_d1 = x_d1;
  P.w = weights;
  P.yhat_out = yhat_out;
  P.depth_out = depth_out;
  P.depth_compact = depth_compact_out;
  P.n_rel = n_rel;
  P.kmax = 3;
  for (int k = 0; k < 8; ++k) {
    P.K[k] = (k <= 3) ? 1 : 0;
    P.cand[0][k] = 0;
    P.nact[k] = 0;
  }
  P.nlev[0] = 3;
  P.doff[0] = 0;
  int dtotal = off_level(4);
  for (int r = 0; r < n_rel; ++r) {
    RDM_REQUIRE(rel[r], "rdm_fuse_tail: rel[%d] is null", r);
    RDM_REQUIRE(is_pow2(sides[r]) && sides[r] >= 2 && sides[r] <= 64, "rdm_fuse_tail: rel side must be a power of two in 2..64 (got %d)", sides[r]);
    P.rel[r] = rel[r];
    P.side[r] = sides[r];
    const int n = ilog2(sides[r]);
    if (n > P.kmax) P.kmax = n;
    for (int k = 1; k <= n; ++k) P.cand[r + 1][k] = P.K[k]++;
    P.nlev[r + 1] = n;
    P.doff[r + 1] = dtotal;
    dtotal += off_level(n + 1);
  }
  P.dtotal = (dtotal + 1) & ~1;
  for (int k = 1; k <= P.kmax; ++k)
    for (int d = 0; d <= n_rel; ++d)
      if (P.nlev[d] >= k) P.act[k][P.nact[k]++] = d;
  int ltotal = 0;
  for (int k = P.kmax; k >= 1; --k) {   // log buffer: level kmax first
    P.lofs[k] = ltotal;
    ltotal += P.nact[k] << (2 * k);
  }
  P.lofs[0] = ltotal;
  P.ltotal = ltotal;
  int woff = 0;
  for (int k = 0; k < 8; ++k) {
    P.w_off[k] = woff;
    woff += P.K[k];
    P.A_out[k] = (A_out && k <= P.kmax) ? A_out[k] : nullptr;
  }
  // bands: a cluster of `bands` CTAs per image shortens a SMALL launch (the f64 div + log stage is split over the
  // cluster), but every band redoes the pyramids and a cluster has to be placed as a whole: with a reference batch of
  // 16 per call and 32 calls in flight, 4 bands (64 CTAs per call) cost 12 % of the path's throughput against 1 band
  // (measured with a temporary override of this count: 1.53 -> 1.72 M maps/s; the lone launch 11 -> 18 us).  So: >= 16 CTAs in total, no more.
  int bands = 1;   // a band must hold whole constant blocks: rows per band >= 2^(7-kmax)
  while (bands < 8 && bands < (1 << P.kmax) && n_images * bands < 16) bands <<= 1;   // one thread-block cluster per image
  if (bands_req) bands = bands_req < (1 << P.kmax) ? bands_req : (1 << P.kmax);          // the caller knows it is alone on the GPU
  P.bands = bands;
  RDM_REQUIRE(n_images * bands < (1ll << 31), "rdm_fuse_tail: too many images");
  const size_t smem = (size_t)P.dtotal * sizeof(double) + ((size_t)off_level(P.kmax + 1) + (size_t)ltotal) * sizeof(float);
  RDM_REQUIRE(smem <= 220 * 1024, "rdm_fuse_tail: decoder pyramids need %zu bytes of shared memory (max 220 KB)", smem);
  cudaError_t e = ensure_dyn_smem(fuse_tail_kernel, smem, smem_set_fuse_tail_kernel);
  if (e != cudaSuccess) {
    set_error("rdm_fuse_tail: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_images * bands));
  cfg.blockDim = dim3(RDM_TAIL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)bands;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, fuse_tail_kernel, P);
  if (e != cudaSuccess) {
    set_error("fuse_tail_kernel: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return launch_status("fuse_tail_kernel");
}

extern "C" int rdm_fuse_tail(const int64_t* x_d1, const float* const* rel, const int32_t* sides, int32_t n_rel,
                             const float* weights, int64_t n_images, float* yhat_out, double* depth_out,
                             double* depth_compact_out, double* const* A_out, rdm_stream_t stream) {
  return fuse_tail_impl(x_d1, rel, sides, n_rel, weights, n_images, yhat_out, depth_out, depth_compact_out, A_out, 0, stream);
}

extern "C" int rdm_fuse_tail_bands(const int64_t* x_d1, const float* const* rel, const int32_t* sides, int32_t n_rel,
                                   const float* weights, int64_t n_images, float* yhat_out, double* depth_out,
                                   double* depth_compact_out, double* const* A_out, int32_t bands, rdm_stream_t stream) {
  return fuse_tail_impl(x_d1, rel, sides, n_rel, weights, n_images, yhat_out, depth_out, depth_compact_out, A_out, bands, stream);
}

static size_t smem_set_gt_prepare_float_[64];
static size_t smem_set_gt_prepare_double_[64];

extern "C" int rdm_gt_prepare(const void* y_raw, int32_t in_is_f64, int64_t n_images, int32_t in_h, int32_t in_w, double sid_K,
                              double sid_alpha, double sid_log_ratio, double* y_out, double* pyramid_out, int32_t* ord_target_out,
                              rdm_stream_t stream) {
  RDM_REQUIRE(y_raw && y_out && pyramid_out && ord_target_out, "rdm_gt_prepare: null pointer");
  RDM_REQUIRE(in_h >= 1 && in_w >= 1 && in_h <= 4096 && in_w <= 4096, "rdm_gt_prepare: bad input size %d x %d", in_h, in_w);
  RDM_REQUIRE(n_images >= 0 && n_images * kBandCluster < (1ll << 31), "rdm_gt_prepare: bad n_images");
  RDM_REQUIRE(sid_alpha > 0.0 && sid_log_ratio != 0.0, "rdm_gt_prepare: bad SID parameters");
  if (n_images == 0) return 0;
  const size_t smem = sizeof(BandSmem);
  cudaError_t e = in_is_f64 ? ensure_dyn_smem(gt_prepare_kernel<double>, smem, smem_set_gt_prepare_double_)
                            : ensure_dyn_smem(gt_prepare_kernel<float>, smem, smem_set_gt_prepare_float_);
  if (e != cudaSuccess) {
    set_error("rdm_gt_prepare: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_images * kBandCluster));
  cfg.blockDim = dim3(RDM_GT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kBandCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (in_is_f64)
    e = cudaLaunchKernelEx(&cfg, gt_prepare_kernel<double>, (const double*)y_raw, (int)in_h, (int)in_w, n_images, sid_K, sid_alpha,
                           sid_log_ratio, y_out, pyramid_out, ord_target_out);
  else
    e = cudaLaunchKernelEx(&cfg, gt_prepare_kernel<float>, (const float*)y_raw, (int)in_h, (int)in_w, n_images, sid_K, sid_alpha,
                           sid_log_ratio, y_out, pyramid_out, ord_target_out);
  if (e != cudaSuccess) {
    set_error("rdm_gt_prepare: launch: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return launch_status("gt_prepare_kernel");
}

// ---------------------------------------------------------------------------------------------
// Training step, the two places where the per-slot ops left a trail of small launches.
namespace rdm {

struct TailBwdLevels {
  const double* A[8];   // (B, K_k, 4^k) f64
  int32_t K[8];
  int32_t first[9];     // first[k] = K_0 + ... + K_{k-1}
};

// grad_w of every slot in one launch: CTA = one weight; the arithmetic (thread-strided f64 FMA, warp tree, warps in
// order) is make_pred_bwd_w_kernel's, so the result is bit-identical to the slot-by-slot chain.
__global__ void __launch_bounds__(512) make_pred_bwd_w_all_kernel(const __grid_constant__ TailBwdLevels lv, int kmax, int64_t batch,
                                                                  const float* __restrict__ gs, float* __restrict__ grad_w) {
  __shared__ double part[16];
  int k = 0;
  while (k < kmax && (int)blockIdx.x >= lv.first[k + 1]) ++k;
  const int j = (int)blockIdx.x - lv.first[k], K = lv.K[k];
  const int64_t M = (int64_t)1 << (2 * k);
  const double* __restrict__ A = lv.A[k];
  const float* __restrict__ g = gs + batch * off_level(k);   // level-major: level k is (B, 4^k)
  double acc = 0.0;
  for (int64_t o = threadIdx.x; o < batch * M; o += blockDim.x) {
    const int64_t b = o / M, m = o - b * M;
    acc = fma((double)(float)A[(b * K + j) * M + m], (double)g[o], acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    grad_w[blockIdx.x] = (float)t;
  }
}

// CP:499-510 (the per-scale component loss of the training step): sum_k mean((yhat_k - target_k)^2) in f64, one CTA,
// levels in order.  yhat is image-major (B, sum 4^k) f32, the targets are the level-major pyramid of rdm_gt_prepare /
// rdm_decompose(relative_map = 0).
__global__ void __launch_bounds__(1024) component_loss_kernel(const float* __restrict__ yhat, const double* __restrict__ target, int64_t batch,
                                                              int kmax, double* __restrict__ out) {
  __shared__ double part[32];
  const int64_t row = off_level(kmax + 1);   // floats of yhat per image
  double total = 0.0;
  for (int k = 0; k <= kmax; ++k) {
    const int64_t M = (int64_t)1 << (2 * k), n = batch * M;
    const double* __restrict__ t = target + batch * off_level(k);
    double acc = 0.0;
    for (int64_t o = threadIdx.x; o < n; o += blockDim.x) {
      const int64_t b = o / M, m = o - b * M;
      const double d = (double)yhat[b * row + off_level(k) + m] - t[o];
      acc = fma(d, d, acc);
    }
    acc = warp_sum(acc);
    __syncthreads();   // part[] of the previous level has been read
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < 32; ++i) s += part[i];
      total += s / (double)n;
    }
  }
  if (threadIdx.x == 0) out[0] = total;
}

}  // namespace rdm

extern "C" int rdm_fuse_tail_bwd(const double* grad_depth, const double* const* A, const int32_t* K, int32_t kmax, int64_t n_images,
                                 float* ws, float* grad_w, rdm_stream_t stream) {
  RDM_REQUIRE(grad_depth && A && K && ws && grad_w, "rdm_fuse_tail_bwd: null pointer");
  RDM_REQUIRE(kmax >= 3 && kmax <= 6, "rdm_fuse_tail_bwd: kmax must be 3..6 (got %d)", kmax);
  RDM_REQUIRE(n_images >= 0 && n_images < (1ll << 31), "rdm_fuse_tail_bwd: bad n_images");
  TailBwdLevels lv{};
  int total = 0;
  for (int k = 0; k <= kmax; ++k) {
    RDM_REQUIRE(A[k] && K[k] >= 1, "rdm_fuse_tail_bwd: slot %d: null A or K < 1", k);
    lv.A[k] = A[k];
    lv.K[k] = K[k];
    lv.first[k] = total;
    total += K[k];
  }
  lv.first[kmax + 1] = total;
  if (n_images == 0) {
    cudaMemsetAsync(grad_w, 0, sizeof(float) * total, (cudaStream_t)stream);
    return launch_status("rdm_fuse_tail_bwd memset");
  }
  // pooled gradients of the recombination (128x128 output, n = 7), f32 like the per-slot chain, level-major in ws
  void* gp[8];
  int32_t sides[8];
  for (int k = 0; k <= kmax; ++k) {
    gp[k] = ws + n_images * off_level(k);
    sides[k] = 1 << k;
  }
  int rc = rdm_recombination_bwd(grad_depth, gp, sides, kmax + 1, 0, n_images, 7, stream);
  if (rc) return rc;
  make_pred_bwd_w_all_kernel<<<(unsigned)total, 512, 0, (cudaStream_t)stream>>>(lv, kmax, n_images, ws, grad_w);
  return launch_status("make_pred_bwd_w_all_kernel");
}

extern "C" int rdm_component_loss(const float* yhat, const double* target, int64_t n_images, int32_t kmax, double* out,
                                  rdm_stream_t stream) {
  RDM_REQUIRE(yhat && target && out, "rdm_component_loss: null pointer");
  RDM_REQUIRE(kmax >= 0 && kmax <= 7, "rdm_component_loss: kmax must be 0..7 (got %d)", kmax);
  RDM_REQUIRE(n_images >= 1 && n_images < (1ll << 31), "rdm_component_loss: bad n_images");
  component_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(yhat, target, n_images, kmax, out);
  return launch_status("component_loss_kernel");
}
