// SURVEY 8f "next", rank 1: the step immediately before the fusion path (it produces x_d1) and the
// other loss of the training step.
//   dorn_regression (+bwd) : RN:313-345 Ordinal_Layer.DornOrdinalRegression
//   ordinal_loss (+bwd)    : loss.py:17-59 Ordinal_Loss.calc
// Both are streaming kernels bounded by HBM bandwidth: x (N,2K,H,W) f32 is read once, the ordinal
// probabilities (N,K,H,W) f64 written once; the reference materialises six (N,K,H,W)-sized
// temporaries (clones, cat, clamp, double, softmax, clone) and, in the loss, an (N,K,H,W) index
// tensor filled by a Python loop over K plus two boolean masks.
#include "rdm_common.cuh"

namespace rdm {


static int grid_cap2(int64_t items) {
  int64_t blocks = (items + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  return blocks < 1 ? 1 : (int)blocks;
}



// torch.clamp(x, min=1e-8, max=1e4) on an f32 tensor: the bounds become f32 scalars, NaN stays NaN
__device__ __forceinline__ float clamp_f32(float v) {
  return v != v ? v : fminf(fmaxf(v, 1e-8f), 1e4f);
}

// softmax over the pair (A, B) as ATen computes it: exp(x - max) / sum, in f64; returns P(B).
__device__ __forceinline__ double pair_softmax_b(double a, double b) {
  if (a != a || b != b) return a + b;   // a NaN logit makes the whole pair NaN, as ATen's softmax does
  const double m = fmax(a, b);
  const double ea = exp(a - m), eb = exp(b - m);
  return eb / (ea + eb);
}

// ord[n,k,hw] = softmax(clamp(x[n,2k,hw]), clamp(x[n,2k+1,hw]))[1] (f64): one element per thread,
// hw fastest in both tensors (coalesced).  Two f64 exp() per element: the kernel is bound by the FP64
// pipe, not by HBM, so it is spread over the whole chip rather than one CTA per image.
__global__ void __launch_bounds__(256) dorn_ord_kernel(const float* __restrict__ x, uint32_t total, uint32_t K, uint32_t HW,
                                                       double* __restrict__ ord) {
  for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gridDim.x * blockDim.x) {
    const uint32_t hw = o % HW, nk = o / HW;
    const uint32_t k = nk % K, n = nk / K;
    const size_t ia = ((size_t)n * 2 * K + 2 * k) * HW + hw;
    // RN:334 clamps the f32 tensor (bounds rounded to f32) and only then widens it; torch.clamp propagates NaN
    const double a = (double)clamp_f32(x[ia]), b = (double)clamp_f32(x[ia + HW]);
    ord[o] = pair_softmax_b(a, b);
  }
}

// decode[n,hw] = #{k : ord[n,k,hw] > 0.5} (RN:342): one thread per (n, hw), k strided by HW.
__global__ void __launch_bounds__(256) dorn_decode_kernel(const double* __restrict__ ord, uint32_t total_nhw, uint32_t K, uint32_t HW,
                                                          int64_t* __restrict__ decode) {
  for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < total_nhw; o += gridDim.x * blockDim.x) {
    const uint32_t hw = o % HW, n = o / HW;
    const double* p = ord + (size_t)n * K * HW + hw;
    int c = 0;
    for (uint32_t k = 0; k < K; ++k) c += (p[(size_t)k * HW] > 0.5) ? 1 : 0;
    decode[o] = c;
  }
}

// backward: d ord / dB = ord (1 - ord), d ord / dA = -ord (1 - ord), gated by the clamp (RN:334).
__global__ void __launch_bounds__(256) dorn_regression_bwd_kernel(const float* __restrict__ x, const double* __restrict__ ord,
                                                                  const double* __restrict__ g_ord, int64_t total, int K, int HW,
                                                                  float* __restrict__ gx) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t o32 = (uint32_t)o;   // total < 2^31 is checked on the host: 32-bit index arithmetic
    const int hw = (int)(o32 % (uint32_t)HW);
    const uint32_t nk = o32 / (uint32_t)HW;
    const int k = (int)(nk % (uint32_t)K);
    const int64_t n = nk / (uint32_t)K;
    const int64_t ia = (n * 2 * K + 2 * k) * HW + hw, ib = ia + HW;
    const double p = ord[o];
    const double g = g_ord[o] * p * (1.0 - p);
    // clamp backward passes the gradient where min <= x <= max (f32 bounds); a NaN logit has a NaN ord, hence a NaN g
    const float a = x[ia], b = x[ib];
    gx[ia] = (a != a) ? (float)(-g) : ((a >= 1e-8f && a <= 1e4f) ? (float)(-g) : 0.f);
    gx[ib] = (b != b) ? (float)g : ((b >= 1e-8f && b <= 1e4f) ? (float)g : 0.f);
  }
}

// loss.py:41-57: -( sum_{k<=t} log(clamp(ord).float()) + sum_{k>t} log(clamp(1-ord).float()) ) / (N H W)
__global__ void __launch_bounds__(256) ordinal_loss_kernel(const double* __restrict__ ord, const int32_t* __restrict__ target, int64_t total,
                                                           int K, int HW, double* __restrict__ partial) {
  __shared__ double part[8];
  double acc = 0.0;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t o32 = (uint32_t)o;   // total < 2^31 is checked on the host: 32-bit index arithmetic
    const int hw = (int)(o32 % (uint32_t)HW);
    const uint32_t nk = o32 / (uint32_t)HW;
    const int k = (int)(nk % (uint32_t)K);
    const int64_t n = nk / (uint32_t)K;
    const int t = target[n * HW + hw];
    const double p = ord[o];
    const double v = (k <= t) ? p : 1.0 - p;
    acc += (double)logf((float)fmin(fmax(v, 1e-8), 1e8));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    partial[blockIdx.x] = t;
  }
}

__global__ void ordinal_loss_finish_kernel(const double* __restrict__ partial, int n_partial, double scale, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n_partial; ++i) t += partial[i];   // fixed order: deterministic
    *loss = (float)(t * scale);
  }
}

__global__ void __launch_bounds__(256) ordinal_loss_bwd_kernel(const double* __restrict__ ord, const int32_t* __restrict__ target,
                                                               const float* __restrict__ g_loss, int64_t total, int K, int HW, double scale,
                                                               double* __restrict__ g_ord) {
  const double g = (double)g_loss[0] * scale;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t o32 = (uint32_t)o;   // total < 2^31 is checked on the host: 32-bit index arithmetic
    const int hw = (int)(o32 % (uint32_t)HW);
    const uint32_t nk = o32 / (uint32_t)HW;
    const int k = (int)(nk % (uint32_t)K);
    const int64_t n = nk / (uint32_t)K;
    const int t = target[n * HW + hw];
    const double p = ord[o];
    const double v = (k <= t) ? p : 1.0 - p;
    // d/dv log(float(clamp(v))) = 1/v inside the clamp range, 0 outside; dv/dp = +1 (k<=t) or -1
    double d = (v >= 1e-8 && v <= 1e8) ? 1.0 / (double)(float)v : 0.0;
    g_ord[o] = g * ((k <= t) ? d : -d);
  }
}

// utils.py:195-211 depth2label_sid: label = K * log(depth / alpha) / log(beta / alpha), max(label, 0), .int().  The
// reference holds K, alpha, beta as 0-dim f32 tensors, so an f64 depth map is processed in f64 with f32-rounded
// scalars and an f32 map entirely in f32; log_ratio = log(beta / alpha) is evaluated by the caller (torch, f32).
template <typename T>
__global__ void __launch_bounds__(256) depth2label_kernel(const T* __restrict__ depth, int64_t n, double K, double alpha, double log_ratio,
                                                          int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if constexpr (sizeof(T) == 8) {
      const double label = __ddiv_rn(__dmul_rn(K, log(__ddiv_rn(depth[i], alpha))), log_ratio);
      out[i] = __double2int_rz(fmax(label, 0.0));
    } else {
      const float label = __fdiv_rn(__fmul_rn((float)K, logf(__fdiv_rn(depth[i], (float)alpha))), (float)log_ratio);
      out[i] = __float2int_rz(fmaxf(label, 0.0f));
    }
  }
}

}  // namespace rdm

using namespace rdm;

extern "C" int rdm_dorn_regression_f32(const float* x, int64_t n_images, int32_t K, int32_t HW, int64_t* decode_out, double* ord_out,
                                       rdm_stream_t stream) {
  RDM_REQUIRE(x && decode_out && ord_out, "rdm_dorn_regression_f32: null pointer");
  RDM_REQUIRE(K >= 1 && HW >= 1 && HW <= 8192, "rdm_dorn_regression_f32: bad K / HW");
  RDM_REQUIRE(n_images >= 0 && n_images < (1ll << 31), "rdm_dorn_regression_f32: bad n_images");
  if (n_images == 0) return 0;
  const int64_t total = n_images * (int64_t)K * HW;
  RDM_REQUIRE(total < (1ll << 31), "rdm_dorn_regression_f32: N*K*H*W must be below 2^31");
  dorn_ord_kernel<<<grid_cap2(total), 256, 0, (cudaStream_t)stream>>>(x, (uint32_t)total, (uint32_t)K, (uint32_t)HW, ord_out);
  int rc = launch_status("dorn_ord_kernel");
  if (rc) return rc;
  dorn_decode_kernel<<<grid_cap2(n_images * HW), 256, 0, (cudaStream_t)stream>>>(ord_out, (uint32_t)(n_images * HW), (uint32_t)K,
                                                                                 (uint32_t)HW, decode_out);
  return launch_status("dorn_decode_kernel");
}

extern "C" int rdm_dorn_regression_bwd(const float* x, const double* ord, const double* grad_ord, int64_t n_images, int32_t K, int32_t HW,
                                       float* grad_x, rdm_stream_t stream) {
  RDM_REQUIRE(x && ord && grad_ord && grad_x, "rdm_dorn_regression_bwd: null pointer");
  RDM_REQUIRE(K >= 1 && HW >= 1 && n_images >= 0, "rdm_dorn_regression_bwd: bad shape");
  if (n_images == 0) return 0;
  const int64_t total = n_images * K * HW;
  RDM_REQUIRE(total < (1ll << 31), "rdm_dorn_regression_bwd: N*K*H*W must be below 2^31");
  dorn_regression_bwd_kernel<<<grid_cap2(total), 256, 0, (cudaStream_t)stream>>>(x, ord, grad_ord, total, K, HW, grad_x);
  return launch_status("dorn_regression_bwd_kernel");
}

extern "C" int64_t rdm_ordinal_loss_ws_doubles(void) { return (int64_t)kNumSMs * 8; }

extern "C" int rdm_ordinal_loss_f64(const double* ord, const int32_t* target, int64_t n_images, int32_t K, int32_t HW, double* ws,
                                    float* loss_out, rdm_stream_t stream) {
  RDM_REQUIRE(ord && target && ws && loss_out, "rdm_ordinal_loss_f64: null pointer");
  RDM_REQUIRE(K >= 1 && HW >= 1 && n_images >= 1, "rdm_ordinal_loss_f64: bad shape");
  const int64_t total = n_images * K * HW;
  RDM_REQUIRE(total < (1ll << 31), "rdm_ordinal_loss_f64: N*K*H*W must be below 2^31");
  const int grid = grid_cap2(total);
  ordinal_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ord, target, total, K, HW, ws);
  int rc = launch_status("ordinal_loss_kernel");
  if (rc) return rc;
  ordinal_loss_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ws, grid, -1.0 / ((double)n_images * HW), loss_out);
  return launch_status("ordinal_loss_finish_kernel");
}

extern "C" int rdm_ordinal_loss_bwd(const double* ord, const int32_t* target, const float* grad_loss, int64_t n_images, int32_t K,
                                    int32_t HW, double* grad_ord, rdm_stream_t stream) {
  RDM_REQUIRE(ord && target && grad_loss && grad_ord, "rdm_ordinal_loss_bwd: null pointer");
  RDM_REQUIRE(K >= 1 && HW >= 1 && n_images >= 1, "rdm_ordinal_loss_bwd: bad shape");
  const int64_t total = n_images * K * HW;
  RDM_REQUIRE(total < (1ll << 31), "rdm_ordinal_loss_bwd: N*K*H*W must be below 2^31");
  ordinal_loss_bwd_kernel<<<grid_cap2(total), 256, 0, (cudaStream_t)stream>>>(ord, target, grad_loss, total, K, HW,
                                                                              -1.0 / ((double)n_images * HW), grad_ord);
  return launch_status("ordinal_loss_bwd_kernel");
}

extern "C" int rdm_depth2label_sid(const void* depth, int32_t is_f64, int64_t n, double sid_K, double sid_alpha, double sid_log_ratio,
                                   int32_t* labels_out, rdm_stream_t stream) {
  RDM_REQUIRE(n >= 0, "rdm_depth2label_sid: negative n");
  if (n == 0) return 0;
  RDM_REQUIRE(depth && labels_out, "rdm_depth2label_sid: null pointer");
  RDM_REQUIRE(sid_alpha > 0.0 && sid_log_ratio != 0.0, "rdm_depth2label_sid: bad SID parameters");
  int64_t blocks = (n + 255) / 256;
  if (blocks > rdm::kNumSMs * 8) blocks = rdm::kNumSMs * 8;
  if (is_f64)
    rdm::depth2label_kernel<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const double*)depth, n, sid_K, sid_alpha, sid_log_ratio, labels_out);
  else
    rdm::depth2label_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)depth, n, sid_K, sid_alpha, sid_log_ratio, labels_out);
  return rdm::launch_status("depth2label_kernel");
}
