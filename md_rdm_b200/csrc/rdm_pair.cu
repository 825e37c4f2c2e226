// Stage 1 (pair-matrix build) and stage 2 (stand-alone Lloyd quantisation) kernels.
//
// All three are streaming kernels bounded by HBM bandwidth: O(1) flop per byte,
// 128-bit coalesced stores, inputs a few hundred bytes per image.
//   pair_v1  : RN:244-252   (N,64) f32            -> (N,64,64) f32      16 KB / image
//   pair_id  : RN:259-280   (N,s,s) f32           -> (N,P,256,64) f64   128 KB / page
//   lloyd    : RN:286-311   n values f32|f64      -> values + u8 bins
#include "rdm_common.cuh"

namespace rdm {

// ---------------------------------------------------------------------------------------------
// pair_v1: raw[i][j] = fl32(d_i * fl32(1/d_j)).  torch.pow(x,-1) is the IEEE reciprocal and the
// K=1 matmul is a plain product, hence __frcp_rn / __fmul_rn and no contraction.
__global__ void __launch_bounds__(256) pair_v1_kernel(const float* __restrict__ d3, float* __restrict__ raw,
                                                      int64_t n_images) {
  __shared__ float d_s[64];
  __shared__ __align__(16) float inv_s[64];
  for (int64_t img = blockIdx.x; img < n_images; img += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < 64) {
      float v = d3[img * 64 + threadIdx.x];
      d_s[threadIdx.x] = v;
      inv_s[threadIdx.x] = __frcp_rn(v);
    }
    __syncthreads();
    float* out = raw + img * 4096;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int c = threadIdx.x + 256 * k;   // float4 chunk id, 1024 per image
      int i = c >> 4, j = (c & 15) << 2;
      float di = d_s[i];
      float4 iv = *reinterpret_cast<const float4*>(&inv_s[j]);
      float4 o;
      o.x = __fmul_rn(di, iv.x);
      o.y = __fmul_rn(di, iv.y);
      o.z = __fmul_rn(di, iv.z);
      o.w = __fmul_rn(di, iv.w);
      stg_stream_f32x4(out + (i << 6) + j, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// resize_half: one thread per output pixel.
template <typename TIn>
__global__ void __launch_bounds__(256) resize_half_kernel(const TIn* __restrict__ in, double* __restrict__ out,
                                                          int64_t n_images, int side) {
  const int half = side >> 1;
  const int64_t per = (int64_t)half * half;
  const int64_t total = n_images * per;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t img = idx / per;
    int rem = (int)(idx - img * per);
    int y = rem / half, x = rem - y * half;
    const TIn* src = in + img * (int64_t)side * side;
    out[idx] = bicubic_half_at([&](int r, int c) { return (double)src[r * side + c]; }, y, x, side);
  }
}

// ---------------------------------------------------------------------------------------------
// pair_id: one CTA per (image, page).  The 8x8 parent page is the bicubic half of the FULL map
// restricted to the page (taps clamp at the full-map border, RN:386 + CP:214), so it is computed
// here from the map directly.  Row r*16+c of the page matrix is f64(dn[r,c]) everywhere except
// the 3x3 window anchored at (min(r/2,5), min(c/2,5)) where it is f64(dn[r,c]) * fl64(1/dn_1).
__global__ void __launch_bounds__(256) pair_id_kernel(const float* __restrict__ dn, double* __restrict__ raw,
                                                      double* __restrict__ parent_out, int side, int ratio) {
  __shared__ double d_s[256];
  __shared__ double inv_s[64];
  const int pages = ratio * ratio;
  const int64_t img = blockIdx.x / pages;
  const int pg = blockIdx.x - (int)(img * pages);
  const int pi = pg / ratio, pj = pg - pi * ratio;
  const float* map = dn + img * (int64_t)side * side;
  const int t = threadIdx.x;
  {
    int r = t >> 4, c = t & 15;
    d_s[t] = (double)map[(16 * pi + r) * side + 16 * pj + c];
  }
  if (t < 64) {
    int y = 8 * pi + (t >> 3), x = 8 * pj + (t & 7);
    double v = bicubic_half_at([&](int r, int c) { return (double)map[r * side + c]; }, y, x, side);
    if (parent_out) parent_out[img * (int64_t)(side / 2) * (side / 2) + y * (side / 2) + x] = v;
    inv_s[t] = 1.0 / v;   // torch.pow(area,-1) == IEEE 1/x (correctly rounded f64 division)
  }
  __syncthreads();
  double* out = raw + ((img * pages + pg) << 14);
  const int warp = t >> 5, lane = t & 31;
  const int col = lane * 2;
  const double i0 = inv_s[col], i1 = inv_s[col + 1];
#pragma unroll 4
  for (int it = 0; it < 32; ++it) {
    int row = warp + 8 * it;
    double d = d_s[row];
    double a = in_window(row, col) ? __dmul_rn(d, i0) : d;
    double b = in_window(row, col + 1) ? __dmul_rn(d, i1) : d;
    stg_stream_f64x2(out + (row << 6) + col, a, b);
  }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone Lloyd quantisation (the fused path quantises inside the ALS kernel instead).
template <typename T>
__global__ void __launch_bounds__(256) lloyd_kernel(const T* x, int64_t n, const double* __restrict__ thr,
                                                    const double* __restrict__ lvl, T* values,   // x may alias values (in place)
                                                    uint8_t* __restrict__ bins) {
  __shared__ T thr_s[kThrPad];
  __shared__ T lvl_s[kLvl];
  __shared__ int sorted_s;
  load_codebook<T>(thr, lvl, thr_s, lvl_s, &sorted_s, threadIdx.x, blockDim.x);
  const int sorted = sorted_s;
  constexpr int V = 16 / sizeof(T);   // elements per 128-bit access
  const int64_t nvec = n / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += stride) {
    T v[V];
    if constexpr (sizeof(T) == 4) {
      float4 q = *reinterpret_cast<const float4*>(x + i * 4);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      double2 q = *reinterpret_cast<const double2*>(x + i * 2);
      v[0] = q.x; v[1] = q.y;
    }
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      int b = lloyd_bin<T>(v[e], thr_s, sorted);
      packed |= (uint32_t)b << (8 * e);
      v[e] = lvl_s[b];
    }
    if (values) {
      if constexpr (sizeof(T) == 4)
        *reinterpret_cast<float4*>(values + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
      else
        *reinterpret_cast<double2*>(values + i * 2) = make_double2(v[0], v[1]);
    }
    if (bins) {
      if constexpr (sizeof(T) == 4)
        *reinterpret_cast<uint32_t*>(bins + i * 4) = packed;
      else
        *reinterpret_cast<uint16_t*>(bins + i * 2) = (uint16_t)packed;
    }
  }
  // scalar tail
  for (int64_t i = nvec * V + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    int b = lloyd_bin<T>(x[i], thr_s, sorted);
    if (values) values[i] = lvl_s[b];
    if (bins) bins[i] = (uint8_t)b;
  }
}

// ---------------------------------------------------------------------------------------------
// pair matrix of caller-supplied pages (the literal signature of RN:259 sparse_comparison_id):
// pages (n,256) f32 and parent pages (n,64) f64, any values.  Same emit as pair_id_kernel.
__global__ void __launch_bounds__(256) pair_pages_kernel(const float* __restrict__ pages, const double* __restrict__ parents,
                                                         double* __restrict__ raw) {
  __shared__ double d_s[256];
  __shared__ double inv_s[64];
  const int64_t pg = blockIdx.x;
  const int t = threadIdx.x;
  d_s[t] = (double)pages[pg * 256 + t];
  if (t < 64) inv_s[t] = 1.0 / parents[pg * 64 + t];
  __syncthreads();
  double* out = raw + (pg << 14);
  const int warp = t >> 5, lane = t & 31;
  const int col = lane * 2;
  const double i0 = inv_s[col], i1 = inv_s[col + 1];
#pragma unroll 4
  for (int it = 0; it < 32; ++it) {
    int row = warp + 8 * it;
    double d = d_s[row];
    double a = in_window(row, col) ? __dmul_rn(d, i0) : d;
    double b = in_window(row, col + 1) ? __dmul_rn(d, i1) : d;
    stg_stream_f64x2(out + (row << 6) + col, a, b);
  }
}

// ---------------------------------------------------------------------------------------------
// CP:357-366 upsample / multi_upsample: `.double()` + nearest x2, `times` times.
template <typename TIn>
__global__ void __launch_bounds__(256) upsample_nearest_kernel(const TIn* __restrict__ in, double* __restrict__ out, int64_t n_maps,
                                                               int side, int times) {
  const int S = side << times;
  const int64_t per = (int64_t)S * S;
  const int64_t total = n_maps * per;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = idx / per;
    int64_t rem = idx - m * per;
    int y = (int)(rem / S), x = (int)(rem - (int64_t)y * S);
    out[idx] = (double)in[m * side * side + (y >> times) * side + (x >> times)];
  }
}

// ---------------------------------------------------------------------------------------------
// CP:308-311 cp.resize for an arbitrary target size (network/module.py:68 resizes 226 -> 128):
// torch bicubic, align_corners=False, A = -0.75, clamped taps, f64 arithmetic, horizontal taps
// summed first.

template <typename TIn>
__global__ void __launch_bounds__(256) resize_bicubic_kernel(const TIn* __restrict__ in, double* __restrict__ out, int64_t n_maps,
                                                             int ih, int iw, int oh, int ow) {
  const double sy = (double)ih / (double)oh, sx = (double)iw / (double)ow;
  const int64_t per = (int64_t)oh * ow;
  const int64_t total = n_maps * per;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = idx / per;
    int rem = (int)(idx - m * per);
    int oy = rem / ow, ox = rem - oy * ow;
    const double fy = sy * ((double)oy + 0.5) - 0.5, fx = sx * ((double)ox + 0.5) - 0.5;
    const double fly = floor(fy), flx = floor(fx);
    double wy[4], wx[4];
    cubic_coeffs(fy - fly, wy);
    cubic_coeffs(fx - flx, wx);
    const int iy = (int)fly, ix = (int)flx;
    const TIn* src = in + m * (int64_t)ih * iw;
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
    out[idx] = acc;
  }
}

static int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)kNumSMs * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace rdm

using namespace rdm;

extern "C" int rdm_pair_v1_f32(const float* d3, int64_t n_images, float* raw_out, rdm_stream_t stream) {
  RDM_REQUIRE(d3 && raw_out, "rdm_pair_v1_f32: null pointer");
  RDM_REQUIRE(n_images >= 0, "rdm_pair_v1_f32: negative n_images");
  RDM_REQUIRE(aligned16(raw_out), "rdm_pair_v1_f32: raw_out must be 16-byte aligned");
  if (n_images == 0) return 0;
  pair_v1_kernel<<<grid_for(n_images, 1, 16), 256, 0, (cudaStream_t)stream>>>(d3, raw_out, n_images);
  return launch_status("pair_v1_kernel");
}

extern "C" int rdm_resize_half(const void* in, int32_t in_is_f64, int64_t n_images, int32_t side, double* out,
                               rdm_stream_t stream) {
  RDM_REQUIRE(in && out, "rdm_resize_half: null pointer");
  RDM_REQUIRE(is_pow2(side) && side >= 2 && side <= 4096, "rdm_resize_half: side must be a power of two >= 2 (got %d)", side);
  if (n_images <= 0) return n_images == 0 ? 0 : (set_error("rdm_resize_half: negative n_images"), -1);
  int64_t total = n_images * (int64_t)(side / 2) * (side / 2);
  int grid = grid_for(total, 256);
  if (in_is_f64)
    resize_half_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)in, out, n_images, side);
  else
    resize_half_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)in, out, n_images, side);
  return launch_status("resize_half_kernel");
}

extern "C" int rdm_pair_id_f64(const float* dn, int64_t n_images, int32_t side, double* raw_out, double* parent_out,
                               rdm_stream_t stream) {
  RDM_REQUIRE(dn && raw_out, "rdm_pair_id_f64: null pointer");
  RDM_REQUIRE(is_pow2(side) && side >= 16 && side <= 128, "rdm_pair_id_f64: side must be 16, 32, 64 or 128 (got %d)", side);
  RDM_REQUIRE(aligned16(raw_out), "rdm_pair_id_f64: raw_out must be 16-byte aligned");
  if (n_images <= 0) return n_images == 0 ? 0 : (set_error("rdm_pair_id_f64: negative n_images"), -1);
  int ratio = side / 16;
  int64_t blocks = n_images * ratio * ratio;
  RDM_REQUIRE(blocks < (1ll << 31), "rdm_pair_id_f64: too many pages");
  pair_id_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dn, raw_out, parent_out, side, ratio);
  return launch_status("pair_id_kernel");
}

extern "C" int rdm_pair_pages_f64(const float* pages, const double* parents, int64_t n_pages, double* raw_out, rdm_stream_t stream) {
  RDM_REQUIRE(pages && parents && raw_out, "rdm_pair_pages_f64: null pointer");
  RDM_REQUIRE(aligned16(raw_out), "rdm_pair_pages_f64: raw_out must be 16-byte aligned");
  RDM_REQUIRE(n_pages >= 0 && n_pages < (1ll << 31), "rdm_pair_pages_f64: bad n_pages");
  if (n_pages == 0) return 0;
  pair_pages_kernel<<<(unsigned)n_pages, 256, 0, (cudaStream_t)stream>>>(pages, parents, raw_out);
  return launch_status("pair_pages_kernel");
}

extern "C" int rdm_upsample_nearest_f64(const void* in, int32_t in_is_f64, int64_t n_maps, int32_t side, int32_t times, double* out,
                                        rdm_stream_t stream) {
  RDM_REQUIRE(in && out, "rdm_upsample_nearest_f64: null pointer");
  RDM_REQUIRE(side >= 1 && times >= 0 && times <= 12 && ((int64_t)side << times) <= 8192, "rdm_upsample_nearest_f64: bad side/times");
  RDM_REQUIRE(n_maps >= 0, "rdm_upsample_nearest_f64: bad n_maps");
  if (n_maps == 0) return 0;
  int64_t total = n_maps * ((int64_t)side << times) * ((int64_t)side << times);
  if (in_is_f64)
    upsample_nearest_kernel<double><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const double*)in, out, n_maps, side, times);
  else
    upsample_nearest_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)in, out, n_maps, side, times);
  return launch_status("upsample_nearest_kernel");
}

extern "C" int rdm_resize_bicubic_f64(const void* in, int32_t in_is_f64, int64_t n_maps, int32_t in_h, int32_t in_w, int32_t out_h,
                                      int32_t out_w, double* out, rdm_stream_t stream) {
  RDM_REQUIRE(in && out, "rdm_resize_bicubic_f64: null pointer");
  RDM_REQUIRE(in_h >= 1 && in_w >= 1 && out_h >= 1 && out_w >= 1 && in_h <= 16384 && in_w <= 16384 && out_h <= 16384 && out_w <= 16384,
              "rdm_resize_bicubic_f64: bad sizes");
  RDM_REQUIRE(n_maps >= 0, "rdm_resize_bicubic_f64: bad n_maps");
  if (n_maps == 0) return 0;
  int64_t total = n_maps * (int64_t)out_h * out_w;
  if (in_is_f64)
    resize_bicubic_kernel<double><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const double*)in, out, n_maps, in_h, in_w, out_h, out_w);
  else
    resize_bicubic_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)in, out, n_maps, in_h, in_w, out_h, out_w);
  return launch_status("resize_bicubic_kernel");
}

template <typename T>
static int lloyd_launch(const T* x, int64_t n, const double* thr, const double* lvl, T* values, uint8_t* bins,
                        rdm_stream_t stream, const char* name) {
  RDM_REQUIRE(n >= 0, "%s: negative n", name);
  if (n == 0) return 0;
  RDM_REQUIRE(x && thr && lvl, "%s: null pointer", name);
  RDM_REQUIRE(aligned16(x) && (!values || aligned16(values)) && (!bins || (reinterpret_cast<uintptr_t>(bins) & 3u) == 0),
              "%s: x/values must be 16-byte aligned and bins 4-byte aligned", name);
  if (n == 0) return 0;
  lloyd_kernel<T><<<grid_for(n, 256 * (16 / (int)sizeof(T)) * 4), 256, 0, (cudaStream_t)stream>>>(x, n, thr, lvl, values, bins);
  return launch_status(name);
}

extern "C" int rdm_lloyd_quantize_f32(const float* x, int64_t n, const double* thr40, const double* lvl41,
                                      float* values_out, uint8_t* bins_out, rdm_stream_t stream) {
  return lloyd_launch<float>(x, n, thr40, lvl41, values_out, bins_out, stream, "rdm_lloyd_quantize_f32");
}
extern "C" int rdm_lloyd_quantize_f64(const double* x, int64_t n, const double* thr40, const double* lvl41,
                                      double* values_out, uint8_t* bins_out, rdm_stream_t stream) {
  return lloyd_launch<double>(x, n, thr40, lvl41, values_out, bins_out, stream, "rdm_lloyd_quantize_f64");
}
