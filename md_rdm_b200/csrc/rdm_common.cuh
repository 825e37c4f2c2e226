// Shared device/host helpers for librdm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cmath>

#include "../../include/rdm_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librdm_b200 is written for sm_100a (B200) only"
#endif

namespace rdm {

// ---- error plumbing (no exceptions cross the ABI) --------------------------------------
void set_error(const char* fmt, ...);
int launch_status(const char* what);   // cudaGetLastError -> return code + message
int als_sparse_launch(const rdm_als_scale_t* scales, int32_t n_scales, int64_t n_images, int32_t group, bool sparsify, bool iterate, cudaStream_t stream);   // rdm_als_sparse.cu

#define RDM_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::rdm::set_error(__VA_ARGS__);      \
      return -1;                          \
    }                                     \
  } while (0)

// Opt-in dynamic shared memory, set once per (kernel, device) and raised only when a larger
// size is needed: cudaFuncSetAttribute costs a few microseconds of host time per call.
// `cache` is a per-kernel array of 64 entries (benign race: at worst the attribute is set twice).
template <typename Kernel>
inline cudaError_t ensure_dyn_smem(Kernel kernel, size_t bytes, size_t* cache) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && cache[dev] >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) cache[dev] = bytes;
  return e;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int ilog2(int v) { int n = 0; while ((1 << n) < v) ++n; return n; }

constexpr int kNumSMs = 148;          // B200
constexpr int kThr = 40;              // Lloyd thresholds (RN:290)
constexpr int kLvl = 41;              // Lloyd levels
constexpr int kThrPad = 64;           // padded table for the branch-free search (NaN fill)

// ---- ALS workspace layout (rdm_als_scale_t.ws), per unit = one (image, page) ------------------
//   256-row units: the 16 KB compact page form (rdm_als_sparse.cu: 256 rows x [f, D[12], 3 pad]), then 4 band
//                  flags (1.0f = the band has the pair-build structure) and 4 pad floats;
//   64-row units:  the SSE record of iterations 0..limit (exchanged between the CTAs of a group's cluster).
// No iterate history: the arg-min is taken inside the iterate kernels (rdm_als_sparse.cu, rdm_als.cu).
constexpr int kCompactRowFloats = 16;   // f, D[12], 3 pad
constexpr int kCompactFloats = 256 * kCompactRowFloats;
__host__ __device__ __forceinline__ int64_t als_ws_stride(int rows, int limit) {
  return rows == 256 ? kCompactFloats + 8 : ((limit + 1 + 3) & ~3);
}

// ---- the normaliser of an ALS map ----------------------------------------------------------------------
// quick_gm(p, H) (CP:76, CP:146, CP:244-255) raises every entry to 1 / rc^2 with rc = H = rows, i.e. to 1 / rows^2 -
// not the geometric mean (that would be 1 / rows).  RDM_ALS_TRUE_GM asks for the geometric mean.
// Reference exponent 2^-12 / 2^-16: p^(e) = exp(x) with |x| = |ln p| e < 3e-3, so 1 + x + x^2/2 + x^3/6 is exact to f32
// rounding (x^4/24 < 4e-12) and p = 1 gives exactly 1; anything else goes through pow() in f64.
__device__ __forceinline__ float gm_factor(float p, int rows, bool true_gm) {
  if (!true_gm) {
    const float x = logf(p) * (1.0f / ((float)rows * (float)rows));
    const float pw = 1.0f + fmaf(fmaf(x, 1.0f / 6.0f, 0.5f) * x, x, x);
    if (p > 0.0f && fabsf(x) < 3e-3f) return pw;
    return (float)pow((double)p, 1.0 / ((double)rows * (double)rows));   // zeros, negatives, NaN, huge ratios
  }
  return (float)pow((double)p, 1.0 / (double)rows);
}

// ---- warp helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_prod(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_stream_f32x4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_stream_f32x2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f64x2(double* p, double a, double b) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void stg_stream_f32x4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// ---- Lloyd bin search (RN:286-311) --------------------------------------------------------
// Table of kThrPad entries: thresholds 0..39 then NaN.  For a non-decreasing table the count
// of thresholds <= x equals the branch-free upper-bound position below; NaN x (and NaN
// padding) compare false, so NaN / negatives / 0 land in bin 0 and +inf in bin 40 exactly
// as the reference's 40 compares do.  `sorted == 0` falls back to the literal 40 compares.
template <typename T>
__device__ __forceinline__ int lloyd_bin(T x, const T* __restrict__ tab, int sorted) {
  if (sorted) {
    int pos = 0;
#pragma unroll
    for (int step = 32; step >= 1; step >>= 1) pos += (x >= tab[pos + step - 1]) ? step : 0;
    return pos;
  }
  int c = 0;
#pragma unroll 8
  for (int i = 0; i < kThr; ++i) c += (x >= tab[i]) ? 1 : 0;
  return c;
}

// Fill shared copies of one codebook.  thr_s: kThrPad entries of T; lvl_s: kLvl entries of T.
// Returns (through *sorted_s, an int in shared memory) whether the thresholds are non-decreasing.
template <typename T>
__device__ __forceinline__ void load_codebook(const double* __restrict__ thr, const double* __restrict__ lvl,
                                              T* thr_s, T* lvl_s, int* sorted_s, int tid, int nthreads) {
  if (tid == 0) *sorted_s = 1;
  __syncthreads();
  for (int i = tid; i < kThrPad; i += nthreads)
    thr_s[i] = (i < kThr) ? static_cast<T>(thr[i]) : static_cast<T>(NAN);
  for (int i = tid; i < kLvl; i += nthreads) lvl_s[i] = static_cast<T>(lvl[i]);
  __syncthreads();
  for (int i = tid; i < kThr - 1; i += nthreads)
    if (!(thr_s[i] <= thr_s[i + 1])) *sorted_s = 0;
  __syncthreads();
}

// ---- Lloyd bin by table lookup ------------------------------------------------------------
// The 6-step search above is a chain of 6 dependent shared-memory loads per element, which made
// quantisation latency-bound (measured: ~35k cycles to quantise one 256x64 page).  The upper bits
// of a positive IEEE number are monotonic in its value, so cell(x) = (top bits of x) >> shift
// orders values; a byte table over the cells between the first and last threshold stores how many
// thresholds lie in LOWER cells.  With at most one threshold per cell (checked when the table is
// built; 2^-10 relative cell width against a minimum threshold ratio of 1.0027 in the shipped
// codebooks) the bin is lut[cell] + (x >= thr[lut[cell]]): two dependent loads, still exactly the
// reference's count of thresholds <= x, in x's dtype.
constexpr int kLutMax = 8192;

template <typename T>
__device__ __forceinline__ int lloyd_cell(T x);
template <>
__device__ __forceinline__ int lloyd_cell<double>(double x) { return __double2hiint(x) >> 10; }
template <>
__device__ __forceinline__ int lloyd_cell<float>(float x) { return __float_as_int(x) >> 13; }

struct LloydLut {
  int base;       // cell of the first threshold
  int ncell;      // cells in the table (<= kLutMax); 0 = table unusable, use lloyd_bin()
  uint8_t lut[kLutMax];
};

// Build the table for thresholds `tab` (kThrPad entries of T, NaN padded, see load_codebook).
// All threads of the CTA call this; `sorted` as returned by load_codebook.
template <typename T>
__device__ __forceinline__ void build_lloyd_lut(LloydLut& L, const T* __restrict__ tab, int sorted, int* cell_s /* >= 40 ints, shared */,
                                                int tid, int nthreads) {
  if (tid < kThr) cell_s[tid] = lloyd_cell<T>(tab[tid]);
  if (tid == 0) L.ncell = 0;
  __syncthreads();
  const int base = cell_s[0];
  const int span = cell_s[kThr - 1] - base + 2;
  bool ok = sorted && tab[0] > (T)0 && span <= kLutMax;
  for (int i = 0; i < kThr - 1; ++i) ok = ok && (cell_s[i] < cell_s[i + 1]);   // one threshold per cell
  if (!ok) return;      // uniform: every thread sees the same table
  for (int c = tid; c < span; c += nthreads) {
    int pos = 0;        // number of thresholds whose cell is < base + c
#pragma unroll
    for (int step = 32; step >= 1; step >>= 1) {
      const int j = pos + step - 1;
      pos += (j < kThr && cell_s[j] < base + c) ? step : 0;
    }
    L.lut[c] = (uint8_t)pos;
  }
  if (tid == 0) {
    L.base = base;
    L.ncell = span;
  }
  __syncthreads();
}

// Branch-free (so that independent look-ups interleave): base/ncell are passed in registers.
template <typename T>
__device__ __forceinline__ int lloyd_bin_lut(T x, const T* __restrict__ tab, const uint8_t* __restrict__ lut, int base, int ncell, T thr0) {
  const int c = max(min(lloyd_cell<T>(x) - base, ncell - 1), 0);
  const int b = lut[c];
  const int r = b + ((x >= tab[b]) ? 1 : 0);   // tab[40..] is NaN: never true
  return (x >= thr0) ? r : 0;   // below the first threshold, zero, negative or NaN: all 40 compares false
}

// Bicubic (A = -0.75) weights at t = 0.5: the only ones an exact halving needs (CP:308-311).
#define RDM_W0 (-0.09375)
#define RDM_W1 (0.59375)

// One output pixel of the stride-2 4-tap separable filter; `at(r, c)` returns the f64 source
// value, indices are clamped here.  Horizontal taps are summed first, then vertical ones.
template <typename F>
__device__ __forceinline__ double bicubic_half_at(F at, int y, int x, int side) {
  const double w[4] = {RDM_W0, RDM_W1, RDM_W1, RDM_W0};
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int r = min(max(2 * y - 1 + a, 0), side - 1);
    double inner = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int c = min(max(2 * x - 1 + b, 0), side - 1);
      double v = at(r, c);
      inner = (b == 0) ? __dmul_rn(v, w[0]) : fma(v, w[b], inner);
    }
    acc = (a == 0) ? __dmul_rn(inner, w[0]) : fma(inner, w[a], acc);
  }
  return acc;
}

// torch bicubic weights (A = -0.75) at fractional position t (CP:308-311 for arbitrary sizes)
__device__ __forceinline__ void cubic_coeffs(double t, double (&w)[4]) {
  const double A = -0.75;
  double x;
  x = t + 1.0; w[0] = ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A;
  x = t;       w[1] = ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0;
  x = 1.0 - t; w[2] = ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0;
  x = 2.0 - t; w[3] = ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A;
}

// RN:266-273 + CP:269-295: is column `col` (parent pixel col/8, col%8) inside the 3x3 window of
// page row `row` (pixel row/16, row%16)?  The window is anchored top-left at (min(r/2,5), min(c/2,5)).
__device__ __forceinline__ bool in_window(int row, int col) {
  int r0 = min((row >> 4) >> 1, 5), c0 = min((row & 15) >> 1, 5);
  int dr = (col >> 3) - r0, dc = (col & 7) - c0;
  return (unsigned)dr < 3u && (unsigned)dc < 3u;
}

}  // namespace rdm
