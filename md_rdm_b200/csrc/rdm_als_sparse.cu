// Rank-1 ALS on page pair matrices in their COMPACT form (CP:95-155 on the matrices RN:259-284 builds).
//
// A page pair matrix is 256 x 64, but RN:266-280 + CP:269-295 fill it with very little information:
// row rho = 16 r + c (pixel (r,c) of the 16x16 page) holds the pixel's own value d in 55 columns and
// d / parent in the 9 columns of a 3x3 window of the 8x8 parent page anchored at
// (r0, c0) = (min(r/2,5), min(c/2,5)).  After Lloyd quantisation (RN:286-311) a row is therefore
//     R[rho][j] = f_rho + D_rho[j],   D_rho[j] = 0 outside the window.
// Both ALS GEMVs and the residual then need 12 multiply-adds per row instead of 64 (the window is
// stored as a 3 x 4 span whose first column is even, so that operands are read as 64-bit words and
// the register indexing is static):
//   p-update (CP:186-192)        s_rho = f_rho * sum_j q_j + sum_span D_rho[j] q_j
//   q-update (CP:133, the reference's R.view(B,W,H) reshape: "row i" is rows 4i..4i+3 of R laid end
//             to end)            q_i = sum_{r'<4} [ f_rho P_r' + sum_span D_rho[j] p[64 r' + j] ],
//                                rho = 4 i + r',  P_r' = sum_{c<64} p[64 r' + c]
//   residual (CP:172-173)        sum_j (R[rho][j] - p q_j)^2
//                                  = 64 g^2 - 2 p g S1 + p^2 V + A_rho - 2 p sum_span D_rho[j] q_j
//                                with g = f - p m, S1 = sum_j (q_j - m), V = sum_j (q_j - m)^2 for ANY
//                                centre m (the mean of the previous q is used: no cancellation even
//                                when q is nearly constant), A_rho = sum_span D (2 f + D).
// These are exact identities: the results differ from the dense evaluation by f32 summation order
// only (measured against an f64 evaluation: iterates and record at least as accurate as the
// reference's own f32 bmm).
//
// Three kernels:
//   als_sparsify_raw_kernel  streams raw (or already quantised) f64 matrices from HBM once (HBM-bound),
//                            CHECKS the structure bit-wise (every column outside the window must hold
//                            the same bits), quantises, writes the 16 KB compact form + 4 band flags to
//                            the workspace, and optionally the full bins / quantised matrix.
//   als_sparsify_map_kernel  the same from the decoder map (pair build fused, nothing else is read).
//   als_sparse_kernel        ONE WARP per unit: a lane owns the 2 x 4 pixel block (rows 2rh..2rh+1,
//                            columns 4kq..4kq+3) = 8 matrix rows = 104 registers of matrix, and with it
//                            the two entries q[8rh+kq], q[8rh+4+kq] of q: no block barrier in the loop,
//                            only warp shuffles and __syncwarp.
// A unit whose matrix fails the check (an arbitrary matrix handed to cp.alternating_least_squares) is
// left to the dense kernel (rdm_als.cu), which in turn skips the units flagged here.
#include "rdm_common.cuh"
#include <cooperative_groups.h>
#include <type_traits>
namespace cg = cooperative_groups;

namespace rdm {

namespace {

constexpr int kMaxSparseScales = 8;
constexpr float kLambda = 0.05f;   // CP:175 regularization_term
constexpr unsigned kFull = 0xffffffffu;

struct SparseScaleDev {
  const void* src;
  const double* thr;
  const double* lvl;
  uint8_t* bins;
  float* values;
  float* ws;
  int32_t kind, pages, side, limit;
  int32_t unit_begin;   // first unit of this scale in the kernel's unit numbering
  int32_t live;         // pages of an image the launch works on (pg < live): `pages`, or side/16 under RDM_ALS_SKIP_UNUSED_PAGES
};

struct SparseParams {
  SparseScaleDev s[kMaxSparseScales];
  int64_t n_images;
  int32_t n_scales;
};

__device__ __forceinline__ const SparseScaleDev& find_scale(const SparseParams& P, int gunit) {
  int si = 0;
#pragma unroll 1
  for (int k = 1; k < P.n_scales; ++k)
    if (gunit >= P.s[k].unit_begin) si = k;
  return P.s[si];
}

// codebook prologue: f64 thresholds (pages compare in f64, RN:376-378 via SURVEY 8a-a5), f32 levels
__device__ __forceinline__ void load_book(const SparseScaleDev& sc, double* thr_d, float* lvl_f, int* sorted, int tid, int nt) {
  if (tid == 0) *sorted = 1;
  __syncthreads();
  for (int i = tid; i < kThrPad; i += nt) thr_d[i] = (i < kThr) ? sc.thr[i] : (double)NAN;
  for (int i = tid; i < kLvl; i += nt) lvl_f[i] = (float)sc.lvl[i];
  __syncthreads();
  for (int i = tid; i < kThr - 1; i += nt)
    if (!(thr_d[i] <= thr_d[i + 1])) *sorted = 0;
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Geometry of a page row rho = 16 r + c (RN:266-273 + CP:269-295): the 3x3 window of the 8x8 parent page is anchored
// at (r0, c0) = (min(r/2,5), min(c/2,5)); the compact row keeps it as a 3 x 4 span starting at the even column
// span0 = min(2 (c/4), 4).
struct RowGeom {
  int r0, c0, span0;
  __host__ __device__ __forceinline__ explicit RowGeom(int rho) {
    const int r = rho >> 4, c = rho & 15;
    r0 = (r >> 1) < 5 ? (r >> 1) : 5;
    c0 = (c >> 1) < 5 ? (c >> 1) : 5;
    span0 = 2 * (c >> 2) < 4 ? 2 * (c >> 2) : 4;
  }
  __host__ __device__ __forceinline__ int window_col(int a, int b) const { return 8 * (r0 + a) + c0 + b; }
  // a matrix column outside every window of this pixel row: parent row 0 unless the window starts there, else row 7
  __host__ __device__ __forceinline__ int fill_col() const { return r0 >= 1 ? 0 : 56; }
};

// Per-row constants of the residual identity (header comment), computed where the row is built so that the iterate
// kernel's prologue does not spend ~400 f64 instructions per lane on them: A_rho = sum_span D (2 f + D) as an
// unevaluated f32 pair (exact to ~2^-48) and the iteration-0 residual sum_j fl(1 - R[rho][j])^2 (p = q = 1, CP:123).
// wl: the nine window levels (row-major), f: the fill level.
__device__ __forceinline__ void row_constants(float f, const float (&wl)[9], float& a_hi, float& a_lo, float& e0) {
  const float u = __fsub_rn(1.0f, f);
  double asum = 0.0, acc = 55.0 * ((double)u * (double)u);   // 64 - 9 columns hold f
#pragma unroll
  for (int w = 0; w < 9; ++w) {
    const float D = wl[w] - f;                                // the compact entry, as stored
    asum = fma((double)D, 2.0 * (double)f + (double)D, asum);
    const float r = __fsub_rn(1.0f, __fadd_rn(f, D));
    acc = fma((double)r, (double)r, acc);
  }
  a_hi = (float)asum;
  a_lo = (float)(asum - (double)a_hi);
  e0 = (float)acc;
}

// ---- mbarrier / bulk-copy (TMA) primitives --------------------------------------------------------------
__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- the HBM-streaming kernel of the path ----------------------------------------------------------------
// Work item = one band of 64 matrix rows (32 KB of f64, contiguous in HBM) = one CTA of 64 threads; six CTAs fit an
// SM, so ~190 KB of loads are in flight per SM.  Thread 0 issues ONE 32 KB bulk async copy (cp.async.bulk = TMA,
// 1-D) that completes on an mbarrier; the codebook is fetched while it flies.  The 64 threads then own ONE ROW
// EACH, with no cross-lane traffic:
//   * read the fill value and the nine window values of the row (their columns follow from the row index);
//   * overwrite the nine window entries of the staged row with the fill value, then compare all 64 entries with
//     the fill value bit-wise (two LOP3 per entry): any difference = the row lacks the pair-build structure;
//   * quantise the 10 informative values (f64 compares, RN:376-378), emit the 16-float compact row, and - if
//     asked - the row of bins / quantised values, which are fill or window values by construction (assembled in
//     the row's own staging bytes and copied out by the warp with 128-bit stores).
// Shared-memory banks: rows are 512 bytes apart, i.e. on the same banks (a padded layout needs one bulk copy per
// row, and 64 small copies per band are request-rate bound: measured 2.6 TB/s).  So (i) thread t walks the 16-byte
// chunks of its row in the order c ^ t, and keeps logical output chunk k at position k ^ t: at every step the
// lanes of a warp touch 32 different chunks; (ii) the rows of a band are dealt to the two warps so that each
// warp holds every residue rho mod 32 once AND at most 3 rows per half-warp share a window position (a warp
// of 32 consecutive rows would have 12): lane l of warp w takes pixel column c = l & 15 of pixel row
// 4 band + (l >> 4) + 2 ((c ^ (l >> 4) ^ w) & 1).
// About 12 warp-instructions per matrix row (the round-1 kernel, in which a warp shared a row: 92).
constexpr int kBandRows = 64;
constexpr int kRowBytes = 512;                 // 64 f64
constexpr int kStageBytes = kBandRows * kRowBytes;

__global__ void __launch_bounds__(kBandRows) als_sparsify_raw_kernel(const __grid_constant__ SparseParams P) {
  extern __shared__ __align__(128) unsigned char stage[];
  __shared__ double thr_d[kThrPad];
  __shared__ float lvl_f[kLvl + 3];
  __shared__ int sorted;
  __shared__ __align__(8) unsigned long long full_bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gunit = blockIdx.x >> 2, band = blockIdx.x & 3;
  const SparseScaleDev& sc = find_scale(P, gunit);
  const int64_t live_unit = gunit - sc.unit_begin;                 // (image, live page)
  const int64_t unit = (live_unit / sc.live) * sc.pages + live_unit % sc.live;   // (image, page) as laid out in memory
  const bool quant = sc.kind == RDM_SRC_RAW_F64;
  if (tid == 0) {
    mbar_init(sm_addr(&full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned char* src = reinterpret_cast<const unsigned char*>(sc.src) + (unit * 256 + band * kBandRows) * (int64_t)kRowBytes;
    mbar_expect_tx(sm_addr(&full_bar), kStageBytes);
    bulk_g2s(sm_addr(stage), src, kStageBytes, sm_addr(&full_bar));
  }
  if (quant) load_book(sc, thr_d, lvl_f, &sorted, tid, kBandRows);   // ends with __syncthreads
  else __syncthreads();                                              // the mbarrier is initialised for everyone
  const int srt = quant ? sorted : 1;
  // ---- this thread's row
  const int c = lane & 15, hi = lane >> 4;
  const int rho = band * kBandRows + 16 * (hi + 2 * ((c ^ hi ^ warp) & 1)) + c;   // matrix row = pixel of the page; rho % 32 == lane
  const RowGeom gm(rho);
  const uint32_t rot = (uint32_t)lane << 4;   // chunk rotation of this thread (bytes)
  unsigned char* rb = stage + (rho & (kBandRows - 1)) * kRowBytes;
  double* row = reinterpret_cast<double*>(rb);
  mbar_wait(sm_addr(&full_bar), 0);
  // ---- informative values: fill + 3x3 window; then the structure check on the neutralised row.  The fill value
  // is read from one of the eight columns of a parent row outside the window (spread over the lanes)
  const double F = row[gm.fill_col() + (lane & 7)];
  double wv[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double* e = row + gm.window_col(a, b);
      wv[3 * a + b] = *e;
      *e = F;
    }
  const unsigned Flo = (unsigned)__double2loint(F), Fhi = (unsigned)__double2hiint(F);
  unsigned diff = 0;
#pragma unroll 8
  for (int k = 0; k < 32; ++k) {
    const uint4 x = *reinterpret_cast<const uint4*>(rb + (((uint32_t)k << 4) ^ rot));
    diff |= (x.x ^ Flo) | (x.y ^ Fhi);
    diff |= (x.z ^ Flo) | (x.w ^ Fhi);
  }
  const bool ok = diff == 0;
  // ---- Lloyd (f64 compares) of the 10 values, or `.float()` of an already quantised matrix (CP:106)
  int bf = 0, wb[9];
  float f, wl[9];
  if (quant) {
    bf = lloyd_bin<double>(F, thr_d, srt);
    f = lvl_f[bf];
#pragma unroll
    for (int w = 0; w < 9; ++w) {
      wb[w] = lloyd_bin<double>(wv[w], thr_d, srt);
      wl[w] = lvl_f[wb[w]];
    }
  } else {
    f = (float)F;
#pragma unroll
    for (int w = 0; w < 9; ++w) {
      wb[w] = 0;
      wl[w] = (float)wv[w];
    }
  }
  // ---- compact row: f, then the 3 x 4 span of (window value - f), zero outside the window
  float* compact = sc.ws + unit * als_ws_stride(256, sc.limit);
  {
    float o[16];
    o[0] = f;
    row_constants(f, wl, o[13], o[14], o[15]);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int b = gm.span0 + g - gm.c0;   // window column of span column g
        float v = 0.f;
#pragma unroll
        for (int bb = 0; bb < 3; ++bb)
          if (b == bb) v = wl[3 * a + bb] - f;
        o[1 + 4 * a + g] = v;
      }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(compact + rho * kCompactRowFloats + 4 * k) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  }
  // ---- optional full-size outputs, assembled in the row's own staging bytes: logical chunks 0..3 = the 64 bins,
  // 4..19 = the 64 quantised values; logical chunk k lives at physical chunk k ^ lane
  if (sc.bins || sc.values) {
    if (sc.bins) {
      const unsigned rep = 0x01010101u * (unsigned)bf;
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(rb + (((uint32_t)k << 4) ^ rot)) = make_uint4(rep, rep, rep, rep);
    }
    if (sc.values) {
#pragma unroll
      for (int k = 4; k < 20; ++k) *reinterpret_cast<float4*>(rb + (((uint32_t)k << 4) ^ rot)) = make_float4(f, f, f, f);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const uint32_t col = (uint32_t)gm.window_col(a, b);
        if (sc.bins) rb[((col & ~15u) ^ rot) | (col & 15u)] = (unsigned char)wb[3 * a + b];
        if (sc.values) *reinterpret_cast<float*>(rb + ((((64u + 4u * col) & ~15u) ^ rot) | ((4u * col) & 15u))) = wl[3 * a + b];
      }
    __syncwarp();
    // copy-out by the warp: the row of lane l' sits at its own rotation l'
    const int64_t mrow0 = unit * 256 + band * kBandRows;   // first matrix row of the band
    auto row_of = [&](uint32_t l2) {                        // band-local row held by lane l2 of this warp
      const uint32_t c2 = l2 & 15, h2 = l2 >> 4;
      return 16 * (h2 + 2 * ((c2 ^ h2 ^ (uint32_t)warp) & 1)) + c2;
    };
    if (sc.bins) {   // 64 B per row: a 128-bit store instruction covers 8 rows
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const uint32_t l2 = 8 * it + (lane >> 2), ch = lane & 3, r2 = row_of(l2);
        const uint4 x = *reinterpret_cast<const uint4*>(stage + r2 * kRowBytes + ((ch ^ l2) << 4));
        *reinterpret_cast<uint4*>(sc.bins + (mrow0 + r2) * 64 + 16 * ch) = x;
      }
    }
    if (sc.values) {   // 256 B per row: an instruction covers 2 rows
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {
        const uint32_t l2 = 2 * it + (lane >> 4), ch = lane & 15, r2 = row_of(l2);
        const float4 x = *reinterpret_cast<const float4*>(stage + r2 * kRowBytes + (((4 + ch) ^ l2) << 4));
        *reinterpret_cast<float4*>(sc.values + (mrow0 + r2) * 64 + 4 * ch) = x;
      }
    }
  }
  const int all_ok = __syncthreads_and(ok ? 1 : 0);
  if (tid == 0) compact[kCompactFloats + band] = all_ok ? 1.0f : 0.0f;
}

// The compact form of ONE page straight from the decoder map (RN:259-284 + CP:269-295 + CP:308-311 fused: the pair
// matrix never exists).  Called by the 256 threads of a CTA, thread = matrix row = pixel of the page; `map` may be
// global or shared memory.  inv_d: 64 doubles of shared scratch.
__device__ __forceinline__ void compact_page_from_map(const SparseScaleDev& sc, int64_t unit, const float* __restrict__ map, int side, int pg,
                                                      const double* thr_d, const float* lvl_f, int srt, double* inv_d, int tid) {
  const int ratio = side >> 4;
  const int pi = pg / ratio, pj = pg - pi * ratio;
  if (tid < 64) {
    const int y = 8 * pi + (tid >> 3), x = 8 * pj + (tid & 7);
    const double v = bicubic_half_at([&](int r, int c) { return (double)map[r * side + c]; }, y, x, side);
    inv_d[tid] = 1.0 / v;   // torch.pow(area,-1): IEEE reciprocal
  }
  __syncthreads();
  const int row = tid;
  const double d = (double)map[(16 * pi + (row >> 4)) * side + 16 * pj + (row & 15)];
  const RowGeom gm(row);
  const int r0 = gm.r0, c0 = gm.c0, span0 = gm.span0;
  const int bf = lloyd_bin<double>(d, thr_d, srt);   // 55 of 64 columns hold d itself
  const float f = lvl_f[bf];
  float* compact = sc.ws + unit * als_ws_stride(256, sc.limit);
  float o[16];
  o[0] = f;
  int wb[9];   // window bins, row-major over the 3x3 window
  float wlv[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      wb[3 * a + b] = lloyd_bin<double>(__dmul_rn(d, inv_d[8 * (r0 + a) + c0 + b]), thr_d, srt);
      wlv[3 * a + b] = lvl_f[wb[3 * a + b]];
    }
  row_constants(f, wlv, o[13], o[14], o[15]);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int b = span0 + g - c0;   // window column of span column g
      float v = 0.f;
#pragma unroll
      for (int bb = 0; bb < 3; ++bb)
        if (b == bb) v = wlv[3 * a + bb] - f;
      o[1 + 4 * a + g] = v;
    }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    *reinterpret_cast<float4*>(compact + row * kCompactRowFloats + 4 * k) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  if (tid < 4) compact[kCompactFloats + tid] = 1.0f;
  if (sc.bins || sc.values) {
    const int64_t off = unit * (int64_t)(256 * 64) + row * 64;
    for (int c4 = 0; c4 < 16; ++c4) {
      uint32_t pk = 0;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = 4 * c4 + e;
        const int a = (c >> 3) - r0, b = (c & 7) - c0;
        int bin = bf;
        if ((unsigned)a < 3u && (unsigned)b < 3u) {
#pragma unroll
          for (int w = 0; w < 9; ++w)
            if (w == 3 * a + b) bin = wb[w];
        }
        pk |= (uint32_t)bin << (8 * e);
        v[e] = lvl_f[bin];
      }
      if (sc.bins) *reinterpret_cast<uint32_t*>(sc.bins + off + 4 * c4) = pk;
      if (sc.values) *reinterpret_cast<float4*>(sc.values + off + 4 * c4) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// grid = units; one page per CTA.
__global__ void __launch_bounds__(256) als_sparsify_map_kernel(const __grid_constant__ SparseParams P) {
  __shared__ double thr_d[kThrPad];
  __shared__ double inv_d[64];
  __shared__ float lvl_f[kLvl + 3];
  __shared__ int sorted;
  const int tid = threadIdx.x;
  const int gunit = blockIdx.x;
  const SparseScaleDev& sc = find_scale(P, gunit);
  const int64_t live_unit = gunit - sc.unit_begin;
  const int side = sc.side;
  const int64_t img = live_unit / sc.live;
  const int pg = (int)(live_unit - img * sc.live);
  const int64_t unit = img * sc.pages + pg;
  const float* map = reinterpret_cast<const float*>(sc.src) + img * (int64_t)side * side;
  load_book(sc, thr_d, lvl_f, &sorted, tid, 256);   // ends with __syncthreads
  compact_page_from_map(sc, unit, map, side, pg, thr_d, lvl_f, sorted, inv_d, tid);
}

// ---- the decoder's 1x1 conv head (RN:146, RN:157) fused with the pair build -----------------
// conv1 maps the (C, s, s) feature block of a relative decoder to its one-channel map: out[p] = bias + sum_c w_c x[c][p].
// That is a C-long reduction per pixel over C s^2 floats per image (0.5 .. 7 MB) - it moves more bytes than the whole
// fusion path - so it is split over a thread-block CLUSTER of 8 CTAs per image: CTA r reduces channels
// [r C/8, (r+1) C/8) for all pixels with coalesced 128-bit streaming loads (eight in flight per thread), the eight
// partial maps are reduce-scattered through distributed shared memory (CTA r sums pixel slice r in rank order
// and adds the bias), and the finished slices are broadcast back.  Every CTA then holds the map in shared memory and
// builds the compact pair-matrix form of its share of the 16x16 pages (compact_page_from_map): for scales >= 16 the
// decoder map never touches HBM (it is written only if the caller asks for it; the 8x8 scale, which the dense ALS
// kernel reads from HBM, always asks).
// Summation order: channels in increasing order within a lane, then lanes, then CTAs - fixed, but not cuDNN's or
// MKL-DNN's: the map agrees with torch.nn.Conv2d to f32 rounding (~1e-6), so a ratio within that distance of a Lloyd
// threshold can land in the neighbouring bin.  Bins are exact with respect to the map THIS kernel produces.
constexpr int kConvCluster = 8;
constexpr int kConvMaxPixels = 64 * 64;

struct ConvSmem {
  float part[kConvMaxPixels];        // this CTA's partial map (its channel range)
  float map[kConvMaxPixels];         // the finished map
  float lanes[4096];                 // cross-lane reduction scratch when pixels/4 < 256
  double thr_d[kThrPad];
  double inv_d[64];
  float lvl_f[kLvl + 3];
  int sorted;
};

__global__ void __launch_bounds__(256) conv_head_kernel(const float* __restrict__ feat, const float* __restrict__ weight,
                                                        const float* __restrict__ bias, int channels, int side, float* __restrict__ map_out,
                                                        const __grid_constant__ SparseScaleDev sc, int build_pages) {
  extern __shared__ __align__(16) unsigned char conv_raw[];
  ConvSmem& sm = *reinterpret_cast<ConvSmem*>(conv_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int64_t img = blockIdx.x / kConvCluster;
  const int tid = threadIdx.x;
  const int M = side * side, quads = M >> 2;
  const int c_per = channels / kConvCluster, c_lo = rank * c_per;
  const float* x = feat + (img * channels + c_lo) * (int64_t)M;
  // ---- this CTA's channel range: thread (quad, lane): channels lane, lane + nl, ...
  const int nl = quads >= 256 ? 1 : 256 / quads;         // channel lanes per pixel quad
  for (int q0 = 0; q0 < quads; q0 += 256) {
    const int quad = quads >= 256 ? q0 + tid : tid % quads;
    const int cl = quads >= 256 ? 0 : tid / quads;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = cl;
    for (; c + 7 * nl < c_per; c += 8 * nl) {
      float4 v[8];
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = ldg_stream_f32x4(x + (int64_t)(c + u * nl) * M + 4 * quad);
        w[u] = weight[c_lo + c + u * nl];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc.x = fmaf(w[u], v[u].x, acc.x);
        acc.y = fmaf(w[u], v[u].y, acc.y);
        acc.z = fmaf(w[u], v[u].z, acc.z);
        acc.w = fmaf(w[u], v[u].w, acc.w);
      }
    }
    for (; c < c_per; c += nl) {
      const float4 v = ldg_stream_f32x4(x + (int64_t)c * M + 4 * quad);
      const float w = weight[c_lo + c];
      acc.x = fmaf(w, v.x, acc.x);
      acc.y = fmaf(w, v.y, acc.y);
      acc.z = fmaf(w, v.z, acc.z);
      acc.w = fmaf(w, v.w, acc.w);
    }
    if (nl == 1) {
      *reinterpret_cast<float4*>(&sm.part[4 * quad]) = acc;
    } else {
      *reinterpret_cast<float4*>(&sm.lanes[4 * (cl * quads + quad)]) = acc;
    }
  }
  if (nl > 1) {
    __syncthreads();
    for (int p = tid; p < M; p += 256) {
      float t = 0.f;
      for (int l = 0; l < nl; ++l) t += sm.lanes[l * M + p];   // lanes in order
      sm.part[p] = t;
    }
  }
  cluster.sync();   // every partial map is complete
  // ---- reduce-scatter: CTA `rank` finishes pixels [rank M/8, (rank+1) M/8), then broadcasts them
  {
    const int per = M / kConvCluster;
    const float b0 = bias ? bias[0] : 0.f;
    for (int i = tid; i < per; i += 256) {
      const int p = rank * per + i;
      float t = 0.f;
      for (int r = 0; r < kConvCluster; ++r) t += *cluster.map_shared_rank(&sm.part[p], r);   // CTAs in rank order
      t += b0;
      if (map_out) map_out[img * M + p] = t;
      for (int r = 0; r < kConvCluster; ++r) *cluster.map_shared_rank(&sm.map[p], r) = t;
    }
  }
  cluster.sync();   // the map is complete everywhere; no remote access after this
  if (!build_pages) return;
  // ---- compact page form of pages rank, rank + 8, ...
  load_book(sc, sm.thr_d, sm.lvl_f, &sm.sorted, tid, 256);
  for (int pg = rank; pg < sc.live; pg += kConvCluster) {
    compact_page_from_map(sc, img * sc.pages + pg, sm.map, side, pg, sm.thr_d, sm.lvl_f, sm.sorted, sm.inv_d, tid);
    __syncthreads();   // inv_d is reused by the next page
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_newton(float x) {   // see rdm_als.cu
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(fmaf(-x, r, 1.0f), r, r);
}


// sum_span D v as one FMA chain (eight independent rows interleave).  FROM = 1 skips the span's first
// column: rows of pixel columns 4kq+2, 4kq+3 have their window in span columns 1..3 for every kq.
template <int FROM>
__device__ __forceinline__ float span_dot(const float (&D)[12], const float (&v)[12]) {
  float acc = D[FROM] * v[FROM];
#pragma unroll
  for (int e = FROM + 1; e < 12; ++e)
    if ((e & 3) >= FROM) acc = fmaf(D[e], v[e], acc);
  return acc;
}
template <int T>   // row slot T = 4 dr + cc
__device__ __forceinline__ float row_dot(const float (&D)[12], const float (&v)[12]) {
  return span_dot<((T & 3) >= 2) ? 1 : 0>(D, v);
}

// ---- the grouped page kernel --------------------------------------------------------------------
// One CTA per (reference batch = "group", page); warp w iterates the page of image w of the group, so the 16
// images whose rmse the reference averages (CP:172-173) meet in ONE CTA and the arg-min (CP:143) is taken while
// the iterations run: nothing but the selected iterate ever leaves the SM.
//
//  * Step k of a warp (iteration k) leaves its 32 per-lane residuals in E[k & 3][warp][lane] and ARRIVES (bar.arrive,
//    no wait) on named barrier 1 + (k & 3).
//  * Warp (j % W) is the reducer of iteration j: at the top of its step j + 2 - when every warp has normally long
//    arrived - it completes the barrier (bar.sync), sums the group's residuals (f64), takes the rmse, compares it
//    with the running minimum (strict <: the FIRST minimum wins, as list.index(min(list)) does) and publishes
//    flag[j & 3] = 2 j + new_minimum.
//    Verdicts are issued in iteration order: the reducer of j waits for the verdict on j - 1 (another warp's,
//    possibly a step behind) before it touches the running minimum.
//  * Every warp reads the verdict on iteration j at the end of its step j + kLag (= j + 3), and on a new minimum
//    copies p_j from its 4-deep ring of iterates into its `best` row.
// Why the 4-deep rings and 4 barriers can be reused: a warp enters step k only after it has read the verdict on
// iteration k - 4 (end of step k - 1), and that verdict is published after the reducer completed barrier (k-4) & 3 and
// read E[(k-4) & 3]; so the arrival on barrier k & 3, the store to E[k & 3] and the store to the p ring slot k & 3
// of step k never meet their previous use.  Every wait refers to a strictly earlier step, so there is no cycle.
// Groups of more than 16 images take several rounds of 16 warps: pass 1 accumulates the group record over the
// rounds (no iterate is kept), pass 2 re-runs the k* selected iterations of every unit (bit-identical
// arithmetic) and emits them.  On noise-like maps k* <= 1, so pass 2 is a few per cent of pass 1.
#ifndef RDM_PAGES_CLUSTER_MINB
#define RDM_PAGES_CLUSTER_MINB 3   // resident CTAs per SM the cluster form is compiled for: 168 registers, no spills in the loop (4: 128 registers, a lone call 81 instead of 66 us; bench throughput equal)
#endif
constexpr int kLag = 3;
constexpr int kRing = 4;
constexpr int kRedDelay = 2;   // the reducer of iteration j works at the top of its step j + kRedDelay
static_assert(kRing == kLag + 1 && (kRing & (kRing - 1)) == 0 && kRedDelay < kLag, "ring arithmetic of pages_iterate");
constexpr int kGroupWarps = 16;
constexpr int kWarpFloats = kRing * 256 + 64 + 256;   // p ring, q, best
constexpr int kMaxLimit = 127;

struct PagesScaleDev {
  float* ws;
  float* pages_out;
  float* map_out;
  float* record_out;
  int32_t* kstar_out;
  int32_t pages, side, limit;
  int32_t cta_begin;   // first (group, page) item of this scale
  int32_t flags;       // RDM_ALS_TRUE_GM | RDM_ALS_CORRECT_TILING
  int32_t live;        // pages of an image the launch works on (pg < live)
};
struct PagesParams {
  PagesScaleDev s[kMaxSparseScales];
  int64_t n_images;
  int32_t n_scales, group;
};

struct PagesShared {            // fixed part of the dynamic shared memory (the per-warp rows follow)
  float E[kRing][kGroupWarps][32];
  double recg[kMaxLimit + 1];   // group record accumulated over the rounds (multi-round groups)
  int flag[kRing];              // verdicts: 2 j + new_minimum, -1 before the first use
  float best_rmse;
  int kstar;
  int all_compact;
};

__device__ __forceinline__ void named_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// shared-memory accesses of the iteration loop by 32-bit shared address (generic pointers make nvcc re-derive the
// shared window base inside the loop, see rdm_als.cu)
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float2 s_ld2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 s_ld4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void s_st4(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void s_st1(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ int s_ld_flag(uint32_t a) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void load_span_u32(uint32_t a, float (&v)[12]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float2 lo = s_ld2(a + 32 * r), hi = s_ld2(a + 32 * r + 8);
    v[4 * r] = lo.x;
    v[4 * r + 1] = lo.y;
    v[4 * r + 2] = hi.x;
    v[4 * r + 3] = hi.y;
  }
}

// The matrix of one page in registers: a lane owns the 2 x 4 pixel block (rows 2rh..2rh+1, columns 4kq..4kq+3)
// = 8 matrix rows, and with them the two entries q[8rh+kq], q[8rh+4+kq].
struct PageRegs {
  float f[8], D[8][12];
  float A, e0;   // sum over the lane's rows of A_rho and of the iteration-0 residual (row_constants)
};

// The reducer's work for iteration j (one warp, all lanes): complete the barrier of iteration j, group rmse,
// first-minimum test (DECIDE) or accumulation into the multi-round record, verdict.  Out of line: it runs once per W
// steps of a warp and would otherwise sit in the instruction stream of the iteration loop.
template <bool DECIDE>
__device__ __noinline__ void pages_reduce(PagesShared& sh, int j, int W, int lane, double inv_cnt, float* record_out) {
  const int slot = j & (kRing - 1);
  named_sync(1 + slot, (W + 1) * 32);
  const int u = lane & 15, h = lane >> 4;
  double t = 0.0;
  if (u < W) {
    const float4* e4 = reinterpret_cast<const float4*>(&sh.E[slot][u][16 * h]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 x = e4[i];
      t += (double)x.x;
      t += (double)x.y;
      t += (double)x.z;
      t += (double)x.w;
    }
  }
  t += __shfl_xor_sync(kFull, t, 16);
  // unit record rounded to f32, then the group sum in f64 (the reference sums f32 values; exact in f64)
  double g = (h == 0 && u < W) ? (double)(float)t : 0.0;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) g += __shfl_xor_sync(kFull, g, o);
  if (lane == 0) {
    int verdict = j << 1;
    // Decisions are taken in iteration order: the reducer of iteration j - 1 is another warp and may be a step
    // behind this one (its arrival on THIS barrier came before its own reduction), so wait for its verdict -
    // published after its update of the running minimum - before reading the minimum.
    if (j > 0) {
      const uint32_t fa = s_u32(&sh.flag[(j - 1) & (kRing - 1)]);
      while (s_ld_flag(fa) < 2 * (j - 1)) {
      }
    }
    if (DECIDE) {
      const float rm = (float)sqrt(g * inv_cnt);           // CP:172-173 over the whole reference batch
      if (rm < sh.best_rmse) {                             // strict: the first minimum wins (CP:143)
        sh.best_rmse = rm;
        sh.kstar = j;
        verdict |= 1;
      }
      if (record_out) record_out[j] = rm;
    } else {
      sh.recg[j] += g;
    }
    __threadfence_block();
    *reinterpret_cast<volatile int*>(&sh.flag[slot]) = verdict;
  }
  __syncwarp();
}

// ---- how the warps of a group meet ------------------------------------------------------------------------
// LocalSync: all warps of the group in ONE CTA (named barriers, shared-memory flags; pages_reduce above).
struct LocalSync {
  PagesShared& sh;
  int warp, lane, W, nbar, next_red;
  double inv_cnt;
  float* record_out;
  uint32_t a_E, a_flag;
  __device__ __forceinline__ LocalSync(PagesShared& s, int warp_, int lane_, int W_, double ic, float* rec)
      : sh(s), warp(warp_), lane(lane_), W(W_), nbar((W_ + 1) * 32), next_red(warp_), inv_cnt(ic), record_out(rec) {
    uint32_t e = s_u32(&sh.E[0][warp][lane]);
    asm volatile("mov.b32 %0, %0;" : "+r"(e));   // keep it in a register (see pages_iterate)
    a_E = e;
    a_flag = s_u32(&sh.flag[0]);
  }
  __device__ __forceinline__ void publish(int k, float e) {   // this lane's residual of iteration k
    const uint32_t slot = (uint32_t)k & (kRing - 1);
    s_st1(a_E + (slot << 11), e);
    named_arrive(1 + (int)slot, nbar);
  }
  template <bool DECIDE>
  __device__ __forceinline__ void duty(int j) {               // called with j = k - kRedDelay at the top of step k, and in the drain
    if (j == next_red) {
      pages_reduce<DECIDE>(sh, j, W, lane, inv_cnt, record_out);
      next_red += W;
    }
  }
  __device__ __forceinline__ bool verdict(int j) {            // blocks until the verdict on iteration j is out
    const uint32_t fa = a_flag + 4 * (j & (kRing - 1));
    int v;
    do {
      v = s_ld_flag(fa);
    } while (v < 2 * j);             // the slot's previous verdict (or -1) is smaller
    return (v & 1) != 0;
  }
};

// ClusterSync: the 16 warps of a group spread over a thread-block CLUSTER of 4 CTAs x 4 warps (a lone call then runs on
// 20 SMs instead of 5, and the small CTAs pack beside other kernels).  Same protocol, different plumbing:
//  * the residuals of iteration j go to the CTA of its reducer (global warp j % 16) through distributed shared memory
//    with st.async, which also counts the bytes on the reducer's mbarrier: fire and forget;
//  * the reducer arms that mbarrier (arrive.expect_tx 16 x 128 bytes), waits on it, takes the group rmse, waits for the verdict on
//    j - 1 in its own copy of the flags, and BROADCASTS one 64-bit word to the flag[j & 3] of all four CTAs:
//    (j + 1) << 33 | new_minimum << 32 | bits of the running minimum after j - so the running minimum travels with the
//    verdicts and nothing else is shared;
//  * every warp spins on its own CTA's copy of the flag.
// Slot s of a CTA's mbarriers / E buffer serves the iterations j = 16 n + 4 rank + s: phase parity n & 1.
constexpr int kClusterCtas = 4;
constexpr int kClusterWarps = kGroupWarps / kClusterCtas;   // warps (images) per CTA

struct PagesClusterShared {
  float E[kRing][kGroupWarps][32];            // residuals of the iterations this CTA reduces
  unsigned long long flag[kRing];             // this CTA's copy of the verdicts
  unsigned long long mbar[kRing];
  int all_compact;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ unsigned long long s_ld_u64_volatile(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}

// The reducer's work for iteration j (ClusterSync), out of line like pages_reduce.
__device__ __noinline__ void cluster_reduce(PagesClusterShared& sh, int j, int lane, double inv_cnt, float* record_out, uint32_t a_flag, uint32_t a_mbar) {
  const uint32_t slot = (uint32_t)j & (kRing - 1), parity = ((uint32_t)j >> 4) & 1u;
  // the phase completes when this one arrival is in and the 16 warps' 16 x 128 bytes have landed (complete_tx may
  // run ahead of expect_tx: the transaction count is signed)
  if (lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a_mbar + 8 * slot), "r"(kGroupWarps * 128) : "memory");
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}"
      ::"r"(a_mbar + 8 * slot), "r"(parity) : "memory");
  const int u = lane & 15, h = lane >> 4;
  double t = 0.0;
  const float4* e4 = reinterpret_cast<const float4*>(&sh.E[slot][u][16 * h]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 x = e4[i];
    t += (double)x.x;
    t += (double)x.y;
    t += (double)x.z;
    t += (double)x.w;
  }
  t += __shfl_xor_sync(kFull, t, 16);
  double g = h == 0 ? (double)(float)t : 0.0;   // unit record rounded to f32, then the group sum in f64
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) g += __shfl_xor_sync(kFull, g, o);
  if (lane == 0) {
    const float rm = (float)sqrt(g * inv_cnt);   // CP:172-173 over the whole reference batch
    float best = __int_as_float(0x7f800000);
    if (j > 0) {                                 // verdicts are issued in iteration order (see LocalSync / pages_reduce)
      const uint32_t fa = a_flag + 8 * ((uint32_t)(j - 1) & (kRing - 1));
      unsigned long long w;
      do {
        w = s_ld_u64_volatile(fa);
      } while ((w >> 33) < (unsigned long long)j);
      best = __uint_as_float((unsigned)w);
    }
    const bool better = rm < best;               // strict: the first minimum wins (CP:143)
    const unsigned long long word = ((unsigned long long)(j + 1) << 33) | ((unsigned long long)(better ? 1 : 0) << 32) |
                                    (unsigned long long)__float_as_uint(better ? rm : best);
    if (record_out) record_out[j] = rm;
#pragma unroll
    for (uint32_t r = 0; r < (uint32_t)kClusterCtas; ++r)
      asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(mapa_u32(a_flag + 8 * slot, r)), "l"(word) : "memory");
  }
  __syncwarp();
}

struct ClusterSync {
  PagesClusterShared& sh;
  int gw, lane, next_red, kstar;
  uint32_t crank;
  double inv_cnt;
  float* record_out;
  uint32_t a_E, a_flag, a_mbar;   // local shared addresses: &E[0][gw][lane], &flag[0], &mbar[0]
  __device__ __forceinline__ ClusterSync(PagesClusterShared& s, int gw_, int lane_, uint32_t crank_, double ic, float* rec)
      : sh(s), gw(gw_), lane(lane_), next_red(gw_), kstar(0), crank(crank_), inv_cnt(ic), record_out(rec) {
    uint32_t e = s_u32(&sh.E[0][gw][lane]);
    asm volatile("mov.b32 %0, %0;" : "+r"(e));
    a_E = e;
    a_flag = s_u32(&sh.flag[0]);
    a_mbar = s_u32(&sh.mbar[0]);
  }
  __device__ __forceinline__ void publish(int k, float e) {
    const uint32_t slot = (uint32_t)k & (kRing - 1), dst = ((uint32_t)k & (kGroupWarps - 1)) / kClusterWarps;   // the reducer's CTA
    // st.async: the value lands in the reducer's E buffer and 4 bytes are counted on its mbarrier - no fence, no
    // round trip on this warp's critical path (a release-arrive per iteration costs a DSMEM round trip: measured 2x)
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(mapa_u32(a_E + (slot << 11), dst)), "r"(__float_as_uint(e)), "r"(mapa_u32(a_mbar + 8 * slot, dst)) : "memory");
  }
  template <bool DECIDE>
  __device__ __forceinline__ void duty(int j) {
    if (j == next_red) {
      cluster_reduce(sh, j, lane, inv_cnt, record_out, a_flag, a_mbar);
      next_red += kGroupWarps;
    }
  }
  __device__ __forceinline__ bool verdict(int j) {
    const uint32_t fa = a_flag + 8 * ((uint32_t)j & (kRing - 1));
    unsigned long long w;
    do {
      w = s_ld_u64_volatile(fa);
    } while ((w >> 33) < (unsigned long long)(j + 1));
    const bool better = ((w >> 32) & 1ull) != 0;
    if (better) kstar = j;
    return better;
  }
};

// MODE 0: record + online arg-min (single-round groups); MODE 1: record only, accumulated into sh.recg
// (multi-round groups, pass 1); MODE 2: no record, n_iter iterations, the last iterate is returned in p_out
// (multi-round groups, pass 2).
template <int MODE, typename Sync>
__device__ __forceinline__ void pages_iterate(const PageRegs& M, Sync& sync, float* __restrict__ ps, float* __restrict__ qs,
                                              float* __restrict__ best, int lane, int n_iter, float (&p_out)[8]) {
  const int rh = lane >> 2, kq = lane & 3;
  const int r0 = min(rh, 5), span0 = min(2 * kq, 4);
  const int sp = 8 * r0 + span0;            // first span column
  const int row_base = 32 * rh + 4 * kq;    // matrix row of pixel (2rh + dr, 4kq + cc): row_base + 16 dr + cc
  // Register slot t = 4 ds + cc holds pixel row dr = ds ^ (rh & 1): lanes of odd rh keep their two pixel rows
  // in the opposite order, so that the 128-bit stores of p (8 lanes = rh, rh+1 per wavefront) hit 32 distinct
  // banks.  Nothing else depends on the order: both pixel rows share the window.
  const int off_s[2] = {row_base + 16 * (rh & 1), row_base + 16 * ((rh & 1) ^ 1)};
  constexpr bool RECORD = MODE != 2;
  const float (&f)[8] = M.f;
  const float (&D)[8][12] = M.D;
  // 32-bit shared addresses of everything the loop touches
  // (made opaque to the compiler: left alone it re-derives them from %tid inside the loop - five S2R and some forty
  // integer instructions per iteration - rather than keep them in registers)
  auto keep = [](uint32_t x) {
    asm volatile("mov.b32 %0, %0;" : "+r"(x));
    return x;
  };
  const uint32_t a_ps = keep(s_u32(ps)), a_qs = s_u32(qs);
  const uint32_t a_qspan = keep(a_qs + 4 * sp);
  const uint32_t o_s0 = keep(4 * off_s[0]), o_s1 = keep(4 * off_s[1]), o_span = 4 * sp;
  const uint32_t a_q0 = keep(a_qs + 4 * (off_s[0] >> 2)), a_q1 = keep(a_qs + 4 * (off_s[1] >> 2));   // q index of row rho is rho >> 2
  const uint32_t a_best = s_u32(best);

  const float A = M.A;     // sum over this lane's rows of A_rho (row_constants, computed by the sparsify kernels)
  if (RECORD) sync.publish(0, M.e0);   // the record of iteration 0 (p = q = 1, CP:123)

  auto keep_if_best = [&](int j) {   // end of step j + kLag: wait for the verdict on iteration j
    const bool better = sync.verdict(j);
    if (MODE == 0 && better) {
      if (j == 0) {
        s_st4(a_best + o_s0, 1.f, 1.f, 1.f, 1.f);
        s_st4(a_best + o_s1, 1.f, 1.f, 1.f, 1.f);
      } else {
        const uint32_t src = a_ps + ((j & (kRing - 1)) << 10);
        const float4 x0 = s_ld4(src + o_s0), x1 = s_ld4(src + o_s1);
        s_st4(a_best + o_s0, x0.x, x0.y, x0.z, x0.w);
        s_st4(a_best + o_s1, x1.x, x1.y, x1.z, x1.w);
      }
    }
  };

  float qa = 1.0f, qb = 1.0f, m = 1.0f;   // this lane's two entries of q (slot rows >> 2); centre of the q statistics
  s_st1(a_qs + 4 * lane, 1.0f);
  s_st1(a_qs + 4 * lane + 128, 1.0f);
#pragma unroll
  for (int t = 0; t < 8; ++t) p_out[t] = 1.0f;
  __syncwarp();
  for (int k = 1; k <= n_iter; ++k) {
    if (RECORD) sync.template duty<MODE == 0>(k - kRedDelay);   // this warp's turn as the reducer of iteration k - kRedDelay?
    const uint32_t slot = (uint32_t)k & (kRing - 1);
    // ---- statistics of q_{k-1} about m: S1 = sum (q - m), V = sum (q - m)^2
    const float da = qa - m, db = qb - m;
    float S1 = da + db, V = fmaf(da, da, db * db);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      S1 += __shfl_xor_sync(kFull, S1, o);
      V += __shfl_xor_sync(kFull, V, o);
    }
    const float Q = fmaf(64.0f, m, S1);                                   // sum q
    const float QQ = fmaf(64.0f * m, m, fmaf(2.0f * m, S1, V));           // |q|^2
    const float invA = rcp_newton(QQ + kLambda);                           // torch.inverse of the 1x1 matrix
    // ---- p-update, and the residual of these rows against q_{k-1} (header comment) summed over the rows:
    //      64 sum g^2 + V sum p^2 + A - 2 (S1 sum p g + sum p sD)
    float v[12];
    load_span_u32(a_qspan, v);
    float p[8], gg = 0.f, pg = 0.f, psd = 0.f, pp = 0.f, psum = 0.f;
    auto p_row = [&](auto tc) {
      constexpr int t = decltype(tc)::value;
      const float sD = row_dot<t>(D[t], v);
      p[t] = fmaf(f[t], Q, sD) * invA;
      if (RECORD) {
        const float g = fmaf(-p[t], m, f[t]);
        gg = fmaf(g, g, gg);
        pg = fmaf(p[t], g, pg);
        psd = fmaf(p[t], sD, psd);
      }
      pp = fmaf(p[t], p[t], pp);
      psum += p[t];
    };
    p_row(std::integral_constant<int, 0>{}); p_row(std::integral_constant<int, 1>{});
    p_row(std::integral_constant<int, 2>{}); p_row(std::integral_constant<int, 3>{});
    p_row(std::integral_constant<int, 4>{}); p_row(std::integral_constant<int, 5>{});
    p_row(std::integral_constant<int, 6>{}); p_row(std::integral_constant<int, 7>{});
    const uint32_t pk = a_ps + (slot << 10);
    s_st4(pk + o_s0, p[0], p[1], p[2], p[3]);
    s_st4(pk + o_s1, p[4], p[5], p[6], p[7]);
    if (RECORD) sync.publish(k, fmaf(-2.0f, fmaf(S1, pg, psd), fmaf(64.0f, gg, fmaf(V, pp, A))));
    __syncwarp();
    if (k == n_iter) {   // the reference's last q-update is never used
      if (!RECORD) {
#pragma unroll
        for (int t = 0; t < 8; ++t) p_out[t] = p[t];
      }
      break;
    }
    // ---- q-update: |p|^2, the four segment sums P_r' (segment r' = rows 64r'..64r'+63 = lanes 8r'..8r'+7)
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
      pp += __shfl_xor_sync(kFull, pp, o);
      psum += __shfl_xor_sync(kFull, psum, o);
    }
    pp += __shfl_xor_sync(kFull, pp, 8);
    pp += __shfl_xor_sync(kFull, pp, 16);
    const float invB = rcp_newton(pp + kLambda);
    float ua[4], ub[4];
    const uint32_t pspan = pk + o_span;
    auto q_seg = [&](auto cc_c) {   // row (dr, cc) has rho % 4 = cc: it meets p[64 cc + j]
      constexpr int cc = decltype(cc_c)::value;
      const float Pseg = __shfl_sync(kFull, psum, 8 * cc);
      load_span_u32(pspan + 256 * cc, v);
      ua[cc] = fmaf(f[cc], Pseg, row_dot<cc>(D[cc], v));
      ub[cc] = fmaf(f[4 + cc], Pseg, row_dot<4 + cc>(D[4 + cc], v));
    };
    q_seg(std::integral_constant<int, 0>{}); q_seg(std::integral_constant<int, 1>{});
    q_seg(std::integral_constant<int, 2>{}); q_seg(std::integral_constant<int, 3>{});
    m = Q * (1.0f / 64.0f);
    qa = ((ua[0] + ua[1]) + (ua[2] + ua[3])) * invB;
    qb = ((ub[0] + ub[1]) + (ub[2] + ub[3])) * invB;
    s_st1(a_q0, qa);
    s_st1(a_q1, qb);
    if (RECORD && k >= kLag) keep_if_best(k - kLag);
    __syncwarp();
  }
  if (RECORD) {
    // drain: the reducers of the last kRedDelay iterations, and the verdicts not yet read (the loop read those up to
    // n_iter - 1 - kLag).  Reducers first, in iteration order: every verdict awaited below is then on its way.
    for (int j = max(n_iter - kRedDelay + 1, 0); j <= n_iter; ++j) sync.template duty<MODE == 0>(j);
    for (int j = max(n_iter - kLag, 0); j <= n_iter; ++j) keep_if_best(j);
  }
}

// Normalise p by quick_gm(p, H) with H = rows = 256: prod_i p_i^(1/H^2) (CP:146, CP:244-255: the exponent is
// 2^-16, so p^(1/H^2) = exp(x) with |x| = |ln p| / H^2 < 3e-3: 1 + x + x^2/2 + x^3/6 is exact to f32 rounding and
// p = 1 gives exactly 1) and scatter the page into pages_out / the re-tiled map (CP:218-238 as written: block-row
// j of every block-column holds page j < ratio).
__device__ __forceinline__ void pages_emit(const PagesScaleDev& sc, int64_t unit_idx, const float (&pv)[8], int lane) {
  const int rh = lane >> 2, kq = lane & 3;
  const int row_base = 32 * rh + 4 * kq;
  const int off_s[2] = {row_base + 16 * (rh & 1), row_base + 16 * ((rh & 1) ^ 1)};
  float prod = 1.0f;
  const bool true_gm = (sc.flags & RDM_ALS_TRUE_GM) != 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) prod *= gm_factor(pv[t], 256, true_gm);
  const float gm = warp_prod(prod);
  const int64_t img = unit_idx / sc.pages;
  const int pg = (int)(unit_idx - img * sc.pages);
#pragma unroll
  for (int ds = 0; ds < 2; ++ds) {
    const float4 o = make_float4(pv[4 * ds] / gm, pv[4 * ds + 1] / gm, pv[4 * ds + 2] / gm, pv[4 * ds + 3] / gm);
    const int row = off_s[ds];   // pixel (row >> 4, row & 15 .. +3)
    if (sc.pages_out) *reinterpret_cast<float4*>(sc.pages_out + unit_idx * 256 + row) = o;
    if (sc.map_out) {
      const int side = sc.side, ratio = side >> 4;
      float* mp = sc.map_out + img * (int64_t)side * side;
      if (sc.flags & RDM_ALS_CORRECT_TILING) {   // page (i, j) to block (i, j): what CP:218-238 evidently intended
        const int pi = pg / ratio, pj = pg - pi * ratio;
        *reinterpret_cast<float4*>(mp + (16 * pi + (row >> 4)) * side + 16 * pj + (row & 15)) = o;
      } else if (pg < ratio) {
        for (int bc = 0; bc < ratio; ++bc) *reinterpret_cast<float4*>(mp + (16 * pg + (row >> 4)) * side + 16 * bc + (row & 15)) = o;
      }
    }
  }
}

__device__ __forceinline__ void load_page(PageRegs& M, const float* __restrict__ compact, int lane) {
  const int rh = lane >> 2, kq = lane & 3;
  const int row_base = 32 * rh + 4 * kq;
  const int off_s[2] = {row_base + 16 * (rh & 1), row_base + 16 * ((rh & 1) ^ 1)};
  double asum = 0.0, e0 = 0.0;
#pragma unroll
  for (int dr = 0; dr < 2; ++dr)
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int t = 4 * dr + cc;
      const float4* c4 = reinterpret_cast<const float4*>(compact + (off_s[dr] + cc) * kCompactRowFloats);
      const float4 a = c4[0], b = c4[1], c = c4[2], d = c4[3];
      M.f[t] = a.x;
      M.D[t][0] = a.y; M.D[t][1] = a.z; M.D[t][2] = a.w;
      M.D[t][3] = b.x; M.D[t][4] = b.y; M.D[t][5] = b.z; M.D[t][6] = b.w;
      M.D[t][7] = c.x; M.D[t][8] = c.y; M.D[t][9] = c.z; M.D[t][10] = c.w;
      M.D[t][11] = d.x;
      asum += (double)d.y + (double)d.z;
      e0 += (double)d.w;
    }
  M.A = (float)asum;
  M.e0 = (float)e0;
}

__device__ __forceinline__ bool unit_is_compact(const float* compact) {
  const float4 fl = *reinterpret_cast<const float4*>(compact + kCompactFloats);
  return fl.x == 1.0f && fl.y == 1.0f && fl.z == 1.0f && fl.w == 1.0f;
}

// grid = (group, page) items of every page scale; block = 32 x min(group, 16).
#ifdef RDM_PAGES_MAXNREG   // A/B builds: leave part of the register file to a co-resident CTA of another kernel
__global__ void __maxnreg__(RDM_PAGES_MAXNREG) als_pages_kernel(const __grid_constant__ PagesParams P) {
#else
__global__ void __launch_bounds__(32 * kGroupWarps, 1) als_pages_kernel(const __grid_constant__ PagesParams P) {
#endif
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PagesShared& sh = *reinterpret_cast<PagesShared*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* wrow = reinterpret_cast<float*>(smem_raw + ((sizeof(PagesShared) + 15) & ~size_t(15))) + warp * kWarpFloats;
  float* ps = wrow;
  float* qs = wrow + kRing * 256;
  float* best = qs + 64;
  int si = 0;
#pragma unroll 1
  for (int k = 1; k < P.n_scales; ++k)
    if ((int)blockIdx.x >= P.s[k].cta_begin) si = k;
  const PagesScaleDev& sc = P.s[si];
  const int item = (int)blockIdx.x - sc.cta_begin;
  const int g = item / sc.live, pg = item - g * sc.live;
  const int group = P.group, limit = sc.limit;
  const int W0 = blockDim.x >> 5;                        // warps per round
  const int rounds = (group + W0 - 1) / W0;
  const int64_t stride = als_ws_stride(256, limit);
  const double inv_cnt = 1.0 / ((double)group * (double)(256 * 64));
  float* record_out = sc.record_out ? sc.record_out + ((int64_t)g * sc.pages + pg) * (limit + 1) : nullptr;
  auto unit_of = [&](int r) { return ((int64_t)g * group + r * W0 + warp) * sc.pages + pg; };

  // every unit of the item must have the pair-build structure; otherwise the dense kernel takes the whole item
  if (threadIdx.x == 0) sh.all_compact = 1;
  if (threadIdx.x < kRing) sh.flag[threadIdx.x] = -1;
  if (threadIdx.x == 0) {
    sh.best_rmse = __int_as_float(0x7f800000);
    sh.kstar = 0;
  }
  for (int k = threadIdx.x; k <= limit; k += blockDim.x) sh.recg[k] = 0.0;
  __syncthreads();
  for (int r = 0; r < rounds; ++r)
    if (r * W0 + warp < group && lane == 0 && !unit_is_compact(sc.ws + unit_of(r) * stride)) sh.all_compact = 0;
  __syncthreads();
  if (!sh.all_compact) return;

  PageRegs M;
  float pv[8];
  if (rounds == 1) {
    load_page(M, sc.ws + unit_of(0) * stride, lane);
#pragma unroll
    for (int dr = 0; dr < 2; ++dr)
      *reinterpret_cast<float4*>(best + 32 * (lane >> 2) + 4 * (lane & 3) + 16 * dr) = make_float4(1.f, 1.f, 1.f, 1.f);
    __syncwarp();
    LocalSync sync(sh, warp, lane, W0, inv_cnt, record_out);
    pages_iterate<0>(M, sync, ps, qs, best, lane, limit, pv);
    {
      const int rh = lane >> 2, kq = lane & 3;
      const int row_base = 32 * rh + 4 * kq;
      const int off_s[2] = {row_base + 16 * (rh & 1), row_base + 16 * ((rh & 1) ^ 1)};
#pragma unroll
      for (int ds = 0; ds < 2; ++ds) {
        const float4 b = *reinterpret_cast<const float4*>(best + off_s[ds]);
        pv[4 * ds] = b.x; pv[4 * ds + 1] = b.y; pv[4 * ds + 2] = b.z; pv[4 * ds + 3] = b.w;
      }
    }
    pages_emit(sc, unit_of(0), pv, lane);
    __syncthreads();
    if (threadIdx.x == 0 && sc.kstar_out) sc.kstar_out[(int64_t)g * sc.pages + pg] = sh.kstar;
    return;
  }
  // ---- multi-round groups: pass 1 = group record, pass 2 = replay of the selected iterations
  for (int r = 0; r < rounds; ++r) {
    const int Wr = min(W0, group - r * W0);
    if (warp < Wr) {
      load_page(M, sc.ws + unit_of(r) * stride, lane);
      LocalSync sync(sh, warp, lane, Wr, inv_cnt, nullptr);
      pages_iterate<1>(M, sync, ps, qs, best, lane, limit, pv);
    }
    __syncthreads();
    if (threadIdx.x < kRing) sh.flag[threadIdx.x] = -1;
    __syncthreads();
  }
  if (warp == 0) {   // first minimum of the group record (CP:143), as a min over (value bits, index) keys
    unsigned long long key = ~0ull;
    for (int k = lane; k <= limit; k += 32) {
      const float rm = (float)sqrt(sh.recg[k] * inv_cnt);
      if (record_out) record_out[k] = rm;
      const unsigned long long kk = ((unsigned long long)__float_as_uint(rm) << 32) | (unsigned)k;
      key = kk < key ? kk : key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(kFull, key, o);
      key = other < key ? other : key;
    }
    if (lane == 0) {
      sh.kstar = (int)(key & 0xffffffffu);
      if (sc.kstar_out) sc.kstar_out[(int64_t)g * sc.pages + pg] = sh.kstar;
    }
  }
  __syncthreads();
  const int kstar = sh.kstar;
  for (int r = 0; r < rounds; ++r) {
    if (r * W0 + warp >= group) break;
    load_page(M, sc.ws + unit_of(r) * stride, lane);
    LocalSync sync(sh, warp, lane, 1, inv_cnt, nullptr);
    pages_iterate<2>(M, sync, ps, qs, best, lane, kstar, pv);
    pages_emit(sc, unit_of(r), pv, lane);
  }
}

// The same work with the 16 warps of a group spread over a cluster of 4 CTAs (ClusterSync).  grid = 4 x (group, page)
// items; group must be 16.
#ifdef RDM_PAGES_CLUSTER_MAXNREG   // A/B builds: leave part of the register file to co-resident CTAs of the sparsify kernel
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __maxnreg__(RDM_PAGES_CLUSTER_MAXNREG)
#else
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(32 * kClusterWarps, RDM_PAGES_CLUSTER_MINB)
#endif
    als_pages_cluster_kernel(const __grid_constant__ PagesParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PagesClusterShared& sh = *reinterpret_cast<PagesClusterShared*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* wrow = reinterpret_cast<float*>(smem_raw + ((sizeof(PagesClusterShared) + 15) & ~size_t(15))) + warp * kWarpFloats;
  float* ps = wrow;
  float* qs = wrow + kRing * 256;
  float* best = qs + 64;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int item_all = (int)(blockIdx.x / kClusterCtas);
  int si = 0;
#pragma unroll 1
  for (int k = 1; k < P.n_scales; ++k)
    if (item_all >= P.s[k].cta_begin) si = k;
  const PagesScaleDev& sc = P.s[si];
  const int item = item_all - sc.cta_begin;
  const int g = item / sc.live, pg = item - g * sc.live;
  const int limit = sc.limit;
  const int gw = (int)crank * kClusterWarps + warp;                 // image of the group this warp iterates
  const int64_t stride = als_ws_stride(256, limit);
  const double inv_cnt = 1.0 / ((double)kGroupWarps * (double)(256 * 64));
  float* record_out = sc.record_out ? sc.record_out + ((int64_t)g * sc.pages + pg) * (limit + 1) : nullptr;
  const int64_t unit = ((int64_t)g * kGroupWarps + gw) * sc.pages + pg;
  if (threadIdx.x < kRing) {
    sh.flag[threadIdx.x] = 0ull;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&sh.mbar[threadIdx.x])), "r"(1) : "memory");
  }
  if (threadIdx.x == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // every CTA checks all 16 units of the item itself, so that the four CTAs decide alike without talking
  if (warp == 0) {
    const bool ok = lane >= kGroupWarps || unit_is_compact(sc.ws + (((int64_t)g * kGroupWarps + lane) * sc.pages + pg) * stride);
    const unsigned all = __ballot_sync(kFull, ok);
    if (lane == 0) sh.all_compact = all == kFull ? 1 : 0;
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // mbarriers and flags are initialised everywhere
  if (!sh.all_compact) return;   // the dense kernel takes the whole item
  PageRegs M;
  float pv[8];
  load_page(M, sc.ws + unit * stride, lane);
#pragma unroll
  for (int dr = 0; dr < 2; ++dr)
    *reinterpret_cast<float4*>(best + 32 * (lane >> 2) + 4 * (lane & 3) + 16 * dr) = make_float4(1.f, 1.f, 1.f, 1.f);
  __syncwarp();
  ClusterSync sync(sh, gw, lane, crank, inv_cnt, record_out);
  pages_iterate<0>(M, sync, ps, qs, best, lane, limit, pv);
  {
    const int rh = lane >> 2, kq = lane & 3;
    const int row_base = 32 * rh + 4 * kq;
    const int off_s[2] = {row_base + 16 * (rh & 1), row_base + 16 * ((rh & 1) ^ 1)};
#pragma unroll
    for (int ds = 0; ds < 2; ++ds) {
      const float4 b = *reinterpret_cast<const float4*>(best + off_s[ds]);
      pv[4 * ds] = b.x; pv[4 * ds + 1] = b.y; pv[4 * ds + 2] = b.z; pv[4 * ds + 3] = b.w;
    }
  }
  pages_emit(sc, unit, pv, lane);
  if (gw == 0 && lane == 0 && sc.kstar_out) sc.kstar_out[(int64_t)g * sc.pages + pg] = sync.kstar;
  // nobody leaves while a sibling may still store into its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace

// Launch the compact path for every eligible scale (256-row units given as f64 matrices or as maps).
int als_sparse_launch(const rdm_als_scale_t* scales, int32_t n_scales, int64_t n_images, int32_t group, bool sparsify, bool iterate,
                      cudaStream_t stream) {
  SparseParams raw{}, map{};
  PagesParams all{};
  raw.n_images = map.n_images = all.n_images = n_images;
  all.group = group;
  int64_t n_raw = 0, n_map = 0, n_items = 0;
  const int64_t n_groups = n_images / group;
  for (int k = 0; k < n_scales; ++k) {
    const rdm_als_scale_t& h = scales[k];
    if (h.rows != 256 || (h.flags & (RDM_ALS_DENSE_ONLY | RDM_ALS_TRUE_TRANSPOSE))) continue;   // the transpose variant lives in the dense kernel
    const bool is_map = h.src_kind == RDM_SRC_MAP_F32;
    if (!(is_map || h.src_kind == RDM_SRC_RAW_F64 || h.src_kind == RDM_SRC_VAL_F64)) continue;
    const bool quant = h.src_kind != RDM_SRC_VAL_F64;
    SparseScaleDev d;
    d.src = h.src;
    d.thr = quant ? h.thresholds : nullptr;
    d.lvl = quant ? h.levels : nullptr;
    d.bins = h.bins_out;
    d.values = h.values_out;
    d.ws = h.ws;
    d.kind = h.src_kind;
    d.pages = h.pages;
    d.side = h.side;
    d.limit = h.limit;
    // CP:218-238 as written copies only pages 0 .. side/16 - 1 into the map; the others are computed by the reference and
    // dropped.  RDM_ALS_SKIP_UNUSED_PAGES leaves them out (their pages_out / record / kstar / bins entries stay unwritten).
    const bool skip = (h.flags & RDM_ALS_SKIP_UNUSED_PAGES) && !(h.flags & RDM_ALS_CORRECT_TILING);
    d.live = skip ? h.side / 16 : h.pages;
    const int64_t units = n_images * d.live;
    SparseParams& part = is_map ? map : raw;
    int64_t& n_part = is_map ? n_map : n_raw;
    d.unit_begin = (int32_t)n_part;
    part.s[part.n_scales++] = d;
    n_part += units;
    PagesScaleDev& a = all.s[all.n_scales++];
    a.ws = h.ws;
    a.pages_out = h.pages_out;
    a.map_out = h.map_out;
    a.record_out = h.record_out;
    a.kstar_out = h.kstar_out;
    a.pages = h.pages;
    a.side = h.side;
    a.limit = h.limit;
    a.cta_begin = (int32_t)n_items;
    a.flags = h.flags;
    a.live = d.live;
    n_items += n_groups * d.live;
  }
  if (n_items == 0) return 0;
  RDM_REQUIRE(n_raw + n_map < (1ll << 28), "rdm_als_fused: too many work units");
  if (n_raw && sparsify) {
    static size_t smem_sp[64];
    cudaError_t es = ensure_dyn_smem(als_sparsify_raw_kernel, kStageBytes, smem_sp);
    if (es != cudaSuccess) {
      set_error("rdm_als_fused: cudaFuncSetAttribute(als_sparsify_raw_kernel): %s", cudaGetErrorString(es));
      return (int)es;
    }
    als_sparsify_raw_kernel<<<(unsigned)(4 * n_raw), kBandRows, kStageBytes, stream>>>(raw);
    int rc = launch_status("als_sparsify_raw_kernel");
    if (rc) return rc;
  }
  if (n_map && sparsify) {
    als_sparsify_map_kernel<<<(unsigned)n_map, 256, 0, stream>>>(map);
    int rc = launch_status("als_sparsify_map_kernel");
    if (rc) return rc;
  }
  if (!iterate) return 0;
  // Two forms of the page kernel, same results bit for bit.  One CTA of 16 warps per (batch, page) has the better
  // throughput once the chip is full; a cluster of 4 CTAs of 4 warps spreads a SMALL launch over 4x the SMs (a lone
  // batch-16 call: 66 vs 131 us) and wins as long as every cluster is resident at once - a second
  // wave of clusters costs a whole extra pass (29 batches per launch: 7.5 vs 4.7 us per batch).  Default: the cluster
  // form while all clusters are resident at once; RDM_ALS_PAGES_ONE_CTA / RDM_ALS_PAGES_CLUSTER force one.
  const size_t dync = ((sizeof(PagesClusterShared) + 15) & ~size_t(15)) + (size_t)kClusterWarps * kWarpFloats * sizeof(float);
  bool one_cta = group != kGroupWarps;
  if (!one_cta) {
    static int max_clusters[64];   // per device, queried once (benign race)
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    if (max_clusters[slot] == 0) {
      static size_t smem_setq[64];
      int n = 0;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kClusterCtas * 1024);
      cfg.blockDim = dim3(32 * kClusterWarps);
      cfg.dynamicSmemBytes = dync;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kClusterCtas;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (ensure_dyn_smem(als_pages_cluster_kernel, dync, smem_setq) != cudaSuccess ||
          cudaOccupancyMaxActiveClusters(&n, als_pages_cluster_kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = kNumSMs / 2;   // conservative
      }
      max_clusters[slot] = n;
      if (getenv("RDM_B200_DEBUG")) fprintf(stderr, "rdm_b200: als_pages_cluster_kernel: %d clusters of %d CTAs resident at once\n", n, kClusterCtas);
    }
    // measured, us per batch (one CTA / cluster): 40 items 16.8 / 11.0, 70: 9.6 / 6.3, 85: 8.0 / 6.5, 100: 6.8 / 5.6
    // (104 clusters resident), 145: 4.7 / 7.5
    one_cta = n_items > max_clusters[slot];
  }
  for (int k = 0; k < n_scales; ++k) {
    if (scales[k].flags & RDM_ALS_PAGES_CLUSTER) one_cta = group != kGroupWarps;
    if (scales[k].flags & RDM_ALS_PAGES_ONE_CTA) one_cta = true;
  }
  if (!one_cta) {
    static size_t smem_setc[64];
    cudaError_t ec = ensure_dyn_smem(als_pages_cluster_kernel, dync, smem_setc);
    if (ec != cudaSuccess) {
      set_error("rdm_als_fused: cudaFuncSetAttribute(als_pages_cluster_kernel): %s", cudaGetErrorString(ec));
      return (int)ec;
    }
    als_pages_cluster_kernel<<<(unsigned)(n_items * kClusterCtas), 32 * kClusterWarps, dync, stream>>>(all);
    return launch_status("als_pages_cluster_kernel");
  }
  const int warps = group < kGroupWarps ? group : kGroupWarps;
  const size_t dyn = ((sizeof(PagesShared) + 15) & ~size_t(15)) + (size_t)warps * kWarpFloats * sizeof(float);
  static size_t smem_set[64];
  cudaError_t e = ensure_dyn_smem(als_pages_kernel, dyn, smem_set);
  if (e != cudaSuccess) {
    set_error("rdm_als_fused: cudaFuncSetAttribute(als_pages_kernel): %s", cudaGetErrorString(e));
    return (int)e;
  }
  als_pages_kernel<<<(unsigned)n_items, 32 * warps, dyn, stream>>>(all);
  return launch_status("als_pages_kernel");
}

}  // namespace rdm

extern "C" int rdm_conv_head_f32(const float* feat, const float* weight, const float* bias, int64_t n_images, int32_t channels, int32_t side,
                                 float* map_out, const rdm_als_scale_t* page_scale, rdm_stream_t stream) {
  using namespace rdm;
  RDM_REQUIRE(feat && weight, "rdm_conv_head_f32: null pointer");
  RDM_REQUIRE(side == 8 || side == 16 || side == 32 || side == 64, "rdm_conv_head_f32: side must be 8, 16, 32 or 64 (got %d)", side);
  RDM_REQUIRE(channels >= kConvCluster && channels % kConvCluster == 0, "rdm_conv_head_f32: channels (%d) must be a multiple of %d", channels, kConvCluster);
  RDM_REQUIRE(aligned16(feat) && (!map_out || aligned16(map_out)), "rdm_conv_head_f32: feat and map_out must be 16-byte aligned");
  RDM_REQUIRE(n_images >= 0 && n_images * kConvCluster < (1ll << 31), "rdm_conv_head_f32: bad n_images");
  RDM_REQUIRE(map_out || page_scale, "rdm_conv_head_f32: nothing to produce (map_out and page_scale are both NULL)");
  SparseScaleDev d{};
  int build = 0;
  if (page_scale) {
    const rdm_als_scale_t& h = *page_scale;
    RDM_REQUIRE(side >= 16, "rdm_conv_head_f32: the compact page form exists for sides >= 16 (the 8x8 scale takes map_out)");
    RDM_REQUIRE(h.rows == 256 && h.side == side && h.pages == (side / 16) * (side / 16), "rdm_conv_head_f32: page_scale does not describe a %dx%d map", side, side);
    RDM_REQUIRE(h.ws && aligned16(h.ws) && h.thresholds && h.levels, "rdm_conv_head_f32: page_scale needs ws (16-byte aligned) and the codebook");
    RDM_REQUIRE(h.limit >= 0 && h.limit <= 127, "rdm_conv_head_f32: limit out of range");
    d.src = nullptr;
    d.thr = h.thresholds;
    d.lvl = h.levels;
    d.bins = h.bins_out;
    d.values = h.values_out;
    d.ws = h.ws;
    d.kind = RDM_SRC_MAP_F32;
    d.pages = h.pages;
    d.side = side;
    d.limit = h.limit;
    d.unit_begin = 0;
    d.live = ((h.flags & RDM_ALS_SKIP_UNUSED_PAGES) && !(h.flags & RDM_ALS_CORRECT_TILING)) ? side / 16 : h.pages;
    build = 1;
  }
  if (n_images == 0) return 0;
  static size_t smem_conv[64];
  cudaError_t e = ensure_dyn_smem(conv_head_kernel, sizeof(ConvSmem), smem_conv);
  if (e != cudaSuccess) {
    set_error("rdm_conv_head_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_images * kConvCluster));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = sizeof(ConvSmem);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kConvCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, conv_head_kernel, feat, weight, bias, (int)channels, (int)side, map_out, d, build);
  if (e != cudaSuccess) {
    set_error("rdm_conv_head_f32: launch: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return launch_status("conv_head_kernel");
}

// Host copy of the compact page form's geometry (computed by the same RowGeom the kernels use), so that the CPU test
// suite can check it against the oracle's window mask without a GPU: per matrix row the nine window columns
// (row-major over the 3x3 window), the fill reference column, and the source of each entry of the 16-float compact
// row (0..8 = window slot minus f, 9 = the fill level f itself, 15 = zero).
extern "C" int rdm_sparsify_geometry(int32_t* window_cols, int32_t* fill_col, uint8_t* compact) {
  RDM_REQUIRE(window_cols && fill_col && compact, "rdm_sparsify_geometry: null pointer");
  for (int rho = 0; rho < 256; ++rho) {
    const rdm::RowGeom g(rho);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) window_cols[rho * 9 + 3 * a + b] = g.window_col(a, b);
    fill_col[rho] = g.fill_col();
    for (int e = 0; e < 16; ++e) {
      uint8_t code = 15;
      if (e == 0) code = 9;
      else if (e <= 12) {
        const int b = g.span0 + ((e - 1) & 3) - g.c0;
        if (b >= 0 && b < 3) code = (uint8_t)(3 * ((e - 1) >> 2) + b);
      }
      compact[rho * 16 + e] = code;
    }
  }
  return 0;
}
