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
#include <type_traits>

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
// Geometry tables of the sparsify kernel (built at compile time).  A warp owns 8 matrix rows = pixels
// (r, cb..cb+7) of the page, cb in {0, 8}; all share r0 = min(r/2, 5); pixel column c has c0 = min(c/2, 5) and
// span0 = min(2 (c/4), 4).  "Slot" numbers the 10 informative values of a row: 0..8 the 3x3 window (row-major),
// 9 the fill value.
struct SparsifyTables {
  // lane_slots[cb/8][r0][lane]: byte i = slot of matrix column 2*lane (low nibble) and 2*lane+1 (high nibble) in row i
  unsigned long long lane_slots[2][6][32];
  // item[cb/8][id], id = 10 i + slot < 80: index into the row's staging line (0..23 parent rows x columns, 24 fill)
  // | i << 5 | slot << 8 | valid << 12
  unsigned short item[2][96];
  // compact[c][e]: source of entry e of the 16-float compact row of pixel column c: 0..8 window slot (value - f),
  // 9 the fill value itself, 15 zero
  unsigned char compact[16][16];
};
constexpr SparsifyTables make_sparsify_tables() {
  SparsifyTables t{};
  for (int h = 0; h < 2; ++h)
    for (int r0 = 0; r0 < 6; ++r0)
      for (int lane = 0; lane < 32; ++lane) {
        unsigned long long w = 0;
        for (int i = 0; i < 8; ++i) {
          const int c = 8 * h + i, c0 = (c >> 1) < 5 ? (c >> 1) : 5;
          const int a = (lane >> 2) - r0, cc = (2 * lane) & 7;
          unsigned long long sl[2] = {9, 9};
          for (int e = 0; e < 2; ++e)
            if (a >= 0 && a < 3 && cc + e - c0 >= 0 && cc + e - c0 < 3) sl[e] = (unsigned long long)(3 * a + cc + e - c0);
          w |= (sl[0] | (sl[1] << 4)) << (8 * i);
        }
        t.lane_slots[h][r0][lane] = w;
      }
  for (int h = 0; h < 2; ++h)
    for (int id = 0; id < 96; ++id) {
      if (id >= 80) { t.item[h][id] = 0; continue; }
      const int i = id / 10, slot = id % 10;
      const int c = 8 * h + i, c0 = (c >> 1) < 5 ? (c >> 1) : 5;
      const int src = slot == 9 ? 24 : 8 * (slot / 3) + c0 + slot % 3;
      t.item[h][id] = (unsigned short)(src | (i << 5) | (slot << 8) | (1 << 12));
    }
  for (int c = 0; c < 16; ++c) {
    const int c0 = (c >> 1) < 5 ? (c >> 1) : 5, span0 = 2 * (c >> 2) < 4 ? 2 * (c >> 2) : 4;
    for (int e = 0; e < 16; ++e) {
      unsigned char code = 15;
      if (e == 0) code = 9;
      else if (e <= 12) {
        const int be = span0 + ((e - 1) & 3) - c0;
        if (be >= 0 && be < 3) code = (unsigned char)(3 * ((e - 1) >> 2) + be);
      }
      t.compact[c][e] = code;
    }
  }
  return t;
}
__device__ const SparsifyTables kSparsifyTables = make_sparsify_tables();

// grid = units x 4 bands of 64 rows; a warp owns 8 rows, a lane two adjacent matrix columns of each.  Only 10
// values per row carry information, so after the bit-wise structure check the warp gathers those 80 values
// through shared memory and quantises them in three full-warp passes instead of quantising 512 values in
// sixteen.  The per-lane geometry comes from kSparsifyTables (three small loads per warp).
__global__ void __launch_bounds__(256) als_sparsify_raw_kernel(const __grid_constant__ SparseParams P) {
  __shared__ double thr_d[kThrPad];
  __shared__ float lvl_f[kLvl + 3];
  __shared__ int sorted;
  __shared__ __align__(16) double stage[8][8][26];   // per warp, per row: 3 parent rows x 8 columns, then the fill value
  __shared__ float resv[8][8][12];                   // quantised values by slot
  __shared__ uint8_t resb[8][8][16];                 // their bins
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gunit = blockIdx.x >> 2, band = blockIdx.x & 3;
  const SparseScaleDev& sc = find_scale(P, gunit);
  const int64_t unit = gunit - sc.unit_begin;
  const bool quant = sc.kind == RDM_SRC_RAW_F64;
  const int64_t mat_off = unit * (int64_t)(256 * 64);
  const int row0 = band * 64 + warp * 8;             // rows row0 .. row0+7: pixel row row0 >> 4, columns cb .. cb+7
  const double* src = reinterpret_cast<const double*>(sc.src) + mat_off + row0 * 64 + 2 * lane;
  double2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = ldg_stream_f64x2(src + i * 64);
  const int r0 = min(row0 >> 5, 5), h = (row0 >> 3) & 1;
  const unsigned long long slots = kSparsifyTables.lane_slots[h][r0][lane];
  const unsigned it0 = kSparsifyTables.item[h][lane], it1 = kSparsifyTables.item[h][32 + lane], it2 = kSparsifyTables.item[h][64 + lane];
  const unsigned centry = *reinterpret_cast<const unsigned*>(&kSparsifyTables.compact[8 * h + (lane >> 2)][4 * (lane & 3)]);
  if (quant) load_book(sc, thr_d, lvl_f, &sorted, tid, 256);
  float* compact = sc.ws + unit * als_ws_stride(256, sc.limit);
  const int srt = quant ? sorted : 1;
  const int lane_f = (r0 >= 1) ? 0 : 28;             // column 0 / 56 is outside every window of this pixel row
  const int a = (lane >> 2) - r0;
  const bool rowin = (unsigned)a < 3u;
  double* my_stage = &stage[warp][0][rowin ? 8 * a + ((2 * lane) & 7) : 24];
  const unsigned slo = (unsigned)slots, shi = (unsigned)(slots >> 32);
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const unsigned sb = ((i < 4 ? slo : shi) >> (8 * (i & 3))) & 0xffu;   // slots of this lane's two columns in row i
    const long long b0 = __double_as_longlong(x[i].x), b1 = __double_as_longlong(x[i].y);
    const long long fb = __shfl_sync(kFull, b0, lane_f);
    ok = ok && ((sb & 15u) != 9u || b0 == fb) && ((sb >> 4) != 9u || b1 == fb);
    if (rowin) *reinterpret_cast<double2*>(my_stage + 26 * i) = x[i];
    else if (lane == lane_f) my_stage[26 * i] = x[i].x;
  }
  __syncwarp();
  auto quantise_item = [&](unsigned g) {
    if (g & 0x1000u) {
      const int i = (g >> 5) & 7, slot = (g >> 8) & 15;
      const double v = stage[warp][i][g & 31u];
      int q = 0;
      float lv;
      if (quant) {
        q = lloyd_bin<double>(v, thr_d, srt);
        lv = lvl_f[q];
      } else {
        lv = (float)v;   // `.float()` CP:106
      }
      resv[warp][i][slot] = lv;
      resb[warp][i][slot] = (uint8_t)q;
    }
  };
  quantise_item(it0);
  quantise_item(it1);
  quantise_item(it2);
  __syncwarp();
  {   // compact form: 8 rows x 16 floats, one float4 per lane
    const int i = lane >> 2;
    const float f = resv[warp][i][9];
    float o[4];
#pragma unroll
    for (int e4 = 0; e4 < 4; ++e4) {
      const unsigned code = (centry >> (8 * e4)) & 15u;
      const float wv = resv[warp][i][code == 15u ? 9 : code];
      o[e4] = code == 9u ? f : (code == 15u ? 0.f : wv - f);
    }
    *reinterpret_cast<float4*>(compact + (row0 + i) * kCompactRowFloats + 4 * (lane & 3)) = make_float4(o[0], o[1], o[2], o[3]);
  }
  if (sc.bins || sc.values) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned sb = ((i < 4 ? slo : shi) >> (8 * (i & 3))) & 0xffu;
      const int s0 = sb & 15u, s1 = sb >> 4;
      const int64_t off = mat_off + (row0 + i) * 64 + 2 * lane;
      if (sc.bins) *reinterpret_cast<uint16_t*>(sc.bins + off) = (uint16_t)(resb[warp][i][s0] | (resb[warp][i][s1] << 8));
      if (sc.values) *reinterpret_cast<float2*>(sc.values + off) = make_float2(resv[warp][i][s0], resv[warp][i][s1]);
    }
  }
  const int all_ok = __syncthreads_and(ok ? 1 : 0);
  if (tid == 0) compact[kCompactFloats + band] = all_ok ? 1.0f : 0.0f;
}

// grid = units; thread = matrix row (pixel of the page).  RN:259-284 + CP:269-295 + CP:308-311 fused.
__global__ void __launch_bounds__(256) als_sparsify_map_kernel(const __grid_constant__ SparseParams P) {
  __shared__ double thr_d[kThrPad];
  __shared__ double inv_d[64];
  __shared__ float lvl_f[kLvl + 3];
  __shared__ int sorted;
  const int tid = threadIdx.x;
  const int gunit = blockIdx.x;
  const SparseScaleDev& sc = find_scale(P, gunit);
  const int64_t unit = gunit - sc.unit_begin;
  const int side = sc.side, ratio = side >> 4;
  const int64_t img = unit / sc.pages;
  const int pg = (int)(unit - img * sc.pages);
  const int pi = pg / ratio, pj = pg - pi * ratio;
  const float* map = reinterpret_cast<const float*>(sc.src) + img * (int64_t)side * side;
  if (tid < 64) {
    const int y = 8 * pi + (tid >> 3), x = 8 * pj + (tid & 7);
    const double v = bicubic_half_at([&](int r, int c) { return (double)map[r * side + c]; }, y, x, side);
    inv_d[tid] = 1.0 / v;   // torch.pow(area,-1): IEEE reciprocal
  }
  load_book(sc, thr_d, lvl_f, &sorted, tid, 256);   // ends with __syncthreads
  const int srt = sorted;
  const int row = tid;
  const double d = (double)map[(16 * pi + (row >> 4)) * side + 16 * pj + (row & 15)];
  const int r0 = min(row >> 5, 5), c0 = min((row & 15) >> 1, 5), span0 = min(((row & 15) >> 2) * 2, 4);
  const int bf = lloyd_bin<double>(d, thr_d, srt);   // 55 of 64 columns hold d itself
  const float f = lvl_f[bf];
  float* compact = sc.ws + unit * als_ws_stride(256, sc.limit);
  float o[16];
  o[0] = f;
  o[13] = o[14] = o[15] = 0.f;
  int wb[9];   // window bins, row-major over the 3x3 window
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) wb[3 * a + b] = lloyd_bin<double>(__dmul_rn(d, inv_d[8 * (r0 + a) + c0 + b]), thr_d, srt);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int b = span0 + g - c0;   // window column of span column g
      float v = 0.f;
#pragma unroll
      for (int bb = 0; bb < 3; ++bb)
        if (b == bb) v = lvl_f[wb[3 * a + bb]] - f;
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

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_newton(float x) {   // see rdm_als.cu
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(fmaf(-x, r, 1.0f), r, r);
}

__device__ __forceinline__ void load_span(const float* base, float (&v)[12]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float2 lo = *reinterpret_cast<const float2*>(base + 8 * a);
    const float2 hi = *reinterpret_cast<const float2*>(base + 8 * a + 2);
    v[4 * a] = lo.x;
    v[4 * a + 1] = lo.y;
    v[4 * a + 2] = hi.x;
    v[4 * a + 3] = hi.y;
  }
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
//  * Iteration k of a warp leaves its 32 per-lane residuals in E[k % 3][warp][lane] and ARRIVES (bar.arrive, no
//    wait) on named barrier 1 + k % 3.
//  * Warp (k % W) is the reducer of iteration k: at the top of its step k+1 it waits for the arrivals
//    (bar.sync on the same barrier), sums the group's residuals (f64), takes the rmse, compares it with the
//    running minimum (strict <: the FIRST minimum wins, as list.index(min(list)) does) and publishes
//    flag[k % 3] = 2 k + new_minimum.
//  * Every warp reads the verdict on iteration k at the end of its step k + kLag, and on a new minimum copies
//    p_k from its 3-deep ring of iterates into its `best` row.  The flag wait also bounds the skew between the
//    warps, which is what makes the 3-deep rings and the 3 barriers safe to reuse (see the ordering argument
//    in DESIGN.md 4.1).
// Groups of more than 16 images take kRounds rounds of 16 warps: pass 1 accumulates the group record over the
// rounds (no iterate is kept), pass 2 re-runs the k* selected iterations of every unit (bit-identical
// arithmetic) and emits them.  On noise-like maps k* <= 1, so pass 2 is a few per cent of pass 1.
constexpr int kLag = 2;
constexpr int kRing = kLag + 1;
static_assert(kLag == 2 && kRing == 3, "the ring slot arithmetic of pages_iterate assumes a lag of 2");
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
};
struct PagesParams {
  PagesScaleDev s[kMaxSparseScales];
  int64_t n_images;
  int32_t n_scales, group;
};

struct PagesShared {            // fixed part of the dynamic shared memory (the per-warp rows follow)
  float E[kRing][kGroupWarps][32];
  double recg[kMaxLimit + 1];   // group record accumulated over the rounds (multi-round groups)
  unsigned flag[kRing];
  float best_rmse;
  int kstar;
  int all_compact;
};

__device__ __forceinline__ void named_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}

// The matrix of one page in registers: a lane owns the 2 x 4 pixel block (rows 2rh..2rh+1, columns 4kq..4kq+3)
// = 8 matrix rows, and with them the two entries q[8rh+kq], q[8rh+4+kq].
struct PageRegs {
  float f[8], D[8][12];
};

// MODE 0: record + online arg-min (single-round groups); MODE 1: record only, accumulated into sh.recg
// (multi-round groups, pass 1); MODE 2: no record, n_iter iterations, the last iterate is returned in p_out
// (multi-round groups, pass 2).
template <int MODE>
__device__ __forceinline__ void pages_iterate(const PageRegs& M, PagesShared& sh, float* __restrict__ ps, float* __restrict__ qs,
                                              float* __restrict__ best, int lane, int warp, int W, int n_iter, double inv_cnt,
                                              float* __restrict__ record_out, float (&p_out)[8]) {
  const int rh = lane >> 2, kq = lane & 3;
  const int r0 = min(rh, 5), span0 = min(2 * kq, 4);
  const int sp = 8 * r0 + span0;            // first span column
  const int row_base = 32 * rh + 4 * kq;    // matrix row of pixel (2rh + dr, 4kq + cc): row_base + 16 dr + cc
  // Register slot t = 4 ds + cc holds pixel row dr = ds ^ (rh & 1): lanes of odd rh keep their two pixel rows
  // in the opposite order, so that the 128-bit stores of p (8 lanes = rh, rh+1 per wavefront) hit 32 distinct
  // banks.  Nothing else depends on the order: both pixel rows share the window.
  const int off_s[2] = {row_base + 16 * (rh & 1), row_base + 16 * ((rh & 1) ^ 1)};
  const int nbar = (W + 1) * 32;
  constexpr bool RECORD = MODE != 2;
  const float (&f)[8] = M.f;
  const float (&D)[8][12] = M.D;

  float A = 0.f;
  if (RECORD) {
    // A = sum over this lane's rows of A_rho, and the record of iteration 0 (p = q = 1, CP:123):
    // sum_j fl(1 - R)^2 with 52 columns outside the span
    double e0 = 0.0, asum = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float u = __fsub_rn(1.0f, f[t]);
      double acc = 52.0 * ((double)u * (double)u);
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        asum = fma((double)D[t][e], 2.0 * (double)f[t] + (double)D[t][e], asum);
        const float w = __fsub_rn(1.0f, __fadd_rn(f[t], D[t][e]));
        acc = fma((double)w, (double)w, acc);
      }
      e0 += acc;
    }
    A = (float)asum;
    sh.E[0][warp][lane] = (float)e0;
    named_arrive(1, nbar);
  }

  // the reducer's work for iteration j (one warp, all lanes): group rmse, first-minimum test, verdict
  auto reduce_iteration = [&](int j) {
    const int slot = j % kRing;   // once per W iterations per warp
    named_sync(1 + slot, nbar);
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
    g = __shfl_sync(kFull, g, 0);
    unsigned verdict = (unsigned)j << 1;
    if (MODE == 0) {
      const float rm = (float)sqrt(g * inv_cnt);           // CP:172-173 over the whole reference batch
      bool better = false;
      if (lane == 0) {
        better = rm < sh.best_rmse;                        // strict: the first minimum wins (CP:143)
        if (better) {
          sh.best_rmse = rm;
          sh.kstar = j;
        }
        if (record_out) record_out[j] = rm;
      }
      verdict |= better ? 1u : 0u;
    } else if (lane == 0) {
      sh.recg[j] += g;
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      *reinterpret_cast<volatile unsigned*>(&sh.flag[slot]) = verdict;
    }
  };
  auto await_verdict = [&](int j, int slot) -> bool {
    const unsigned* fp = &sh.flag[slot];
    unsigned v;
    do {
      v = ld_volatile_u32(fp);
    } while ((v >> 1) != (unsigned)j);
    return (v & 1u) != 0u;
  };
  auto keep_if_best = [&](int j, int slot) {   // end of step j + kLag; slot = j % kRing
    if (await_verdict(j, slot) && MODE == 0) {
      if (j == 0) {
#pragma unroll
        for (int dr = 0; dr < 2; ++dr) *reinterpret_cast<float4*>(best + off_s[dr]) = make_float4(1.f, 1.f, 1.f, 1.f);
      } else {
        const float* src = ps + slot * 256;
#pragma unroll
        for (int dr = 0; dr < 2; ++dr) *reinterpret_cast<float4*>(best + off_s[dr]) = *reinterpret_cast<const float4*>(src + off_s[dr]);
      }
    }
  };

  float qa = 1.0f, qb = 1.0f, m = 1.0f;   // this lane's two entries of q (slot rows >> 2); centre of the q statistics
  qs[lane] = 1.0f;
  qs[lane + 32] = 1.0f;
#pragma unroll
  for (int t = 0; t < 8; ++t) p_out[t] = 1.0f;
  __syncwarp();
  int next_red = warp;   // next iteration this warp is the reducer of (j % W == warp)
  int sk = 0;            // k % kRing
  for (int k = 1; k <= n_iter; ++k) {
    sk = sk == kRing - 1 ? 0 : sk + 1;
    if (RECORD && k - 1 == next_red) {
      reduce_iteration(k - 1);
      next_red += W;
    }
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
    load_span(qs + sp, v);
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
    float* pk = ps + sk * 256;
#pragma unroll
    for (int dr = 0; dr < 2; ++dr)
      *reinterpret_cast<float4*>(pk + off_s[dr]) = make_float4(p[4 * dr], p[4 * dr + 1], p[4 * dr + 2], p[4 * dr + 3]);
    if (RECORD) {
      sh.E[sk][warp][lane] = fmaf(-2.0f, fmaf(S1, pg, psd), fmaf(64.0f, gg, fmaf(V, pp, A)));
      named_arrive(1 + sk, nbar);
    }
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
    auto q_seg = [&](auto cc_c) {   // row (dr, cc) has rho % 4 = cc: it meets p[64 cc + j]
      constexpr int cc = decltype(cc_c)::value;
      const float Pseg = __shfl_sync(kFull, psum, 8 * cc);
      load_span(pk + 64 * cc + sp, v);
      ua[cc] = fmaf(f[cc], Pseg, row_dot<cc>(D[cc], v));
      ub[cc] = fmaf(f[4 + cc], Pseg, row_dot<4 + cc>(D[4 + cc], v));
    };
    q_seg(std::integral_constant<int, 0>{}); q_seg(std::integral_constant<int, 1>{});
    q_seg(std::integral_constant<int, 2>{}); q_seg(std::integral_constant<int, 3>{});
    m = Q * (1.0f / 64.0f);
    qa = ((ua[0] + ua[1]) + (ua[2] + ua[3])) * invB;
    qb = ((ub[0] + ub[1]) + (ub[2] + ub[3])) * invB;
    qs[off_s[0] >> 2] = qa;   // q index of row rho is rho >> 2
    qs[off_s[1] >> 2] = qb;
    if (RECORD && k >= kLag) keep_if_best(k - kLag, sk == kRing - 1 ? 0 : sk + 1);   // (k - 2) % 3 == (k + 1) % 3
    __syncwarp();
  }
  if (RECORD) {
    // drain: the last iteration's reducer, and the verdicts not yet read (iterations n_iter-kLag+1 .. n_iter; the
    // loop read those up to n_iter-1-kLag)
    if (n_iter == next_red) reduce_iteration(n_iter);
    for (int j = max(n_iter - kLag, 0); j <= n_iter; ++j) keep_if_best(j, j % kRing);
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
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float p = pv[t];
    const float x = logf(p) * (1.0f / 65536.0f);
    float pw = 1.0f + fmaf(fmaf(x, 1.0f / 6.0f, 0.5f) * x, x, x);
    if (!(p > 0.0f) || !(fabsf(x) < 3e-3f)) pw = (float)pow((double)p, 1.0 / 65536.0);   // zeros, negatives, NaN, huge ratios
    prod *= pw;
  }
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
      if (pg < ratio)
        for (int bc = 0; bc < ratio; ++bc) *reinterpret_cast<float4*>(mp + (16 * pg + (row >> 4)) * side + 16 * bc + (row & 15)) = o;
    }
  }
}

__device__ __forceinline__ void load_page(PageRegs& M, const float* __restrict__ compact, int lane) {
  const int rh = lane >> 2, kq = lane & 3;
  const int row_base = 32 * rh + 4 * kq;
  const int off_s[2] = {row_base + 16 * (rh & 1), row_base + 16 * ((rh & 1) ^ 1)};
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
    }
}

__device__ __forceinline__ bool unit_is_compact(const float* compact) {
  const float4 fl = *reinterpret_cast<const float4*>(compact + kCompactFloats);
  return fl.x == 1.0f && fl.y == 1.0f && fl.z == 1.0f && fl.w == 1.0f;
}

// grid = (group, page) items of every page scale; block = 32 x min(group, 16).
__global__ void __launch_bounds__(32 * kGroupWarps, 1) als_pages_kernel(const __grid_constant__ PagesParams P) {
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
  const int g = item / sc.pages, pg = item - g * sc.pages;
  const int group = P.group, limit = sc.limit;
  const int W0 = blockDim.x >> 5;                        // warps per round
  const int rounds = (group + W0 - 1) / W0;
  const int64_t stride = als_ws_stride(256, limit);
  const double inv_cnt = 1.0 / ((double)group * (double)(256 * 64));
  float* record_out = sc.record_out ? sc.record_out + ((int64_t)g * sc.pages + pg) * (limit + 1) : nullptr;
  auto unit_of = [&](int r) { return ((int64_t)g * group + r * W0 + warp) * sc.pages + pg; };

  // every unit of the item must have the pair-build structure; otherwise the dense kernel takes the whole item
  if (threadIdx.x == 0) sh.all_compact = 1;
  if (threadIdx.x < kRing) sh.flag[threadIdx.x] = 0xffffffffu;
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
    pages_iterate<0>(M, sh, ps, qs, best, lane, warp, W0, limit, inv_cnt, record_out, pv);
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
      pages_iterate<1>(M, sh, ps, qs, best, lane, warp, Wr, limit, inv_cnt, nullptr, pv);
    }
    __syncthreads();
    if (threadIdx.x < kRing) sh.flag[threadIdx.x] = 0xffffffffu;
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
    pages_iterate<2>(M, sh, ps, qs, best, lane, warp, 1, kstar, inv_cnt, nullptr, pv);
    pages_emit(sc, unit_of(r), pv, lane);
  }
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
    if (h.rows != 256 || (h.flags & RDM_ALS_DENSE_ONLY)) continue;
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
    const int64_t units = n_images * h.pages;
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
    n_items += n_groups * h.pages;
  }
  if (n_items == 0) return 0;
  RDM_REQUIRE(n_raw + n_map < (1ll << 28), "rdm_als_fused: too many work units");
  if (n_raw && sparsify) {
    als_sparsify_raw_kernel<<<(unsigned)(4 * n_raw), 256, 0, stream>>>(raw);
    int rc = launch_status("als_sparsify_raw_kernel");
    if (rc) return rc;
  }
  if (n_map && sparsify) {
    als_sparsify_map_kernel<<<(unsigned)n_map, 256, 0, stream>>>(map);
    int rc = launch_status("als_sparsify_map_kernel");
    if (rc) return rc;
  }
  if (!iterate) return 0;
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

// Host copy of the sparsify kernel's geometry tables (built by the same constexpr function), so that the CPU
// test suite can check them against the oracle's window mask without a GPU.
extern "C" int rdm_sparsify_geometry(uint64_t* lane_slots, uint16_t* item, uint8_t* compact) {
  RDM_REQUIRE(lane_slots && item && compact, "rdm_sparsify_geometry: null pointer");
  static constexpr rdm::SparsifyTables t = rdm::make_sparsify_tables();
  for (int h = 0; h < 2; ++h)
    for (int r0 = 0; r0 < 6; ++r0)
      for (int l = 0; l < 32; ++l) lane_slots[(h * 6 + r0) * 32 + l] = t.lane_slots[h][r0][l];
  for (int h = 0; h < 2; ++h)
    for (int i = 0; i < 96; ++i) item[h * 96 + i] = t.item[h][i];
  for (int c = 0; c < 16; ++c)
    for (int e = 0; e < 16; ++e) compact[c * 16 + e] = t.compact[c][e];
  return 0;
}
