// Stage 2+3 fused: Lloyd quantisation + rank-1 ALS + batch-wide arg-min + geometric
// normalisation + page re-tiling (RN:286-311, RN:358-396, CP:38-85, CP:95-155, CP:175-193,
// CP:218-238, CP:244-255), optionally with the pair build in front (RN:244-280) so the pair
// matrix never exists in HBM.
//
// Work unit = one pair matrix = one (image, page): rows x 64 with rows = 256 (16x16 page against
// its 8x8 parent, 100 iterations) or rows = 64 (the 8x8 map against itself, 30 iterations).  A
// unit is owned by `rows` threads and the whole matrix lives in REGISTERS for all iterations:
// every thread holds a 4-row x 16-column tile (64 values as 32 FFMA2 register pairs).  The four
// lanes cb = 0..3 of a lane group share the same four rows and split the 64 columns.
//   p-update  p_i = (sum_c R[i][c] q_c) / (|q|^2 + lambda)              CP:186-192
//   q-update  q_i = (sum_m Rflat[i*rows + m] p_m) / (|p|^2 + lambda)    CP:64 / CP:133: the
//             reference passes R.view(B,W,H) - a reshape, not a transpose - so "row i" of the
//             second operand is rows G*i .. G*i+G-1 of R laid end to end (G = rows/64), i.e.
//             q_i = sum_{r'<G} sum_c R[G i + r'][c] p[64 r' + c].
// A warp's 32 rows share r' = row % G, so both GEMVs read a 16-float operand slice per lane
// (4 LDS.128, four distinct addresses per warp), do 64 FMAs (32 FFMA2), and finish with a 2-level
// reduce-scatter over the lane group (3 SHFL) that leaves each lane with the full sum of ONE
// row.  The operand norms |q|^2 and |p segment|^2 come from the same loaded slice (8 FFMA2 + 2
// SHFL levels).  Measured on B200 the earlier row-per-thread layout (16 broadcast LDS.128 per
// GEMV) was bound by shared-memory operand delivery, not by FMA issue (profiles/README.md).
// Per iteration: 2 named barriers.
//
// rmse record (CP:53-61, CP:121-130): see the comment above als_iterate.  Per-unit SSE goes to
// the workspace; the arg-min is over the mean of the whole reference batch ("group",
// CP:172-173), so phase 1 (second launch) sums the group's records, picks the first minimum
// (CP:74, CP:143) and emits p_k*.  Phase 0 streams every iterate p_1..p_limit to the workspace (one
// 4-byte store per thread per iteration, ~100 KB per unit that mostly lives in L2): on noise-like
// inputs k* is 0 or 1, but on the smooth maps a real decoder produces the record
// plateaus and k* lands anywhere up to `limit` (tests/golden/README.md), so no iterate can be
// dropped.  No kernel waits on another: two plain launches.
//
// Bound: per unit 2*limit*rows*64 FMA against rows*64*(4..8) input bytes = 25..100 FMA/byte:
// FP32-issue / dependency-latency bound, not HBM bound (DESIGN.md "K3").
#include "rdm_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace rdm {

constexpr int kCols = 64;
constexpr int kMaxScales = 8;
constexpr float kLambda = 0.05f;   // CP:175 regularization_term
constexpr int kAlsThreads = 256;
constexpr int kTileFloats = 256 * 64;   // 64 KB matrix staging tile (dynamic shared memory)

struct AlsScaleDev {
  const void* src;
  const double* thr;
  const double* lvl;
  uint8_t* bins;
  float* values;
  float* ws;          // per unit: [limit+1] SSE record, then [limit][rows] iterates p_1..p_limit
  float* pages_out;
  float* map_out;
  float* record_out;
  int32_t* kstar_out;
  int32_t kind, rows, pages, side, limit;
  int32_t cta_begin;   // first blockIdx.x of this scale
  int32_t cta_count;   // working CTAs of this scale (64-row: groups x cluster; pages: CTAs that walk the (group, page) items with this stride)
  int32_t dense_only;  // page scale: every item is iterated here (RDM_ALS_DENSE_ONLY / TRUE_TRANSPOSE, or a source kind the compact kernel does not take)
  int32_t flags;       // RDM_ALS_* bits
};

struct AlsParams {
  AlsScaleDev s[kMaxScales];
  int64_t n_images;
  int32_t n_scales;
  int32_t group;
};

struct AlsSmem {
  __align__(16) float p_s[256];       // p by row index (64-row units: 4 x 64)
  __align__(16) float q_w[2][8][64];  // per-warp copies of q, double buffered (q_{k-1} stays readable for the record)
  __align__(16) float qpart[4][64];   // q partial sums (256-row: by r'; 64-row: by unit)
  __align__(16) float part_pp[8];     // |p segment|^2 by r'
  double thr_d[kThrPad];
  double inv_d[64];                   // 1/parent (pair build fused, 256-row units)
  float thr_f[kThrPad];
  float lvl_f[kLvl + 3];
  float inv_f[4][64];                 // 1/d (pair build fused, 64-row units)
  float qT[8][64];                    // RDM_ALS_TRUE_TRANSPOSE: per-warp column sums of R^T p
  float rm[256];                      // group rmse record (64-row: one row of 64 per unit team)
  float recu[128];                    // SSE record of the unit just iterated (page fallback)
  double recg[128];                   // group SSE record accumulated over the units (page fallback)
  int cell_s[kThr];
  int sorted;
  LloydLut lut;                       // bin lookup table in the dtype this CTA compares in
};

// Bin look-up context held in registers for the duration of a load (USE_LUT is hoisted out of
// the element loops so that the look-ups are straight-line code and interleave).
template <typename T, bool USE_LUT>
struct BinCtx {
  const T* tab;
  const uint8_t* lut;
  int base, ncell, sorted;
  T thr0;
  __device__ __forceinline__ BinCtx(const AlsSmem& sm, const T* t) : tab(t), lut(sm.lut.lut), base(sm.lut.base), ncell(sm.lut.ncell), sorted(sm.sorted), thr0(t[0]) {}
  __device__ __forceinline__ int operator()(T x) const {
    if constexpr (USE_LUT) return lloyd_bin_lut<T>(x, tab, lut, base, ncell, thr0);
    else return lloyd_bin<T>(x, tab, sorted);
  }
};

__device__ __forceinline__ void unit_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- shared-memory accesses of the iteration loop by 32-bit shared address.  With generic
// pointers nvcc 12.9 re-derived the shared window base (S2UR SR_CgaCtaId + ULEA, ~75 cycles of
// scoreboard wait) inside every iteration (profiles/r1_als_kernel_stalls.txt).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_v4f32(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ ulonglong2 lds_v2u64(uint32_t a) {
  ulonglong2 v;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a) : "memory");
  return v;
}

// ---- packed f32x2 arithmetic (sm_100a FFMA2)
using u64 = unsigned long long;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 as_u64(float2 v) { return *reinterpret_cast<u64*>(&v); }
__device__ __forceinline__ float2 as_f2(u64 v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float hsum2(u64 a, u64 b) {
  const float2 fa = as_f2(a), fb = as_f2(b);
  return (fa.x + fa.y) + (fb.x + fb.y);
}

// 1/x for x in the normal range (|q|^2 + lambda, |p|^2 + lambda): MUFU.RCP plus one Newton step, the
// same sequence the compiler emits for 1.0f / x minus its range-check branch and slow path.
__device__ __forceinline__ float rcp_newton(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(fmaf(-x, r, 1.0f), r, r);
}

// Thread <-> tile mapping of a unit.  lt = thread index inside the unit (0 .. 64G-1).
// The four lanes cb = 0..3 of a lane group share the rows G (ib + 0..3) + rp and hold the columns
// 16 cb .. 16 cb + 15 of each.  A lane keeps them in the lane-specific slot order
// slot t <-> row(t) = G (ib + (t ^ cb)) + rp, so that slot 0 is the row the lane owns after a
// reduce-scatter and the "keep"/"send" halves of every exchange step are fixed register slots
// (no selects): see reduce_scatter4.
template <int G>
struct TileMap {
  int lane, lw, cb, rp, ib, row_own;
  __device__ __forceinline__ explicit TileMap(int lt) {
    lane = lt & 31;
    lw = lt >> 5;
    cb = lane & 3;
    rp = lw % G;
    ib = (lw / G) * 32 + 4 * (lane >> 2);
    row_own = G * (ib + cb) + rp;
  }
  __device__ __forceinline__ int row(int t) const { return G * (ib + (t ^ cb)) + rp; }
};

// Sum the per-slot partials over the 4 lanes of a lane group; every lane receives the total of its
// slot 0, i.e. of the row it owns (3 shuffles).  With slot t <-> row (t ^ cb): the partner cb^2 holds our
// rows of slots 0,1 in its slots 2,3, and the partner cb^1 holds our slot-0 row in its slot 1.
template <typename T>
__device__ __forceinline__ T reduce_scatter4(T v0, T v1, T v2, T v3, int /*cb*/) {
  const T r0 = v0 + __shfl_xor_sync(0xffffffffu, v2, 2);
  const T r1 = v1 + __shfl_xor_sync(0xffffffffu, v3, 2);
  return r0 + __shfl_xor_sync(0xffffffffu, r1, 1);
}
template <typename T>
__device__ __forceinline__ T group_sum4(T v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// Partial GEMV of the tile with the lane's 16-float operand slice `op` (4 LDS.128) and the
// squared norm of the slice: own = full dot of the row this lane owns, nrm = |whole operand|^2.
// Measured on B200: packed FFMA2 (fma.rn.f32x2) issues only on the fmaheavy sub-pipe, scalar FFMA on both
// halves; for these GEMVs the scalar form is 12 % faster per launch (78.8 vs 89.3 us) despite twice the
// FMA instructions.  -DRDM_FFMA2 builds the packed variant for comparison.
#ifndef RDM_FFMA2
__device__ __forceinline__ void tile_dot(const float2 (&R)[4][8], uint32_t op, int cb, float& own, float& nrm) {
  float4 x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = lds_v4f32(op + 16 * k);
  float a[4][2], n0 = 0.f, n1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j][0] = a[j][1] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j][0] = fmaf(R[j][2 * k].x, x[k].x, a[j][0]);
      a[j][1] = fmaf(R[j][2 * k].y, x[k].y, a[j][1]);
      a[j][0] = fmaf(R[j][2 * k + 1].x, x[k].z, a[j][0]);
      a[j][1] = fmaf(R[j][2 * k + 1].y, x[k].w, a[j][1]);
    }
    n0 = fmaf(x[k].x, x[k].x, fmaf(x[k].z, x[k].z, n0));
    n1 = fmaf(x[k].y, x[k].y, fmaf(x[k].w, x[k].w, n1));
  }
  own = reduce_scatter4(a[0][0] + a[0][1], a[1][0] + a[1][1], a[2][0] + a[2][1], a[3][0] + a[3][1], cb);
  nrm = group_sum4(n0 + n1);
}
#else
__device__ __forceinline__ void tile_dot(const float2 (&R)[4][8], uint32_t op, int cb, float& own, float& nrm) {
  ulonglong2 x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = lds_v2u64(op + 16 * k);
  u64 a[4][2];
  u64 n0 = 0, n1 = 0;   // all-zero bits = (0.f, 0.f)
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j][0] = a[j][1] = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j][0] = ffma2(as_u64(R[j][2 * k]), x[k].x, a[j][0]);
      a[j][1] = ffma2(as_u64(R[j][2 * k + 1]), x[k].y, a[j][1]);
    }
    n0 = ffma2(x[k].x, x[k].x, n0);
    n1 = ffma2(x[k].y, x[k].y, n1);
  }
  own = reduce_scatter4(hsum2(a[0][0], a[0][1]), hsum2(a[1][0], a[1][1]), hsum2(a[2][0], a[2][1]), hsum2(a[3][0], a[3][1]), cb);
  nrm = group_sum4(hsum2(n0, n1));
}
#endif

// Direct residual of the tile rows: sum_c (p_j q_c - R[j][c])^2 in f32 (what CP:172-173
// evaluates); lane cb receives the row it owns.
__device__ __forceinline__ float tile_sse(const float2 (&R)[4][8], uint32_t qop, const float (&pj)[4], int cb) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 x = lds_v4f32(qop + 16 * k);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // fl(fl(p q_c) - R): the outer product is rounded before the subtraction in the reference
      // (matmul, then sub); a fused multiply-subtract would differ when the fit is nearly exact
      const float t0 = __fsub_rn(__fmul_rn(pj[j], x.x), R[j][2 * k].x), t1 = __fsub_rn(__fmul_rn(pj[j], x.y), R[j][2 * k].y);
      const float t2 = __fsub_rn(__fmul_rn(pj[j], x.z), R[j][2 * k + 1].x), t3 = __fsub_rn(__fmul_rn(pj[j], x.w), R[j][2 * k + 1].y);
      acc[j] = fmaf(t3, t3, fmaf(t2, t2, fmaf(t1, t1, fmaf(t0, t0, acc[j]))));
    }
  }
  return reduce_scatter4(acc[0], acc[1], acc[2], acc[3], cb);
}

// The same residual with fused arithmetic: t = R - p q (one rounding), acc += t^2.  Used for
// the record of iterations k >= 1: as accurate as the reference's own f32 evaluation (~1e-7 after
// averaging), which matters on smooth maps where the record declines by ~1 f32 ulp per iteration and
// the arg-min follows that decline.
__device__ __forceinline__ float tile_sse_fused(const float2 (&R)[4][8], uint32_t qop, const float (&pj)[4], int cb) {
  float4 x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = lds_v4f32(qop + 16 * k);
  float acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float t0 = fmaf(-pj[j], x[k].x, R[j][2 * k].x), t1 = fmaf(-pj[j], x[k].y, R[j][2 * k].y);
      const float t2 = fmaf(-pj[j], x[k].z, R[j][2 * k + 1].x), t3 = fmaf(-pj[j], x[k].w, R[j][2 * k + 1].y);
      a0 = fmaf(t2, t2, fmaf(t0, t0, a0));
      a1 = fmaf(t3, t3, fmaf(t1, t1, a1));
    }
    acc[j] = a0 + a1;
  }
  return reduce_scatter4(acc[0], acc[1], acc[2], acc[3], cb);
}

// ---------------------------------------------------------------------------------------------
// Load (and, for RAW_* / MAP kinds, build + quantise) the unit's matrix into register tiles.
template <int G, bool USE_LUT>
__device__ __forceinline__ void load_unit_impl(float2 (&R)[4][8], const AlsScaleDev& sc, AlsSmem& sm, float* tile,
                                               int64_t unit_idx, int unit, int lt, bool emit) {
  constexpr int NT = 64 * G;
  constexpr int ROWS = 64 * G;
  const TileMap<G> m(lt);
  const int bar_id = (G == 4) ? 0 : 1 + unit;
  const BinCtx<double, USE_LUT> bin_d(sm, sm.thr_d);
  const BinCtx<float, USE_LUT> bin_f(sm, sm.thr_f);
  uint8_t* bins = emit ? sc.bins : nullptr;
  float* values = emit ? sc.values : nullptr;
  const int64_t mat_off = unit_idx * (int64_t)(ROWS * kCols);

  if (sc.kind == RDM_SRC_MAP_F32) {
    // ---- pair build fused: nothing but the decoder map is read from HBM
    const float* map;
    int pi = 0, pj = 0, side = 8;
    if constexpr (G == 4) {
      side = sc.side;
      const int ratio = side >> 4;
      const int64_t img = unit_idx / sc.pages;
      const int pg = (int)(unit_idx - img * sc.pages);
      pi = pg / ratio;
      pj = pg - pi * ratio;
      map = reinterpret_cast<const float*>(sc.src) + img * (int64_t)side * side;
      if (lt < 64) {
        int y = 8 * pi + (lt >> 3), x = 8 * pj + (lt & 7);
        double v = bicubic_half_at([&](int r, int c) { return (double)map[r * side + c]; }, y, x, side);
        sm.inv_d[lt] = 1.0 / v;   // torch.pow(area,-1): IEEE reciprocal
      }
    } else {
      map = reinterpret_cast<const float*>(sc.src) + unit_idx * 64;
      sm.inv_f[unit][lt] = __frcp_rn(map[lt]);   // RN:248
    }
    unit_barrier(bar_id, NT);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = m.row(j);
      uint32_t pk[4] = {0u, 0u, 0u, 0u};
      if constexpr (G == 4) {
        const double d = (double)map[(16 * pi + (row >> 4)) * side + 16 * pj + (row & 15)];
        const int r0 = min((row >> 4) >> 1, 5), c0 = min((row & 15) >> 1, 5);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int c = 16 * m.cb + e;
          const bool win = (unsigned)((c >> 3) - r0) < 3u && (unsigned)((c & 7) - c0) < 3u;
          const int b = bin_d(win ? __dmul_rn(d, sm.inv_d[c]) : d);   // 55 of 64 columns hold d itself
          const float v = sm.lvl_f[b];
          if (e & 1) R[j][e >> 1].y = v; else R[j][e >> 1].x = v;
          pk[e >> 2] |= (uint32_t)b << (8 * (e & 3));
        }
      } else {
        const float dv = map[row];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int c = 16 * m.cb + e;
          const int b = bin_f(__fmul_rn(dv, sm.inv_f[unit][c]));   // RN:252
          const float v = sm.lvl_f[b];
          if (e & 1) R[j][e >> 1].y = v; else R[j][e >> 1].x = v;
          pk[e >> 2] |= (uint32_t)b << (8 * (e & 3));
        }
      }
      if (bins) *reinterpret_cast<uint4*>(bins + mat_off + row * 64 + 16 * m.cb) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (values) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<float4*>(values + mat_off + row * 64 + 16 * m.cb + 4 * k) =
              make_float4(R[j][2 * k].x, R[j][2 * k].y, R[j][2 * k + 1].x, R[j][2 * k + 1].y);
      }
    }
    return;
  }

  // ---- matrix given in HBM: coalesced 128-bit loads, quantise on the fly, transpose through an
  // XOR-swizzled shared tile into the register tiles.
#ifdef RDM_TIMING
  const long long tq0 = clock64();
#endif
  if (sc.kind == RDM_SRC_RAW_F64 || sc.kind == RDM_SRC_VAL_F64) {
    const double* src = reinterpret_cast<const double*>(sc.src) + mat_off;
    const bool quant = sc.kind == RDM_SRC_RAW_F64;
    constexpr int ITERS = ROWS * 32 / NT;   // element pairs per thread
    // 16 independent 128-bit loads in flight per thread (the matrix tile is not live yet, so the
    // registers are free): two round trips to HBM per page instead of four
#pragma unroll 1
    for (int k0 = 0; k0 < ITERS; k0 += 16) {
     double2 xs[16];
#pragma unroll
     for (int kk = 0; kk < 16; ++kk) xs[kk] = ldg_stream_f64x2(src + 2 * (lt + NT * (k0 + kk)));
#pragma unroll
     for (int kk = 0; kk < 16; ++kk) {
      const int k = k0 + kk;
      const int e2 = lt + NT * k;
      const int r = e2 >> 5, cp = e2 & 31;
      double2 x = xs[kk];
      float v0, v1;
      if (quant) {
        int b0 = bin_d(x.x), b1 = bin_d(x.y);
        v0 = sm.lvl_f[b0];
        v1 = sm.lvl_f[b1];
        if (bins) *reinterpret_cast<uint16_t*>(bins + mat_off + 2 * e2) = (uint16_t)(b0 | (b1 << 8));
      } else {
        v0 = (float)x.x;   // `.float()` CP:40 / CP:106
        v1 = (float)x.y;
      }
      if (values) *reinterpret_cast<float2*>(values + mat_off + 2 * e2) = make_float2(v0, v1);
      const int swz = (r / G) & 7;
      *reinterpret_cast<float2*>(tile + r * 64 + (((cp >> 1) ^ swz) << 2) + 2 * (cp & 1)) = make_float2(v0, v1);
     }
    }
  } else {
    const float* src = reinterpret_cast<const float*>(sc.src) + mat_off;
    const bool quant = sc.kind == RDM_SRC_RAW_F32;
    constexpr int ITERS = ROWS * 16 / NT;   // float4 chunks per thread
#pragma unroll 8
    for (int k = 0; k < ITERS; ++k) {
      const int e4 = lt + NT * k;
      const int r = e4 >> 4, c4 = e4 & 15;
      float4 x = ldg_stream_f32x4(src + 4 * e4);
      if (quant) {
        int b0 = bin_f(x.x), b1 = bin_f(x.y), b2 = bin_f(x.z), b3 = bin_f(x.w);
        x = make_float4(sm.lvl_f[b0], sm.lvl_f[b1], sm.lvl_f[b2], sm.lvl_f[b3]);
        if (bins)
          *reinterpret_cast<uint32_t*>(bins + mat_off + 4 * e4) = (uint32_t)(b0 | (b1 << 8) | (b2 << 16) | (b3 << 24));
      }
      if (values) *reinterpret_cast<float4*>(values + mat_off + 4 * e4) = x;
      const int swz = (r / G) & 7;
      *reinterpret_cast<float4*>(tile + r * 64 + ((c4 ^ swz) << 2)) = x;
    }
  }
#ifdef RDM_TIMING
  const long long tq1 = clock64();
#endif
  unit_barrier(bar_id, NT);
#ifdef RDM_TIMING
  if (lt == 0 && unit == 0 && blockIdx.x % 37 == 0) printf("  block %d kind %d: global load+quantise %lld cycles, barrier wait %lld\n", blockIdx.x, sc.kind, tq1 - tq0, clock64() - tq1);
  if (lt == 255 && blockIdx.x % 37 == 0) printf("  block %d (thread 255): global load+quantise %lld cycles\n", blockIdx.x, tq1 - tq0);
#endif
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = m.row(j);
    const int swz = (row / G) & 7;
    const float4* t4 = reinterpret_cast<const float4*>(tile) + row * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 x = t4[(4 * m.cb + k) ^ swz];
      R[j][2 * k] = make_float2(x.x, x.y);
      R[j][2 * k + 1] = make_float2(x.z, x.w);
    }
  }
}

template <int G>
__device__ __forceinline__ void load_unit(float2 (&R)[4][8], const AlsScaleDev& sc, AlsSmem& sm, float* tile,
                                          int64_t unit_idx, int unit, int lt, bool emit) {
  if (sm.lut.ncell)   // CTA-uniform
    load_unit_impl<G, true>(R, sc, sm, tile, unit_idx, unit, lt, emit);
  else
    load_unit_impl<G, false>(R, sc, sm, tile, unit_idx, unit, lt, emit);
}

// Record (CP:53-61, CP:121-130).  Two evaluations of row i's residual sum_j (p_i q_j - R_ij)^2:
//  * direct (tile_sse / tile_sse_fused): the third pass over the matrix the reference makes; as
//    accurate as the reference's own f32 evaluation (~1e-7 after averaging);
//  * algebraic: |R_i|^2 + p_i (p_i |q|^2 - 2 s_i), s_i = R_i . q being a by-product of the p-update;
//    O(1) per row, absolute error ~1e-7 |R_i|^2 (s_i and |q|^2 are f32), i.e. ~1e-6 relative once the
//    residual is a few per cent of the energy.
// On smooth maps the record plateaus and declines by about one f32 ulp per iteration; the arg-min
// follows that decline, so only the direct form is good enough there.  The choice is made per unit
// and per iteration from the residual estimate |R|^2 - |p_k|^2 (|q_{k-1}|^2 + 2 lambda) (scalars every
// thread holds after barrier B): below kDirectFrac of |R|^2 -> direct.  Iteration 0 is always direct,
// with the reference's unfused rounding so that constant maps give exactly 0.
// Each thread parks the value of the row it owns for iteration k in shared memory E[k][thread]; the
// unit reduces all iterations once, after the loop (two rows per slot: the scratch fits in the dead
// staging tile).
constexpr float kDirectFrac = 0.015f;
// n_iter alternating iterations; returns the iterate p_{n_iter} of the row this thread owns (1 for n_iter = 0).
// RECORD: also writes the SSE of iterations 0..n_iter to rec[] (shared or global); E: (n_iter+1) x (NT/2+1)
// floats of shared scratch.  Without RECORD the same iterates are computed (bit-identical: the record never
// feeds back into p or q) and nothing else: this is how the selected iterate p_k* is re-materialised once the
// group-wide arg-min is known, instead of keeping every iterate.
// tt (RDM_ALS_TRUE_TRANSPOSE, off by default): the q-update uses R^T - q_c = sum_i R[i][c] p_i / (|p|^2 + lambda) - instead of
// the reference's R.view(B,W,H) reshape (CP:64, CP:133).
template <int G, bool RECORD>
__device__ __forceinline__ float als_iterate(const float2 (&R)[4][8], AlsSmem& sm, float* __restrict__ E, int unit, int lt, int n_iter,
                                             float* __restrict__ rec, bool tt) {
  constexpr int NW = 2 * G;
  constexpr int NT = 64 * G;
  constexpr int EH = NT / 2;                     // two rows (lanes l, l^4) share one slot
  constexpr int ES = EH + 1;                     // odd stride: the final column sums are conflict-free
  const TileMap<G> m(lt);
  const int gw = unit * NW + m.lw;
  const int bar_id = (G == 4) ? 0 : 1 + unit;
  const int qp_row = (G == 4) ? m.rp : unit;
  uint32_t qw = smem_u32(sm.q_w[0][gw]);                       // this warp's current copy of q
  const uint32_t qw_flip = smem_u32(sm.q_w[0][gw]) ^ smem_u32(sm.q_w[1][gw]);
  const uint32_t ps = smem_u32(sm.p_s + unit * 64);            // p of this unit
  uint32_t qop = qw + 64 * m.cb;                               // operand slices (16 floats)
  const uint32_t pop = ps + 256 * m.rp + 64 * m.cb;
  const uint32_t qpart = smem_u32(sm.qpart);
  const uint32_t ppart = smem_u32(sm.part_pp);
  const int eslot = ((lt >> 3) << 2) | (lt & 3);
  const uint32_t Es = RECORD ? smem_u32(E) + 4 * eslot : 0;

  // q_0 = 1
  sts_f32(qw + 4 * m.lane, 1.0f);
  sts_f32(qw + 4 * m.lane + 128, 1.0f);
  __syncwarp();
  float s, Q;
  tile_dot(R, qop, m.cb, s, Q);                  // s = row sum, Q = 64
  float invA = rcp_newton(Q + kLambda);          // torch.inverse of the 1x1 matrix |q|^2 + lambda
  double r2 = 0.0;
  float r2h = 0.f, r2l = 0.f, r2tot = 0.f;
  if (RECORD) {
    double t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t[j] = 0.0;
#pragma unroll
      for (int e = 0; e < 8; ++e) t[j] = fma((double)R[j][e].y, (double)R[j][e].y, fma((double)R[j][e].x, (double)R[j][e].x, t[j]));
    }
    r2 = reduce_scatter4(t[0], t[1], t[2], t[3], m.cb);
    r2h = (float)r2;                             // |R_i|^2 as an unevaluated f32 pair (exact to ~2^-48)
    r2l = (float)(r2 - (double)r2h);
    {                                            // |R|^2 of the whole unit (decision threshold only)
      const float w = warp_sum(r2h);
      if (m.lane == 0) sts_f32(smem_u32(sm.rm) + 4 * gw, w);
      unit_barrier(bar_id, NT);
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) r2tot += lds_f32(smem_u32(sm.rm) + 4 * (unit * NW + w2));
    }
    const float ones[4] = {1.f, 1.f, 1.f, 1.f};
    const float e0 = tile_sse(R, qop, ones, m.cb);   // k = 0: p = q = 1 (CP:55, CP:123)
    unit_barrier(bar_id, NT);                    // the staging tile (aliased by E) is dead from here on
    const float e0p = e0 + __shfl_xor_sync(0xffffffffu, e0, 4);
    if (!(m.lane & 4)) sts_f32(Es, e0p);
  }
  float p = 1.0f;
  for (int k = 1; k <= n_iter; ++k) {
    p = s * invA;                                // (R q) @ inverse(A)
    if (!RECORD && k == n_iter) break;           // replay: p_k* is all that is wanted
    sts_f32(ps + 4 * m.row_own, p);
    unit_barrier(bar_id, NT);                    // A: p visible
    float ef = 0.f;
    if (RECORD) {   // algebraic residual of this row (f32 pair arithmetic: error-free product and sum); straight-line
                    // code next to the operand loads so that it fills their latency
      const float t = fmaf(p, Q, -2.0f * s);
      const float ph = p * t, pl = fmaf(p, t, -ph);
      const float sm1 = r2h + ph, bb = sm1 - r2h;
      const float er = (r2h - (sm1 - bb)) + (ph - bb);
      ef = fmaxf(sm1 + (er + (r2l + pl)), 0.f);
    }
    float npp;
    if (tt) {
      // column sums of this thread's 4 x 16 tile against its four p entries, reduced over the 8 lanes of the warp that
      // hold the same columns, then over the unit's warps through shared memory (fixed order: deterministic)
      float pj[4], c16[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) pj[j] = lds_f32(ps + 4 * m.row(j));
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        c16[2 * e] = fmaf(R[3][e].x, pj[3], fmaf(R[2][e].x, pj[2], fmaf(R[1][e].x, pj[1], R[0][e].x * pj[0])));
        c16[2 * e + 1] = fmaf(R[3][e].y, pj[3], fmaf(R[2][e].y, pj[2], fmaf(R[1][e].y, pj[1], R[0][e].y * pj[0])));
      }
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1)
#pragma unroll
        for (int e = 0; e < 16; ++e) c16[e] += __shfl_xor_sync(0xffffffffu, c16[e], o);
      if ((m.lane >> 2) == 0)
#pragma unroll
        for (int e = 0; e < 16; ++e) sm.qT[gw][16 * m.cb + e] = c16[e];
      const float w = warp_sum(p * p);
      if (m.lane == 0) sts_f32(ppart + 4 * gw, w);
      unit_barrier(bar_id, NT);                  // B: column sums and |p|^2 parts visible
      npp = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) npp += lds_f32(ppart + 4 * (unit * NW + w2));
    } else {
      float u, pseg;
      tile_dot(R, pop, m.cb, u, pseg);           // this row's share of q_{ib+cb}; |p segment r'|^2
      sts_f32(qpart + 4 * (64 * qp_row + m.ib + m.cb), u);
      if constexpr (G == 4) {
        if (m.lw < 4 && m.lane == 0) sts_f32(ppart + 4 * m.rp, pseg);
      }
      unit_barrier(bar_id, NT);                  // B: q partials (and |p|^2 segments) visible
      npp = pseg;
      if constexpr (G == 4) {
        const float4 a = lds_v4f32(ppart);
        npp = (a.x + a.y) + (a.z + a.w);
      }
    }
    // the direct evaluation (rare on noise-like maps) needs p_k (ps) and q_{k-1} (this warp's old q copy)
    const uint32_t qop_k = qop;
    const bool direct = RECORD && (r2tot - npp * (Q + 2.0f * kLambda)) < kDirectFrac * r2tot;   // unit-uniform
    if (k < n_iter) {
      // every warp finalises q for itself (no further barrier), into its other q buffer
      float u0, u1;
      if (tt) {
        u0 = u1 = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2) {
          u0 += sm.qT[unit * NW + w2][m.lane];
          u1 += sm.qT[unit * NW + w2][m.lane + 32];
        }
      } else if constexpr (G == 4) {
        const uint32_t q0a = qpart + 4 * m.lane;
        const float a0 = lds_f32(q0a), a1 = lds_f32(q0a + 256), a2 = lds_f32(q0a + 512), a3 = lds_f32(q0a + 768);
        const float b0 = lds_f32(q0a + 128), b1 = lds_f32(q0a + 384), b2 = lds_f32(q0a + 640), b3 = lds_f32(q0a + 896);
        u0 = (a0 + a1) + (a2 + a3);
        u1 = (b0 + b1) + (b2 + b3);
      } else {
        u0 = lds_f32(qpart + 4 * (64 * unit + m.lane));
        u1 = lds_f32(qpart + 4 * (64 * unit + m.lane + 32));
      }
      const float invB = rcp_newton(npp + kLambda);
      qw ^= qw_flip;
      qop = qw + 64 * m.cb;
      sts_f32(qw + 4 * m.lane, u0 * invB);
      sts_f32(qw + 4 * m.lane + 128, u1 * invB);
      __syncwarp();
      tile_dot(R, qop, m.cb, s, Q);
      invA = rcp_newton(Q + kLambda);
    }
    if (RECORD) {
      if (direct) {   // reads only this warp's own data (its rows of p in ps, its old q copy): no barrier needed
        float pj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pj[j] = lds_f32(ps + 4 * m.row(j));
        ef = tile_sse_fused(R, qop_k, pj, m.cb);
      }
      ef += __shfl_xor_sync(0xffffffffu, ef, 4);
      if (!(m.lane & 4)) sts_f32(Es + 4 * k * ES, ef);
    }
  }
  if (RECORD) {
    unit_barrier(bar_id, NT);
    for (int k = lt; k <= n_iter; k += NT) {
      const float* col = E + k * ES;
      double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll 4
      for (int j = 0; j < EH; j += 4) {
        t0 += (double)col[j];
        t1 += (double)col[j + 1];
        t2 += (double)col[j + 2];
        t3 += (double)col[j + 3];
      }
      rec[k] = (float)((t0 + t1) + (t2 + t3));
    }
  }
  return p;
}

// ---------------------------------------------------------------------------------------------
// First minimum of a unit team's rmse record rm[0..limit] (CP:74, CP:143) as a parallel min over (value bits,
// index) keys: rmse values are >= 0, so their bit patterns order like the values and ties resolve to the
// smaller index (limit <= 127: threads 0..127 hold one key each, teams of 64 threads take two).
template <int G>
__device__ __forceinline__ int first_argmin(const float* rm, int limit, AlsSmem& sm, int unit, int lt) {
  constexpr int NT = 64 * G;
  const TileMap<G> m(lt);
  const int bar_id = (G == 4) ? 0 : 1 + unit;
  unsigned long long key = ~0ull;
  for (int k = lt; k <= limit; k += NT) {
    const unsigned long long kk = ((unsigned long long)__float_as_uint(rm[k]) << 32) | (unsigned)k;
    key = kk < key ? kk : key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other < key ? other : key;
  }
  unsigned long long* kscr = reinterpret_cast<unsigned long long*>(sm.qpart) + unit * 8;   // 8 keys per unit
  if (m.lane == 0) kscr[m.lw] = key;
  unit_barrier(bar_id, NT);
#pragma unroll
  for (int w = 0; w < 2 * G; ++w) {
    const unsigned long long other = kscr[w];
    key = other < key ? other : key;
  }
  unit_barrier(bar_id, NT);   // kscr (= qpart) is reused by the replay
  return (int)(key & 0xffffffffu);
}

// quick_gm(p, H) with H = rows: prod_i p_i^(1/H^2)  (CP:76, CP:146, CP:244-255), then p / gm into pages_out and the
// re-tiled map (CP:218-238).  The exponent is 2^-12 or 2^-16, so p^(1/H^2) = exp(x) with |x| = |ln p| / H^2 < 3e-3:
// 1 + x + x^2/2 + x^3/6 is exact to f32 rounding (x^4/24 < 4e-12), and p = 1 gives exactly 1.
template <int G>
__device__ __forceinline__ void emit_unit(const AlsScaleDev& sc, AlsSmem& sm, int64_t unit_idx, int unit, int lt, float p) {
  constexpr int NT = 64 * G;
  constexpr int ROWS = 64 * G;
  const TileMap<G> m(lt);
  const int row = m.row_own;
  const int bar_id = (G == 4) ? 0 : 1 + unit;
  const int64_t img = unit_idx / sc.pages;
  const int pg = (int)(unit_idx - img * sc.pages);
  const float pw = gm_factor(p, ROWS, (sc.flags & RDM_ALS_TRUE_GM) != 0);
  const float prod = warp_prod(pw);
  unit_barrier(bar_id, NT);
  float* scratch = sm.p_s + unit * 64;
  if (m.lane == 0) scratch[m.lw] = prod;
  unit_barrier(bar_id, NT);
  float gm = scratch[0];
#pragma unroll
  for (int w = 1; w < 2 * G; ++w) gm *= scratch[w];
  unit_barrier(bar_id, NT);   // scratch (= p_s) is reused by the next unit
  const float out = p / gm;
  if (sc.pages_out) sc.pages_out[unit_idx * ROWS + row] = out;
  if (sc.map_out) {
    if constexpr (G == 1) {
      sc.map_out[img * 64 + row] = out;
    } else {
      const int side = sc.side, ratio = side >> 4;
      float* mp = sc.map_out + img * (int64_t)side * side;
      if (sc.flags & RDM_ALS_CORRECT_TILING) {   // page (i, j) to block (i, j): what CP:218-238 evidently intended
        const int pi = pg / ratio, pj = pg - pi * ratio;
        mp[(16 * pi + (row >> 4)) * side + 16 * pj + (row & 15)] = out;
      } else if (pg < ratio) {   // CP:218-238 as written: block-row j of every block-column holds page j (< ratio)
        for (int bc = 0; bc < ratio; ++bc) mp[(16 * pg + (row >> 4)) * side + 16 * bc + (row & 15)] = out;
      }
    }
  }
}

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The dense kernel: one launch, thread-block clusters of `cluster` CTAs.
//  * 64-row units (the 8x8 maps, CP:38-85): the `group` images of a reference batch belong to ONE cluster, four
//    units per CTA and round.  Pass 1 iterates with the record and leaves each unit's SSE record in the
//    workspace; one cluster barrier later every team sums the group's records (CP:172-173), takes the first
//    minimum (CP:74) and replays its own k* iterations on the matrix it still holds in registers.
//  * 256-row units (pages): only the (group, page) items the compact kernel left alone - a matrix without the
//    pair-build structure, or a scale flagged RDM_ALS_DENSE_ONLY.  One CTA walks the units of an item one after
//    the other: pass 1 accumulates the group record, pass 2 reloads and replays k* iterations.
__global__ void __launch_bounds__(kAlsThreads, 2) als_kernel(const __grid_constant__ AlsParams P) {
  extern __shared__ __align__(16) float tile[];
  __shared__ AlsSmem sm;
  int si = 0;
#pragma unroll 1
  for (int k = 1; k < P.n_scales; ++k)
    if ((int)blockIdx.x >= P.s[k].cta_begin) si = k;
  const AlsScaleDev& sc = P.s[si];
  const int tid = threadIdx.x;
  const int local_cta = (int)blockIdx.x - sc.cta_begin;
  const int group = P.group;
  const bool pages = sc.rows == 256;
  const int64_t n_items = (P.n_images / group) * sc.pages;
  const int64_t stride256 = als_ws_stride(256, sc.limit);
  // does the compact kernel own this (group, page) item?  (all of its units carry the four band flags)
  auto item_is_compact = [&](int64_t item) {
    const int64_t g = item / sc.pages, pg = item - g * sc.pages;
    // RDM_ALS_SKIP_UNUSED_PAGES: pages CP:218-238 never copies into the map are nobody's work
    if ((sc.flags & RDM_ALS_SKIP_UNUSED_PAGES) && !(sc.flags & RDM_ALS_CORRECT_TILING) && pg >= (sc.side >> 4)) return true;
    if (sc.dense_only) return false;
    bool all = true;
    for (int b = 0; b < group; ++b) {
      const float4 fl = *reinterpret_cast<const float4*>(sc.ws + ((g * group + b) * sc.pages + pg) * stride256 + kCompactFloats);
      all = all && fl.x == 1.0f && fl.y == 1.0f && fl.z == 1.0f && fl.w == 1.0f;
    }
    return all;
  };
  if (pages) {   // nothing to do for this CTA?  leave before the codebook prologue (CTA-uniform)
    if (local_cta >= sc.cta_count) return;           // padding CTA of the last cluster
    bool any = false;
    for (int64_t it = local_cta; it < n_items; it += sc.cta_count) any |= !item_is_compact(it);
    if (!any) return;
  }
  if (sc.thr) {
    if (tid == 0) sm.sorted = 1;
    __syncthreads();
    if (tid < kThrPad) {
      double t = (tid < kThr) ? sc.thr[tid] : (double)NAN;
      sm.thr_d[tid] = t;
      sm.thr_f[tid] = (float)t;
    }
    if (tid < kLvl) sm.lvl_f[tid] = (float)sc.lvl[tid];
    __syncthreads();
    // the branch-free search needs non-decreasing thresholds in BOTH dtypes
    if (tid < kThr - 1 && (!(sm.thr_d[tid] <= sm.thr_d[tid + 1]) || !(sm.thr_f[tid] <= sm.thr_f[tid + 1]))) sm.sorted = 0;
    __syncthreads();
    // f32 compares for the 8x8 path (RAW_F32, or MAP with 64 rows), f64 compares for pages
    const bool f32cmp = sc.kind == RDM_SRC_RAW_F32 || (sc.kind == RDM_SRC_MAP_F32 && sc.rows == 64);
    if (f32cmp)
      build_lloyd_lut<float>(sm.lut, sm.thr_f, sm.sorted, sm.cell_s, tid, kAlsThreads);
    else
      build_lloyd_lut<double>(sm.lut, sm.thr_d, sm.sorted, sm.cell_s, tid, kAlsThreads);
  } else if (tid == 0) {
    sm.lut.ncell = 0;
  }
  __syncthreads();
  const int limit = sc.limit;
  const bool tt = (sc.flags & RDM_ALS_TRUE_TRANSPOSE) != 0;

  if (pages) {
    // ---- page items without pair structure: sequential units, the record scratch E aliases the staging tile
    const double inv_cnt = 1.0 / ((double)group * (double)(256 * kCols));
    for (int64_t item = local_cta; item < n_items; item += sc.cta_count) {
      if (item_is_compact(item)) continue;
      const int64_t g = item / sc.pages, pg = item - g * sc.pages;
      for (int k = tid; k <= limit; k += kAlsThreads) sm.recg[k] = 0.0;
      __syncthreads();
      float2 R[4][8];
      for (int b = 0; b < group; ++b) {
        const int64_t unit_idx = (g * group + b) * sc.pages + pg;
        load_unit<4>(R, sc, sm, tile, unit_idx, 0, tid, true);
        als_iterate<4, true>(R, sm, tile, 0, tid, limit, sm.recu, tt);
        __syncthreads();
        for (int k = tid; k <= limit; k += kAlsThreads) sm.recg[k] += (double)sm.recu[k];   // images in order, as CP:172-173 sums them
        __syncthreads();
      }
      for (int k = tid; k <= limit; k += kAlsThreads) {
        sm.rm[k] = (float)sqrt(sm.recg[k] * inv_cnt);
        if (sc.record_out) sc.record_out[item * (limit + 1) + k] = sm.rm[k];
      }
      __syncthreads();
      const int kstar = first_argmin<4>(sm.rm, limit, sm, 0, tid);
      if (sc.kstar_out && tid == 0) sc.kstar_out[item] = kstar;
      for (int b = 0; b < group; ++b) {
        const int64_t unit_idx = (g * group + b) * sc.pages + pg;
        load_unit<4>(R, sc, sm, tile, unit_idx, 0, tid, false);
        const float p = als_iterate<4, false>(R, sm, tile, 0, tid, kstar, nullptr, tt);
        emit_unit<4>(sc, sm, unit_idx, 0, tid, p);
        __syncthreads();
      }
    }
    return;
  }

  // ---- 64-row units: cluster = one reference batch
  const unsigned crank = cluster_ctarank(), csize = cluster_nctarank();
  const int64_t g = local_cta / (int)csize;
  const int unit = tid >> 6, lt = tid & 63;
  const int per_round = 4 * (int)csize;
  const int rounds = (group + per_round - 1) / per_round;
  const int64_t stride64 = als_ws_stride(64, limit);
  float* tile_u = tile + unit * (64 * 64);
  float* E_u = tile + kTileFloats + unit * ((limit + 1) * 33);
  float2 R[4][8];
  for (int r = 0; r < rounds; ++r) {
    const int b = r * per_round + (int)crank * 4 + unit;
    if (b < group) {
      const int64_t unit_idx = g * group + b;
      load_unit<1>(R, sc, sm, tile_u, unit_idx, unit, lt, true);
      als_iterate<1, true>(R, sm, E_u, unit, lt, limit, sc.ws + unit_idx * stride64, tt);
    }
  }
  __threadfence();
  cluster_sync_all();
  // every team sums the group's records in image order (f64), takes the rmse and its first minimum
  float* rm = sm.rm + unit * 64;
  const double inv_cnt = 1.0 / ((double)group * (double)(64 * kCols));
  for (int k = lt; k <= limit; k += 64) {
    const float* col = sc.ws + g * group * stride64 + k;
    double t = 0.0;
    int b = 0;
    for (; b + 8 <= group; b += 8) {   // 8 independent loads in flight, summed in image order
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldcg(col + (b + j) * stride64);
#pragma unroll
      for (int j = 0; j < 8; ++j) t += (double)v[j];
    }
    for (; b < group; ++b) t += (double)__ldcg(col + b * stride64);
    rm[k] = (float)sqrt(t * inv_cnt);
  }
  unit_barrier(1 + unit, 64);
  const int kstar = first_argmin<1>(rm, limit, sm, unit, lt);
  if (crank == 0 && unit == 0) {
    if (sc.record_out)
      for (int k = lt; k <= limit; k += 64) sc.record_out[g * (limit + 1) + k] = rm[k];
    if (sc.kstar_out && lt == 0) sc.kstar_out[g] = kstar;
  }
  for (int r = 0; r < rounds; ++r) {
    const int b = r * per_round + (int)crank * 4 + unit;
    if (b < group) {
      const int64_t unit_idx = g * group + b;
      if (rounds > 1) load_unit<1>(R, sc, sm, tile_u, unit_idx, unit, lt, false);   // single round: the tile is still in registers
      const float p = als_iterate<1, false>(R, sm, E_u, unit, lt, kstar, nullptr, tt);
      emit_unit<1>(sc, sm, unit_idx, unit, lt, p);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// CP:175-193 als_step as a stand-alone op: one warp per output row.
__global__ void __launch_bounds__(256) als_step_kernel(const float* __restrict__ ratings, const float* __restrict__ fixed,
                                                       int64_t batch, int rows, int cols, float reg, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t o = warp; o < batch * rows; o += nwarps) {
    const int64_t b = o / rows;
    const float* r = ratings + o * cols;
    const float* f = fixed + b * cols;
    float acc = 0.f;
    double nn = 0.0;
    for (int c = lane; c < cols; c += 32) {
      float fv = f[c];
      acc = fmaf(r[c], fv, acc);
      nn = fma((double)fv, (double)fv, nn);
    }
    acc = warp_sum(acc);
    nn = warp_sum(nn);
    if (lane == 0) out[o] = acc * (1.0f / ((float)nn + reg));
  }
}

}  // namespace rdm

using namespace rdm;

extern "C" int rdm_als_fused_phases(const rdm_als_scale_t*, int32_t, int64_t, int32_t, int32_t, rdm_stream_t);

extern "C" int64_t rdm_als_ws_floats(int32_t rows, int32_t pages, int32_t limit) {
  return (int64_t)pages * als_ws_stride(rows, limit);
}

extern "C" int rdm_als_fused(const rdm_als_scale_t* scales, int32_t n_scales, int64_t n_images, int32_t group,
                             rdm_stream_t stream) {
  return rdm_als_fused_phases(scales, n_scales, n_images, group, RDM_ALS_PHASE_ALL, stream);
}

extern "C" int rdm_als_fused_phases(const rdm_als_scale_t* scales, int32_t n_scales, int64_t n_images, int32_t group,
                                    int32_t phase_mask, rdm_stream_t stream) {
  RDM_REQUIRE(scales, "rdm_als_fused: null scales");
  RDM_REQUIRE(phase_mask >= 1 && phase_mask <= RDM_ALS_PHASE_ALL, "rdm_als_fused_phases: phase_mask must be in 1..%d", RDM_ALS_PHASE_ALL);
  RDM_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "rdm_als_fused: n_scales must be 1..%d (got %d)", kMaxScales, n_scales);
  RDM_REQUIRE(n_images >= 0, "rdm_als_fused: negative n_images");
  RDM_REQUIRE(group >= 1, "rdm_als_fused: group must be >= 1");
  RDM_REQUIRE(n_images % group == 0, "rdm_als_fused: n_images (%lld) must be a multiple of group (%d)", (long long)n_images, group);
  if (n_images == 0) return 0;
  AlsParams P;
  P.n_scales = n_scales;
  P.group = group;
  P.n_images = n_images;
  const int64_t n_groups = n_images / group;
  // cluster size of the dense launch: the 64-row units of one reference batch share a cluster, four per CTA
  int cluster = 1;
  for (int k = 0; k < n_scales; ++k)
    if (scales[k].rows == 64)
      while (cluster < 8 && 4 * cluster < group) cluster *= 2;
  int64_t ctas = 0;
  for (int k = 0; k < n_scales; ++k) {
    const rdm_als_scale_t& h = scales[k];
    RDM_REQUIRE(h.src && h.ws, "rdm_als_fused: scale %d: src and ws are required", k);
    RDM_REQUIRE(aligned16(h.ws), "rdm_als_fused: scale %d: ws must be 16-byte aligned", k);
    RDM_REQUIRE(h.rows == 64 || h.rows == 256, "rdm_als_fused: scale %d: rows must be 64 or 256 (got %d)", k, h.rows);
    RDM_REQUIRE(h.src_kind >= RDM_SRC_RAW_F64 && h.src_kind <= RDM_SRC_MAP_F32, "rdm_als_fused: scale %d: bad src_kind %d", k, h.src_kind);
    RDM_REQUIRE(h.limit >= 0 && h.limit <= (h.rows == 64 ? 63 : 127), "rdm_als_fused: scale %d: limit %d out of range", k, h.limit);
    RDM_REQUIRE(h.pages >= 1, "rdm_als_fused: scale %d: pages must be >= 1", k);
    RDM_REQUIRE((h.flags & ~RDM_ALS_FLAGS_ALL) == 0, "rdm_als_fused: scale %d: unknown flag bits 0x%x", k, h.flags);
    const bool needs_side = h.src_kind == RDM_SRC_MAP_F32 || h.map_out;
    if (needs_side) {
      if (h.rows == 64)
        RDM_REQUIRE(h.side == 8 && h.pages == 1, "rdm_als_fused: scale %d: rows=64 maps are 8x8, one page", k);
      else
        RDM_REQUIRE(is_pow2(h.side) && h.side >= 16 && h.side <= 128 && h.pages == (h.side / 16) * (h.side / 16),
                    "rdm_als_fused: scale %d: side %d / pages %d inconsistent", k, h.side, h.pages);
    }
    const bool quant = h.src_kind == RDM_SRC_RAW_F64 || h.src_kind == RDM_SRC_RAW_F32 || h.src_kind == RDM_SRC_MAP_F32;
    RDM_REQUIRE(!quant || (h.thresholds && h.levels), "rdm_als_fused: scale %d: codebook required for this src_kind", k);
    RDM_REQUIRE(quant || !h.bins_out, "rdm_als_fused: scale %d: bins_out needs a quantising src_kind", k);
    RDM_REQUIRE(h.src_kind == RDM_SRC_MAP_F32 || aligned16(h.src), "rdm_als_fused: scale %d: src must be 16-byte aligned", k);
    RDM_REQUIRE((!h.values_out || aligned16(h.values_out)) && (!h.bins_out || (reinterpret_cast<uintptr_t>(h.bins_out) & 3u) == 0),
                "rdm_als_fused: scale %d: values_out must be 16-byte and bins_out 4-byte aligned", k);
    RDM_REQUIRE((!h.pages_out || aligned16(h.pages_out)) && (!h.map_out || aligned16(h.map_out)),
                "rdm_als_fused: scale %d: pages_out and map_out must be 16-byte aligned", k);
    AlsScaleDev& d = P.s[k];
    d.src = h.src;
    d.thr = quant ? h.thresholds : nullptr;
    d.lvl = quant ? h.levels : nullptr;
    d.bins = h.bins_out;
    d.values = h.values_out;
    d.ws = h.ws;
    d.pages_out = h.pages_out;
    d.map_out = h.map_out;
    d.record_out = h.record_out;
    d.kstar_out = h.kstar_out;
    d.kind = h.src_kind;
    d.rows = h.rows;
    d.pages = h.pages;
    d.side = h.side;
    d.limit = h.limit;
    d.cta_begin = (int32_t)ctas;
    // page scales the compact kernel can take (rdm_als_sparse.cu) keep a few strided fallback CTAs here, each of
    // which normally finds nothing to do; every other page scale gets one CTA per (group, page) item
    const bool compact_eligible = h.rows == 256 && !(h.flags & (RDM_ALS_DENSE_ONLY | RDM_ALS_TRUE_TRANSPOSE)) &&
                                  (h.src_kind == RDM_SRC_RAW_F64 || h.src_kind == RDM_SRC_VAL_F64 || h.src_kind == RDM_SRC_MAP_F32);
    d.dense_only = (h.rows == 256 && !compact_eligible) ? 1 : 0;
    d.flags = h.flags;
    int64_t n;
    if (h.rows == 64) {
      n = n_groups * cluster;
      d.cta_count = (int32_t)n;
    } else {
      const int64_t items = n_groups * h.pages;
      n = compact_eligible ? std::min<int64_t>(items, std::max<int64_t>(8, (items + 7) / 8)) : items;
      d.cta_count = (int32_t)n;
      n = (n + cluster - 1) / cluster * cluster;   // whole clusters
    }
    ctas += n;
    RDM_REQUIRE(ctas < (1ll << 30), "rdm_als_fused: too many work units");
  }
  // dynamic shared memory: staging tile, plus the record scratch E
  size_t dyn1 = kTileFloats * sizeof(float), dyn = dyn1;
  for (int k = 0; k < n_scales; ++k) {
    const size_t need = (scales[k].rows == 256) ? (size_t)(scales[k].limit + 1) * 129 * sizeof(float)
                                                : dyn1 + (size_t)4 * (scales[k].limit + 1) * 33 * sizeof(float);
    if (need > dyn) dyn = need;
  }
  static size_t smem_set0[64];
  cudaError_t e = ensure_dyn_smem(als_kernel, dyn, smem_set0);
  if (e != cudaSuccess) {
    set_error("rdm_als_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (phase_mask & (RDM_ALS_PHASE_SPARSIFY | RDM_ALS_PHASE_PAGES)) {
    int rc = als_sparse_launch(scales, n_scales, n_images, group, (phase_mask & RDM_ALS_PHASE_SPARSIFY) != 0,
                               (phase_mask & RDM_ALS_PHASE_PAGES) != 0, (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (phase_mask & RDM_ALS_PHASE_DENSE) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kAlsThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, als_kernel, P);
    if (e != cudaSuccess) {
      set_error("rdm_als_fused: als_kernel launch: %s", cudaGetErrorString(e));
      return (int)e;
    }
    return launch_status("als_kernel");
  }
  return 0;
}

extern "C" int rdm_als_step_f32(const float* ratings, const float* fixed, int64_t batch, int32_t rows, int32_t cols,
                                float reg, float* out, rdm_stream_t stream) {
  RDM_REQUIRE(ratings && fixed && out, "rdm_als_step_f32: null pointer");
  RDM_REQUIRE(batch >= 0 && rows >= 1 && cols >= 1, "rdm_als_step_f32: bad shape");
  if (batch == 0) return 0;
  int64_t warps = batch * rows;
  int64_t blocks = (warps + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  als_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ratings, fixed, batch, rows, cols, reg, out);
  return launch_status("als_step_kernel");
}
