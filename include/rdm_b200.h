/* rdm_b200.h -- C ABI of librdm_b200.so: the B200 (sm_100a) implementation of the
 * MD_RDM depth-map fusion path.
 *
 * The reference (az16/MD_RDM) has no FFI layer: the path is a set of plain Python
 * functions in network/computations.py ("CP") and methods of Ordinal_Layer /
 * Weights in network/RDM_Net.py ("RN").  Each entry point below names the
 * reference function(s) it replaces (file:line).  The Python host code
 * (md_rdm_b200/ops.py) binds these symbols with ctypes and registers torch custom
 * ops on top; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer on the current device unless noted; buffers
 *    are dense row-major; the caller owns every buffer including workspaces;
 *    nothing is allocated, freed or retained by the library;
 *  - work is enqueued on `stream`; no host synchronisation, no host callbacks
 *    (CUDA-graph capturable);
 *  - return 0 = enqueued, <0 = argument error (nothing launched, see
 *    rdm_last_error()), >0 = cudaError_t of the failed launch;
 *  - "image" = one element of the reference batch dimension B; "group" = the B
 *    images of one reference call, which share the ALS arg-min (CP:74, CP:143).
 */
#ifndef RDM_B200_H
#define RDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDM_ABI_VERSION 2

typedef void* rdm_stream_t; /* cudaStream_t */

int rdm_abi_version(void);
/* thread-local, valid until the next failing call on this thread */
const char* rdm_last_error(void);

/* ------------------------------------------------------------------ stage 1: pair build */

/* RN:244-252 Ordinal_Layer.sparse_comparison_v1 (before its Lloyd call):
 * raw[n,i,j] = fl32(d[n,i] * fl32(1/d[n,j])), d = (N,64) f32 (the 8x8 map, row-major). */
int rdm_pair_v1_f32(const float* d3, int64_t n_images, float* raw_out, rdm_stream_t stream);

/* CP:308-311 cp.resize for newsize == side/2 (bicubic, align_corners=False, f64 out).
 * in: (N,side,side) f32 (in_is_f64=0) or f64; out: (N,side/2,side/2) f64. side in 2..128. */
int rdm_resize_half(const void* in, int32_t in_is_f64, int64_t n_images, int32_t side,
                    double* out, rdm_stream_t stream);

/* RN:259-280 sparse_comparison_id (before its Lloyd call) + CP:201-216 split_matrix +
 * the CP:308 resize that feeds them (RN:373, RN:386), for every 16x16 page of a
 * (N,side,side) f32 map, side in {16,32,64,128}, P=(side/16)^2 pages in row-major order.
 * raw_out: (N,P,256,64) f64.  parent_out (optional, may be NULL): (N,side/2,side/2) f64. */
int rdm_pair_id_f64(const float* dn, int64_t n_images, int32_t side, double* raw_out,
                    double* parent_out, rdm_stream_t stream);

/* RN:259-280 with caller-supplied pages (the literal sparse_comparison_id(dn, dn_1) signature):
 * pages (n,256) f32 = 16x16 pages, parents (n,64) f64 = their 8x8 parent pages (any values);
 * raw_out (n,256,64) f64. */
int rdm_pair_pages_f64(const float* pages, const double* parents, int64_t n_pages, double* raw_out,
                       rdm_stream_t stream);

/* CP:308-311 cp.resize for an arbitrary target size (network/module.py:68 resizes the 226x226
 * ground truth to 128x128): torch bicubic, align_corners=False, A=-0.75, f64 out.
 * in: (n,in_h,in_w) f32|f64; out: (n,out_h,out_w) f64. */
int rdm_resize_bicubic_f64(const void* in, int32_t in_is_f64, int64_t n_maps, int32_t in_h,
                           int32_t in_w, int32_t out_h, int32_t out_w, double* out,
                           rdm_stream_t stream);

/* CP:357-366 cp.upsample / cp.multi_upsample: `.double()` + nearest x2 applied `times` times.
 * in: (n,side,side) f32|f64; out: (n, side<<times, side<<times) f64. */
int rdm_upsample_nearest_f64(const void* in, int32_t in_is_f64, int64_t n_maps, int32_t side,
                             int32_t times, double* out, rdm_stream_t stream);

/* ------------------------------------------------------------------ stage 2: Lloyd quantisation */

/* RN:286-311 Ordinal_Layer.LloydQuantization: bin = #{i<40 : x >= thr[i]} evaluated in the
 * dtype of x (f32: thresholds rounded to f32 first), value = lvl[bin] rounded to that dtype.
 * thr40/lvl41 are DEVICE f64 tables (RN:397-418 Quantization). values_out may alias x
 * (the reference quantises in place); values_out and bins_out may each be NULL. */
int rdm_lloyd_quantize_f32(const float* x, int64_t n, const double* thr40, const double* lvl41,
                           float* values_out, uint8_t* bins_out, rdm_stream_t stream);
int rdm_lloyd_quantize_f64(const double* x, int64_t n, const double* thr40, const double* lvl41,
                           double* values_out, uint8_t* bins_out, rdm_stream_t stream);

/* ------------------------------------------------------------------ stage 3 (+2 fused): ALS */

enum {
  RDM_SRC_RAW_F64 = 0, /* raw ratios f64 -> Lloyd (f64 compare) -> ALS   (pages, RN:376-378) */
  RDM_SRC_RAW_F32 = 1, /* raw ratios f32 -> Lloyd (f32 compare) -> ALS   (8x8,   RN:362-364) */
  RDM_SRC_VAL_F32 = 2, /* already quantised / arbitrary matrix, f32        (CP:38, CP:95)     */
  RDM_SRC_VAL_F64 = 3, /* same, f64: `.float()` applied on load            (CP:40, CP:106)    */
  RDM_SRC_MAP_F32 = 4  /* decoder map f32: pair build + Lloyd + ALS fused, the pair matrix is
                          never written (rows=64: src=(N,64); rows=256: src=(N,side,side)) */
};

/* rdm_als_scale_t.flags */
enum {
  RDM_ALS_DENSE_ONLY = 1,     /* page matrices take the dense kernel even when they have the pair-build structure
                                 (A/B measurements, tests of the dense path) */
  RDM_ALS_TRUE_TRANSPOSE = 2, /* "paper-correct" knobs of SURVEY 8f rank 4, OFF by default (see below) */
  RDM_ALS_TRUE_GM = 4,
  RDM_ALS_CORRECT_TILING = 8,
  RDM_ALS_PAGES_ONE_CTA = 16, /* page ALS: the 16 images of a batch in ONE CTA per page (best when the chip is full) */
  RDM_ALS_PAGES_CLUSTER = 32, /* page ALS: a cluster of 4 CTAs per (batch, page) (lowest latency of a small launch);
                                 neither bit = chosen from the launch size; same results bit for bit either way */
  RDM_ALS_SKIP_UNUSED_PAGES = 64, /* opt-in: leave out the pages CP:218-238 (`reconstruct`, as written) never copies into
                                     the map - pages >= side/16 of every image; the reference computes and drops them.  The
                                     map and everything after it are unchanged bit for bit; pages_out / record_out /
                                     kstar_out / bins_out entries of those pages are not written.  Ignored together with
                                     RDM_ALS_CORRECT_TILING (then every page is used). */
  RDM_ALS_FLAGS_ALL = 127
};
/* rdm_als_fused_phases phase_mask bits: the launches of rdm_als_fused, selectable one by one (profiling) */
enum {
  RDM_ALS_PHASE_SPARSIFY = 1, /* compact page form: structure check + Lloyd of the page scales (the HBM-streaming kernel) */
  RDM_ALS_PHASE_PAGES = 2,    /* ALS on the compact pages, one CTA per (group, page), arg-min + normalise + re-tile inside */
  RDM_ALS_PHASE_DENSE = 4,    /* dense ALS: the 8x8 maps (one cluster per group) and page items without pair structure */
  RDM_ALS_PHASE_ALL = 7
};

/* One relative decoder scale.  Matrices are (N, pages, rows, 64), rows = 64 (8x8 map, square
 * case CP:38-85, limit 30) or 256 (16x16 page, CP:95-155, limit 100). */
typedef struct rdm_als_scale {
  const void* src;          /* see src_kind */
  int32_t src_kind;
  int32_t rows;             /* 64 or 256 */
  int32_t pages;            /* P = 1 (side 8, 16) or (side/16)^2 */
  int32_t side;             /* map side s */
  int32_t limit;            /* ALS iterations (rows 64: <= 63, rows 256: <= 127) */
  int32_t flags;            /* RDM_ALS_* bits below; 0 = the reference's behaviour on the fast path */
  const double* thresholds; /* device f64[40]; required for RAW_* and MAP_F32 */
  const double* levels;     /* device f64[41]; required for RAW_* and MAP_F32 */
  uint8_t* bins_out;        /* optional (N,P,rows,64) u8 Lloyd bins */
  float* values_out;        /* optional (N,P,rows,64) f32 quantised matrix as ALS sees it */
  float* pages_out;         /* optional (N,P,rows) f32: per-page ALS maps before re-tiling */
  float* map_out;           /* optional (N,side,side) f32: CP:218-238 `reconstruct` re-tiling
                               (bug-compatible: only pages 0..side/16-1 reach the map) */
  float* ws;                /* REQUIRED workspace, N * rdm_als_ws_floats(rows, pages, limit) f32, 16-byte
                               aligned: per (image, page) the 16 KB compact page form and its structure flags
                               (256-row units) or the unit's SSE record (64-row units).  No iterate is kept:
                               the batch-wide arg-min is taken inside the iterate kernels. */
  float* record_out;        /* optional (N/group,P,limit+1) f32 rmse record (CP:53-61) */
  int32_t* kstar_out;       /* optional (N/group,P) i32 selected iteration (CP:74, CP:143) */
} rdm_als_scale_t;

/* Lloyd + rank-1 ALS + batch-wide arg-min + geometric normalisation + page re-tiling for `n_scales`
 * scales in three launches (two without page scales).  Replaces Ordinal_Layer.forward (non-DORN branch, RN:358-396)
 * minus the pair build (unless src_kind == RDM_SRC_MAP_F32), i.e. LloydQuantization,
 * cp.quadratic_als, cp.alternating_least_squares, cp.als_step, cp.quick_gm, cp.reconstruct.
 * `scales` is a HOST array.  n_images must be a multiple of group. */
int rdm_als_fused(const rdm_als_scale_t* scales, int32_t n_scales, int64_t n_images,
                  int32_t group, rdm_stream_t stream);
/* The same, launching only the phases selected by RDM_ALS_PHASE_* bits.  PAGES and DENSE read the per-page
 * structure flags SPARSIFY leaves in ws, so they must follow a SPARSIFY launch on the same inputs (same stream, or
 * ordered by an event); they are independent of each other. */
int rdm_als_fused_phases(const rdm_als_scale_t* scales, int32_t n_scales, int64_t n_images,
                         int32_t group, int32_t phase_mask, rdm_stream_t stream);
/* SURVEY 8f rank 3 - RN:146, RN:157 `Decoder.conv1` (1x1 conv, C channels -> the one-channel relative map) fused with
 * the pair build of RN:259-284: feat (N,C,side,side) f32, weight (C) f32 (conv1.weight.view(-1)), bias (1) f32 or NULL.
 * map_out (optional, NULL allowed when page_scale is given): (N,side,side) f32 decoder map.  page_scale (HOST pointer,
 * optional, side >= 16): the scale descriptor a following rdm_als_fused_phases(..., RDM_ALS_PHASE_PAGES |
 * RDM_ALS_PHASE_DENSE, ...) call will use - its ws receives the compact page form exactly as RDM_ALS_PHASE_SPARSIFY
 * would leave it (so that phase is skipped and the map never touches HBM), bins_out / values_out are honoured.
 * C must be a multiple of 8, side in {8,16,32,64}.  The map agrees with torch's conv to f32 rounding (different
 * summation order); Lloyd bins are exact for the map this kernel produces. */
int rdm_conv_head_f32(const float* feat, const float* weight, const float* bias, int64_t n_images,
                      int32_t channels, int32_t side, float* map_out,
                      const rdm_als_scale_t* page_scale, rdm_stream_t stream);

/* Host-side copy of the compact page form's geometry (RN:266-273 + CP:269-295 window anchoring), for tests:
 * window_cols i32[256*9] = the nine window columns of every matrix row (row-major over the 3x3 window),
 * fill_col i32[256] = a column outside every window of that pixel row, compact u8[256*16] = source of each
 * entry of the 16-float compact row (0..8 window slot minus f, 9 = f, 15 = zero).  No device work. */
int rdm_sparsify_geometry(int32_t* window_cols, int32_t* fill_col, uint8_t* compact);
/* f32 workspace elements per image for one scale (rdm_als_scale_t.ws) */
int64_t rdm_als_ws_floats(int32_t rows, int32_t pages, int32_t limit);

/* CP:175-193 cp.als_step: out[b,i] = (sum_j ratings[b,i,j] * fixed[b,j]) * (1/(|fixed_b|^2+reg)).
 * ratings (B,rows,cols) f32, fixed (B,cols) f32, out (B,rows) f32. */
int rdm_als_step_f32(const float* ratings, const float* fixed, int64_t batch, int32_t rows,
                     int32_t cols, float reg, float* out, rdm_stream_t stream);

/* ------------------------------------------------------------------ stage 4: decomposition */

/* CP:244-255 cp.quick_gm over dim 1: out[b] = prod_i pow(t[b,i], 1/rc^2).
 * dtype: 0 = f32, 1 = f64, 2 = i64 (computed in f32 like torch.pow(int64, float)). */
int rdm_quick_gm(const void* t, int32_t dtype, int64_t batch, int64_t n, int32_t rc, void* out,
                 rdm_stream_t stream);

/* RN:117 / network/module.py:145-149: x / quick_gm(x.view(B,HW,1), H) for a (N,side,side) map.
 * dtype as above; out is f32 for f32/i64 input and f64 for f64 input. */
int rdm_gm_normalize(const void* x, int32_t dtype, int64_t n_images, int32_t side, void* out,
                     rdm_stream_t stream);

/* CP:368-392 cp.decompose_depth_map (+ callers' [::-1]): whole pyramid of a (N,side,side) map
 * in one launch.  n = log2(side) levels, side <= 128.  pyramid_out: f64, LEVEL-MAJOR
 * [D_0: (N,1), only if relative_map == 0] [F_1: (N,4)] [F_2: (N,16)] ... [F_n: (N,side^2)], so
 * every component is a dense (N,1,2^k,2^k) tensor.  in_is_f64: 0 = f32 input, 1 = f64 input. */
int rdm_decompose(const void* in, int32_t in_is_f64, int64_t n_images, int32_t side,
                  int32_t relative_map, double* pyramid_out, rdm_stream_t stream);
/* number of f64 values per image written by rdm_decompose (total = N times this) */
int64_t rdm_pyramid_len(int32_t side, int32_t relative_map);
/* backward of rdm_decompose (autograd through CP:368-392): grad_pyramid in the same level-major
 * layout, grad_in (N,side,side) in the dtype of `in`. */
int rdm_decompose_bwd(const void* in, int32_t in_is_f64, int64_t n_images, int32_t side,
                      int32_t relative_map, const double* grad_pyramid, void* grad_in,
                      rdm_stream_t stream);
/* backward of rdm_quick_gm (grad_gm (B), optional) and/or rdm_gm_normalize (grad_norm (B,n),
 * optional) for f32 (is_f64 = 0) or f64 x (B,n); all gradients in x's dtype. */
int rdm_gm_bwd(const void* x, int32_t is_f64, int64_t batch, int64_t n, int32_t rc,
               const void* grad_gm, const void* grad_norm, void* grad_x, rdm_stream_t stream);

/* Backward of rdm_fuse_tail to `weights` in two launches (the training step, network/module.py:89-95: only the MSE on the
 * final depth reaches Weights): grad_w[first_k + j] = sum_b sum_m f32(A_k[b,j,m]) * f32(pool_k(grad_depth)[b,m]), pool_k =
 * the successive 2x2 sums of rdm_recombination_bwd (n = 7).  Bit-identical to rdm_recombination_bwd followed by
 * rdm_make_pred_bwd slot by slot.  A: HOST array of kmax+1 device pointers (A_out of rdm_fuse_tail); K: HOST, candidates
 * per slot; ws: n_images * (4^(kmax+1) - 1) / 3 floats of scratch; grad_w: sum K floats, overwritten. */
int rdm_fuse_tail_bwd(const double* grad_depth, const double* const* A, const int32_t* K, int32_t kmax,
                      int64_t n_images, float* ws, float* grad_w, rdm_stream_t stream);

/* network/computations.py:499-510 (the detached per-scale component loss of the training step) in one launch:
 * out[0] = sum_{k=0..kmax} mean((yhat_k - target_k)^2), f64.  yhat: (N, sum 4^k) f32 as rdm_fuse_tail packs it;
 * target: level-major pyramid (rdm_gt_prepare / rdm_decompose with relative_map = 0). */
int rdm_component_loss(const float* yhat, const double* target, int64_t n_images, int32_t kmax, double* out,
                       rdm_stream_t stream);

/* Ground-truth preparation of the training step in ONE launch: network/module.py:68 cp.resize(y, 128)
 * (bicubic), :74-78 mask (+1e-4 everywhere, invalid -> 1.0001), :145-149 normalize, :123 decompose n = 7, and the ordinal
 * target of :126-127 / :134-143: utils.depth2label_sid(cp.resize(y, 8)) (utils.py:195-211) whose normalised
 * decomposition replaces D_0.  y_raw: (N,in_h,in_w) f32|f64.  y_out: (N,128,128) f64 masked map (what the MSE uses);
 * pyramid_out: N * rdm_pyramid_len(128, 0) f64, level-major like rdm_decompose, slot 0 = the ORDINAL D_0;
 * ord_target_out: (N,64) i32 SID labels.  sid_K, sid_alpha and sid_log_ratio = log(beta / alpha) are passed as the
 * f32-rounded scalars the reference's torch code holds (K = 90, alpha = 0.02, beta = 10). */
int rdm_gt_prepare(const void* y_raw, int32_t in_is_f64, int64_t n_images, int32_t in_h, int32_t in_w,
                   double sid_K, double sid_alpha, double sid_log_ratio, double* y_out,
                   double* pyramid_out, int32_t* ord_target_out, rdm_stream_t stream);

/* ------------------------------------------------------------------ stage 5: weighted reconstruction */

/* CP:464-484 cp.make_matrix: out[b,k,:] = log(cand_k[b,:]) for K candidates of M values each.
 * cands: HOST array of K device pointers to (B,M) f64. out (B,K,M) f64. */
int rdm_log_stack_f64(const double* const* cands, int32_t K, int64_t batch, int64_t M, double* out,
                      rdm_stream_t stream);

/* backward of rdm_log_stack_f64: grad_cands[k][b,m] = grad_out[b,k,m] / cands[k][b,m]. */
int rdm_log_stack_bwd(const double* const* cands, int32_t K, int64_t batch, int64_t M,
                      const double* grad_out, double* const* grad_cands, rdm_stream_t stream);

/* CP:512-528 cp.make_pred for one slot: out[b,m] = sum_k f32(A[b,k,m]) * w[k] (f32). */
int rdm_make_pred_f32(const double* A, const float* w, int64_t batch, int32_t K, int64_t M,
                      float* out, rdm_stream_t stream);
/* backward of the above: grad_A (B,K,M) f64 (optional), grad_w (K) f32 (optional, overwritten). */
int rdm_make_pred_bwd(const double* A, const float* w, const float* grad_out, int64_t batch,
                      int32_t K, int64_t M, double* grad_A, float* grad_w, rdm_stream_t stream);

/* CP:394-421 cp.recombination (+ CP:357-366 upsample/multi_upsample):
 * out[b,y,x] = sum_i comp_i[b, y >> sh_i, x >> sh_i] in f64, out side = 2^n, summed in the
 * reference's order (d_0 added last).  comps: HOST array of n_comps device pointers in list
 * order: optional d_0 (side 1) then sides 2, 4, ... (any prefix); comp i is (B,side_i,side_i)
 * f32 (comps_are_f64 = 0) or f64; sides[] HOST. */
int rdm_recombination_f64(const void* const* comps, const int32_t* sides, int32_t n_comps,
                          int32_t comps_are_f64, int64_t batch, int32_t n, double* out,
                          rdm_stream_t stream);
/* backward: grad_comp_i[b,u,v] = sum over the 2^sh x 2^sh block of grad_out, by successive 2x2
 * pooling (what autograd does through the upsample chain).  Entries of grad_comps may be NULL.
 * n <= 7. */
int rdm_recombination_bwd(const double* grad_out, void* const* grad_comps, const int32_t* sides,
                          int32_t n_comps, int32_t comps_are_f64, int64_t batch, int32_t n,
                          rdm_stream_t stream);

/* Fused stages 4+5 for the configuration RN:96-133 intends (decoder 1 + relative decoders):
 *   f_d1 = decompose(x_d1 / gm(x_d1), 3); f_dk = decompose(rel_k, log2 side_k, relative);
 *   A = relative_fine_detail_matrix; y_hat = make_pred(weights, A); depth = recombination(y_hat)
 * x_d1: (N,64) i64; rel[k]: (N,side_k,side_k) f32, side_k <= 64, HOST array of device ptrs;
 * weights: device f32, slots concatenated in slot order [d0 | f1 | ... ], within a slot in
 * decoder order (decoder 1 first); weight count per slot is implied by `sides`.
 * yhat_out (optional): (N, sum_{k<=kmax} 4^k) f32 packed by slot; depth_out: (N,128,128) f64 (may be NULL if
 * depth_compact_out is given); depth_compact_out (optional): (N,2^kmax,2^kmax) f64, kmax = log2 of the largest side
 * (>= 3) - the log-depth map is constant on blocks of 2^(7-kmax) pixels (no component is finer), so this holds every
 * distinct value of it: depth[b,y,x] == compact[b, y >> (7-kmax), x >> (7-kmax)] bit for bit;
 * A_out (optional): HOST array of 8 device pointers (entries may be NULL), A_out[k] receives the
 * slot-k fine-detail matrix (N, K_k, 4^k) f64 exactly as cp.relative_fine_detail_matrix builds it
 * (what the backward needs). */
int rdm_fuse_tail(const int64_t* x_d1, const float* const* rel, const int32_t* sides,
                  int32_t n_rel, const float* weights, int64_t n_images, float* yhat_out,
                  double* depth_out, double* depth_compact_out, double* const* A_out,
                  rdm_stream_t stream);
/* The same with the number of CTAs per image chosen by the caller: bands = 1, 2, 4 or 8 (a thread-block cluster per
 * image; clamped to 2^kmax), 0 = as rdm_fuse_tail (one CTA per image from 16 images up: best with many calls in flight).
 * A caller whose launch is alone on the GPU (a training step) gets a shorter launch from 4 bands (batch 16: 11 vs 18 us).
 * Results are bit-identical for every band count. */
int rdm_fuse_tail_bands(const int64_t* x_d1, const float* const* rel, const int32_t* sides,
                  int32_t n_rel, const float* weights, int64_t n_images, float* yhat_out,
                  double* depth_out, double* depth_compact_out, double* const* A_out,
                  int32_t bands, rdm_stream_t stream);
/* number of f32 weights rdm_fuse_tail expects for these relative decoder sides (-1 = bad sides) */
int64_t rdm_fuse_tail_weight_count(const int32_t* sides, int32_t n_rel);

/* ------------------------------------------------------------------ SURVEY 8f "next": DORN head + ordinal loss */

/* RN:313-345 Ordinal_Layer.DornOrdinalRegression: x (N,2K,H,W) f32 ->
 * ord_out (N,K,H,W) f64 = softmax over each (even, odd) channel pair of clamp(x, 1e-8, 1e4), odd member;
 * decode_out (N,H*W) i64 = #{k : ord > 0.5}.  HW = H*W <= 8192. */
int rdm_dorn_regression_f32(const float* x, int64_t n_images, int32_t K, int32_t HW,
                            int64_t* decode_out, double* ord_out, rdm_stream_t stream);
/* backward of the above w.r.t. x (grad_x (N,2K,H,W) f32, fully overwritten). */
int rdm_dorn_regression_bwd(const float* x, const double* ord, const double* grad_ord,
                            int64_t n_images, int32_t K, int32_t HW, float* grad_x,
                            rdm_stream_t stream);

/* utils.py:195-211 depth2label_sid (the ordinal target of network/module.py:126, 134-143): depth (n) f32|f64 ->
 * labels (n) i32 = int(max(K * log(depth / alpha) / log(beta / alpha), 0)).  K, alpha and log_ratio = log(beta / alpha)
 * are the f32-rounded scalars the reference's torch code holds (K = 90, alpha = 0.02, beta = 10). */
int rdm_depth2label_sid(const void* depth, int32_t is_f64, int64_t n, double sid_K, double sid_alpha,
                        double sid_log_ratio, int32_t* labels_out, rdm_stream_t stream);

/* loss.py:17-59 Ordinal_Loss.calc: ord (N,K,H,W) f64, target (N,H*W) i32 SID labels ->
 * loss_out[0] = -(sum_{k<=t} log(f32(clamp(ord))) + sum_{k>t} log(f32(clamp(1-ord)))) / (N H W), f32.
 * ws: caller workspace of rdm_ordinal_loss_ws_doubles() f64 values. */
int rdm_ordinal_loss_f64(const double* ord, const int32_t* target, int64_t n_images, int32_t K,
                         int32_t HW, double* ws, float* loss_out, rdm_stream_t stream);
int64_t rdm_ordinal_loss_ws_doubles(void);
/* backward: grad_ord (N,K,H,W) f64 from grad_loss (device f32 scalar). */
int rdm_ordinal_loss_bwd(const double* ord, const int32_t* target, const float* grad_loss,
                         int64_t n_images, int32_t K, int32_t HW, double* grad_ord,
                         rdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RDM_B200_H */
