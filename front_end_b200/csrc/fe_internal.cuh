// Internal declarations shared by the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/fe_abi.h"

namespace fe {

constexpr int STRIP_ROWS = 8;          // FAST strip height of the generic kernel: one CTA = full image width x 8 rows
constexpr int WIDE_STRIP_ROWS = 30;    // strip height of the FAST-9_16 rolling-window kernel (fast16_strip_kernel)
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ inline int div_up(int v, int m) { return (v + m - 1) / m; }

// Geometry of one batch launch (all images of a batch share width/height).
struct Geom {
    int w, h, pitch;       // pitch = round_up(w, 16) bytes
    int n_images;
    int n_strips;          // div_up(h, STRIP_ROWS); also the per-image stride of strip_raw / strip_sel
    int slab_cap;          // candidate capacity per 8-row strip (u32 records)
    size_t slab_img;       // slab records per image: room for either strip layout (see strip_view)
    int kp_cap;            // per-image keypoint capacity
    size_t img_stride;     // bytes between consecutive images (pitch * h)
    int rs_h;              // rows of the row tables (Buffers::rowstart): fe_config.max_height, whatever image is resident
};

// Strip geometry of an image of g.w x g.h (pitch set): shared by set_geom and the pyramid levels.
inline void set_strip_geometry(Geom &g, bool nonmax) {
    g.n_strips = div_up(g.h, STRIP_ROWS);
    g.slab_cap = nonmax ? g.pitch * STRIP_ROWS / 4 : g.pitch * STRIP_ROWS;
    const size_t narrow = (size_t)g.n_strips * (size_t)g.slab_cap;
    const size_t wide = (size_t)div_up(g.h, WIDE_STRIP_ROWS) * (size_t)(g.pitch * WIDE_STRIP_ROWS / 4);
    g.slab_img = narrow > wide ? narrow : wide;
}

struct DetectParams {
    int threshold;         // FAST t
    int ps;                // 16 / 12 / 8
    int nonmax;
    int n_features;        // < 0: keep all
    int edge;              // border filter
    const int *thr_img = nullptr;   // device, optional: per-image FAST thresholds (grid cells: live_stereo.cpp:293,306)
};

// cv::fastAtan2 (degrees).  Host+device so the exact polynomial can be unit-tested on the CPU.
__host__ __device__ inline float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = 2.2204460492503131e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
#ifdef __CUDA_ARCH__
    if (ax >= ay) c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    else c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fadd_rn(__fmul_rn(p7, c2), p5);
    a = __fadd_rn(__fmul_rn(a, c2), p3);
    a = __fadd_rn(__fmul_rn(a, c2), p1);
    a = __fmul_rn(a, c);
    if (!(ax >= ay)) a = __fsub_rn(90.f, a);
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
#else
    if (ax >= ay) c = ay / (ax + eps);
    else c = ax / (ay + eps);
    c2 = c * c;
    a = p7 * c2; a = a + p5;
    a = a * c2;  a = a + p3;
    a = a * c2;  a = a + p1;
    a = a * c;
    if (!(ax >= ay)) a = 90.f - a;
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
#endif
    return a;
}

// Warp-cooperative exact FP32 squared L2 distance between a query held in registers (lane l owns dims
// [D/32 * l, D/32 * (l + 1))) and a descriptor row in global memory: coalesced row load, per-lane FMA chain,
// xor-butterfly sum -- a fixed, deterministic summation order; every lane returns the same value.
template <int D>
struct WarpRow {
    float q[D / 32];
    __device__ __forceinline__ void load(const float *row, int lane) {
        if (D == 128) { const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + lane); q[0] = v.x; q[1] = v.y; q[D == 128 ? 2 : 0] = v.z; q[D == 128 ? 3 : 1] = v.w; }
        else { const float2 v = __ldg(reinterpret_cast<const float2 *>(row) + lane); q[0] = v.x; q[1] = v.y; }
    }
    __device__ __forceinline__ float dist2(const float *row, int lane) const {
        float t[D / 32];
        if (D == 128) { const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + lane); t[0] = v.x; t[1] = v.y; t[D == 128 ? 2 : 0] = v.z; t[D == 128 ? 3 : 1] = v.w; }
        else { const float2 v = __ldg(reinterpret_cast<const float2 *>(row) + lane); t[0] = v.x; t[1] = v.y; }
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) { const float df = __fsub_rn(q[i], t[i]); acc = __fmaf_rn(df, df, acc); }
#pragma unroll
        for (int off = 16; off; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        return acc;
    }
};

// ---- device buffers (one set per ctx, sized for cfg.max_*) ------------------------------------
struct Buffers {
    uint8_t *img = nullptr;        // [n_images][h][pitch]
    uint8_t *blur = nullptr;       // same layout
    uint8_t *respmap = nullptr;    // same layout: NMS-surviving FAST responses (s - t), 0 elsewhere
    uint32_t *slab = nullptr;      // [n_images][slab_img]: per strip (StripView) records score<<24 | ylocal<<16 | x
    uint32_t *strip_raw = nullptr; // [n_images][n_strips] candidates emitted per strip
    uint32_t *strip_sel = nullptr; // [n_images][n_strips] survivors of the top-N cut per strip
    uint32_t *hist = nullptr;      // [n_images][256] response histogram (inside the border)
    uint32_t *n_kp = nullptr;      // [n_images] keypoints found (may exceed kp_cap)
    uint32_t *kp_key = nullptr;    // [n_images][kp_cap]  (y<<16 | x), ascending = raster order
    uint8_t *kp_score = nullptr;   // [n_images][kp_cap]
    float *harris = nullptr;       // [n_images][kp_cap] Harris responses (cv::ORB HARRIS_SCORE), lazily allocated
    fe_kpoint *kp = nullptr;       // [n_images][kp_cap]  wire layout
    float *kx = nullptr, *ky = nullptr;   // [n_images][kp_cap]
    float2 *kcs = nullptr;         // [n_images][kp_cap]  (cos, sin) of the steering angle
    uint8_t *desc = nullptr;       // [n_images][kp_cap][32]
    float *fdesc = nullptr;        // [n_images][kp_cap][128] float descriptors (SURF), lazily allocated
    int32_t *integral = nullptr;   // [n_images][(h+1)][(w+1)] CV_32S integral image (SURF orientation), lazy
    // matching, per pair (image 2p = query/left, 2p+1 = train/right)
    uint32_t *best = nullptr, *second = nullptr, *allbest = nullptr, *colbest = nullptr; // [n_pairs][kp_cap]
    unsigned long long *best64 = nullptr, *second64 = nullptr, *allbest64 = nullptr, *colbest64 = nullptr;  // L2 keys, lazy
    // tensor-core L2 path, lazy: bf16 operands in UMMA core-matrix layout, |x|^2, per-row candidates, error flag
    uint16_t *bf16desc = nullptr;  // [n_images][tiles][dim/8][128][8]
    float *fnorm = nullptr;        // [n_images][tiles * 128]
    uint32_t *cand = nullptr;      // [n_pairs][2][kp_cap][4]
    int *tc_error = nullptr;
    // exact tensor-core cross-check (l2verify.cu), lazy: band candidates, thresholds, flagged-element lists
    unsigned long long *vf_candL = nullptr, *vf_candR = nullptr;    // [n_pairs][kp_cap]
    float *vf_limq = nullptr, *vf_limt = nullptr, *vf_limqd = nullptr, *vf_limtd = nullptr;   // [n_pairs][round_up(kp_cap, 128)]
    uint32_t *vf_list = nullptr, *vf_npush = nullptr, *vf_maxnorm = nullptr;   // [n_pairs][32 kp_cap], [n_pairs], [n_images]
    fe_match *match_a = nullptr, *match_b = nullptr;   // [n_pairs][kp_cap]
    uint32_t *n_a = nullptr, *n_b = nullptr;           // [n_pairs]
    int *rowstart = nullptr;                           // [n_images][rs_h + 2] first keypoint whose floor(y) >= row (raster-ordered lists)
    uint32_t *n_override = nullptr;                    // [n_images] counts for externally supplied kps
    int *thr_img = nullptr;                            // [n_images] per-image FAST thresholds (grid detector)
    int *umax = nullptr;                               // [128] OpenCV's umax table of the general orientation kernel
    int8_t *brief_pat = nullptr;                       // [512][4] (y1, x1, y2, x2) tests of cv::BriefDescriptorExtractor, lazy
    uint8_t *brief_desc = nullptr;                     // [n_images][kp_cap][64] BRIEF-16 / 32 / 64 rows, lazy
    int8_t *pattern = nullptr;                         // [512][2] rBRIEF points when patch size != 31 or WTA_K != 2 (else the built-in table)
    // WindowMatcher sequence buffers, lazy: landmark lists as virtual pairs (cur = slot 2v, prev = slot 2v + 1)
    uint8_t *wdesc = nullptr;      // [n_images][kp_cap][32]
    float *wkx = nullptr, *wky = nullptr;              // [n_images][kp_cap]
    uint32_t *wcount = nullptr;                        // [n_images]
    uint32_t *wbest = nullptr, *wsecond = nullptr;     // [n_pairs][kp_cap]
    fe_match *wmatch = nullptr;                        // [n_pairs][kp_cap]
    uint32_t *wn = nullptr;                            // [n_pairs]
    double *wq = nullptr, *wxyz = nullptr;             // [16], [n_pairs][kp_cap][3]
    uint8_t *wdesc_r = nullptr;                        // [n_images][kp_cap][32] right descriptors of the landmarks (liveGraph mode), lazy
    uint32_t *wbest_r = nullptr, *wcol_r = nullptr;    // [n_pairs][kp_cap] row / column arg-mins of the right-descriptor cross-check
    // fe_window_update: the current (slot 0) and previous (slot 1) frame's landmark lists, lazy
    fe_kpoint *wu_kp = nullptr;                        // [2][kp_cap]
    uint8_t *wu_desc = nullptr, *wu_rdesc = nullptr;   // [2][kp_cap][32]
    float *wu_kx = nullptr, *wu_ky = nullptr;          // [2][kp_cap]
    float2 *wu_kcs = nullptr;                          // [2][kp_cap] (scratch of the keypoint unpack)
    uint32_t *wu_n = nullptr;                          // [2]
    // pruned cross-check, lazy: band candidates, thresholds, easy / hard partitions
    uint32_t *cx_bestL = nullptr, *cx_bestR = nullptr, *cx_dummy = nullptr;   // [n_pairs][kp_cap]
    int *cx_thrq = nullptr, *cx_thrt = nullptr;                                // [n_pairs][kp_cap]
    uint16_t *cx_qperm = nullptr, *cx_tperm = nullptr;                         // [n_pairs][kp_cap]
    uint32_t *cx_n = nullptr;                                                  // [n_pairs][8]: class sizes A, B, C of the queries, then of the trains
    uint16_t *cx_half = nullptr, *cx_star = nullptr;   // [n_images][16][kp_cap] halves in permutation order, [n_images][kp_cap] own candidate
    // Fast-Hessian scale space (single image), lazy
    float *hes_det = nullptr, *hes_trace = nullptr;    // [image chunk][all layers back to back]
    uint32_t *hes_count = nullptr;                     // [n_images] accepted maxima, then [n_images] largest size (float bits)
    fe_kpoint *hes_kp = nullptr;                       // [n_images][kp_cap] unsorted maxima
    // stereoLandmarks packing, lazy
    fe_kpoint *lm_lkp = nullptr, *lm_rkp = nullptr;    // [n_pairs][kp_cap]
    uint8_t *lm_ldesc = nullptr, *lm_rdesc = nullptr;  // [n_pairs][kp_cap][32]
    fe_match *lm_match = nullptr;                      // [n_pairs][kp_cap]
    // multi-level ORB, lazy: two ping-pong level images, coefficient table, accumulated results
    uint8_t *pyr_img[2] = {nullptr, nullptr};          // [n_images][h1][pitch1] (level-1 geometry is the largest)
    int *pyr_tab = nullptr;                            // [2 * (max_width + max_height)]
    fe_kpoint *pyr_kp = nullptr;                       // [n_images][kp_cap]
    uint8_t *pyr_desc = nullptr;                       // [n_images][kp_cap][32]
    uint32_t *pyr_n = nullptr;                         // [n_images]
};

// The strip layout the FAST kernel of this (geometry, parameters) writes and the top-N selection reads: strip s of an image
// starts at slab + image * g.slab_img + s * cap and holds its candidates in raster order (ylocal < rows).
struct StripView { int rows, n, cap; };
StripView strip_view(const Geom &g, const DetectParams &p);

// ---- kernel launchers (each returns the number of kernels it launched) -------------------------
int launch_fast(const Geom &g, const DetectParams &p, const Buffers &b, cudaStream_t s);
int launch_select(const Geom &g, const DetectParams &p, const Buffers &b, cudaStream_t s);
int launch_orient_pack(const Geom &g, const DetectParams &p, const Buffers &b, bool orientation,
                       float kp_size, int half_patch, cudaStream_t s);
int launch_blur(const Geom &g, const Buffers &b, cudaStream_t s);
// cv::ORB HARRIS_SCORE (harris.cu): score the selected FAST corners, keep the n_features best (ties kept, raster order);
// then, after the keypoints are packed, store the score as their response
int launch_harris_select(const Geom &g, int n_features, const Buffers &b, cudaStream_t s);
int launch_harris_store(const Geom &g, const Buffers &b, cudaStream_t s);

// cv::cornerSubPix (win x win half-size, zeroZone -1) for the keypoints of every image of the batch.
// Sampling image of batch image i: src[i] (w[i] x h[i], row pitch[i]); the point is  (kp + pre) -> refine ->
// (+ post1) + post2  with float adds in that order (src/live_stereo.cpp:321-350, features.py:623-640).
constexpr int SUBPIX_MAX_IMAGES = 16;
struct SubpixParams {
    const uint8_t *src[SUBPIX_MAX_IMAGES];
    int w[SUBPIX_MAX_IMAGES], h[SUBPIX_MAX_IMAGES], pitch[SUBPIX_MAX_IMAGES];
    float pre_x[SUBPIX_MAX_IMAGES], pre_y[SUBPIX_MAX_IMAGES];
    float post1_x[SUBPIX_MAX_IMAGES], post1_y[SUBPIX_MAX_IMAGES], post2_x[SUBPIX_MAX_IMAGES], post2_y[SUBPIX_MAX_IMAGES];
    int refine;              // 0: only apply the offsets
    int max_iters;           // 40
    float epsilon;           // 0.001
};
int launch_subpix(const Geom &g, const Buffers &b, const uint32_t *counts, const SubpixParams &sp, cudaStream_t s);
int launch_brief(const Geom &g, const Buffers &b, const uint32_t *counts, cudaStream_t s);
// rBRIEF with the ctx's own pattern (b.pattern; ORB::setPatchSize != 31), reflect-101 raw pixels outside the image
int launch_brief_general(const Geom &g, const Buffers &b, const uint32_t *counts, cudaStream_t s);
// ORB WTA_K = 3 / 4: 128 tuples of wta_k points from b.pattern, two bits per tuple
int launch_brief_wta(const Geom &g, const Buffers &b, const uint32_t *counts, int wta_k, cudaStream_t s);
int launch_unpack_kps(const Geom &g, const Buffers &b, const uint32_t *counts, cudaStream_t s);
// cv::BriefDescriptorExtractor(bytes) at the keypoints in b.kp with the caller's test table (brief.cu); needs b.integral.
// out rows are 64 bytes apart (bytes used)
int launch_brief_ext(const Geom &g, const Buffers &b, const uint32_t *counts, const int8_t *pattern, int bytes, int use_orientation,
                     uint8_t *out, cudaStream_t s);

// cv::FREAK at the keypoints in b.kp (freak.cu); needs b.integral and b.img.  table: [64][256][43] (x, y, sigma); opairs: [45]
// (i, j, weight_dx, weight_dy); dpairs: [512] (i, j); scale_idx: [n_images][kp_cap].  Writes kp.angle and 64-byte rows.
int launch_freak(const Geom &g, const Buffers &b, const uint32_t *counts, const int32_t *scale_idx, const float *table,
                 const int4 *opairs, const uchar2 *dpairs, int orientation_normalized, uint8_t *out, cudaStream_t s);

// SURF / SURF_EXTENDED descriptors at the keypoints in b.kp (x, y, size); writes b.fdesc rows of
// `128` floats (64 used when !extended), kp.angle, and kp.size = -1 for keypoints the reference drops.
// max_win: upper bound of (int)(21 * size * 1.2 / 9) over the batch's keypoints (picks the shared-memory variant).
int launch_surf(const Geom &g, const Buffers &b, const uint32_t *counts, bool extended, bool upright, int max_win,
                cudaStream_t s);
constexpr int SURF_DIRECT_MAX_WIN = 1024;   // windows above SURF_MAX_WIN are produced on the fly (never staged), up to this size
constexpr int SURF_MAX_WIN = 88;   // largest supported (int)(21 * size * 1.2 / 9): keypoint size <= 31.4 (ORB: 31 -> 86)

// multi-level ORB (pyramid.cu)
int launch_resize_linear_exact(const uint8_t *src, int sw, int sh, int spitch, size_t sstride, uint8_t *dst, int dw, int dh,
                               int dpitch, size_t dstride, const int *tab, int n_images, cudaStream_t s);
int launch_pyr_append(const Geom &g, int level, float scale, float kp_size, const Buffers &b, fe_kpoint *akp, uint8_t *adesc,
                      uint32_t *n_acc, bool with_desc, cudaStream_t s);
int launch_pyr_coords(const Geom &g, const Buffers &b, cudaStream_t s);

// SURF Fast-Hessian detector (surf_detect.cu)
struct HaarBoxI { int dx1, dy1, dx2, dy2; float w; };
struct HessianLayer {
    int size, step, margin, samples_i, samples_j, valid;
    HaarBoxI box[10];          // 3 Dxx, 3 Dyy, 4 Dxy boxes of the 9 x 9 pattern resized to `size`
};
// all three take a batch of n_images (blockIdx.z / .y): image i reads S + i * s_img_stride, layer arrays det + i * l_img_stride
int launch_hessian_layer(const int32_t *S, size_t s_img_stride, int stride, int R, int C, const HessianLayer &hl, float *det,
                         float *trace, size_t l_img_stride, int n_images, cudaStream_t s);
int launch_hessian_maxima(const float *d0, const float *d1, const float *d2, const float *tr, size_t l_img_stride, int rows,
                          int cols, int margin, int size, int size_prev, int step, int octave, float threshold, fe_kpoint *out,
                          int cap, uint32_t *count, int n_images, cudaStream_t s);
// std::sort(keypoints, KeypointGreater()) on the device; also the largest keypoint size per image (float bits)
int launch_surf_rank_sort(const fe_kpoint *in, fe_kpoint *out, const uint32_t *count, int cap, uint32_t *max_size_bits, int n_images,
                          cudaStream_t s);
int launch_integral(const Geom &g, const Buffers &b, cudaStream_t s);

// WindowMatcher over a resident sequence (window.cu)
// side 0: left keypoints / descriptors of the landmarks (queryIdx of matches_a); side 1: right descriptors (trainIdx), wkx / wky unused
int launch_gather_landmarks(const Geom &g, int n_frames, const Buffers &b, int side, uint8_t *wdesc, float *wkx, float *wky,
                            uint32_t *wcount, cudaStream_t s);
int launch_triangulate(const Geom &g, int n_frames, const Buffers &b, const double *Q, double *xyz, cudaStream_t s);
int launch_pack_landmarks(const Geom &g, int n_pairs, const Buffers &b, const uint32_t *n_m, const fe_match *matches,
                          fe_kpoint *lkp, fe_kpoint *rkp, uint8_t *ldesc, uint8_t *rdesc, fe_match *out, cudaStream_t s);

// Row table of raster-ordered keypoint lists (match.cu): rowstart[image][r] = first index whose floor(y), clamped to
// [0, h], is >= r, for r in [0, h + 1] (rowstart[h + 1] = n); h = Geom::rs_h.  The banded matchers read their candidate range from it instead of searching.
int launch_rowstart(const Geom &g, const Buffers &b, const uint32_t *counts, cudaStream_t s);
// Candidate index range [lo, hi) of a query whose allowed trains have raw y within `reach` of c (one row of slack on both
// sides for float rounding; the caller still applies the exact predicate).
__device__ __forceinline__ void band_range(const int *rows, int h, float c, float reach, int &lo, int &hi) {
    const int r0 = (int)floorf(c - reach) - 1, r1 = (int)floorf(c + reach) + 2;
    lo = rows[min(max(r0, 0), h)];          // rows 0 and h also hold the keypoints whose y was clamped into the table
    hi = rows[min(max(r1, 1), h + 1)];
}

// Trim the table's candidate range to the exact allowed run with warp-wide ballots from both ends (keypoint y is
// non-decreasing, so `started(t)` -- t is at or after the first allowed index -- and `past(t)` -- t is beyond the last one --
// are monotone false -> true).  One round per end unless a row holds more than 32 keypoints in the slack.
template <typename PA, typename PB>
__device__ __forceinline__ void band_trim(int &lo, int &hi, int lane, PA started, PB past) {
    while (lo < hi) {
        const int t = lo + lane;
        const uint32_t m = __ballot_sync(0xffffffffu, t < hi ? started(t) : true);
        if (m) { lo = min(lo + __ffs(m) - 1, hi); break; }
        lo += 32;
    }
    while (hi > lo) {
        const int t = hi - 32 + lane;
        const uint32_t m = __ballot_sync(0xffffffffu, t >= lo ? past(t) : false);
        if (m == 0) break;
        hi = hi - 32 + __ffs(m) - 1;                       // first index of this window that is past the run
        if (m != 0xffffffffu) break;                       // (all 32 past: look at the window before it)
    }
}

struct MatchParams {
    int mask;                  // fe_mask_kind for the (best, second) pair
    float epi_threshold, q_off, t_off;
    float half_w, half_h;
    int h2 = 0;                // 1: cv::NORM_HAMMING2 (two-bit symbols, ORB WTA_K 3 / 4)
};
// unmasked row arg-min + column arg-min (cross-check)
// register-only POPC throughput probe; returns the number of POPCs issued
double launch_popc_peak(int sms, int iters, uint32_t *sink, cudaStream_t s);
int launch_hamming_cross(const Geom &g, int n_pairs, bool h2, const Buffers &b, const uint32_t *counts, cudaStream_t s);
// cross-check + |dy| <= max_dy by band candidates + pruned verification (raster-ordered keypoints on both sides);
// writes match_b / n_b itself (no separate finalize)
// fused_ratio >= 0: mode A's Lowe-ratio finalize (b.best / b.second -> match_a / n_a) rides in the same finalize launch
int launch_hamming_cross_pruned(const Geom &g, int n_pairs, float max_dy, bool have_band, bool use_join, const Buffers &b,
                                const uint32_t *counts, double fused_ratio, cudaStream_t s, cudaStream_t aux = nullptr,
                                cudaEvent_t ev_fork = nullptr, cudaEvent_t ev_join = nullptr);      // aux: side stream for the join
// masked kNN-2; train_sorted = train keypoints are in raster order (enables the banded kernel)
// inner_thr >= 0 (banded path only): additionally produce the cross-check's band candidates b.cx_bestL / b.cx_bestR for
// |dy| <= inner_thr in the same pass
int launch_hamming_knn2(const Geom &g, int n_pairs, const MatchParams &mp, bool train_sorted, const Buffers &b,
                        const uint32_t *counts, float inner_thr, cudaStream_t s);
// 512-bit rows (BRIEF-64, FREAK; match512.cu): desc64 = [n_images][kp_cap][64]; same keys and output arrays as the 256-bit kernels
int launch_hamming512_knn2(const Geom &g, int n_pairs, const MatchParams &mp, const uint8_t *desc64, const Buffers &b,
                           const uint32_t *counts, cudaStream_t s);
int launch_hamming512_cross(const Geom &g, int n_pairs, const uint8_t *desc64, const Buffers &b, const uint32_t *counts, cudaStream_t s);
// float descriptors (b.fdesc, 128-float rows; dim = 64 or 128), keys (float bits of d^2 << 32 | index)
int launch_l2_match(const Geom &g, int n_pairs, int dim, const MatchParams &mp, bool masked, bool all, const Buffers &b,
                    const uint32_t *counts, cudaStream_t s);
int launch_l2_band(const Geom &g, int n_pairs, int dim, const MatchParams &mp, const Buffers &b, const uint32_t *counts,
                   cudaStream_t s);
int launch_l2_tensor(const Geom &g, int n_pairs, int dim, bool need_second, const Buffers &b, const uint32_t *counts,
                     int phase, cudaStream_t s);
// exact cross-check + |dy| <= max_dy by band candidates + ONE tcgen05 GEMM + verification (l2verify.cu)
int launch_l2_band_cand(const Geom &g, int n_pairs, int dim, const MatchParams &mp, float inner, bool write_knn, const Buffers &b,
                        const uint32_t *counts, cudaStream_t s);
int launch_l2_verify(const Geom &g, int n_pairs, int dim, const Buffers &b, const uint32_t *counts, int phase, cudaStream_t s);
size_t l2_verify_list_entries(int kp_cap);
int launch_l2_finalize_ratio(const Geom &g, int n_pairs, double ratio, const Buffers &b, const uint32_t *counts, cudaStream_t s);
int launch_l2_finalize_cross(const Geom &g, int n_pairs, float max_dy, const Buffers &b, const uint32_t *counts, cudaStream_t s);
int launch_finalize_ratio(const Geom &g, int n_pairs, double ratio, const Buffers &b,
                          const uint32_t *counts, cudaStream_t s);
int launch_finalize_cross(const Geom &g, int n_pairs, float max_dy, const Buffers &b,
                          const uint32_t *counts, cudaStream_t s);
// liveGraph: mutual matches of the left descriptors AND of the right descriptors with the same partner (algorithm.py:1160-1190)
int launch_finalize_cross_both(const Geom &g, int n_pairs, const uint32_t *counts, const uint32_t *abL, const uint32_t *cbL,
                               const uint32_t *abR, const uint32_t *cbR, fe_match *out, uint32_t *n_out, cudaStream_t s);

}  // namespace fe
