// FAST corner detection + score + 3x3 NMS + ordered candidate emission + response histogram.
//
// Replaces cv::FASTX(img, kps, thr, nms, type) as called at /root/reference
// src/live_stereo.cpp:293,306, src/utils.cpp:30, src/front_end/features.py:62-67,595-621 and the
// FAST-9_16 stage inside cv::ORB::detect (features.py:378-387).  Semantics: SURVEY.md A.1
// (ring offsets, class bits, OpenCV's literal-index quick test for 12/8 rings, score =
// max(t, max_arc min d, max_arc min -d) - 1, strict 3x3 NMS, 3-px border, raster order).
//
// Layout: one CTA owns a STRIP of 8 image rows at full width.  The strip plus a 4-row/col halo is
// staged once in shared memory with 16-byte loads; scores for the strip +-1 row live in shared
// memory only (never written to HBM); surviving corners are emitted IN RASTER ORDER into the
// strip's slab (score<<24 | ylocal<<16 | x) via one block-wide scan, so strip order x slab order is
// the canonical raster order with no sort.  A 256-bin histogram of responses inside the ORB border
// is accumulated for the top-N cut (select.cu).
#include <cstdlib>

#include "fe_internal.cuh"
#include <type_traits>

namespace fe {

constexpr int FAST_THREADS = 256;
constexpr int IN_ROWS = STRIP_ROWS + 8;    // strip + 4 above + 4 below (3 ring + 1 NMS)
constexpr int SC_ROWS = STRIP_ROWS + 2;    // scores for strip +- 1 row
constexpr int XPAD = 16;                   // smem column of image x = 0 (keeps 16-B alignment)

template <int PS>
struct Ring;
template <>
struct Ring<16> {
    static __device__ __forceinline__ void off(int k, int &dx, int &dy) {
        const int X[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
        const int Y[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
        dx = X[k]; dy = Y[k];
    }
};
template <>
struct Ring<12> {
    static __device__ __forceinline__ void off(int k, int &dx, int &dy) {
        const int X[12] = {0, 1, 2, 2, 2, 1, 0, -1, -2, -2, -2, -1};
        const int Y[12] = {2, 2, 1, 0, -1, -2, -2, -2, -1, 0, 1, 2};
        dx = X[k]; dy = Y[k];
    }
};
template <>
struct Ring<8> {
    static __device__ __forceinline__ void off(int k, int &dx, int &dy) {
        const int X[8] = {0, 1, 1, 1, 0, -1, -1, -1};
        const int Y[8] = {1, 1, 0, -1, -1, -1, 0, 1};
        dx = X[k]; dy = Y[k];
    }
};

// Any circular run of ARC set bits in the PS-bit mask m?
template <int PS, int ARC>
__device__ __forceinline__ bool has_arc(uint32_t m) {
    uint32_t e = m | (m << PS);        // 2*PS <= 32 bits
    uint32_t r = e;
#pragma unroll
    for (int i = 1; i < ARC; ++i) r &= (e >> i);
    return r != 0;
}

// FAST response of the pixel at c (shared memory, row pitch sp); 0 if not a corner.
// With NMS off the response is not needed and 1 is returned for corners.
template <int PS>
__device__ __forceinline__ int fast_pixel(const uint8_t *c, int sp, int t, bool want_score) {
    constexpr int K = PS / 2, ARC = K + 1;
    const int v = c[0];
    int d[PS];
    uint32_t dark = 0, bright = 0;
#pragma unroll
    for (int k = 0; k < PS; ++k) {
        int dx, dy;
        Ring<PS>::off(k, dx, dy);
        d[k] = v - (int)c[dy * sp + dx];
        dark |= (uint32_t)(d[k] > t) << k;
        bright |= (uint32_t)(d[k] < -t) << k;
    }
    bool qd = true, qb = true;
    if (PS != 16) {
        // OpenCV's quick test uses the 16-ring's literal index pairs for every ring size.
        const int A[8] = {0, 2, 4, 6, 1, 3, 5, 7};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int a = A[i] % PS, b = (A[i] + 8) % PS;
            qd = qd && (((dark >> a) | (dark >> b)) & 1u);
            qb = qb && (((bright >> a) | (bright >> b)) & 1u);
        }
    }
    const bool corner = (qd && has_arc<PS, ARC>(dark)) || (qb && has_arc<PS, ARC>(bright));
    if (!corner) return 0;
    if (!want_score) return 1;
    // score: max over the PS cyclic arcs of min(d) and of min(-d) == -max(d)
    int lo3[PS], hi3[PS];
#pragma unroll
    for (int k = 0; k < PS; ++k) {
        lo3[k] = __vimin3_s32(d[k], d[(k + 1) % PS], d[(k + 2) % PS]);
        hi3[k] = __vimax3_s32(d[k], d[(k + 1) % PS], d[(k + 2) % PS]);
    }
    int best_pos = -512, best_neg = 512;   // max of window-min ; min of window-max
#pragma unroll
    for (int k = 0; k < PS; ++k) {
        int wmin, wmax;
        if (ARC == 9) {
            wmin = __vimin3_s32(lo3[k], lo3[(k + 3) % PS], lo3[(k + 6) % PS]);
            wmax = __vimax3_s32(hi3[k], hi3[(k + 3) % PS], hi3[(k + 6) % PS]);
        } else if (ARC == 7) {
            wmin = __vimin3_s32(lo3[k], lo3[(k + 3) % PS], d[(k + 6) % PS]);
            wmax = __vimax3_s32(hi3[k], hi3[(k + 3) % PS], d[(k + 6) % PS]);
        } else {  // ARC == 5
            wmin = min(lo3[k], lo3[(k + 2) % PS]);
            wmax = max(hi3[k], hi3[(k + 2) % PS]);
        }
        best_pos = max(best_pos, wmin);
        best_neg = min(best_neg, wmax);
    }
    return __vimax3_s32(t, best_pos, -best_neg) - 1;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

template <int PS>
__global__ void __launch_bounds__(FAST_THREADS)
fast_strip_kernel(const uint8_t *__restrict__ img, Geom g, DetectParams p,
                  uint32_t *__restrict__ slab, uint32_t *__restrict__ strip_raw,
                  uint32_t *__restrict__ hist) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int sp = g.pitch + 2 * XPAD;                 // shared row pitch (bytes), multiple of 16
    uint8_t *s_in = smem;                              // IN_ROWS x sp
    uint8_t *s_sc = smem + IN_ROWS * sp;               // SC_ROWS x sp
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_warp[FAST_THREADS / 32];
    __shared__ uint32_t s_base;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int strip = blockIdx.x, image = blockIdx.y;
    const int y0 = strip * STRIP_ROWS;
    const uint8_t *src = img + (size_t)image * g.img_stride;

    // ---- stage the strip (+halo) : 16-byte loads, zero outside the image -----------------------
    {
        const int vec_per_row = sp / 16;
        for (int i = tid; i < IN_ROWS * vec_per_row; i += FAST_THREADS) {
            const int r = i / vec_per_row, c = i - r * vec_per_row;
            const int gy = y0 - 4 + r;
            const int gx = c * 16 - XPAD;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (gy >= 0 && gy < g.h && gx >= 0 && gx < g.pitch)
                val = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)gy * g.pitch + gx));
            *reinterpret_cast<uint4 *>(s_in + r * sp + c * 16) = val;
        }
        for (int i = tid; i < 256; i += FAST_THREADS) s_hist[i] = 0;
        if (tid == 0) s_base = 0;
    }
    __syncthreads();

    // ---- scores for rows y0-1 .. y0+STRIP_ROWS (shared memory only) ----------------------------
    const int threshold = p.thr_img ? p.thr_img[image] : p.threshold;   // per-cell thresholds of the grid detector
    {
        for (int r = 0; r < SC_ROWS; ++r) {
            const int y = y0 - 1 + r;
            const bool row_ok = y >= 3 && y < g.h - 3;
            for (int x = tid; x < g.pitch; x += FAST_THREADS) {
                int s = 0;
                if (row_ok && x >= 3 && x < g.w - 3)
                    s = fast_pixel<PS>(s_in + (r + 3) * sp + XPAD + x, sp, threshold, p.nonmax != 0);
                s_sc[r * sp + XPAD + x] = (uint8_t)s;
            }
        }
        // zero guard columns x = -1 and x = pitch (read by the NMS of x = 0 / x = pitch-1)
        for (int r = tid; r < SC_ROWS; r += FAST_THREADS) {
            s_sc[r * sp + XPAD - 1] = 0;
            s_sc[r * sp + XPAD + g.pitch] = 0;
        }
    }
    __syncthreads();

    // ---- NMS + ordered emission -------------------------------------------------------------------
    const int rows_here = min(STRIP_ROWS, g.h - y0);
    const int total = rows_here * g.w;
    uint32_t *out = slab + (size_t)image * g.slab_img + (size_t)strip * g.slab_cap;
    for (int pass0 = 0; pass0 < total; pass0 += 64 * FAST_THREADS) {
        const int span = min(total - pass0, 64 * FAST_THREADS);
        const int chunk = div_up(span, FAST_THREADS);        // <= 64 pixels per thread, contiguous
        const int first = pass0 + tid * chunk;
        const int last = min(first + chunk, pass0 + span);
        unsigned long long keep = 0ull;
        int r = first / g.w, x = first - r * g.w;
        for (int i = first; i < last; ++i, ++x) {
            if (x == g.w) { x = 0; ++r; }
            const uint8_t *c = s_sc + (r + 1) * sp + XPAD + x;
            const int s = c[0];
            bool k = s > 0;
            if (k && p.nonmax) {
                k = s > c[-1] && s > c[1] && s > c[-sp - 1] && s > c[-sp] && s > c[-sp + 1] &&
                    s > c[sp - 1] && s > c[sp] && s > c[sp + 1];
            }
            if (k) keep |= 1ull << (i - first);
        }
        const uint32_t cnt = __popcll(keep);
        const uint32_t incl = warp_incl_scan(cnt, lane);
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < FAST_THREADS / 32; ++w) {
            const uint32_t c = s_warp[w];
            if (w < wid) wbase += c;
            tot += c;
        }
        uint32_t pos = s_base + wbase + incl - cnt;
        while (keep) {
            const int b = __ffsll((long long)keep) - 1;
            keep &= keep - 1;
            const int i = first + b;
            const int rr = i / g.w, xx = i - rr * g.w;
            const int y = y0 + rr;
            const uint32_t s = p.nonmax ? s_sc[(rr + 1) * sp + XPAD + xx] : 0u;
            if (pos < (uint32_t)g.slab_cap) out[pos] = (s << 24) | ((uint32_t)rr << 16) | (uint32_t)xx;
            ++pos;
            if (xx >= p.edge && xx < g.w - p.edge && y >= p.edge && y < g.h - p.edge)
                atomicAdd(&s_hist[s], 1u);
        }
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    if (tid == 0) strip_raw[image * g.n_strips + strip] = min(s_base, (uint32_t)g.slab_cap);
    for (int i = tid; i < 256; i += FAST_THREADS) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[image * 256 + i], c);
    }
}

// =================================================================================================
// FAST-9_16 fast path (the ring ORB uses): packed 16-bit SIMD, two pixels per lane-op.
//
// For the 16-ring the quick test is implied by the arc test, so corner-ness and response both follow
// from  s = max( max_arc min_k (v - p_k),  max_arc min_k (p_k - v) )  over the 16 cyclic 9-arcs:
// corner <=> s > t, response = s - 1.  Because v is constant over an arc,
//     max_arc min_k (v - p_k) = v - min_arc max_k p_k      and
//     max_arc min_k (p_k - v) = max_arc min_k p_k - v,
// so the kernel needs only sliding 9-window maxima and minima of the raw ring values: no per-position
// subtraction or compare.  Ring values are held as u16 pairs (two horizontally adjacent pixels per
// 32-bit register) and the windows are built from 3-input VIMNMX3.S16x2: 16 + 16 ops per polarity.
//
// Stage A widens the tile (+halo) to u16 pairs in shared memory twice, once aligned to even and once
// to odd x, so that every ring offset is one aligned LDS.32.  Stage B computes s'' = max(s - t, 0)
// for a 128 x 32 region.  Stage C is the strict 3x3 NMS on the 16-bit scores (non-corners are 0, so
// "greater than all 8 neighbours" is the same test OpenCV makes on score-1 vs 0) and writes the
// surviving s'' as one byte per pixel to a response map in HBM.  fast16_emit_kernel then scans the
// map strip by strip and emits candidates in raster order + the response histogram, exactly like
// the generic kernel above.
constexpr int FT_OW = 124, FT_OH = 30;        // pixels written per tile
constexpr int FT_CW = 128, FT_CH = 32;        // score region: x0-2 .. x0+125, y0-1 .. y0+30
constexpr int FT_IH = 38;                     // staged rows y0-4 .. y0+33
constexpr int FT_IWORDS = 35;                 // staged bytes x0-8 .. x0+131 as 35 words
constexpr int FT_PW = 72;                     // u16-pair words per staged row (70 used)
constexpr int FT_THREADS = 256;

__device__ __forceinline__ uint32_t max3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t min3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }

__global__ void __launch_bounds__(FT_THREADS)
fast16_tile_kernel(const uint8_t *__restrict__ img, uint8_t *__restrict__ respmap, Geom g, int threshold,
                   int nonmax) {
    __shared__ uint32_t sE[FT_IH][FT_PW];     // sE[r][j] = pixels (2j, 2j+1) of the staged row
    __shared__ uint32_t sO[FT_IH][FT_PW];     // sO[r][j] = pixels (2j+1, 2j+2)
    __shared__ uint32_t sS[FT_CH][FT_CW / 2]; // s'' pairs
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * FT_OW, y0 = blockIdx.y * FT_OH, image = blockIdx.z;
    const uint8_t *src = img + (size_t)image * g.img_stride;

    // ---- A: stage + widen ------------------------------------------------------------------------
    for (int i = tid; i < FT_IH * FT_IWORDS; i += FT_THREADS) {
        const int r = i / FT_IWORDS, wi = i - r * FT_IWORDS;
        const int gy = y0 - 4 + r, gx = x0 - 8 + 4 * wi;
        uint32_t w0 = 0, w1 = 0;
        if (gy >= 0 && gy < g.h) {
            const uint8_t *row = src + (size_t)gy * g.pitch;
            if (gx >= 0 && gx < g.pitch) w0 = __ldg(reinterpret_cast<const uint32_t *>(row + gx));
            if (gx + 4 >= 0 && gx + 4 < g.pitch) w1 = __ldg(reinterpret_cast<const uint32_t *>(row + gx + 4));
        }
        sE[r][2 * wi] = __byte_perm(w0, 0, 0x4140);
        sE[r][2 * wi + 1] = __byte_perm(w0, 0, 0x4342);
        sO[r][2 * wi] = __byte_perm(w0, 0, 0x4241);
        sO[r][2 * wi + 1] = __byte_perm(__byte_perm(w0, w1, 0x0043), 0, 0x4140);
    }
    __syncthreads();

    // ---- B: s'' for the 128 x 32 score region, one warp per 64-pixel row segment --------------------
    const uint32_t bias = 0x01000100u;
    const uint32_t sub = (uint32_t)(0x10000 - (256 + threshold)) * 0x00010001u;   // -(256 + t) per lane
    for (int task = warp; task < FT_CH * 2; task += FT_THREADS / 32) {
        const int r = task >> 1, c = ((task & 1) << 5) + lane;
        const uint32_t *e = &sE[r + 3][c + 3], *o = &sO[r + 3][c + 3];
        uint32_t p[16];
        p[0] = e[3 * FT_PW];       p[1] = o[3 * FT_PW];       p[2] = e[2 * FT_PW + 1];   p[3] = o[FT_PW + 1];
        p[4] = o[1];               p[5] = o[-FT_PW + 1];      p[6] = e[-2 * FT_PW + 1];  p[7] = o[-3 * FT_PW];
        p[8] = e[-3 * FT_PW];      p[9] = o[-3 * FT_PW - 1];  p[10] = e[-2 * FT_PW - 1]; p[11] = o[-FT_PW - 2];
        p[12] = o[-2];             p[13] = o[FT_PW - 2];      p[14] = e[2 * FT_PW - 1];  p[15] = o[3 * FT_PW - 1];
        const uint32_t v = e[0];
        // The 9-arcs starting at 2j and 2j+1 share the 8 ring pixels 2j+1 .. 2j+8, so
        //   max(min arc(2j), min arc(2j+1)) = min( min(p[2j+1 .. 2j+8]), max(p[2j], p[2j+9]) )
        // and the 8-windows at odd starts come from pair and quad minima at odd positions: 36 operations per polarity
        // instead of 40 for sixteen separate 9-windows.
        uint32_t A, B;
        {   // A = min over the arcs of the arc maximum
            uint32_t w2[8], w4[8], P[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w2[j] = __vmaxs2(p[2 * j + 1], p[(2 * j + 2) & 15]);
#pragma unroll
            for (int j = 0; j < 8; ++j) w4[j] = __vmaxs2(w2[j], w2[(j + 1) & 7]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                P[j] = max3_s16x2(w4[j], w4[(j + 2) & 7], __vmins2(p[2 * j], p[(2 * j + 9) & 15]));
            A = min3_s16x2(min3_s16x2(P[0], P[1], P[2]), min3_s16x2(P[3], P[4], P[5]), __vmins2(P[6], P[7]));
        }
        {   // B = max over the arcs of the arc minimum
            uint32_t w2[8], w4[8], P[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w2[j] = __vmins2(p[2 * j + 1], p[(2 * j + 2) & 15]);
#pragma unroll
            for (int j = 0; j < 8; ++j) w4[j] = __vmins2(w2[j], w2[(j + 1) & 7]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                P[j] = min3_s16x2(w4[j], w4[(j + 2) & 7], __vmaxs2(p[2 * j], p[(2 * j + 9) & 15]));
            B = max3_s16x2(max3_s16x2(P[0], P[1], P[2]), max3_s16x2(P[3], P[4], P[5]), __vmaxs2(P[6], P[7]));
        }
        // 256 + (v - A) and 256 + (B - v): lanes stay in [1, 511], so plain 32-bit adds cannot borrow
        const uint32_t pos = (v | bias) - A, neg = (B | bias) - v;
        const uint32_t sb = __vmaxs2(pos, neg);
        uint32_t s2 = __viaddmax_s16x2_relu(sb, sub, 0u);              // max(s - t, 0) per lane
        const int x = x0 - 2 + 2 * c, y = y0 - 1 + r;
        const bool row_ok = y >= 3 && y < g.h - 3;
        uint32_t mask = 0;
        if (row_ok && x >= 3 && x < g.w - 3) mask |= 0x0000FFFFu;
        if (row_ok && x + 1 >= 3 && x + 1 < g.w - 3) mask |= 0xFFFF0000u;
        sS[r][c] = s2 & mask;
    }
    __syncthreads();

    // ---- C: strict 3x3 NMS, four pixels per thread, byte map out -----------------------------------
    uint8_t *dst = respmap + (size_t)image * g.img_stride;
    for (int i = tid; i < FT_OH * (FT_OW / 4); i += FT_THREADS) {
        const int orow = i / (FT_OW / 4), ow = i - orow * (FT_OW / 4);
        const int y = y0 + orow, x = x0 + 4 * ow;
        if (y >= g.h || x >= g.pitch) continue;
        const int r = orow + 1, c1 = 1 + 2 * ow;
        uint32_t outp[2];
        if (nonmax) {
            uint32_t w[3][4];
#pragma unroll
            for (int dr = 0; dr < 3; ++dr)
#pragma unroll
                for (int j = 0; j < 4; ++j) w[dr][j] = sS[r - 1 + dr][c1 - 1 + j];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t lp[3], rp[3];
#pragma unroll
                for (int dr = 0; dr < 3; ++dr) {
                    lp[dr] = __byte_perm(w[dr][q], w[dr][q + 1], 0x5432);       // pixels (x-1, x)
                    rp[dr] = __byte_perm(w[dr][q + 1], w[dr][q + 2], 0x5432);   // pixels (x+1, x+2)
                }
                const uint32_t cc = w[1][q + 1];
                const uint32_t up = max3_s16x2(lp[0], w[0][q + 1], rp[0]);
                const uint32_t dn = max3_s16x2(lp[2], w[2][q + 1], rp[2]);
                const uint32_t m = max3_s16x2(up, dn, __vmaxs2(lp[1], rp[1]));
                // k = max(c - m, 0); keep c where k > 0:  min_u16(c, k << 8)  (c <= 255 < 256 <= k << 8)
                const uint32_t k = __viaddmax_s16x2_relu((cc | bias) - m, 0xFF00FF00u, 0u);
                outp[q] = __vminu2(cc, k << 8);
            }
        } else {
            outp[0] = sS[r][c1];
            outp[1] = sS[r][c1 + 1];
        }
        *reinterpret_cast<uint32_t *>(dst + (size_t)y * g.pitch + x) = __byte_perm(outp[0], outp[1], 0x6420);
    }
}

// Scan the response map of one 8-row strip in raster order: ordered candidate emission + histogram.
__global__ void __launch_bounds__(FAST_THREADS)
fast16_emit_kernel(const uint8_t *__restrict__ respmap, Geom g, DetectParams p, uint32_t *__restrict__ slab,
                   uint32_t *__restrict__ strip_raw, uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_warp[FAST_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int strip = blockIdx.x, image = blockIdx.y;
    const int y0 = strip * STRIP_ROWS;
    const int rows_here = min(STRIP_ROWS, g.h - y0);
    const int vec_per_row = g.pitch / 16;
    const int items = rows_here * vec_per_row;
    const uint8_t *src = respmap + (size_t)image * g.img_stride + (size_t)y0 * g.pitch;
    uint32_t *out = slab + (size_t)image * g.slab_img + (size_t)strip * g.slab_cap;
    for (int i = tid; i < 256; i += FAST_THREADS) s_hist[i] = 0;
    __syncthreads();
    uint32_t base = 0;
    for (int i0 = 0; i0 < items; i0 += FAST_THREADS) {
        const int i = i0 + tid;
        uint4 v = make_uint4(0, 0, 0, 0);
        int rr = 0, xx0 = 0;
        if (i < items) {
            rr = i / vec_per_row;
            xx0 = (i - rr * vec_per_row) * 16;
            v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)rr * g.pitch + xx0));
        }
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
        uint32_t set = 0;                        // bit j <=> byte j of the 16-byte vector is a surviving corner
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t t = ((((wv[j] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | wv[j]) >> 7) & 0x01010101u;
            set |= ((t | (t >> 7) | (t >> 14) | (t >> 21)) & 0xFu) << (4 * j);
        }
        const uint32_t cnt = __popc(set);
        const uint32_t incl = warp_incl_scan(cnt, lane);
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < FAST_THREADS / 32; ++w) {
            const uint32_t c = s_warp[w];
            if (w < wid) wbase += c;
            tot += c;
        }
        uint32_t pos = base + wbase + incl - cnt;
        // one iteration per corner of the lane (not per byte position): the warp runs max-over-lanes(cnt) rounds
        const int y = y0 + rr;
        const bool y_in = y >= p.edge && y < g.h - p.edge;
        while (set) {
            const int j = __ffs(set) - 1;
            set &= set - 1;
            const uint32_t w = (j >> 2) == 0 ? wv[0] : (j >> 2) == 1 ? wv[1] : (j >> 2) == 2 ? wv[2] : wv[3];
            const uint32_t b = __byte_perm(w, 0, 0x4440u | (uint32_t)(j & 3));
            const int xx = xx0 + j;
            const uint32_t s = p.nonmax ? b + (uint32_t)p.threshold - 1u : 0u;
            if (pos < (uint32_t)g.slab_cap) out[pos] = (s << 24) | ((uint32_t)rr << 16) | (uint32_t)xx;
            ++pos;
            if (y_in && xx >= p.edge && xx < g.w - p.edge) atomicAdd(&s_hist[s], 1u);
        }
        base += tot;
        __syncthreads();
    }
    if (tid == 0) strip_raw[image * g.n_strips + strip] = min(base, (uint32_t)g.slab_cap);
    for (int i = tid; i < 256; i += FAST_THREADS) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[image * 256 + i], c);
    }
}

// =================================================================================================
// FAST-9_16 + NMS + raster-ordered emission in ONE kernel (the path ORB and BASELINE configs 1-5 run).
//
// One CTA owns a strip of WIDE_STRIP_ROWS = 30 image rows at full width; a warp owns a column segment of
// 60 pixels (32 lanes x one pixel PAIR, one halo pair on each side) and walks DOWN the strip:
//   * the 7 image rows a score row needs live in a per-warp ring buffer of 8 rows in shared memory, widened to
//     u16 pairs at even and odd alignment (every ring offset is one aligned LDS.32 with an immediate offset: the
//     row loop is unrolled by 8 so ring slots are static).  Each step loads ONE new image row (19 lanes x 4 bytes,
//     prefetched one step ahead), so the image is read once from HBM/L2 plus an 8-row halo per strip;
//   * the score s'' = max(s - t, 0) of the lane's pixel pair is computed exactly as in fast16_tile_kernel
//     (sliding 9-window minima / maxima in VIMNMX3.S16x2: 36 operations per polarity);
//   * the last three score rows stay in REGISTERS: the strict 3x3 NMS is a column max3 + two lane shuffles +
//     one max3, and a surviving pixel (at most one per pair, at most 30 per segment row) is compacted by
//     ballot / popc into the 30-entry slot of (row, segment) of a shared-memory queue (u16: x-in-segment, s'');
//   * after the walk the CTA scans the slot counts in raster order and writes the strip's candidates
//     (score<<24 | ylocal<<16 | x) to the slab in RASTER ORDER, plus the response histogram of the top-N cut.
// Nothing but the image is read and nothing but the candidate list (~7 % of the pixels x 4 B) is written:
// the one-byte response map of the tile kernel (written, then re-read by fast16_emit_kernel) is gone.
constexpr int FS_R = WIDE_STRIP_ROWS;        // output rows per strip; score rows = FS_R + 2 = 32 = 4 x 8
constexpr int FS_SEG = 60;                   // useful pixels per warp segment
constexpr int FS_ROWW = 72;                  // words per ring-buffer row: 36 even-aligned + 36 odd-aligned pairs
constexpr int FS_SLOT = 30;                  // queue entries per (segment, row): NMS keeps <= 1 pixel per pair
constexpr int FS_QROWS = 32;                 // queue rows per segment (FS_R used)
constexpr int FS_MAX_WARPS = 12;
constexpr int FS_MAX_PITCH = 2048;           // shared-memory budget (queue = 32 rows x pitch / 2 entries x 2 B)

struct FastStripSmem {                       // byte offsets into dynamic shared memory
    int ring, queue, cnt, hist, wsum, total;
};
__host__ __device__ inline FastStripSmem fast_strip_smem(int nwarps, int nseg) {
    FastStripSmem m;
    m.ring = 0;
    m.queue = m.ring + nwarps * 8 * FS_ROWW * 4;
    m.cnt = m.queue + nseg * FS_QROWS * FS_SLOT * 2;       // u16 entries [seg][row][30]
    m.hist = m.cnt + nseg * FS_QROWS;                      // u8 counts [seg][row]
    m.wsum = m.hist + 256 * 4;
    m.total = m.wsum + (FS_MAX_WARPS + 1) * 4;
    return m;
}

// FMA = true: pixels are staged as the fp16 bit patterns 0x6400 | p (the value 1024 + p, exact), which order like the
// integers they hold, so the VIMNMX ladder works on them unchanged; the sixteen first-level (min, max) PAIRS of the
// ladder are then taken on the otherwise idle FMA pipe as r = relu(b - a) (HFMA2.RELU), max = a + r, min = b - r
// (HADD2) -- 3 FMA-pipe operations instead of 2 ALU-pipe ones (tools/pipe_probe.cu: VIMNMX, VIMNMX3 and HMNMX2 all
// issue at 64 lanes / clk / SM on the ALU pipe; HFMA2 / HADD2 run beside them on the FMA pipe).
__device__ __forceinline__ void minmax_fma(uint32_t a, uint32_t b, uint32_t &mn, uint32_t &mx) {
    uint32_t r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0xBC00BC00u), "r"(b));     // relu(b - a)
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(mx) : "r"(a), "r"(r));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(mn) : "r"(b), "r"(r));
}

// max / min alone on the FMA pipe: 2 operations instead of 1 ALU-pipe one (FMA level 2: the second ladder level too)
__device__ __forceinline__ uint32_t max_fma(uint32_t a, uint32_t b) {
    uint32_t r, m;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0xBC00BC00u), "r"(b));
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(r));
    return m;
}
__device__ __forceinline__ uint32_t min_fma(uint32_t a, uint32_t b) {
    uint32_t r, m;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0xBC00BC00u), "r"(b));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(m) : "r"(b), "r"(r));
    return m;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// value the compiler must keep in a register (it otherwise re-derives lane constants and shared-window addresses
// from the special registers inside the row loop: ~14 of ~215 instructions per row in the first version)
#define FS_KEEP(x) asm volatile("" : "+r"(x))

template <int FMA, int MINB>
__global__ void __launch_bounds__(FS_MAX_WARPS * 32, MINB)
fast16_strip_kernel(const uint8_t *__restrict__ img, Geom g, int threshold, int edge, int nseg, int slab_cap,
                    uint32_t *__restrict__ slab, uint32_t *__restrict__ strip_raw, uint32_t *__restrict__ hist) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const FastStripSmem lay = fast_strip_smem(nwarps, nseg);
    uint16_t *s_queue = reinterpret_cast<uint16_t *>(smem + lay.queue);
    uint8_t *s_cnt = smem + lay.cnt;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + lay.hist);
    uint32_t *s_wsum = reinterpret_cast<uint32_t *>(smem + lay.wsum);
    const int strip = blockIdx.x, image = blockIdx.y;
    const int y0 = strip * FS_R;
    const uint8_t *src = img + (size_t)image * g.img_stride;

    for (int i = tid; i < nseg * FS_QROWS / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(s_cnt)[i] = 0;
    for (int i = tid; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();

    const uint32_t FULL = 0xffffffffu;
    const uint32_t bias = 0x01000100u;
    const uint32_t sub = (uint32_t)(0x10000 - (256 + threshold)) * 0x00010001u;   // -(256 + t) per 16-bit lane
    uint32_t keepmask = (lane >= 1 && lane <= 30) ? 0x01000100u : 0u;              // lanes 0 / 31 are the halo pairs
    uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t st_lane = lane < 18 ? 1u : 0u;                                        // lanes that store staged words
    uint32_t xin0 = (uint32_t)(2 * lane - 2);                                      // x-in-segment of the pair's left pixel
    // byte offsets into the dynamic shared memory, opaque to the compiler (it keeps them in registers; the accesses
    // below are smem + offset + immediate)
    uint32_t ring_off = (uint32_t)(lay.ring + warp * 8 * FS_ROWW * 4 + 8 * lane);        // this lane's staging stores
    uint32_t e_off = (uint32_t)(lay.ring + warp * 8 * FS_ROWW * 4 + 4 * (3 + lane));      // this lane's centre pair
    uint32_t is_lane0 = lane == 0 ? 1u : 0u;
    uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    FS_KEEP(keepmask); FS_KEEP(lt_mask); FS_KEEP(st_lane); FS_KEEP(xin0); FS_KEEP(ring_off); FS_KEEP(e_off); FS_KEEP(is_lane0);
    FS_KEEP(smem_base);
    uint32_t *const ring_p = reinterpret_cast<uint32_t *>(smem + ring_off);
    const uint32_t *const E = reinterpret_cast<const uint32_t *>(smem + e_off);
    // score rows with 3 <= y < h - 3 (y = y0 - 1 + r) and output rows with y0 + orow < h, as unsigned ranges of r
    const int r_lo = max(0, 4 - y0), r_hi = min(FS_R + 2, g.h - 3 - (y0 - 1));
    const uint32_t r_span = (uint32_t)max(r_hi - r_lo, 0);
    const uint32_t o_span = (uint32_t)min(FS_R, g.h - y0);
    const int hmax = g.h - 1;
    const uint32_t pitch = (uint32_t)g.pitch;

    // Strips whose 39 staged rows and 32 score rows all lie inside the image (22 of the 24 strips of a 720-row image) take a
    // specialised walk: no row clamps, no row-range tests, a running row pointer -- fewer instructions and, more important,
    // fewer live registers (the general walk spills / re-derives lane constants at 56 registers).
    const bool interior = y0 >= 4 && y0 + FS_R + 5 <= g.h;
    auto walk = [&](auto interior_tag) {
    constexpr bool INT = decltype(interior_tag)::value;
    for (int seg = warp; seg < nseg; seg += nwarps) {
        const int xs = seg * FS_SEG;
        const int x = xs - 2 + 2 * lane;            // left pixel of this lane's pair
        uint32_t xmask = 0;
        if (x >= 3 && x < g.w - 3) xmask |= 0x0000FFFFu;
        if (x + 1 >= 3 && x + 1 < g.w - 3) xmask |= 0xFFFF0000u;
        // staging: lanes 0 .. 18 own the image word at x = xs - 8 + 4 * lane of every row; staged row q <-> image row
        // y0 - 4 + q.  Rows and words outside the image are not zero-filled but CLAMPED to the nearest row / word inside it:
        // they only ever feed scores that the border masks (xmask, r_lo / r_span) zero, and an unconditional load leaves
        // nothing between the load and its use one row later that waits for it (a predicated load compiles to LDG + a
        // select that stalls on the long scoreboard right away: 29 % of the warp samples of the first version).
        uint32_t xoff = (uint32_t)min(max(xs - 8 + 4 * lane, 0), g.pitch - 4);
        FS_KEEP(xoff);
        const uint8_t *rp = src + ((uint32_t)(INT ? y0 - 4 : 0) * pitch + xoff);      // INT: rows are loaded in order, one pitch apart
        auto load_row = [&](int q) -> uint32_t {      // row index clamp(y, 0, h - 1) is ONE instruction (VIMNMX.RELU)
            if constexpr (INT) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(rp));
                rp += pitch;
                return w;
            } else {
                const uint32_t yi = (uint32_t)__vimin_s32_relu(y0 - 4 + q, hmax);
                return __ldg(reinterpret_cast<const uint32_t *>(src + (yi * pitch + xoff)));      // 32-bit offset inside one image
            }
        };
        auto store_row = [&](int slot, uint32_t w) {
            const uint32_t wn = __shfl_down_sync(FULL, w, 1);
            const uint32_t hib = FMA ? 0x64646464u : 0u;                                          // high byte of every u16
            const uint32_t e0 = __byte_perm(w, hib, 0x4140), e1 = __byte_perm(w, hib, 0x4342);    // pixels (0,1) (2,3)
            const uint32_t en = __byte_perm(wn, hib, 0x4140);                                     // pixels (4,5)
            const uint32_t o0 = __funnelshift_r(e0, e1, 16), o1 = __funnelshift_r(e1, en, 16);    // (1,2) (3,4)
            if (st_lane) {
                *reinterpret_cast<uint2 *>(ring_p + slot * FS_ROWW) = make_uint2(e0, e1);
                *reinterpret_cast<uint2 *>(ring_p + slot * FS_ROWW + 36) = make_uint2(o0, o1);
            }
        };
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 6; ++q) store_row(q, load_row(q));
        uint32_t wpre = load_row(6);
        uint32_t S0 = 0, S1 = 0;                    // score rows r - 2, r - 1
        // queue slot / counter of (seg, output row rb - 2): the row offset k is an immediate
        uint32_t qrow_off = smem_base + (uint32_t)(lay.queue + (seg * FS_QROWS - 2) * (FS_SLOT * 2));
        uint32_t crow_off = smem_base + (uint32_t)(lay.cnt + seg * FS_QROWS - 2);
        for (int rb = 0; rb < FS_R + 2; rb += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = rb + k;               // score row r <-> image row y0 - 1 + r; needs staged rows r .. r + 6
                store_row((k + 6) & 7, wpre);
                wpre = load_row(r + 7);
                __syncwarp();
#define FS_E(dy, j) E[((k + 3 + (dy)) & 7) * FS_ROWW + (j)]
#define FS_O(dy, j) E[((k + 3 + (dy)) & 7) * FS_ROWW + 36 + (j)]
                uint32_t p[16];
                p[0] = FS_E(3, 0);    p[1] = FS_O(3, 0);    p[2] = FS_E(2, 1);    p[3] = FS_O(1, 1);
                p[4] = FS_O(0, 1);    p[5] = FS_O(-1, 1);   p[6] = FS_E(-2, 1);   p[7] = FS_O(-3, 0);
                p[8] = FS_E(-3, 0);   p[9] = FS_O(-3, -1);  p[10] = FS_E(-2, -1); p[11] = FS_O(-1, -2);
                p[12] = FS_O(0, -2);  p[13] = FS_O(1, -2);  p[14] = FS_E(2, -1);  p[15] = FS_O(3, -1);
                const uint32_t v = FS_E(0, 0);
#undef FS_E
#undef FS_O
                uint32_t A, B;
                {
                    // first level: (min, max) of the ring pairs (2j+1, 2j+2) and (2j, 2j+9) -- shared by both polarities
                    uint32_t w2x[8], w2n[8], ex[8], en_[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (FMA && !(FMA == 4 && j >= 4) && !(FMA == 5 && j >= 2)) {
                            minmax_fma(p[2 * j + 1], p[(2 * j + 2) & 15], w2n[j], w2x[j]);
                            minmax_fma(p[2 * j], p[(2 * j + 9) & 15], en_[j], ex[j]);
                        } else if (FMA) {           // FMA 4 / 5: 12 / 10 of the 16 first-level pairs on the FMA pipe
                            minmax_fma(p[2 * j + 1], p[(2 * j + 2) & 15], w2n[j], w2x[j]);
                            ex[j] = __vmaxs2(p[2 * j], p[(2 * j + 9) & 15]);      en_[j] = __vmins2(p[2 * j], p[(2 * j + 9) & 15]);
                        } else {
                            w2x[j] = __vmaxs2(p[2 * j + 1], p[(2 * j + 2) & 15]); w2n[j] = __vmins2(p[2 * j + 1], p[(2 * j + 2) & 15]);
                            ex[j] = __vmaxs2(p[2 * j], p[(2 * j + 9) & 15]);      en_[j] = __vmins2(p[2 * j], p[(2 * j + 9) & 15]);
                        }
                    }
                    // A = min over the arcs of the arc maximum, B = max over the arcs of the arc minimum (see fast16_tile_kernel
                    // for the pairing of the arcs that share eight ring pixels)
                    uint32_t w4[8], P[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) w4[j] = FMA >= 2 ? max_fma(w2x[j], w2x[(j + 1) & 7]) : __vmaxs2(w2x[j], w2x[(j + 1) & 7]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) P[j] = max3_s16x2(w4[j], w4[(j + 2) & 7], en_[j]);
                    A = min3_s16x2(min3_s16x2(P[0], P[1], P[2]), min3_s16x2(P[3], P[4], P[5]), __vmins2(P[6], P[7]));
#pragma unroll
                    for (int j = 0; j < 8; ++j) w4[j] = FMA >= 3 ? min_fma(w2n[j], w2n[(j + 1) & 7]) : __vmins2(w2n[j], w2n[(j + 1) & 7]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) P[j] = min3_s16x2(w4[j], w4[(j + 2) & 7], ex[j]);
                    B = max3_s16x2(max3_s16x2(P[0], P[1], P[2]), max3_s16x2(P[3], P[4], P[5]), __vmaxs2(P[6], P[7]));
                }
                // 256 + (v - A), 256 + (B - v): no borrows; bit 8 of a staged pixel is clear (0x00pp or 0x64pp), so + is |
                const uint32_t pos = v + bias - A, neg = B + bias - v;
                const uint32_t s2 = __viaddmax_s16x2_relu(__vmaxs2(pos, neg), sub, 0u);
                const uint32_t S2 = (INT || (uint32_t)(r - r_lo) < r_span) ? (s2 & xmask) : 0u;
                // ---- strict 3 x 3 NMS of score row r - 1 (output row r - 2 of the strip), all in registers ----
                if (INT ? (k >= 2 || rb > 0) : ((uint32_t)(r - 2) < o_span)) {
                    const uint32_t V = max3_s16x2(S0, S1, S2), Vn = __vmaxs2(S0, S2);
                    const uint32_t Vl = __shfl_up_sync(FULL, V, 1), Vr = __shfl_down_sync(FULL, V, 1);
                    // neighbours of the pair's left pixel: columns x - 1 (hi of Vl), x (Vn lo), x + 1 (V hi); right pixel alike
                    const uint32_t m = max3_s16x2(Vn, __byte_perm(Vl, V, 0x5432), __byte_perm(V, Vr, 0x5432));
                    const uint32_t t = (S1 + 0x00FF00FFu - m) & keepmask;        // bit 8 / 24 set <=> S1 > m in that lane
                    const uint32_t bal = __ballot_sync(FULL, t != 0u);
                    // branch-free: a predicated store per lane + the row count from lane 0 (also when it is 0)
                    const uint32_t hi = t >> 24;                                   // 1: the right pixel of the pair survived
                    // entry = (x-in-segment << 8) | s'': byte 0 or 2 of S1, byte 0 of xin
                    const uint32_t ent = __byte_perm(S1, xin0 + hi, 0x0040u + 2u * hi);
                    if (t) sts_u16(qrow_off + 2u * (uint32_t)__popc(bal & lt_mask) + k * FS_SLOT * 2, ent);
                    if (is_lane0) sts_u8(crow_off + k, (uint32_t)__popc(bal));
                }
                S0 = S1; S1 = S2;
            }
            qrow_off += 8 * FS_SLOT * 2;
            crow_off += 8;
        }
    }
    };
    if (interior) walk(std::true_type{});
    else walk(std::false_type{});
    __syncthreads();

    // ---- raster-order write-out: every thread owns `ipt` consecutive slots of the raster order (row, then segment) -------
    const int rows_here = (int)o_span, nslots = rows_here * nseg;
    const int nthreads = blockDim.x;
    const int ipt = div_up(nslots, nthreads);
    const int t_begin = min(tid * ipt, nslots), t_end = min(t_begin + ipt, nslots);
    const uint32_t inv_nseg = (65536u + (uint32_t)nseg - 1u) / (uint32_t)nseg;       // t / nseg == (t * inv) >> 16 for t < 2048
    int orow0 = (int)(((uint32_t)t_begin * inv_nseg) >> 16), seg0 = t_begin - orow0 * nseg;
    uint32_t mine = 0;
    {
        int orow = orow0, sg = seg0;
        for (int t = t_begin; t < t_end; ++t) {
            mine += s_cnt[sg * FS_QROWS + orow];
            if (++sg == nseg) { sg = 0; ++orow; }
        }
    }
    const uint32_t incl = warp_incl_scan(mine, lane);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
        const uint32_t c = s_wsum[w];
        if (w < warp) wbase += c;
        total += c;
    }
    uint32_t *out = slab + (size_t)image * g.slab_img + (size_t)strip * slab_cap;
    {
        uint32_t o = wbase + incl - mine;
        int orow = orow0, sg = seg0;
        for (int t = t_begin; t < t_end; ++t) {
            const int sl = sg * FS_QROWS + orow;
            const uint32_t c = s_cnt[sl];
            const uint32_t base_rec = ((uint32_t)orow << 16) | (uint32_t)(sg * FS_SEG);
            const int yy = y0 + orow;
            const bool y_in = yy >= edge && yy < g.h - edge;
            for (uint32_t j = 0; j < c; ++j) {
                const uint32_t ent = s_queue[sl * FS_SLOT + j];
                const uint32_t sc = (ent & 0xFFu) + (uint32_t)threshold - 1u;
                const uint32_t rec = (sc << 24) | (base_rec + (ent >> 8));
                if (o < (uint32_t)slab_cap) out[o] = rec;
                ++o;
                const int xx = (int)(rec & 0xFFFFu);
                if (y_in && xx >= edge && xx < g.w - edge) atomicAdd(&s_hist[sc], 1u);
            }
            if (++sg == nseg) { sg = 0; ++orow; }
        }
    }
    __syncthreads();
    if (tid == 0) strip_raw[image * g.n_strips + strip] = min(total, (uint32_t)slab_cap);
    for (int i = tid; i < 256; i += nthreads) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[image * 256 + i], c);
    }
}
#undef FS_KEEP

static bool use_wide_strips(const Geom &g, const DetectParams &p) {
    return p.ps == 16 && !p.thr_img && p.nonmax && g.pitch <= FS_MAX_PITCH;
}

StripView strip_view(const Geom &g, const DetectParams &p) {
    if (use_wide_strips(g, p)) return StripView{FS_R, div_up(g.h, FS_R), g.pitch * FS_R / 4};
    return StripView{STRIP_ROWS, g.n_strips, g.slab_cap};
}

int launch_fast(const Geom &g, const DetectParams &p, const Buffers &b, cudaStream_t s) {
    cudaMemsetAsync(b.hist, 0, sizeof(uint32_t) * 256 * g.n_images, s);
    if (use_wide_strips(g, p)) {
        const StripView sv = strip_view(g, p);
        const int nseg = div_up(g.w, FS_SEG);
        const int rounds = div_up(nseg, FS_MAX_WARPS);
        const int nwarps = div_up(nseg, rounds);
        const int smem = fast_strip_smem(nwarps, nseg).total;
        static int smem_set[64] = {0};          // per device: the attribute belongs to the device's context
        int dev = 0;
        cudaGetDevice(&dev);
        if (smem > smem_set[dev & 63]) {
#define FE_STRIP_ATTR(L, M) cudaFuncSetAttribute(fast16_strip_kernel<L, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
            FE_STRIP_ATTR(0, 3); FE_STRIP_ATTR(1, 3); FE_STRIP_ATTR(2, 3); FE_STRIP_ATTR(3, 3);
            FE_STRIP_ATTR(1, 2); FE_STRIP_ATTR(2, 2); FE_STRIP_ATTR(3, 2); FE_STRIP_ATTR(4, 3); FE_STRIP_ATTR(5, 3);
#undef FE_STRIP_ATTR
            smem_set[dev & 63] = smem;
        }
        // A/B knobs: ladder levels taken on the FMA pipe (0 .. 3), CTAs per SM the kernel is compiled for (3: 56 registers, 2: 85)
        static const int fma_level = [] { const char *e = getenv("FE_FAST_FMA"); return e ? atoi(e) : 1; }();
        static const int minb = [] { const char *e = getenv("FE_FAST_MINB"); return e ? atoi(e) : 3; }();
        dim3 grid(sv.n, g.n_images);
#define FE_LAUNCH_STRIP(L, M) fast16_strip_kernel<L, M><<<grid, nwarps * 32, smem, s>>>(b.img, g, p.threshold, p.edge, nseg, sv.cap, b.slab, b.strip_raw, b.hist)
        if (minb == 2) {
            if (fma_level <= 1) FE_LAUNCH_STRIP(1, 2);
            else if (fma_level == 2) FE_LAUNCH_STRIP(2, 2);
            else FE_LAUNCH_STRIP(3, 2);
        } else {
            if (fma_level <= 0) FE_LAUNCH_STRIP(0, 3);
            else if (fma_level == 1) FE_LAUNCH_STRIP(1, 3);
            else if (fma_level == 2) FE_LAUNCH_STRIP(2, 3);
            else if (fma_level == 3) FE_LAUNCH_STRIP(3, 3);
            else if (fma_level == 4) FE_LAUNCH_STRIP(4, 3);
            else FE_LAUNCH_STRIP(5, 3);
        }
#undef FE_LAUNCH_STRIP
        return 1;
    }
    if (p.ps == 16 && !p.thr_img) {
        dim3 tgrid(div_up(g.pitch, FT_OW), div_up(g.h, FT_OH), g.n_images);   // covers the padded row
        fast16_tile_kernel<<<tgrid, FT_THREADS, 0, s>>>(b.img, b.respmap, g, p.threshold, p.nonmax);
        dim3 egrid(g.n_strips, g.n_images);
        fast16_emit_kernel<<<egrid, FAST_THREADS, 0, s>>>(b.respmap, g, p, b.slab, b.strip_raw, b.hist);
        return 2;
    }
    const int sp = g.pitch + 2 * XPAD;
    const size_t smem = (size_t)(IN_ROWS + SC_ROWS) * sp;
    dim3 grid(g.n_strips, g.n_images);
#define FE_LAUNCH_FAST(PS)                                                                        \
    do {                                                                                          \
        cudaFuncSetAttribute(fast_strip_kernel<PS>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem);                                                          \
        fast_strip_kernel<PS><<<grid, FAST_THREADS, smem, s>>>(b.img, g, p, b.slab, b.strip_raw,  \
                                                              b.hist);                            \
    } while (0)
    if (p.ps == 16) FE_LAUNCH_FAST(16);      // only with per-image thresholds (the 16-ring fast path takes one t)
    else if (p.ps == 12) FE_LAUNCH_FAST(12);
    else FE_LAUNCH_FAST(8);
#undef FE_LAUNCH_FAST
    return 1;
}

}  // namespace fe
