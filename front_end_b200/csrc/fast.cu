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
#include "fe_internal.cuh"

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
    uint32_t *out = slab + ((size_t)image * g.n_strips + strip) * g.slab_cap;
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
    uint32_t *out = slab + ((size_t)image * g.n_strips + strip) * g.slab_cap;
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

int launch_fast(const Geom &g, const DetectParams &p, const Buffers &b, cudaStream_t s) {
    cudaMemsetAsync(b.hist, 0, sizeof(uint32_t) * 256 * g.n_images, s);
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
