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
    {
        for (int r = 0; r < SC_ROWS; ++r) {
            const int y = y0 - 1 + r;
            const bool row_ok = y >= 3 && y < g.h - 3;
            for (int x = tid; x < g.pitch; x += FAST_THREADS) {
                int s = 0;
                if (row_ok && x >= 3 && x < g.w - 3)
                    s = fast_pixel<PS>(s_in + (r + 3) * sp + XPAD + x, sp, p.threshold, p.nonmax != 0);
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

int launch_fast(const Geom &g, const DetectParams &p, const Buffers &b, cudaStream_t s) {
    const int sp = g.pitch + 2 * XPAD;
    const size_t smem = (size_t)(IN_ROWS + SC_ROWS) * sp;
    dim3 grid(g.n_strips, g.n_images);
    cudaMemsetAsync(b.hist, 0, sizeof(uint32_t) * 256 * g.n_images, s);
#define FE_LAUNCH_FAST(PS)                                                                        \
    do {                                                                                          \
        cudaFuncSetAttribute(fast_strip_kernel<PS>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem);                                                          \
        fast_strip_kernel<PS><<<grid, FAST_THREADS, smem, s>>>(b.img, g, p, b.slab, b.strip_raw,  \
                                                              b.hist);                            \
    } while (0)
    if (p.ps == 16) FE_LAUNCH_FAST(16);
    else if (p.ps == 12) FE_LAUNCH_FAST(12);
    else FE_LAUNCH_FAST(8);
#undef FE_LAUNCH_FAST
    return 1;
}

}  // namespace fe
