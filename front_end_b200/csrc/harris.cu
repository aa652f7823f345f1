// cv::ORB with scoreType = HARRIS_SCORE (the score field of front_end/setDetector, /root/reference
// src/StereoCamera.cpp:445,462; src/utils.cpp:86-90; the commented-out half of features.py:297).
//
// OpenCV's computeKeyPoints keeps the 2N best FAST corners (ties kept), scores them with HarrisResponses(block 7,
// k = 0.04) on the unblurred level image, then keeps the N best by that score (ties kept) -- per pyramid level.
// Here: the existing histogram cut runs with 2N, `harris_response_kernel` scores the survivors (integer Sobel sums,
// then OpenCV's float expression operation by operation; the library is built with -fmad=false), and
// `harris_retain_kernel` finds the N-th largest float by a 4-pass radix select and compacts IN PLACE, which keeps the
// canonical raster order.  `harris_store_kernel` writes the score into the wire-format keypoints.
#include "fe_internal.cuh"

namespace fe {

__device__ __forceinline__ int reflect101_h(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(128)
harris_response_kernel(const uint8_t *__restrict__ img, Geom g, const uint32_t *__restrict__ n_kp,
                       const uint32_t *__restrict__ kp_key, float harris_k, float scale_sq_sq, float *__restrict__ resp) {
    const int image = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min((int)n_kp[image], g.kp_cap);
    if (i >= n) return;
    const size_t o = (size_t)image * g.kp_cap + i;
    const uint32_t key = kp_key[o];
    const int x0 = key & 0xFFFF, y0 = key >> 16;
    const uint8_t *src = img + (size_t)image * g.img_stride;
    const bool inside = x0 >= 4 && y0 >= 4 && x0 + 4 < g.w && y0 + 4 < g.h;
    int a = 0, b = 0, c = 0;
    // three rows of the 9-wide neighbourhood slide down the 7 x 7 block
    int r0[9], r1[9], r2[9];
    auto load = [&](int *r, int y) {
        if (inside) {
            const uint8_t *p = src + (size_t)y * g.pitch + x0 - 4;
#pragma unroll
            for (int j = 0; j < 9; ++j) r[j] = p[j];
        } else {
            const uint8_t *p = src + (size_t)reflect101_h(y, g.h) * g.pitch;
#pragma unroll
            for (int j = 0; j < 9; ++j) r[j] = p[reflect101_h(x0 - 4 + j, g.w)];
        }
    };
    load(r0, y0 - 4);
    load(r1, y0 - 3);
#pragma unroll 1
    for (int dy = -3; dy <= 3; ++dy) {
        load(r2, y0 + dy + 1);
#pragma unroll
        for (int j = 1; j <= 7; ++j) {
            const int Ix = (r1[j + 1] - r1[j - 1]) * 2 + (r0[j + 1] - r0[j - 1]) + (r2[j + 1] - r2[j - 1]);
            const int Iy = (r2[j] - r0[j]) * 2 + (r2[j - 1] - r0[j - 1]) + (r2[j + 1] - r0[j + 1]);
            a += Ix * Ix;
            b += Iy * Iy;
            c += Ix * Iy;
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) { r0[j] = r1[j]; r1[j] = r2[j]; }
    }
    // ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float t = __fadd_rn(fa, fb);
    const float det = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
    resp[o] = __fmul_rn(__fsub_rn(det, __fmul_rn(__fmul_rn(harris_k, t), t)), scale_sq_sq);
}

constexpr int HR_THREADS = 1024;

__device__ __forceinline__ uint32_t ordered_key(float f) {          // larger float <=> larger unsigned
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ uint32_t block_incl_scan_1024(uint32_t v, uint32_t *s_warp, uint32_t &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < HR_THREADS / 32; ++w) {
        const uint32_t cnt = s_warp[w];
        if (w < wid) base += cnt;
        total += cnt;
    }
    __syncthreads();
    return incl + base;
}

// KeyPointsFilter::retainBest(n_features) on the float responses of one image per block, ties kept, order kept.
__global__ void __launch_bounds__(HR_THREADS)
harris_retain_kernel(Geom g, int n_features, uint32_t *__restrict__ n_kp, uint32_t *__restrict__ kp_key,
                     uint8_t *__restrict__ kp_score, float *__restrict__ resp) {
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_warp[HR_THREADS / 32];
    __shared__ uint32_t s_prefix, s_remaining;
    const int image = blockIdx.x;
    const uint32_t n = min(n_kp[image], (uint32_t)g.kp_cap);
    if (n_features < 0 || n <= (uint32_t)n_features) return;
    uint32_t *key = kp_key + (size_t)image * g.kp_cap;
    uint8_t *score = kp_score + (size_t)image * g.kp_cap;
    float *r = resp + (size_t)image * g.kp_cap;
    if (n_features == 0) { if (threadIdx.x == 0) n_kp[image] = 0; return; }
    // radix select of the n_features-th largest ordered key, 8 bits per pass from the top
    if (threadIdx.x == 0) { s_prefix = 0; s_remaining = (uint32_t)n_features; }
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t himask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (uint32_t i = threadIdx.x; i < n; i += HR_THREADS) {
            const uint32_t k = ordered_key(r[i]);
            if ((k & himask) == prefix) atomicAdd(&s_hist[(k >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t rem = s_remaining;
            int bin = 255;
            for (; bin > 0; --bin) {
                if (s_hist[bin] >= rem) break;
                rem -= s_hist[bin];
            }
            s_prefix = prefix | ((uint32_t)bin << shift);
            s_remaining = rem;
        }
        __syncthreads();
    }
    const uint32_t cut = s_prefix;
    uint32_t offset = 0;
    for (uint32_t base = 0; base < n; base += HR_THREADS) {
        const uint32_t i = base + threadIdx.x;
        uint32_t kk = 0;
        uint8_t sc = 0;
        float rr = 0.f;
        bool keep = false;
        if (i < n) {
            kk = key[i]; sc = score[i]; rr = r[i];
            keep = ordered_key(rr) >= cut;
        }
        uint32_t total;
        const uint32_t incl = block_incl_scan_1024(keep ? 1u : 0u, s_warp, total);   // syncs: all reads of this chunk done
        if (keep) {
            const uint32_t pos = offset + incl - 1;                                   // pos <= i: never ahead of the reads
            key[pos] = kk; score[pos] = sc; r[pos] = rr;
        }
        offset += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_kp[image] = offset;
}

__global__ void harris_store_kernel(Geom g, const uint32_t *__restrict__ n_kp, const float *__restrict__ resp,
                                    fe_kpoint *__restrict__ kp) {
    const int image = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min((int)n_kp[image], g.kp_cap)) return;
    const size_t o = (size_t)image * g.kp_cap + i;
    kp[o].response = resp[o];
}

int launch_harris_select(const Geom &g, int n_features, const Buffers &b, cudaStream_t s) {
    const float scale = 1.f / ((1 << 2) * 7 * 255.f);
    const float scale_sq_sq = scale * scale * scale * scale;
    harris_response_kernel<<<dim3(div_up(g.kp_cap, 128), g.n_images), 128, 0, s>>>(b.img, g, b.n_kp, b.kp_key, 0.04f, scale_sq_sq, b.harris);
    harris_retain_kernel<<<g.n_images, HR_THREADS, 0, s>>>(g, n_features, b.n_kp, b.kp_key, b.kp_score, b.harris);
    return 2;
}

int launch_harris_store(const Geom &g, const Buffers &b, cudaStream_t s) {
    harris_store_kernel<<<dim3(div_up(g.kp_cap, 256), g.n_images), 256, 0, s>>>(g, b.n_kp, b.harris, b.kp);
    return 1;
}

}  // namespace fe
