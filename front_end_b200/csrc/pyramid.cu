// Multi-level ORB (cv::ORB with nlevels > 1): the image pyramid and the per-level result accumulation.
//
// Replaces the pyramid part of cv::ORB::detectAndCompute as the reference configures it at
// /root/reference src/front_end/features.py:292-352,378-387 (nLevels 2 / 4 sweeps), src/utils.cpp:84-94,
// src/StereoCamera.cpp:504-511 and bin/detect_node:50 (ORB_create() -> 8 levels).  Semantics pinned bit-exactly against
// cv2 4.13 by oracle/orb.py:
//   * level l = resize(level l-1, (cvRound(W / s^l), cvRound(H / s^l)), INTER_LINEAR_EXACT) -- from the PREVIOUS level;
//     INTER_LINEAR_EXACT for u8 is a horizontal pass in 8.8 fixed point (weights cvRound(f * 256)), a vertical pass in
//     16.16 and round-half-up: out = ((s00*ax0 + s01*ax1) * ay0 + (s10*ax0 + s11*ax1) * ay1 + 2^15) >> 16;
//   * per level: FAST-9_16 -> border 31 -> retainBest(quota_l, ties kept) -> IC angle -> 7x7 blur -> rBRIEF (the
//     single-level kernels, run on the level geometry);
//   * output pt = pt_l * s^l (float multiply), size = 31 * s^l, octave = l; level-major order.
#include "fe_internal.cuh"

namespace fe {

// one thread per destination pixel; tab = [ofs_x(dw) | a1_x(dw) | ofs_y(dh) | a1_y(dh)] (weights in 1/256)
__global__ void __launch_bounds__(256)
resize_linear_exact_kernel(const uint8_t *__restrict__ src, int sw, int sh, int spitch, size_t sstride,
                           uint8_t *__restrict__ dst, int dw, int dh, int dpitch, size_t dstride,
                           const int *__restrict__ tab) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6), image = blockIdx.z;
    if (x >= dpitch || y >= dh) return;
    uint8_t out = 0;
    if (x < dw) {
        const int ox = tab[x], ax1 = tab[dw + x], oy = tab[2 * dw + y], ay1 = tab[2 * dw + dh + y];
        const int ax0 = 256 - ax1, ay0 = 256 - ay1;
        const int x1 = min(ox + 1, sw - 1), y1 = min(oy + 1, sh - 1);
        const uint8_t *s0 = src + (size_t)image * sstride + (size_t)oy * spitch, *s1 = src + (size_t)image * sstride + (size_t)y1 * spitch;
        const int h0 = (int)s0[ox] * ax0 + (int)s0[x1] * ax1, h1 = (int)s1[ox] * ax0 + (int)s1[x1] * ax1;
        const int v = (h0 * ay0 + h1 * ay1 + (1 << 15)) >> 16;
        out = (uint8_t)min(max(v, 0), 255);
    }
    dst[(size_t)image * dstride + (size_t)y * dpitch + x] = out;     // padding columns are written as 0
}

int launch_resize_linear_exact(const uint8_t *src, int sw, int sh, int spitch, size_t sstride, uint8_t *dst, int dw, int dh,
                               int dpitch, size_t dstride, const int *tab, int n_images, cudaStream_t s) {
    dim3 grid(div_up(dpitch, 64), div_up(dh, 4), n_images);
    resize_linear_exact_kernel<<<grid, 256, 0, s>>>(src, sw, sh, spitch, sstride, dst, dw, dh, dpitch, dstride, tab);
    return 1;
}

// append this level's keypoints / descriptors (scaled to level-0 coordinates) behind what the lower levels produced
__global__ void __launch_bounds__(256)
pyr_append_kernel(Geom g, int level, float scale, float kp_size, const uint32_t *__restrict__ n_level,
                  const fe_kpoint *__restrict__ kp, const uint8_t *__restrict__ desc, const uint32_t *__restrict__ n_acc,
                  fe_kpoint *__restrict__ akp, uint8_t *__restrict__ adesc, int with_desc) {
    const int image = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min((int)n_level[image], g.kp_cap);
    if (i >= n) return;
    const int pos = (int)n_acc[image] + i;
    if (pos >= g.kp_cap) return;                        // counts still report the required size
    const size_t src = (size_t)image * g.kp_cap + i, dst = (size_t)image * g.kp_cap + pos;
    fe_kpoint k = kp[src];
    k.x = __fmul_rn(k.x, scale); k.y = __fmul_rn(k.y, scale);
    k.size = kp_size; k.octave = level;
    akp[dst] = k;
    if (with_desc) {
        const uint4 *sd = reinterpret_cast<const uint4 *>(desc + src * 32);
        uint4 *dd = reinterpret_cast<uint4 *>(adesc + dst * 32);
        dd[0] = sd[0]; dd[1] = sd[1];
    }
}

__global__ void pyr_add_counts_kernel(int n_images, int kp_cap, const uint32_t *__restrict__ n_level, uint32_t *__restrict__ n_acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_images) n_acc[i] += min(n_level[i], (uint32_t)kp_cap);
}

int launch_pyr_append(const Geom &g, int level, float scale, float kp_size, const Buffers &b, fe_kpoint *akp, uint8_t *adesc,
                      uint32_t *n_acc, bool with_desc, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, 256), g.n_images);
    pyr_append_kernel<<<grid, 256, 0, s>>>(g, level, scale, kp_size, b.n_kp, b.kp, b.desc, n_acc, akp, adesc, with_desc ? 1 : 0);
    pyr_add_counts_kernel<<<div_up(g.n_images, 128), 128, 0, s>>>(g.n_images, g.kp_cap, b.n_kp, n_acc);
    return 2;
}

// kx / ky for the matcher from the accumulated wire records
__global__ void pyr_coords_kernel(Geom g, const uint32_t *__restrict__ n_acc, const fe_kpoint *__restrict__ kp,
                                  float *__restrict__ kx, float *__restrict__ ky) {
    const int image = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min((int)n_acc[image], g.kp_cap)) return;
    const size_t o = (size_t)image * g.kp_cap + i;
    kx[o] = kp[o].x; ky[o] = kp[o].y;
}

int launch_pyr_coords(const Geom &g, const Buffers &b, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, 256), g.n_images);
    pyr_coords_kernel<<<grid, 256, 0, s>>>(g, b.n_kp, b.kp, b.kx, b.ky);
    return 1;
}

}  // namespace fe
