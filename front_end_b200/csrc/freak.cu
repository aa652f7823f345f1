// FREAK descriptors (cv::FREAK, opencv_contrib xfeatures2d/src/freak.cpp) with caller-supplied selected pairs.
//
// Replaces  cv2.xfeatures2d.FREAK_create().compute(img, kps)  of the reference's descriptor comparison
// (/root/reference bin/detect_node:43-45; bin/result_ONE:25, result_TWO:29, result_THREE:23 list "FREAK" beside BRIEF /
// SURF / ORB): orientationNormalized = scaleNormalized = true, patternScale 22, nOctaves 4.
//
// Algorithm: 43 receptive fields (7 staggered rings of 6 + the centre), tabulated for 64 scales x 256 orientations as float
// (x, y, sigma) -- the table is built on the host in double exactly as buildPattern does (abi.cu: fe_set_freak).  Field value =
// rounded box mean on the CV_32S integral image over [int(c - s + 0.5), int(c + s + 1.5)) (sigma < 0.5: 10-bit bilinear
// sample).  Orientation: 45 field pairs, integer weights, per-pair truncating division by 2048, atan2 -> kp.angle, quantised to
// 256 steps.  Descriptor: 512 selected pairs, bit = value[i] >= value[j], stored in the order of OpenCV's SSE path
// (pair 128 b + 16 u + t -> byte 16 b + 15 - t, bit u).  PARITY UNPINNED: the default selection (FREAK_DEF_PAIRS) is a table
// of opencv_contrib's sources, absent here; the caller passes cv::FREAK's own `selectedPairs` argument.
//
// One warp per keypoint.  Lane l owns fields l and l + 32 (eight integral-image corners, L1 / L2 resident: the pattern of
// a size-31 keypoint spans ~130 px); the 43 values go through shared memory; lane m owns orientation pairs m and m + 32
// (integer sums, so the warp reduction is order-free); for the descriptor lane p of round w evaluates the pair that lands
// on bit p of output word w, so a ballot IS the output word.
#include "fe_internal.cuh"

namespace fe {

constexpr int FK_WARPS = 8, FK_KPW = 2;      // warps per CTA, keypoints per warp
constexpr int FK_POINTS = 43, FK_ORIENT = 256, FK_PAIRS = 512, FK_OPAIRS = 45;

__device__ __forceinline__ int freak_mean(const uint8_t *__restrict__ img, int pitch, const int32_t *__restrict__ S, int stride,
                                          const float *__restrict__ pt, float kx, float ky) {
    const float px = __ldg(pt), py = __ldg(pt + 1), radius = __ldg(pt + 2);
    const float xf = __fadd_rn(px, kx), yf = __fadd_rn(py, ky);
    if (radius < 0.5f) {
        const int x = (int)xf, y = (int)yf;
        const int r_x = (int)__fmul_rn(__fsub_rn(xf, (float)x), 1024.f), r_y = (int)__fmul_rn(__fsub_rn(yf, (float)y), 1024.f);
        const int r_x_1 = 1024 - r_x, r_y_1 = 1024 - r_y;
        const uint8_t *p = img + (size_t)y * pitch + x;
        unsigned v = (unsigned)(r_x_1 * r_y_1 * (int)p[0] + r_x * r_y_1 * (int)p[1] + r_x_1 * r_y * (int)p[pitch] + r_x * r_y * (int)p[pitch + 1]);
        v += 2u * 1024u * 1024u;
        return (int)((v / (4u * 1024u * 1024u)) & 0xffu);      // (sic) a quarter of the mean, as the source computes it
    }
    const int x_left = (int)((double)__fsub_rn(xf, radius) + 0.5), y_top = (int)((double)__fsub_rn(yf, radius) + 0.5);
    const int x_right = (int)((double)__fadd_rn(xf, radius) + 1.5), y_bottom = (int)((double)__fadd_rn(yf, radius) + 1.5);
    const int32_t *pt_ = S + (size_t)y_top * stride, *pb = S + (size_t)y_bottom * stride;
    const int sum = __ldg(pb + x_right) - __ldg(pb + x_left) + __ldg(pt_ + x_left) - __ldg(pt_ + x_right);
    const int area = (x_right - x_left) * (y_bottom - y_top);
    return ((sum + area / 2) / area) & 0xff;
}

// opairs: [45] (i, j, weight_dx, weight_dy); dpairs: [512] (i, j) in extraction order; scale_idx: per keypoint
__global__ void __launch_bounds__(FK_WARPS * 32)
freak_kernel(const uint8_t *__restrict__ imgs, const int32_t *__restrict__ integ, Geom g, const uint32_t *__restrict__ counts,
             fe_kpoint *__restrict__ kp, const int32_t *__restrict__ scale_idx, const float *__restrict__ table,
             const int4 *__restrict__ opairs, const uchar2 *__restrict__ dpairs, int orientation_normalized, uint8_t *__restrict__ out) {
    __shared__ uchar2 s_pairs[FK_PAIRS];
    __shared__ int4 s_opairs[FK_OPAIRS];
    __shared__ int s_val[FK_WARPS][FK_POINTS + 1];
    for (int i = threadIdx.x; i < FK_PAIRS; i += blockDim.x) s_pairs[i] = dpairs[i];
    for (int i = threadIdx.x; i < FK_OPAIRS; i += blockDim.x) s_opairs[i] = opairs[i];
    __syncthreads();
    const int image = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = min((int)counts[image], g.kp_cap);
    const int stride = g.w + 1;
    const int32_t *S = integ + (size_t)image * (g.h + 1) * stride;
    const uint8_t *img = imgs + (size_t)image * g.img_stride;
    int *val = s_val[warp];
    for (int kk = 0; kk < FK_KPW; ++kk) {
        const int k = (blockIdx.x * FK_WARPS + warp) * FK_KPW + kk;
        if (k >= n) return;
        fe_kpoint *key = kp + (size_t)image * g.kp_cap + k;
        const float kx = key->x, ky = key->y;
        const float *tab = table + (size_t)scale_idx[(size_t)image * g.kp_cap + k] * FK_ORIENT * FK_POINTS * 3;
        int theta = 0;
        float angle = 0.f;
        if (orientation_normalized) {
            val[lane] = freak_mean(img, g.pitch, S, stride, tab + lane * 3, kx, ky);
            if (lane + 32 < FK_POINTS) val[lane + 32] = freak_mean(img, g.pitch, S, stride, tab + (lane + 32) * 3, kx, ky);
            __syncwarp();
            int d0 = 0, d1 = 0;
            for (int m = lane; m < FK_OPAIRS; m += 32) {
                const int4 p = s_opairs[m];
                const int delta = val[p.x] - val[p.y];
                d0 += delta * p.z / 2048;
                d1 += delta * p.w / 2048;
            }
            d0 = __reduce_add_sync(0xffffffffu, d0);
            d1 = __reduce_add_sync(0xffffffffu, d1);
            __syncwarp();
            // atan2f((float)d1, (float)d0) taken as correctly rounded, then * (180 / pi) in double
            const float a = (float)atan2((double)(float)d1, (double)(float)d0);
            angle = (float)((double)a * (180.0 / 3.1415926535897932384626433832795));
            const double t = (double)__fmul_rn((float)FK_ORIENT, angle) * (1 / 360.0);
            theta = angle < 0.f ? (int)(t - 0.5) : (int)(t + 0.5);
            if (theta < 0) theta += FK_ORIENT;
            if (theta >= FK_ORIENT) theta -= FK_ORIENT;
        }
        if (lane == 0) key->angle = angle;
        const float *rt = tab + (size_t)theta * FK_POINTS * 3;
        val[lane] = freak_mean(img, g.pitch, S, stride, rt + lane * 3, kx, ky);
        if (lane + 32 < FK_POINTS) val[lane + 32] = freak_mean(img, g.pitch, S, stride, rt + (lane + 32) * 3, kx, ky);
        __syncwarp();
        uint32_t *dst = reinterpret_cast<uint32_t *>(out + ((size_t)image * g.kp_cap + k) * 64);
#pragma unroll 4
        for (int w = 0; w < 16; ++w) {
            const int B = 4 * w + (lane >> 3), u = lane & 7;
            const uchar2 p = s_pairs[128 * (B >> 4) + 16 * u + 15 - (B & 15)];
            const uint32_t bal = __ballot_sync(0xffffffffu, val[p.x] >= val[p.y]);
            if (lane == 0) dst[w] = bal;
        }
        __syncwarp();
    }
}

int launch_freak(const Geom &g, const Buffers &b, const uint32_t *counts, const int32_t *scale_idx, const float *table,
                 const int4 *opairs, const uchar2 *dpairs, int orientation_normalized, uint8_t *out, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, FK_WARPS * FK_KPW), g.n_images);
    freak_kernel<<<grid, FK_WARPS * 32, 0, s>>>(b.img, b.integral, g, counts, b.kp, scale_idx, table, opairs, dpairs,
                                                 orientation_normalized, out);
    return 1;
}

}  // namespace fe
