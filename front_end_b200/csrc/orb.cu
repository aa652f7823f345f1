// ORB orientation (intensity centroid), the descriptor's 7x7 float Gaussian, and rBRIEF-256.
//
// Replaces the tail of cv::ORB::detect (IC_Angle) and DescriptorExtractor::compute with ORB, as the
// reference reaches them at /root/reference bin/feature_node:50-54, src/front_end/features.py:450-451,
// 721-722, bin/detect_node:50-51, src/StereoCamera.cpp:84-89,123-128.  Semantics SURVEY.md A.3/A.4:
//   * m10 = sum u*I, m01 = sum v*I over |u| <= umax[|v|], radius 15; angle = fastAtan2(m01, m10) with
//     OpenCV's f32 degree-7 polynomial, evaluated WITHOUT fused multiply-add (pinned against cv2);
//   * blur = separable 7-tap sigma=2 in f32, BORDER_REFLECT_101, row pass as a left-to-right FMA
//     chain, column pass centre-out with pair sums, round-half-even to u8;
//   * bit i of the descriptor = blurred(p_2i rotated) < blurred(p_2i+1 rotated), rotation by
//     (cos, sin) of the angle in f32 (cos/sin taken in double, rounded to f32), coordinates rounded
//     half-even.
// The library is compiled with -fmad=false; every FMA below is explicit.
#include "fe_internal.cuh"

namespace fe {

__constant__ int8_t c_pattern[256][4] = {
#include "orb_pattern.inc"
};

__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

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

constexpr int ORI_WARPS = 8;

// One warp per keypoint: lanes span u = -15..15, loop over the 31 rows.  Writes the wire-format
// keypoint, float coordinates and the steering (cos, sin).
__global__ void __launch_bounds__(ORI_WARPS * 32)
orient_pack_kernel(const uint8_t *__restrict__ img, Geom g, const uint32_t *__restrict__ n_kp,
                   const uint32_t *__restrict__ kp_key, const uint8_t *__restrict__ kp_score,
                   int orientation, int report_score, float kp_size, fe_kpoint *__restrict__ kp,
                   float *__restrict__ kx, float *__restrict__ ky, float2 *__restrict__ kcs) {
    const int image = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * ORI_WARPS + (threadIdx.x >> 5);
    const int n = min((int)n_kp[image], g.kp_cap);
    if (i >= n) return;
    const size_t o = (size_t)image * g.kp_cap + i;
    const uint32_t key = kp_key[o];
    const int x = key & 0xFFFF, y = key >> 16;
    float angle = -1.f;
    float2 cs = make_float2(1.f, 0.f);
    if (orientation) {
        const uint8_t *c = img + (size_t)image * g.img_stride + (size_t)y * g.pitch + x;
        const int u = lane - 15;
        int m10 = 0, m01 = 0;
        if (lane < 31) {
#pragma unroll 1
            for (int v = -15; v <= 15; ++v) {
                const int d = c_umax[v < 0 ? -v : v];
                if (u >= -d && u <= d) {
                    const int val = c[v * g.pitch + u];
                    m10 += u * val;
                    m01 += v * val;
                }
            }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, off);
            m01 += __shfl_xor_sync(0xffffffffu, m01, off);
        }
        angle = fast_atan2_deg((float)m01, (float)m10);
    }
    if (lane == 0) {
        if (orientation) {
            // orb.cpp: float angle = kpt.angle * (float)(CV_PI/180.f); a = (float)cos(angle), b = (float)sin(angle)
            const float th = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.0));
            cs = make_float2((float)cos((double)th), (float)sin((double)th));
        }
        fe_kpoint k;
        k.x = (float)x; k.y = (float)y; k.size = kp_size; k.angle = angle;
        k.response = report_score ? (float)kp_score[o] : 0.f;
        k.octave = 0; k.class_id = -1;
        kp[o] = k;
        kx[o] = (float)x; ky[o] = (float)y;
        kcs[o] = cs;
    }
}

int launch_orient_pack(const Geom &g, const DetectParams &p, const Buffers &b, bool orientation,
                       float kp_size, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, ORI_WARPS), g.n_images);
    orient_pack_kernel<<<grid, ORI_WARPS * 32, 0, s>>>(b.img, g, b.n_kp, b.kp_key, b.kp_score,
                                                       orientation ? 1 : 0, p.nonmax ? 1 : 0, kp_size,
                                                       b.kp, b.kx, b.ky, b.kcs);
    return 1;
}

// Externally supplied keypoints (fe_describe / fe_stereo_match): derive kx, ky, (cos, sin) from the
// uploaded wire-format records.  kp.angle is used literally (ORB.compute does not recompute it).
__global__ void unpack_kps_kernel(Geom g, const uint32_t *__restrict__ counts,
                                  const fe_kpoint *__restrict__ kp, float *__restrict__ kx,
                                  float *__restrict__ ky, float2 *__restrict__ kcs) {
    const int image = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min((int)counts[image], g.kp_cap)) return;
    const size_t o = (size_t)image * g.kp_cap + i;
    const fe_kpoint k = kp[o];
    kx[o] = k.x; ky[o] = k.y;
    const float th = __fmul_rn(k.angle, (float)(3.14159265358979323846 / 180.0));
    kcs[o] = make_float2((float)cos((double)th), (float)sin((double)th));
}

int launch_unpack_kps(const Geom &g, const Buffers &b, const uint32_t *counts, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, 256), g.n_images);
    unpack_kps_kernel<<<grid, 256, 0, s>>>(g, counts, b.kp, b.kx, b.ky, b.kcs);
    return 1;
}

// ---- 7x7 Gaussian ----------------------------------------------------------------------------
constexpr int BL_TW = 128, BL_TH = 16, BL_THREADS = 256;
constexpr int BL_IW = BL_TW + 6, BL_IH = BL_TH + 6;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(BL_THREADS)
gauss7_kernel(const uint8_t *__restrict__ img, uint8_t *__restrict__ out, Geom g) {
    __shared__ uint8_t s_in[BL_IH][BL_IW + 2];
    __shared__ float s_row[BL_IH][BL_TW];
    // getGaussianKernel(7, 2, CV_32F)
    const float g0 = 0.07015932351350784f, g1 = 0.13107487559318542f, g2 = 0.1907128244638443f,
                g3 = 0.21610593795776367f;
    const float gk[7] = {g0, g1, g2, g3, g2, g1, g0};
    const int image = blockIdx.z;
    const int x0 = blockIdx.x * BL_TW, y0 = blockIdx.y * BL_TH;
    const uint8_t *src = img + (size_t)image * g.img_stride;
    for (int i = threadIdx.x; i < BL_IH * BL_IW; i += BL_THREADS) {
        const int r = i / BL_IW, c = i - r * BL_IW;
        const int gy = reflect101(y0 + r - 3, g.h), gx = reflect101(x0 + c - 3, g.w);
        // tiles that hang over the right/bottom edge clamp their reads; those outputs are discarded
        s_in[r][c] = src[(size_t)min(max(gy, 0), g.h - 1) * g.pitch + min(max(gx, 0), g.w - 1)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BL_IH * BL_TW; i += BL_THREADS) {
        const int r = i / BL_TW, c = i - r * BL_TW;
        float s = __fmul_rn((float)s_in[r][c], gk[0]);
#pragma unroll
        for (int k = 1; k < 7; ++k) s = __fmaf_rn((float)s_in[r][c + k], gk[k], s);
        s_row[r][c] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BL_TH * BL_TW; i += BL_THREADS) {
        const int r = i / BL_TW, c = i - r * BL_TW;
        const int x = x0 + c, y = y0 + r;
        if (x >= g.w || y >= g.h) continue;
        float s = __fmul_rn(s_row[r + 3][c], g3);
        s = __fmaf_rn(__fadd_rn(s_row[r + 4][c], s_row[r + 2][c]), g2, s);
        s = __fmaf_rn(__fadd_rn(s_row[r + 5][c], s_row[r + 1][c]), g1, s);
        s = __fmaf_rn(__fadd_rn(s_row[r + 6][c], s_row[r + 0][c]), g0, s);
        int v = __float2int_rn(s);
        v = min(max(v, 0), 255);
        out[(size_t)image * g.img_stride + (size_t)y * g.pitch + x] = (uint8_t)v;
    }
}

int launch_blur(const Geom &g, const Buffers &b, cudaStream_t s) {
    dim3 grid(div_up(g.w, BL_TW), div_up(g.h, BL_TH), g.n_images);
    gauss7_kernel<<<grid, BL_THREADS, 0, s>>>(b.img, b.blur, g);
    return 1;
}

// ---- rBRIEF-256 ------------------------------------------------------------------------------
constexpr int BR_WARPS = 8;

// One warp per keypoint.  Lane l evaluates tests l, l+32, ..., l+224; the ballot of test j*32+l is
// exactly little-endian word j of the 32-byte descriptor (bit i of byte b <-> test 8b+i).
__global__ void __launch_bounds__(BR_WARPS * 32)
rbrief_kernel(const uint8_t *__restrict__ blur, Geom g, const uint32_t *__restrict__ counts,
              const float *__restrict__ kx, const float *__restrict__ ky,
              const float2 *__restrict__ kcs, uint8_t *__restrict__ desc) {
    __shared__ char4 s_pat[256];
    for (int i = threadIdx.x; i < 256; i += BR_WARPS * 32)
        s_pat[i] = make_char4(c_pattern[i][0], c_pattern[i][1], c_pattern[i][2], c_pattern[i][3]);
    __syncthreads();
    const int image = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * BR_WARPS + (threadIdx.x >> 5);
    if (i >= min((int)counts[image], g.kp_cap)) return;
    const size_t o = (size_t)image * g.kp_cap + i;
    const int cx = __float2int_rn(kx[o]), cy = __float2int_rn(ky[o]);
    const float2 cs = kcs[o];
    const float a = cs.x, b = cs.y;
    const uint8_t *c = blur + (size_t)image * g.img_stride + (size_t)cy * g.pitch + cx;
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const char4 pt = s_pat[j * 32 + lane];
        const float x0 = (float)pt.x, y0 = (float)pt.y, x1 = (float)pt.z, y1 = (float)pt.w;
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int t0 = c[iy0 * g.pitch + ix0];
        const int t1 = c[iy1 * g.pitch + ix1];
        const uint32_t w = __ballot_sync(0xffffffffu, t0 < t1);
        if (lane == j) word = w;
    }
    if (lane < 8) reinterpret_cast<uint32_t *>(desc + o * 32)[lane] = word;
}

int launch_brief(const Geom &g, const Buffers &b, const uint32_t *counts, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, BR_WARPS), g.n_images);
    rbrief_kernel<<<grid, BR_WARPS * 32, 0, s>>>(b.blur, g, counts, b.kx, b.ky, b.kcs, b.desc);
    return 1;
}

}  // namespace fe
