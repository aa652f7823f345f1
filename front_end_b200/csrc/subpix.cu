// cv::cornerSubPix(win 5x5, zeroZone -1, 40 iterations / eps 1e-3) on the GPU: one warp per keypoint.
//
// Replaces the per-keypoint refinement loops of the live nodes:
//   /root/reference src/live_stereo.cpp:235-237,321-337 (on the cell sub-image, offsets added afterwards
//   at :340-350) and src/front_end/features.py:600-601,637-640 (on the full image, after the offsets).
// The arithmetic is OpenCV's (imgproc cornersubpix.cpp + samplers.cpp getRectSubPix), pinned bit-for-bit
// against cv2 4.13 by the CPU restatement in oracle/subpix.py; this kernel follows the same operations:
//   * 13 x 13 float patch around the current estimate by bilinear interpolation of the u8 image with float
//     weights, border replicated.  Association as pinned: (p00*a11 + p01*a12) + (p10*a21 + p11*a22) without FMA;
//     rows replicated above / below the image are fma(p01, a, p00 * (1 - a)); replicated columns are
//     p_row * (1 - b) + p_row2 * b; and the cv2 quirk that rows replicated ABOVE the image fill the right-hand
//     side from column W - 2.
//   * central differences in float, then double: gxx = tgx*tgx*m ..., a, b, c, bb1, bb2 accumulated in double.
//     The 121 terms are summed lane-parallel + xor-butterfly instead of sequentially -- the only deviation from
//     the CPU order (a double-rounding-level difference; the tests accept 1e-4 px and report the exact fraction).
//   * 2x2 solve in double, float update, stop when the squared step <= eps^2 or after max_iters, reset to the
//     start point when it moved more than the window half-size.
// The library is compiled with -fmad=false, so none of the float / double expressions below is contracted.
#include "fe_internal.cuh"

namespace fe {

constexpr int SPX_WIN = 5;
constexpr int SPX_N = 2 * SPX_WIN + 1;       // 11
constexpr int SPX_P = SPX_N + 2;             // 13
constexpr int SPX_WARPS = 4;

// exp(-((k - 5) / 5)^2) as float, k = 0..10: correctly rounded expf of the float argument (what glibc's expf
// returns for cornersubpix.cpp's std::exp(float); numpy's SIMD float exp differs in the last bit)
__constant__ float c_spx_e[SPX_N] = {0.36787945f, 0.527292371f, 0.697676301f, 0.852143764f, 0.960789442f, 1.f,
                                     0.960789442f, 0.852143764f, 0.697676301f, 0.527292371f, 0.36787945f};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

__global__ void __launch_bounds__(SPX_WARPS * 32)
subpix_kernel(Geom g, SubpixParams sp, const uint32_t *__restrict__ counts, fe_kpoint *__restrict__ kp,
              float *__restrict__ kx, float *__restrict__ ky) {
    __shared__ float s_patch[SPX_WARPS][SPX_P * SPX_P + 3];
    const int image = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i_kp = blockIdx.x * SPX_WARPS + warp;
    const int n = min((int)counts[image], g.kp_cap);
    if (i_kp >= n) return;
    const size_t o = (size_t)image * g.kp_cap + i_kp;
    const uint8_t *src = sp.src[image];
    const int W = sp.w[image], H = sp.h[image], pitch = sp.pitch[image];
    float *patch = s_patch[warp];

    const float cTx = __fadd_rn(kp[o].x, sp.pre_x[image]), cTy = __fadd_rn(kp[o].y, sp.pre_y[image]);
    float cIx = cTx, cIy = cTy;
    if (sp.refine) {
        const double eps2 = (double)sp.epsilon * (double)sp.epsilon;
        int iter = 0;
        while (true) {
            // ---- getRectSubPix(src, 13 x 13, cI) ----------------------------------------------------
            const float x = __fsub_rn(cIx, 6.f), y = __fsub_rn(cIy, 6.f);
            const int ipx = (int)floorf(x), ipy = (int)floorf(y);
            const float a = __fsub_rn(x, (float)ipx), b = __fsub_rn(y, (float)ipy);
            const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
            const float a11 = __fmul_rn(oma, omb), a12 = __fmul_rn(a, omb), a21 = __fmul_rn(oma, b), a22 = __fmul_rn(a, b);
            // adjustRect
            int sx, rx, rw, sy, ry, rh;
            if (ipx >= 0) { sx = ipx; rx = 0; } else { sx = 0; rx = min(-ipx, SPX_P); }
            if (ipx < W - SPX_P) rw = SPX_P; else { rw = W - ipx - 1; if (rw < 0) { sx += rw; rw = 0; } }
            if (ipy >= 0) { sy = ipy; ry = 0; } else { sy = 0; ry = -ipy; }
            if (ipy < H - SPX_P) rh = SPX_P; else { rh = H - ipy - 1; if (rh < 0) { sy += rh; rh = 0; } }
            const int base_x = sx - rx;
            __syncwarp();
            for (int e = lane; e < SPX_P * SPX_P; e += 32) {
                const int i = e / SPX_P, j = e - i * SPX_P;
                const bool adv = i >= ry && i < rh;
                const int row = min(max(sy + max(0, min(i, rh) - ry), 0), H - 1);
                const int row2 = min(adv ? row + 1 : row, H - 1);
                const uint8_t *r0 = src + (size_t)row * pitch, *r1 = src + (size_t)row2 * pitch;
                float v;
                if (j >= rx && j < rw) {
                    const int c0 = min(max(base_x + j, 0), W - 1), c1 = min(max(base_x + j + 1, 0), W - 1);
                    const float p00 = (float)r0[c0], p01 = (float)r0[c1];
                    if (adv) {
                        const float p10 = (float)r1[c0], p11 = (float)r1[c1];
                        v = __fadd_rn(__fadd_rn(__fmul_rn(p00, a11), __fmul_rn(p01, a12)),
                                      __fadd_rn(__fmul_rn(p10, a21), __fmul_rn(p11, a22)));
                    } else {
                        v = __fmaf_rn(p01, a, __fmul_rn(p00, oma));
                    }
                } else {
                    const int jj = j < rx ? rx : (i < ry ? rw - 1 : rw);
                    const int c = min(max(base_x + jj, 0), W - 1);
                    v = __fadd_rn(__fmul_rn((float)r0[c], omb), __fmul_rn((float)r1[c], b));
                }
                patch[e] = v;
            }
            __syncwarp();
            // ---- gradient moments -------------------------------------------------------------------
            double sa = 0, sb = 0, sc = 0, s1 = 0, s2 = 0;
            for (int e = lane; e < SPX_N * SPX_N; e += 32) {
                const int i = e / SPX_N, j = e - i * SPX_N;
                const float *c = patch + (i + 1) * SPX_P + (j + 1);
                const double m = (double)__fmul_rn(c_spx_e[i], c_spx_e[j]);
                const double tgx = (double)__fsub_rn(c[1], c[-1]);
                const double tgy = (double)__fsub_rn(c[SPX_P], c[-SPX_P]);
                const double gxx = __dmul_rn(__dmul_rn(tgx, tgx), m);
                const double gxy = __dmul_rn(__dmul_rn(tgx, tgy), m);
                const double gyy = __dmul_rn(__dmul_rn(tgy, tgy), m);
                const double px = (double)(j - SPX_WIN), py = (double)(i - SPX_WIN);
                sa = __dadd_rn(sa, gxx); sb = __dadd_rn(sb, gxy); sc = __dadd_rn(sc, gyy);
                s1 = __dadd_rn(s1, __dadd_rn(__dmul_rn(gxx, px), __dmul_rn(gxy, py)));
                s2 = __dadd_rn(s2, __dadd_rn(__dmul_rn(gxy, px), __dmul_rn(gyy, py)));
            }
            sa = warp_sum(sa); sb = warp_sum(sb); sc = warp_sum(sc); s1 = warp_sum(s1); s2 = warp_sum(s2);
            const double det = __dsub_rn(__dmul_rn(sa, sc), __dmul_rn(sb, sb));
            if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
            const double scale = __ddiv_rn(1.0, det);
            const float nx = (float)__dsub_rn(__dadd_rn((double)cIx, __dmul_rn(__dmul_rn(sc, scale), s1)),
                                              __dmul_rn(__dmul_rn(sb, scale), s2));
            const float ny = (float)__dadd_rn(__dsub_rn((double)cIy, __dmul_rn(__dmul_rn(sb, scale), s1)),
                                              __dmul_rn(__dmul_rn(sa, scale), s2));
            const float dx = __fsub_rn(nx, cIx), dy = __fsub_rn(ny, cIy);
            const double err = (double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            // cv2 4.13 (the pin): a step that leaves the image is discarded, the estimate stays in bounds
            if (nx < 0.f || nx >= (float)W || ny < 0.f || ny >= (float)H) break;
            cIx = nx; cIy = ny;
            ++iter;
            if (!(iter < sp.max_iters && err > eps2)) break;
        }
        if (fabsf(__fsub_rn(cIx, cTx)) > (float)SPX_WIN || fabsf(__fsub_rn(cIy, cTy)) > (float)SPX_WIN) { cIx = cTx; cIy = cTy; }
    }
    if (lane == 0) {
        const float ox = __fadd_rn(__fadd_rn(cIx, sp.post1_x[image]), sp.post2_x[image]);
        const float oy = __fadd_rn(__fadd_rn(cIy, sp.post1_y[image]), sp.post2_y[image]);
        kp[o].x = ox; kp[o].y = oy;
        kx[o] = ox; ky[o] = oy;
    }
}

int launch_subpix(const Geom &g, const Buffers &b, const uint32_t *counts, const SubpixParams &sp, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, SPX_WARPS), g.n_images);
    subpix_kernel<<<grid, SPX_WARPS * 32, 0, s>>>(g, sp, counts, b.kp, b.kx, b.ky);
    return 1;
}

}  // namespace fe
