// Brute-force Hamming matching on 256-bit descriptors: LOP3 (xor) + POPC on 32-bit words, with the
// reference's mask predicates evaluated in-register instead of materialising an Nq x Nt mask.
//
// Replaces, in one pass over all (query, train) pairs of a stereo pair:
//   * the O(N^2) host mask loops  /root/reference src/StereoCamera.cpp:182-196 (epipolar band),
//     src/front_end/algorithm.py:825-836, src/WindowMatcher.cpp:104-128 (search box);
//   * BFMatcher::knnMatch(q, t, k=2, mask)   StereoCamera.cpp:199-201, WindowMatcher.cpp:150-153,
//     algorithm.py:848-853      -> (best, second) per query among mask-allowed trains;
//   * BFMatcher(crossCheck=true)::match      src/live_stereo.cpp:240,364, features.py:670,724
//     -> unmasked row arg-min (allbest) and column arg-min (colbest).
// Keys are (distance << 16 | index): an unsigned min yields the smallest distance and, among ties,
// the lowest index -- OpenCV's stable ordering.  Requires index < 65536.
//
// Tiling: a CTA owns QT = threads*QPT queries held in registers (8 words each) and streams the
// train descriptors through shared memory in tiles; every train word is a warp-wide broadcast
// LDS.128 amortised over QPT queries per lane.  Column minima are reduced per warp with REDUX
// (__reduce_min_sync) and merged through shared then global atomicMin.
#include "fe_internal.cuh"

namespace fe {

constexpr int TT = 256;   // train descriptors per shared-memory tile

template <int MASK>
__device__ __forceinline__ bool allowed(float qx, float qy, float tx, float ty, const MatchParams &mp) {
    if (MASK == FE_MASK_EPIPOLAR) return fabsf(__fsub_rn(qy, ty)) <= mp.epi_threshold;
    if (MASK == FE_MASK_WINDOW)
        return fabsf(__fsub_rn(qx, tx)) < mp.half_w && fabsf(__fsub_rn(qy, ty)) < mp.half_h;
    return true;
}

template <int QPT, int THREADS, int MASK, bool WANT_ALL>
__global__ void __launch_bounds__(THREADS)
hamming_match_kernel(Geom g, MatchParams mp, const uint32_t *__restrict__ counts,
                     const uint8_t *__restrict__ desc, const float *__restrict__ kx,
                     const float *__restrict__ ky, uint32_t *__restrict__ best_out,
                     uint32_t *__restrict__ second_out, uint32_t *__restrict__ allbest_out,
                     uint32_t *__restrict__ colbest) {
    __shared__ uint4 s_desc[TT * 2];
    __shared__ float s_tx[TT], s_ty[TT];
    __shared__ uint32_t s_col[TT];

    const int pair = blockIdx.y;
    const int qi = 2 * pair, ti = 2 * pair + 1;
    const int nq = min((int)counts[qi], g.kp_cap), nt = min((int)counts[ti], g.kp_cap);
    const int q0 = blockIdx.x * (THREADS * QPT);
    if (q0 >= nq) return;
    const int lane = threadIdx.x & 31;

    uint32_t q[QPT][8];
    float qx[QPT], qy[QPT];
    uint32_t best[QPT], second[QPT], allb[QPT];
    int qidx[QPT];
    bool valid[QPT];
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        // queries of one lane are THREADS apart so that a warp's loads stay coalesced
        qidx[j] = q0 + j * THREADS + threadIdx.x;
        valid[j] = qidx[j] < nq;
        const int src = valid[j] ? qidx[j] : nq - 1;
        const uint4 *p = reinterpret_cast<const uint4 *>(desc + ((size_t)qi * g.kp_cap + src) * 32);
        const uint4 a = __ldg(p), b = __ldg(p + 1);
        q[j][0] = a.x; q[j][1] = a.y; q[j][2] = a.z; q[j][3] = a.w;
        q[j][4] = b.x; q[j][5] = b.y; q[j][6] = b.z; q[j][7] = b.w;
        qx[j] = kx[(size_t)qi * g.kp_cap + src];
        qy[j] = __fadd_rn(ky[(size_t)qi * g.kp_cap + src], mp.q_off);
        best[j] = second[j] = allb[j] = KEY_NONE;
    }

    const uint4 *tdesc = reinterpret_cast<const uint4 *>(desc + (size_t)ti * g.kp_cap * 32);
    const float *tkx = kx + (size_t)ti * g.kp_cap, *tky = ky + (size_t)ti * g.kp_cap;

    for (int t0 = 0; t0 < nt; t0 += TT) {
        const int tn = min(TT, nt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 2; i += THREADS) s_desc[i] = __ldg(tdesc + (size_t)t0 * 2 + i);
        for (int i = threadIdx.x; i < tn; i += THREADS) {
            s_tx[i] = tkx[t0 + i];
            s_ty[i] = __fadd_rn(tky[t0 + i], mp.t_off);
            if (WANT_ALL) s_col[i] = KEY_NONE;
        }
        __syncthreads();
#pragma unroll 2
        for (int t = 0; t < tn; ++t) {
            const uint4 ta = s_desc[2 * t], tb = s_desc[2 * t + 1];
            const float tx = s_tx[t], ty = s_ty[t];
            const uint32_t tidx = (uint32_t)(t0 + t);
            uint32_t cmin = KEY_NONE;
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                const uint32_t d = __popc(q[j][0] ^ ta.x) + __popc(q[j][1] ^ ta.y) +
                                   __popc(q[j][2] ^ ta.z) + __popc(q[j][3] ^ ta.w) +
                                   __popc(q[j][4] ^ tb.x) + __popc(q[j][5] ^ tb.y) +
                                   __popc(q[j][6] ^ tb.z) + __popc(q[j][7] ^ tb.w);
                const uint32_t key = (d << 16) | tidx;
                if (allowed<MASK>(qx[j], qy[j], tx, ty, mp)) {
                    second[j] = min(second[j], max(best[j], key));
                    best[j] = min(best[j], key);
                }
                if (WANT_ALL) {
                    allb[j] = min(allb[j], key);
                    if (valid[j]) cmin = min(cmin, (d << 16) | (uint32_t)qidx[j]);
                }
            }
            if (WANT_ALL) {
                cmin = __reduce_min_sync(0xffffffffu, cmin);
                if (lane == 0) atomicMin(&s_col[t], cmin);
            }
        }
        if (WANT_ALL) {
            __syncthreads();
            uint32_t *col = colbest + (size_t)pair * g.kp_cap + t0;
            for (int i = threadIdx.x; i < tn; i += THREADS) atomicMin(&col[i], s_col[i]);
        }
    }
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        if (!valid[j]) continue;
        const size_t o = (size_t)pair * g.kp_cap + qidx[j];
        best_out[o] = best[j];
        second_out[o] = second[j];
        if (WANT_ALL) allbest_out[o] = allb[j];
    }
}

template <int QPT, int THREADS>
static void launch_variant(const Geom &g, int n_pairs, const MatchParams &mp, const Buffers &b,
                           const uint32_t *counts, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, QPT * THREADS), n_pairs);
#define FE_MATCH_GO(MASK, ALL)                                                                     \
    hamming_match_kernel<QPT, THREADS, MASK, ALL><<<grid, THREADS, 0, s>>>(                        \
        g, mp, counts, b.desc, b.kx, b.ky, b.best, b.second, b.allbest, b.colbest)
    if (mp.want_all) {
        if (mp.mask == FE_MASK_EPIPOLAR) FE_MATCH_GO(FE_MASK_EPIPOLAR, true);
        else if (mp.mask == FE_MASK_WINDOW) FE_MATCH_GO(FE_MASK_WINDOW, true);
        else FE_MATCH_GO(FE_MASK_NONE, true);
    } else {
        if (mp.mask == FE_MASK_EPIPOLAR) FE_MATCH_GO(FE_MASK_EPIPOLAR, false);
        else if (mp.mask == FE_MASK_WINDOW) FE_MATCH_GO(FE_MASK_WINDOW, false);
        else FE_MATCH_GO(FE_MASK_NONE, false);
    }
#undef FE_MATCH_GO
}

int launch_hamming_match(const Geom &g, int n_pairs, const MatchParams &mp, const Buffers &b,
                         const uint32_t *counts, cudaStream_t s) {
    if (mp.want_all)
        cudaMemsetAsync(b.colbest, 0xFF, sizeof(uint32_t) * (size_t)n_pairs * g.kp_cap, s);
    // Enough pairs in flight: 4 queries per lane (train words amortised 4x).  A lone pair would
    // leave most of the 148 SMs idle with 512-query CTAs, so use small CTAs there.
    if (n_pairs >= 8) launch_variant<4, 128>(g, n_pairs, mp, b, counts, s);
    else launch_variant<1, 64>(g, n_pairs, mp, b, counts, s);
    return 1;
}

// ---- finalisation: Lowe ratio (mode A) and mutual check + |dy| filter (mode B) ----------------
constexpr int FIN_THREADS = 1024;

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t *s_warp, uint32_t &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += n;
        }
        s_warp[lane] = wi - w;            // exclusive warp bases
        if (lane == 31) s_warp[32] = wi;  // grand total
    }
    __syncthreads();
    const uint32_t r = s_warp[wid] + incl - v;
    total = s_warp[32];
    __syncthreads();
    return r;
}

// Mode A: accept a row with one candidate, or with d0 < ratio*d1 (double arithmetic, strict) --
// StereoCamera.cpp:206-264, algorithm.py:838-846.  The C++ query-unique de-dup loop is a no-op
// because knnMatch returns one row per query.  Output ordered by queryIdx.
__global__ void __launch_bounds__(FIN_THREADS)
finalize_ratio_kernel(Geom g, double ratio, const uint32_t *__restrict__ counts,
                      const uint32_t *__restrict__ best, const uint32_t *__restrict__ second,
                      fe_match *__restrict__ out, uint32_t *__restrict__ n_out) {
    __shared__ uint32_t s_warp[33];
    const int pair = blockIdx.x;
    const int nq = min((int)counts[2 * pair], g.kp_cap);
    const uint32_t *b = best + (size_t)pair * g.kp_cap, *s2 = second + (size_t)pair * g.kp_cap;
    fe_match *o = out + (size_t)pair * g.kp_cap;
    uint32_t offset = 0;
    for (int base = 0; base < nq; base += FIN_THREADS) {
        const int i = base + threadIdx.x;
        bool good = false;
        uint32_t kb = KEY_NONE;
        if (i < nq) {
            kb = b[i];
            const uint32_t ks = s2[i];
            if (kb != KEY_NONE) {
                if (ks == KEY_NONE) good = true;
                else good = (double)(kb >> 16) < ratio * (double)(ks >> 16);
            }
        }
        uint32_t total;
        const uint32_t pos = offset + block_excl_scan_1024(good ? 1u : 0u, s_warp, total);
        if (good) {
            fe_match m;
            m.queryIdx = (uint32_t)i; m.trainIdx = kb & 0xFFFF; m.imgIdx = 0; m.distance = (float)(kb >> 16);
            o[pos] = m;
        }
        offset += total;
    }
    if (threadIdx.x == 0) n_out[pair] = offset;
}

// Mode B: keep (q, s) when s = argmin_t D[q,t], q = argmin_q' D[q',s] (first minima) and
// |yq - ys| <= max_dy -- live_stereo.cpp:364-377, features.py:724-733.  Ordered by queryIdx.
__global__ void __launch_bounds__(FIN_THREADS)
finalize_cross_kernel(Geom g, float max_dy, const uint32_t *__restrict__ counts,
                      const uint32_t *__restrict__ allbest, const uint32_t *__restrict__ colbest,
                      const float *__restrict__ ky, fe_match *__restrict__ out,
                      uint32_t *__restrict__ n_out) {
    __shared__ uint32_t s_warp[33];
    const int pair = blockIdx.x;
    const int nq = min((int)counts[2 * pair], g.kp_cap);
    const int nt = min((int)counts[2 * pair + 1], g.kp_cap);
    const uint32_t *ab = allbest + (size_t)pair * g.kp_cap, *cb = colbest + (size_t)pair * g.kp_cap;
    const float *qy = ky + (size_t)(2 * pair) * g.kp_cap, *ty = ky + (size_t)(2 * pair + 1) * g.kp_cap;
    fe_match *o = out + (size_t)pair * g.kp_cap;
    uint32_t offset = 0;
    for (int base = 0; base < nq; base += FIN_THREADS) {
        const int i = base + threadIdx.x;
        bool good = false;
        uint32_t kb = KEY_NONE;
        if (i < nq && nt > 0) {
            kb = ab[i];
            const uint32_t t = kb & 0xFFFF;
            good = (cb[t] & 0xFFFF) == (uint32_t)i;
            if (good && max_dy >= 0.f) good = fabsf(__fsub_rn(qy[i], ty[t])) <= max_dy;
        }
        uint32_t total;
        const uint32_t pos = offset + block_excl_scan_1024(good ? 1u : 0u, s_warp, total);
        if (good) {
            fe_match m;
            m.queryIdx = (uint32_t)i; m.trainIdx = kb & 0xFFFF; m.imgIdx = 0; m.distance = (float)(kb >> 16);
            o[pos] = m;
        }
        offset += total;
    }
    if (threadIdx.x == 0) n_out[pair] = offset;
}

int launch_finalize_ratio(const Geom &g, int n_pairs, double ratio, const Buffers &b,
                          const uint32_t *counts, cudaStream_t s) {
    finalize_ratio_kernel<<<n_pairs, FIN_THREADS, 0, s>>>(g, ratio, counts, b.best, b.second,
                                                          b.match_a, b.n_a);
    return 1;
}

int launch_finalize_cross(const Geom &g, int n_pairs, float max_dy, const Buffers &b,
                          const uint32_t *counts, cudaStream_t s) {
    finalize_cross_kernel<<<n_pairs, FIN_THREADS, 0, s>>>(g, max_dy, counts, b.allbest, b.colbest,
                                                          b.ky, b.match_b, b.n_b);
    return 1;
}

}  // namespace fe
