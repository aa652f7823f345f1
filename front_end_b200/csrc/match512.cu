// Hamming matching of 512-bit rows (BRIEF-64, FREAK): the 64-byte descriptors of the reference's descriptor comparison
// (/root/reference bin/result_ONE:25: "BRIEF_64", "FREAK"; bin/detect_node:29,45) through the same BFMatcher call sites
// as the 256-bit rows -- knnMatch(q, t, 2, mask) (src/StereoCamera.cpp:199-201, src/WindowMatcher.cpp:150-153,
// src/front_end/algorithm.py:848-853) and match() with crossCheck (src/live_stereo.cpp:240,364, features.py:724-733).
//
// Same key convention as match.cu (distance << 16 | index, unsigned min = smallest distance, ties -> lowest index; a
// 512-bit distance is <= 512, so the keys still fit), so the finalize kernels of match.cu are reused unchanged.
// All-pairs form: one query per thread with its 16 words in registers, trains broadcast from a shared-memory tile,
// 16 x (XOR, POPC) per distance.  These rows are not on the benchmark path (the headline workloads are ORB-256 and
// SURF-128); the kernel is exact and POPC-pipe bound like the 256-bit all-pairs kernel.
#include "fe_internal.cuh"

namespace fe {

namespace {

constexpr int W_TT = 128;        // train rows per shared-memory tile (8 KB)
constexpr int W_THREADS = 128;

template <int MASK>
__device__ __forceinline__ bool allowed512(float qx, float qy, float tx, float ty, const MatchParams &mp) {
    if (MASK == FE_MASK_EPIPOLAR) return fabsf(__fsub_rn(qy, ty)) <= mp.epi_threshold;
    if (MASK == FE_MASK_WINDOW) return fabsf(__fsub_rn(qx, tx)) < mp.half_w && fabsf(__fsub_rn(qy, ty)) < mp.half_h;
    return true;
}

__device__ __forceinline__ uint32_t hamming512(const uint32_t (&q)[16], const uint4 *t) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint4 v = t[i];
        d += __popc(q[4 * i] ^ v.x) + __popc(q[4 * i + 1] ^ v.y) + __popc(q[4 * i + 2] ^ v.z) + __popc(q[4 * i + 3] ^ v.w);
    }
    return d;
}

// CROSS = false: (best, second) among the mask-allowed trains.  CROSS = true: unmasked row arg-min + column arg-min.
template <int MASK, bool CROSS>
__global__ void __launch_bounds__(W_THREADS)
hamming512_kernel(Geom g, MatchParams mp, const uint32_t *__restrict__ counts, const uint8_t *__restrict__ desc,
                  const float *__restrict__ kx, const float *__restrict__ ky, uint32_t *__restrict__ out0,
                  uint32_t *__restrict__ out1) {
    __shared__ uint4 s_desc[W_TT * 4];
    __shared__ float s_tx[W_TT], s_ty[W_TT];
    __shared__ uint32_t s_col[W_TT];
    const int pair = blockIdx.y, qi = 2 * pair, ti = 2 * pair + 1;
    const int nq = min((int)counts[qi], g.kp_cap), nt = min((int)counts[ti], g.kp_cap);
    const int q0 = blockIdx.x * W_THREADS;
    if (q0 >= nq) return;
    const int qidx = q0 + threadIdx.x;
    const bool valid = qidx < nq;
    const int src = valid ? qidx : nq - 1;
    uint32_t q[16];
    {
        const uint4 *p = reinterpret_cast<const uint4 *>(desc + ((size_t)qi * g.kp_cap + src) * 64);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 v = __ldg(p + i);
            q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
        }
    }
    const float qx = kx[(size_t)qi * g.kp_cap + src], qy = __fadd_rn(ky[(size_t)qi * g.kp_cap + src], mp.q_off);
    uint32_t best = KEY_NONE, second = KEY_NONE;
    const uint4 *tdesc = reinterpret_cast<const uint4 *>(desc + (size_t)ti * g.kp_cap * 64);
    const float *tkx = kx + (size_t)ti * g.kp_cap, *tky = ky + (size_t)ti * g.kp_cap;
    for (int t0 = 0; t0 < nt; t0 += W_TT) {
        const int tn = min(W_TT, nt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 4; i += W_THREADS) s_desc[i] = __ldg(tdesc + (size_t)t0 * 4 + i);
        for (int i = threadIdx.x; i < tn; i += W_THREADS) {
            if (CROSS) s_col[i] = KEY_NONE;
            else { s_tx[i] = tkx[t0 + i]; s_ty[i] = __fadd_rn(tky[t0 + i], mp.t_off); }
        }
        __syncthreads();
        for (int t = 0; t < tn; ++t) {
            const uint32_t d = hamming512(q, s_desc + 4 * t);
            if (CROSS) {
                best = min(best, (d << 16) | (uint32_t)(t0 + t));
                // column arg-min: the warp's smallest (distance, query) key, one shared-memory atomic per warp
                const uint32_t ck = __reduce_min_sync(0xffffffffu, valid ? ((d << 16) | (uint32_t)qidx) : KEY_NONE);
                if ((threadIdx.x & 31) == 0) atomicMin(&s_col[t], ck);
            } else if (allowed512<MASK>(qx, qy, s_tx[t], s_ty[t], mp)) {
                const uint32_t key = (d << 16) | (uint32_t)(t0 + t);
                second = min(second, max(best, key));
                best = min(best, key);
            }
        }
        if (CROSS) {
            __syncthreads();
            for (int i = threadIdx.x; i < tn; i += W_THREADS)
                if (s_col[i] != KEY_NONE) atomicMin(&out1[(size_t)pair * g.kp_cap + t0 + i], s_col[i]);
        }
    }
    if (!valid) return;
    out0[(size_t)pair * g.kp_cap + qidx] = best;
    if (!CROSS) out1[(size_t)pair * g.kp_cap + qidx] = second;
}

}  // namespace

int launch_hamming512_knn2(const Geom &g, int n_pairs, const MatchParams &mp, const uint8_t *desc64, const Buffers &b,
                           const uint32_t *counts, cudaStream_t s) {
    dim3 grid(div_up(g.kp_cap, W_THREADS), n_pairs);
    if (mp.mask == FE_MASK_EPIPOLAR)
        hamming512_kernel<FE_MASK_EPIPOLAR, false><<<grid, W_THREADS, 0, s>>>(g, mp, counts, desc64, b.kx, b.ky, b.best, b.second);
    else if (mp.mask == FE_MASK_WINDOW)
        hamming512_kernel<FE_MASK_WINDOW, false><<<grid, W_THREADS, 0, s>>>(g, mp, counts, desc64, b.kx, b.ky, b.best, b.second);
    else
        hamming512_kernel<FE_MASK_NONE, false><<<grid, W_THREADS, 0, s>>>(g, mp, counts, desc64, b.kx, b.ky, b.best, b.second);
    return 1;
}

int launch_hamming512_cross(const Geom &g, int n_pairs, const uint8_t *desc64, const Buffers &b, const uint32_t *counts, cudaStream_t s) {
    cudaMemsetAsync(b.colbest, 0xFF, sizeof(uint32_t) * (size_t)n_pairs * g.kp_cap, s);
    dim3 grid(div_up(g.kp_cap, W_THREADS), n_pairs);
    hamming512_kernel<FE_MASK_NONE, true><<<grid, W_THREADS, 0, s>>>(g, MatchParams{}, counts, desc64, b.kx, b.ky, b.allbest, b.colbest);
    return 1;
}

}  // namespace fe
