// Unmasked float-L2 matching as a tcgen05 tensor-core GEMM (the only dense contraction on the path).
//
// |q - t|^2 = |q|^2 + |t|^2 - 2 q.t : the N x N inner products of a stereo pair's SURF descriptors are a
// K = 64/128 GEMM.  For BFMatcher(NORM_L2, crossCheck=true)::match (/root/reference src/live_stereo.cpp:240,364
// with float descriptors; features.py:463-467, 670, 724) and unmasked knnMatch, the tensor cores produce the
// CANDIDATES and CUDA cores decide: bf16 operands (8-bit mantissa) perturb d^2 by ~1e-3, so each row keeps
// its 4 smallest approximate distances, and l2_rerank_kernel re-evaluates those 4 exactly in FP32 with the
// same k-ascending sum((q_k - t_k)^2) as the all-pairs FP32 kernel (l2match.cu).  The final arg-min / second
// min and the reported distances are therefore FP32-exact; only a true neighbour ranked below 4th by the
// bf16 pass could be missed.  Column arg-mins (cross-check) come from the same kernel with the roles of the
// two images swapped.
//
// Kernel anatomy (one CTA = 128 query rows = the 128 TMEM lanes, 4 warps):
//   l2_prep_kernel  : fp32 rows -> bf16 in the UMMA canonical K-major, no-swizzle core-matrix layout, tile by
//                     tile ([tile][k/8][128 rows][8 elements]), plus |x|^2.  A tile is 32 KB contiguous, so the
//                     GEMM kernel fills shared memory with a linear 16-byte copy.
//   l2_tc_topk_kernel: A tile resident in shared memory; for every 128-train tile: copy B, fence.proxy.async,
//                     one thread issues K/16 tcgen05.mma (M128 N128 K16, kind::f16, fp32 accumulate in 128 TMEM
//                     columns) + tcgen05.commit -> mbarrier; then each warp tcgen05.ld's its 32 lanes and updates
//                     the per-row top-4 of t_norm - 2 s.  Two or three CTAs are resident per SM, so one CTA's copy /
//                     epilogue overlaps another's MMAs.
// Every mbarrier wait is bounded: on timeout the kernel raises an error flag and exits instead of hanging.
#include <cuda_bf16.h>

#include "fe_internal.cuh"

namespace fe {

constexpr int TC_M = 128;          // rows per tile (queries per CTA, trains per stage)
constexpr int TC_TOPK = 4;
constexpr uint32_t TC_LBO = TC_M * 16;   // bytes between core matrices adjacent in K (one 8-element chunk of all rows)
constexpr uint32_t TC_SBO = 128;         // bytes between core matrices adjacent in M/N (8 rows x 16 B)

// ---- operand preparation -------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(TC_M)
l2_prep_kernel(Geom g, const uint32_t *__restrict__ counts, const float *__restrict__ fdesc,
               uint4 *__restrict__ bf, float *__restrict__ norms, int tiles_per_image) {
    const int image = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
    const int n = min((int)counts[image], g.kp_cap);
    const int row = tile * TC_M + r;
    uint4 *dst = bf + ((size_t)image * tiles_per_image + tile) * (D / 8) * TC_M;
    float nrm = 0.f;
    if (row < n) {
        const float4 *src = reinterpret_cast<const float4 *>(fdesc + ((size_t)image * g.kp_cap + row) * 128);
#pragma unroll 4
        for (int kc = 0; kc < D / 8; ++kc) {
            const float4 a = __ldg(src + 2 * kc), b = __ldg(src + 2 * kc + 1);
            const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat16 lo = __float2bfloat16_rn(v[2 * e]), hi = __float2bfloat16_rn(v[2 * e + 1]);
                w[e] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
                // norms of the ROUNDED operands, so that t_norm - 2 s is a consistent distance estimate
                const float fl = __bfloat162float(lo), fh = __bfloat162float(hi);
                nrm = __fmaf_rn(fl, fl, nrm);
                nrm = __fmaf_rn(fh, fh, nrm);
            }
            dst[kc * TC_M + r] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        for (int kc = 0; kc < D / 8; ++kc) dst[kc * TC_M + r] = make_uint4(0, 0, 0, 0);
    }
    if (row < tiles_per_image * TC_M) norms[(size_t)image * tiles_per_image * TC_M + row] = nrm;
}

// ---- PTX helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46),
    // version = 1 [46,48), layout_type = SWIZZLE_NONE (0) [61,64)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) |
           (1ull << 46);
}

// InstrDescriptor: c_format F32 (1) [4,6), a/b_format BF16 (1) [7,10)/[10,13), K-major A and B, N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_M >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t mbar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ---- GEMM + per-row top-4 ----------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(TC_M)
l2_tc_topk_kernel(Geom g, const uint32_t *__restrict__ counts, const uint4 *__restrict__ bf,
                  const float *__restrict__ norms, int tiles_per_image, uint32_t *__restrict__ cand,
                  int *__restrict__ error_flag) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    constexpr int TILE_BYTES = D * TC_M * 2;                // 32 KB (D = 128) / 16 KB (D = 64)
    uint4 *sA = reinterpret_cast<uint4 *>(tc_smem);
    uint4 *sB = reinterpret_cast<uint4 *>(tc_smem + TILE_BYTES);
    __shared__ float s_tn[TC_M];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;

    const int pair = blockIdx.y, dir = blockIdx.z;
    const int qi = 2 * pair + dir, ti = 2 * pair + 1 - dir;
    const int nq = min((int)counts[qi], g.kp_cap), nt = min((int)counts[ti], g.kp_cap);
    const int q0 = blockIdx.x * TC_M;
    if (q0 >= nq) return;                                    // uniform: before any allocation
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(TC_M));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&s_mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    {
        const uint4 *gA = bf + ((size_t)qi * tiles_per_image + blockIdx.x) * (D / 8) * TC_M;
        for (int i = tid; i < TILE_BYTES / 16; i += TC_M) sA[i] = __ldg(gA + i);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t mbar = smem_u32(&s_mbar);
    const uint64_t adesc0 = umma_desc(smem_u32(sA)), bdesc0 = umma_desc(smem_u32(sB));

    float v[TC_TOPK];
    uint32_t vi[TC_TOPK];
#pragma unroll
    for (int e = 0; e < TC_TOPK; ++e) { v[e] = __int_as_float(0x7f800000); vi[e] = 0xFFFFFFFFu; }

    const int n_tiles = div_up(nt, TC_M);
    bool ok = true;
    for (int j = 0; j < n_tiles && ok; ++j) {
        const uint4 *gB = bf + ((size_t)ti * tiles_per_image + j) * (D / 8) * TC_M;
        for (int i = tid; i < TILE_BYTES / 16; i += TC_M) sB[i] = __ldg(gB + i);
        s_tn[tid] = norms[(size_t)ti * tiles_per_image * TC_M + j * TC_M + tid];
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
                // one K = 16 step = two 8-element core-matrix columns = 2 * LBO bytes
                const uint64_t step = (uint64_t)((k * 2 * TC_LBO) >> 4);
                umma_bf16(tmem, adesc0 + step, bdesc0 + step, k > 0 ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
        }
        ok = mbar_wait_bounded(mbar, (uint32_t)(j & 1));
        ok = __syncthreads_and(ok ? 1 : 0) != 0;
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < TC_M; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int col = j * TC_M + c0 + i;
                const float x = __fmaf_rn(-2.f, __uint_as_float(r[i]), s_tn[c0 + i]);
                if (col < nt && x < v[TC_TOPK - 1]) {
                    // insert into the ascending list (strict <: equal values keep the earlier, lower index)
                    v[TC_TOPK - 1] = x; vi[TC_TOPK - 1] = (uint32_t)col;
#pragma unroll
                    for (int e = TC_TOPK - 1; e > 0; --e)
                        if (v[e] < v[e - 1]) {
                            const float tv = v[e]; v[e] = v[e - 1]; v[e - 1] = tv;
                            const uint32_t tix = vi[e]; vi[e] = vi[e - 1]; vi[e - 1] = tix;
                        }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();            // TMEM and sB are free for the next tile
    }
    if (!ok && tid == 0) atomicExch(error_flag, 1);
    const int q = q0 + tid;
    if (ok && q < nq) {
        uint32_t *o = cand + (((size_t)pair * 2 + dir) * g.kp_cap + q) * TC_TOPK;
#pragma unroll
        for (int e = 0; e < TC_TOPK; ++e) o[e] = vi[e];
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TC_M));
}

// ---- exact FP32 re-rank of the candidates -------------------------------------------------------------------
// 8 rows per warp: lane = (row slot, candidate); each lane evaluates one exact distance sequentially.
template <int D>
__global__ void __launch_bounds__(256)
l2_rerank_kernel(Geom g, const uint32_t *__restrict__ counts, const float *__restrict__ fdesc,
                 const uint32_t *__restrict__ cand, unsigned long long *__restrict__ best64,
                 unsigned long long *__restrict__ second64, unsigned long long *__restrict__ allbest64,
                 unsigned long long *__restrict__ colbest64) {
    const int pair = blockIdx.y, dir = blockIdx.z;
    const int qi = 2 * pair + dir, ti = 2 * pair + 1 - dir;
    const int nq = min((int)counts[qi], g.kp_cap);
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 8 + (lane >> 2);
    const int e = lane & 3;
    unsigned long long key = 0xFFFFFFFFFFFFFFFFull;
    if (q < nq) {
        const uint32_t t = cand[(((size_t)pair * 2 + dir) * g.kp_cap + q) * TC_TOPK + e];
        if (t != 0xFFFFFFFFu) {
            const float4 *a = reinterpret_cast<const float4 *>(fdesc + ((size_t)qi * g.kp_cap + q) * 128);
            const float4 *b = reinterpret_cast<const float4 *>(fdesc + ((size_t)ti * g.kp_cap + t) * 128);
            float acc = 0.f;
#pragma unroll 4
            for (int k = 0; k < D / 4; ++k) {
                const float4 x = __ldg(a + k), y = __ldg(b + k);
                float df = __fsub_rn(x.x, y.x); acc = __fmaf_rn(df, df, acc);
                df = __fsub_rn(x.y, y.y); acc = __fmaf_rn(df, df, acc);
                df = __fsub_rn(x.z, y.z); acc = __fmaf_rn(df, df, acc);
                df = __fsub_rn(x.w, y.w); acc = __fmaf_rn(df, df, acc);
            }
            key = ((unsigned long long)__float_as_uint(acc) << 32) | t;
        }
    }
    // best / second over the 4 lanes of a row
    unsigned long long best = key, second = 0xFFFFFFFFFFFFFFFFull;
#pragma unroll
    for (int off = 1; off < 4; off <<= 1) {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, off);
        const unsigned long long os = __shfl_xor_sync(0xffffffffu, second, off);
        second = min(min(second, os), max(best, ob));
        best = min(best, ob);
    }
    if (q < nq && e == 0) {
        const size_t o = (size_t)pair * g.kp_cap + q;
        if (dir == 0) {
            best64[o] = best; second64[o] = second; allbest64[o] = best;
        } else {
            // column arg-min: key carries the QUERY (left) index, i.e. this pass's candidate
            colbest64[o] = best;
        }
    }
}

template <int D>
static int launch_l2_tc_d(const Geom &g, int n_pairs, const Buffers &b, const uint32_t *counts, cudaStream_t s) {
    const int tiles = div_up(g.kp_cap, TC_M);
    dim3 pgrid(tiles, 2 * n_pairs);
    l2_prep_kernel<D><<<pgrid, TC_M, 0, s>>>(g, counts, b.fdesc, reinterpret_cast<uint4 *>(b.bf16desc), b.fnorm, tiles);
    const size_t smem = (size_t)2 * D * TC_M * 2;
    cudaFuncSetAttribute(l2_tc_topk_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(tiles, n_pairs, 2);
    l2_tc_topk_kernel<D><<<grid, TC_M, smem, s>>>(g, counts, reinterpret_cast<const uint4 *>(b.bf16desc), b.fnorm, tiles,
                                                  b.cand, b.tc_error);
    dim3 rgrid(div_up(g.kp_cap, 64), n_pairs, 2);
    l2_rerank_kernel<D><<<rgrid, 256, 0, s>>>(g, counts, b.fdesc, b.cand, b.best64, b.second64, b.allbest64, b.colbest64);
    return 3;
}

// Unmasked L2: best64 / second64 / allbest64 (row side) and colbest64 (column side) for every pair.
int launch_l2_tensor(const Geom &g, int n_pairs, int dim, const Buffers &b, const uint32_t *counts, cudaStream_t s) {
    return dim == 64 ? launch_l2_tc_d<64>(g, n_pairs, b, counts, s) : launch_l2_tc_d<128>(g, n_pairs, b, counts, s);
}

}  // namespace fe
