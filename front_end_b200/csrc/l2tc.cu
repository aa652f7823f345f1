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
#include <cstdio>
#include <cstdlib>

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
    // rows past the last keypoint: zero operand, +inf norm -> never a candidate
    norms[(size_t)image * tiles_per_image * TC_M + row] = row < n ? nrm : __int_as_float(0x7f800000);
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
// CTA = 256 threads = 256 query rows: two 128-row A tiles resident in shared memory, two 128-column fp32
// accumulators in TMEM (256 columns).  Each 128-train B tile is copied once and used by both A tiles
// (halves the L2 -> shared traffic per FLOP).  Warp w reads accumulator w / 4, TMEM lanes 32 (w % 4)...

__device__ __forceinline__ void topk_insert(float (&v)[TC_TOPK], uint32_t (&vi)[TC_TOPK], float x, uint32_t col) {
    // ascending list; strict <: equal values keep the earlier, lower index
    v[TC_TOPK - 1] = x; vi[TC_TOPK - 1] = col;
#pragma unroll
    for (int e = TC_TOPK - 1; e > 0; --e)
        if (v[e] < v[e - 1]) {
            const float tv = v[e]; v[e] = v[e - 1]; v[e - 1] = tv;
            const uint32_t tix = vi[e]; vi[e] = vi[e - 1]; vi[e - 1] = tix;
        }
}

// 512 threads: 16 warps.  Warp w reads TMEM lanes 32 (w % 4)... of accumulator (w / 4) % 2, columns
// 64 (w / 8)... : two threads share a query row (one per column half) and merge their top-4 lists at the end.
// Two CTAs are resident per SM (96 KB of shared memory, 256 TMEM columns each), so one CTA's copy / epilogue
// overlaps the other's MMAs, and 32 warps hide the epilogue's dependent-issue latency.
constexpr int TC_THREADS2 = 512;

template <int D>
__global__ void __launch_bounds__(TC_THREADS2, 2)
l2_tc_topk_kernel(Geom g, const uint32_t *__restrict__ counts, const uint4 *__restrict__ bf,
                  const float *__restrict__ norms, int tiles_per_image, uint32_t *__restrict__ cand,
                  int *__restrict__ error_flag) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    constexpr int TILE_BYTES = D * TC_M * 2;                // 32 KB (D = 128) / 16 KB (D = 64)
    uint4 *sA = reinterpret_cast<uint4 *>(tc_smem);         // two A tiles back to back
    uint4 *sB = reinterpret_cast<uint4 *>(tc_smem + 2 * TILE_BYTES);
    __shared__ __align__(16) float s_tn[TC_M];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;

    const int pair = blockIdx.y, dir = blockIdx.z;
    const int qi = 2 * pair + dir, ti = 2 * pair + 1 - dir;
    const int nq = min((int)counts[qi], g.kp_cap), nt = min((int)counts[ti], g.kp_cap);
    const int q0 = blockIdx.x * (2 * TC_M);
    if (q0 >= nq) return;                                    // uniform: before any allocation
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = div_up(nt, TC_M);
    const uint4 *gBall = bf + (size_t)ti * tiles_per_image * (D / 8) * TC_M;
    const float *gnorm = norms + (size_t)ti * tiles_per_image * TC_M;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(2 * TC_M));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&s_mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    {
        // tiles 2 * blockIdx.x and 2 * blockIdx.x + 1 are contiguous in global memory (the second may lie past
        // the last valid row: prep zero-fills every tile up to tiles_per_image, which is even)
        const uint4 *gA = bf + ((size_t)qi * tiles_per_image + 2 * blockIdx.x) * (D / 8) * TC_M;
        for (int i = tid; i < 2 * TILE_BYTES / 16; i += TC_THREADS2) sA[i] = __ldg(gA + i);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t mbar = smem_u32(&s_mbar);
    const uint64_t adesc0 = umma_desc(smem_u32(sA)), adesc1 = umma_desc(smem_u32(sA) + TILE_BYTES),
                   bdesc0 = umma_desc(smem_u32(sB));
    const int a_tile = (warp >> 2) & 1, chalf = warp >> 3;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a_tile * TC_M + chalf * 64);

    float v[TC_TOPK];
    uint32_t vi[TC_TOPK];
#pragma unroll
    for (int e = 0; e < TC_TOPK; ++e) { v[e] = __int_as_float(0x7f800000); vi[e] = 0xFFFFFFFFu; }

    bool ok = true;
#ifdef FE_TC_TIMING
    long long t_copy = 0, t_mma = 0, t_epi = 0, t_sync = 0, t0 = clock64();
#define TC_TICK(acc) do { const long long t1_ = clock64(); acc += t1_ - t0; t0 = t1_; } while (0)
#else
#define TC_TICK(acc) do { } while (0)
#endif
    for (int j = 0; j < n_tiles; ++j) {
        const uint4 *gB = gBall + (size_t)j * (D / 8) * TC_M;
        for (int i = tid; i < TILE_BYTES / 16; i += TC_THREADS2) sB[i] = __ldg(gB + i);
        if (tid < TC_M) s_tn[tid] = gnorm[(size_t)j * TC_M + tid];          // +inf past nt
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        TC_TICK(t_copy);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
                // one K = 16 step = two 8-element core-matrix columns = 2 * LBO bytes
                const uint64_t step = (uint64_t)((k * 2 * TC_LBO) >> 4);
                umma_bf16(tmem, adesc0 + step, bdesc0 + step, k > 0 ? 1u : 0u);
                umma_bf16(tmem + TC_M, adesc1 + step, bdesc0 + step, k > 0 ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
        }
        ok = mbar_wait_bounded(mbar, (uint32_t)(j & 1));
        if (__syncthreads_and(ok ? 1 : 0) == 0) { ok = false; break; }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        TC_TICK(t_mma);
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)c0, r);
            const float *tn = s_tn + chalf * 64 + c0;
#pragma unroll
            for (int g8 = 0; g8 < 32; g8 += 8) {
                // x = |t|^2 - 2 q.t (the row constant |q|^2 does not change the ranking); columns past nt are +inf
                const float4 ta = *reinterpret_cast<const float4 *>(&tn[g8]);
                const float4 tb = *reinterpret_cast<const float4 *>(&tn[g8 + 4]);
                float x[8];
                x[0] = __fmaf_rn(-2.f, __uint_as_float(r[g8 + 0]), ta.x); x[1] = __fmaf_rn(-2.f, __uint_as_float(r[g8 + 1]), ta.y);
                x[2] = __fmaf_rn(-2.f, __uint_as_float(r[g8 + 2]), ta.z); x[3] = __fmaf_rn(-2.f, __uint_as_float(r[g8 + 3]), ta.w);
                x[4] = __fmaf_rn(-2.f, __uint_as_float(r[g8 + 4]), tb.x); x[5] = __fmaf_rn(-2.f, __uint_as_float(r[g8 + 5]), tb.y);
                x[6] = __fmaf_rn(-2.f, __uint_as_float(r[g8 + 6]), tb.z); x[7] = __fmaf_rn(-2.f, __uint_as_float(r[g8 + 7]), tb.w);
                const float m = fminf(fminf(fminf(x[0], x[1]), fminf(x[2], x[3])), fminf(fminf(x[4], x[5]), fminf(x[6], x[7])));
                if (m < v[TC_TOPK - 1]) {          // rare after the first tiles: one compare per 8 columns otherwise
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (x[e] < v[TC_TOPK - 1])
                            topk_insert(v, vi, x[e], (uint32_t)(j * TC_M + chalf * 64 + c0 + g8 + e));
                }
            }
        }
        TC_TICK(t_epi);
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();            // TMEM and sB are free for the next tile
        TC_TICK(t_sync);
    }
#ifdef FE_TC_TIMING
    if (blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && (tid == 0 || tid == 300))
        printf("tc timing tid %d tiles %d: copy %lld mma %lld epi %lld sync %lld cycles/tile\n", tid, n_tiles,
               t_copy / n_tiles, t_mma / n_tiles, t_epi / n_tiles, t_sync / n_tiles);
#endif
    if (!ok && tid == 0) atomicExch(error_flag, 1);
    // merge the two column halves of every row (the lower half holds the lower indices: on equal values it wins)
    float *mv = reinterpret_cast<float *>(sB);                       // [256][4]
    uint32_t *mi = reinterpret_cast<uint32_t *>(sB) + 256 * TC_TOPK;  // [256][4]
    const int row = a_tile * TC_M + (warp & 3) * 32 + lane;
    if (chalf == 1) {
#pragma unroll
        for (int e = 0; e < TC_TOPK; ++e) { mv[row * TC_TOPK + e] = v[e]; mi[row * TC_TOPK + e] = vi[e]; }
    }
    __syncthreads();
    const int q = q0 + row;
    if (ok && chalf == 0 && q < nq) {
#pragma unroll
        for (int e = 0; e < TC_TOPK; ++e) {
            const float x = mv[row * TC_TOPK + e];
            if (x < v[TC_TOPK - 1]) topk_insert(v, vi, x, mi[row * TC_TOPK + e]);
        }
        uint32_t *o = cand + (((size_t)pair * 2 + dir) * g.kp_cap + q) * TC_TOPK;
#pragma unroll
        for (int e = 0; e < TC_TOPK; ++e) o[e] = vi[e];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(2 * TC_M));
}

// ---- exact FP32 re-rank of the candidates -------------------------------------------------------------------
// One warp per row: the 4 candidate rows are read coalesced and measured with the warp-cooperative exact
// FP32 distance (WarpRow, the same definition the banded kernel uses).
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
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    WarpRow<D> qr;
    qr.load(fdesc + ((size_t)qi * g.kp_cap + q) * 128, lane);
    const uint32_t *cq = cand + (((size_t)pair * 2 + dir) * g.kp_cap + q) * TC_TOPK;
    unsigned long long best = 0xFFFFFFFFFFFFFFFFull, second = 0xFFFFFFFFFFFFFFFFull;
#pragma unroll
    for (int e = 0; e < TC_TOPK; ++e) {
        const uint32_t t = cq[e];
        if (t == 0xFFFFFFFFu) continue;
        const float d2 = qr.dist2(fdesc + ((size_t)ti * g.kp_cap + t) * 128, lane);
        const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | t;
        second = min(second, max(best, key));
        best = min(best, key);
    }
    if (lane == 0) {
        const size_t o = (size_t)pair * g.kp_cap + q;
        if (dir == 0) {
            best64[o] = best; second64[o] = second; allbest64[o] = best;
        } else {
            colbest64[o] = best;      // column arg-min: the key carries the QUERY (left) index
        }
    }
}

// =====================================================================================================
// Pipelined variant (default): warp-specialised, bulk-async copies, double-buffered TMEM, norm folded into K.
//
//   * Operand layout (l2_prep2_kernel): per 128-row tile  [D/8 data chunks | B-aug | 0 | A-aug | 0]  of 2 KB core-
//     matrix columns.  B-aug holds -|t|^2/2 split into three bf16 (hi, mid, lo: ~24 bits), A-aug holds (1, 1, 1): one
//     extra K = 16 MMA step pairing A's A-aug with B's B-aug makes the accumulator  q.t - |t|^2/2, i.e. the ranking
//     key of |q - t|^2 up to the row constant -- the epilogue is a pure running max, no FMA and no norm reads.  Rows
//     past the last keypoint carry -1e30 in B-aug (never selected) and zeros in A-aug.
//   * One CTA per SM (512 TMEM columns = 2 accumulator stages x 2 A tiles x 128 columns; ~200 KB shared memory):
//       warp 0 lane 0   producer: cp.async.bulk global -> shared of whole tiles (a tile in HBM is the shared image),
//                       mbarrier complete_tx, 3-stage ring (4 for D = 64);
//       warp 1 lane 0   MMA issuer: (D/16 + 1) x 2 tcgen05.mma per tile step, tcgen05.commit frees the shared stage
//                       and publishes the accumulator stage;
//       warps 4..19     epilogue: warp w owns TMEM lanes 32 (w % 4).., A tile (w - 4) / 4 % 2, column half (w - 4) / 8;
//                       tcgen05.ld 64 columns to registers, release the TMEM stage at once, then the top-4 update;
//                       the MMAs of tile j + 1 run under the epilogue of tile j.
//   * Every mbarrier wait is bounded; a timeout raises the error flag and the CTA drains.
constexpr int TP_THREADS = 640;
constexpr int TP_EPI_WARPS = 16;

template <int D>
__global__ void __launch_bounds__(TC_M)
l2_prep2_kernel(Geom g, const uint32_t *__restrict__ counts, const float *__restrict__ fdesc,
                uint4 *__restrict__ bf, int tiles_per_image) {
    constexpr int KC = D / 8 + 4;
    const int image = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
    const int n = min((int)counts[image], g.kp_cap);
    const int row = tile * TC_M + r;
    uint4 *dst = bf + ((size_t)image * tiles_per_image + tile) * KC * TC_M;
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
                const float fl = __bfloat162float(lo), fh = __bfloat162float(hi);   // norm of the ROUNDED operand
                nrm = __fmaf_rn(fl, fl, nrm);
                nrm = __fmaf_rn(fh, fh, nrm);
            }
            dst[kc * TC_M + r] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        for (int kc = 0; kc < D / 8; ++kc) dst[kc * TC_M + r] = make_uint4(0, 0, 0, 0);
    }
    // B-aug: -|t|^2 / 2 as hi + mid + lo bf16; A-aug: (1, 1, 1)
    const float sv = row < n ? __fmul_rn(-0.5f, nrm) : -1e30f;
    const __nv_bfloat16 h = __float2bfloat16_rn(sv);
    const float r1 = __fsub_rn(sv, __bfloat162float(h));
    const __nv_bfloat16 m = __float2bfloat16_rn(row < n ? r1 : 0.f);
    const float r2 = __fsub_rn(r1, __bfloat162float(m));
    const __nv_bfloat16 l = __float2bfloat16_rn(row < n ? r2 : 0.f);
    dst[(D / 8) * TC_M + r] = make_uint4((uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(m) << 16),
                                         (uint32_t)__bfloat16_as_ushort(l), 0u, 0u);
    dst[(D / 8 + 1) * TC_M + r] = make_uint4(0, 0, 0, 0);
    const uint32_t one = 0x3F80u;      // bf16 1.0
    dst[(D / 8 + 2) * TC_M + r] = row < n ? make_uint4(one | (one << 16), one, 0u, 0u) : make_uint4(0, 0, 0, 0);
    dst[(D / 8 + 3) * TC_M + r] = make_uint4(0, 0, 0, 0);
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t *r) {
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
}

// descending list (largest first); strict >: equal values keep the earlier, lower index
__device__ __forceinline__ void topk_insert_max(float (&v)[TC_TOPK], uint32_t (&vi)[TC_TOPK], float x, uint32_t col) {
    v[TC_TOPK - 1] = x; vi[TC_TOPK - 1] = col;
#pragma unroll
    for (int e = TC_TOPK - 1; e > 0; --e)
        if (v[e] > v[e - 1]) {
            const float tv = v[e]; v[e] = v[e - 1]; v[e - 1] = tv;
            const uint32_t tix = vi[e]; vi[e] = vi[e - 1]; vi[e - 1] = tix;
        }
}

template <int D>
__global__ void __launch_bounds__(TP_THREADS, 1)
l2_tc_pipe_kernel(Geom g, const uint32_t *__restrict__ counts, const uint4 *__restrict__ bf, int tiles_per_image,
                  uint32_t *__restrict__ cand, int *__restrict__ error_flag) {
    constexpr int KC = D / 8 + 4;
    constexpr uint32_t TILE_BYTES = KC * TC_M * 16;           // 40 KB (D = 128) / 24 KB (D = 64)
    constexpr int NST = D == 128 ? 3 : 4;
    extern __shared__ __align__(1024) uint8_t tp_smem[];
    __shared__ __align__(8) uint64_t s_full[NST], s_empty[NST], s_tfull[2], s_tempty[2], s_afull;
    __shared__ uint32_t s_tmem;

    const int pair = blockIdx.y, dir = blockIdx.z;
    const int qi = 2 * pair + dir, ti = 2 * pair + 1 - dir;
    const int nq = min((int)counts[qi], g.kp_cap), nt = min((int)counts[ti], g.kp_cap);
    const int q0 = blockIdx.x * (2 * TC_M);
    if (q0 >= nq) return;                                    // uniform: before any allocation
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = div_up(nt, TC_M);
    const uint32_t sA = smem_u32(tp_smem), sB = sA + 2 * TILE_BYTES;

    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&s_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_tfull[i]), 1); mbar_init(smem_u32(&s_tempty[i]), TP_EPI_WARPS); }
        mbar_init(smem_u32(&s_afull), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    bool ok = true;

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer =====
            const uint8_t *gA = reinterpret_cast<const uint8_t *>(bf) + ((size_t)qi * tiles_per_image + 2 * blockIdx.x) * TILE_BYTES;
            const uint8_t *gB = reinterpret_cast<const uint8_t *>(bf) + (size_t)ti * tiles_per_image * TILE_BYTES;
            mbar_expect_tx(smem_u32(&s_afull), 2 * TILE_BYTES);
            bulk_g2s(sA, gA, TILE_BYTES, smem_u32(&s_afull));
            bulk_g2s(sA + TILE_BYTES, gA + TILE_BYTES, TILE_BYTES, smem_u32(&s_afull));
#ifdef FE_TC_TIMING
            long long tw = 0, t00 = clock64();
#endif
            for (int j = 0; j < n_tiles && ok; ++j) {
                const int st = j % NST;
                const uint32_t ph = (uint32_t)(j / NST) & 1u;
#ifdef FE_TC_TIMING
                const long long ta = clock64();
#endif
                ok = mbar_wait_bounded(smem_u32(&s_empty[st]), ph ^ 1u);
#ifdef FE_TC_TIMING
                tw += clock64() - ta;
#endif
                if (!ok) break;
                mbar_expect_tx(smem_u32(&s_full[st]), TILE_BYTES);
                bulk_g2s(sB + st * TILE_BYTES, gB + (size_t)j * TILE_BYTES, TILE_BYTES, smem_u32(&s_full[st]));
            }
#ifdef FE_TC_TIMING
            if (blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0)
                printf("producer: tiles %d total %lld wait_empty %lld cycles/tile\n", n_tiles, (clock64() - t00) / n_tiles, tw / n_tiles);
#endif
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            ok = mbar_wait_bounded(smem_u32(&s_afull), 0);
            const uint64_t adesc0 = umma_desc(sA), adesc1 = umma_desc(sA + TILE_BYTES);
            constexpr uint64_t KSTEP = (uint64_t)((2 * TC_LBO) >> 4);          // one K = 16 step = two core-matrix columns
            constexpr uint64_t AUG_B = (uint64_t)(((D / 8) * TC_LBO) >> 4), AUG_A = (uint64_t)(((D / 8 + 2) * TC_LBO) >> 4);
#ifdef FE_TC_TIMING
            long long twf = 0, twe = 0, t00 = clock64();
#endif
            for (int j = 0; j < n_tiles && ok; ++j) {
                const int st = j % NST, acc = j & 1;
#ifdef FE_TC_TIMING
                const long long ta = clock64();
#endif
                ok = mbar_wait_bounded(smem_u32(&s_full[st]), (uint32_t)(j / NST) & 1u);
#ifdef FE_TC_TIMING
                const long long tb = clock64();
                twf += tb - ta;
#endif
                if (ok) ok = mbar_wait_bounded(smem_u32(&s_tempty[acc]), ((uint32_t)(j >> 1) & 1u) ^ 1u);
#ifdef FE_TC_TIMING
                twe += clock64() - tb;
#endif
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint64_t bdesc = umma_desc(sB + st * TILE_BYTES);
                const uint32_t d0 = tmem + (uint32_t)(acc * 256), d1 = d0 + TC_M;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    umma_bf16(d0, adesc0 + k * KSTEP, bdesc + k * KSTEP, k > 0 ? 1u : 0u);
                    umma_bf16(d1, adesc1 + k * KSTEP, bdesc + k * KSTEP, k > 0 ? 1u : 0u);
                }
                umma_bf16(d0, adesc0 + AUG_A, bdesc + AUG_B, 1u);      // + 1 * (-|t|^2 / 2)
                umma_bf16(d1, adesc1 + AUG_A, bdesc + AUG_B, 1u);
                umma_commit(smem_u32(&s_empty[st]));                   // shared stage free once these MMAs retire
                umma_commit(smem_u32(&s_tfull[acc]));                  // accumulator stage ready
            }
#ifdef FE_TC_TIMING
            if (blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0)
                printf("mma: total %lld wait_full %lld wait_tempty %lld cycles/tile\n", (clock64() - t00) / n_tiles, twf / n_tiles, twe / n_tiles);
#endif
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int ew = warp - 4, a_tile = (ew >> 2) & 1, chalf = ew >> 3;
        const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a_tile * TC_M + chalf * 64);
        // Running top-4 per (row, column half) as a branch-free max/min network on PACKED floats: the low 14 mantissa
        // bits of every value are replaced by its column index, so fmax / fmin move value and index together (9
        // mantissa bits remain -- the bf16 products carry 8).  Per tile and thread: 64 LOP3 to tag the in-tile
        // column, four 16-wide FMNMX3 trees, 4 re-tags with the global column, 4 insertions of 8 min/max each --
        // no branches, no divergence.  (A data-dependent "insert if above the 4th best" is rare per lane but fires
        // for some lane of the warp in most groups: measured 7000 cycles per tile instead of ~300.)
        float v[TC_TOPK];
#pragma unroll
        for (int e = 0; e < TC_TOPK; ++e) v[e] = -1e30f;
        const uint32_t keep6 = 0xFFFFFFC0u;
#ifdef FE_TC_TIMING
        long long twt = 0, tld = 0, tcmp = 0, t00 = clock64();
#endif
        for (int j = 0; j < n_tiles && ok; ++j) {
            const int acc = j & 1;
#ifdef FE_TC_TIMING
            const long long ta = clock64();
#endif
            ok = mbar_wait_bounded(smem_u32(&s_tfull[acc]), (uint32_t)(j >> 1) & 1u);
#ifdef FE_TC_TIMING
            const long long tb = clock64();
            twt += tb - ta;
#endif
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            uint32_t r[64];
            tmem_ld32_nowait(tbase + (uint32_t)(acc * 256), r);
            tmem_ld32_nowait(tbase + (uint32_t)(acc * 256 + 32), r + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_tempty[acc]));      // TMEM stage released: values are in registers
#ifdef FE_TC_TIMING
            const long long tc = clock64();
            tld += tc - tb;
#endif
            const uint32_t colbase = (uint32_t)(j * TC_M + chalf * 64);
#pragma unroll
            for (int g16 = 0; g16 < 64; g16 += 16) {
                float p[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) p[e] = __uint_as_float((r[g16 + e] & keep6) | (uint32_t)(g16 + e));
                const float m = fmaxf(fmaxf(fmaxf(fmaxf(p[0], p[1]), p[2]), fmaxf(fmaxf(p[3], p[4]), p[5])),
                                      fmaxf(fmaxf(fmaxf(fmaxf(p[6], p[7]), p[8]), fmaxf(fmaxf(p[9], p[10]), p[11])),
                                            fmaxf(fmaxf(fmaxf(p[12], p[13]), p[14]), p[15])));
                const uint32_t mb = __float_as_uint(m);
                float t = __uint_as_float((mb & 0xFFFFC000u) | (colbase + (mb & 63u)));
#pragma unroll
                for (int e = 0; e < TC_TOPK; ++e) {
                    const float hi = fmaxf(v[e], t);
                    t = fminf(v[e], t);
                    v[e] = hi;
                }
            }
#ifdef FE_TC_TIMING
            tcmp += clock64() - tc;
#endif
        }
#ifdef FE_TC_TIMING
        if (blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && (tid == 128 || tid == 639))
            printf("epi tid %d: total %lld wait_tfull %lld ld %lld compute %lld cycles/tile\n", tid, (clock64() - t00) / n_tiles,
                   twt / n_tiles, tld / n_tiles, tcmp / n_tiles);
#endif
        // merge the two column halves of every row.  All MMAs have retired (the last tfull was consumed), so the B
        // ring can be reused as scratch.
        asm volatile("bar.sync 1, %0;\n" ::"n"(TP_EPI_WARPS * 32) : "memory");
        float *mv = reinterpret_cast<float *>(tp_smem + 2 * TILE_BYTES);          // [256][4]
        const int row = a_tile * TC_M + (warp & 3) * 32 + lane;
        if (chalf == 1) {
#pragma unroll
            for (int e = 0; e < TC_TOPK; ++e) mv[row * TC_TOPK + e] = v[e];
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(TP_EPI_WARPS * 32) : "memory");
        const int q = q0 + row;
        if (ok && chalf == 0 && q < nq) {
#pragma unroll
            for (int k = 0; k < TC_TOPK; ++k) {
                float t = mv[row * TC_TOPK + k];
#pragma unroll
                for (int e = 0; e < TC_TOPK; ++e) {
                    const float hi = fmaxf(v[e], t);
                    t = fminf(v[e], t);
                    v[e] = hi;
                }
            }
            uint32_t *o = cand + (((size_t)pair * 2 + dir) * g.kp_cap + q) * TC_TOPK;
#pragma unroll
            for (int e = 0; e < TC_TOPK; ++e) {
                const uint32_t idx = __float_as_uint(v[e]) & 0x3FFFu;
                o[e] = (v[e] > -1e29f && (int)idx < nt) ? idx : 0xFFFFFFFFu;
            }
        }
    }
    if (!ok) atomicExch(error_flag, 1);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

template <int D>
static int launch_l2_tp_d(const Geom &g, int n_pairs, const Buffers &b, const uint32_t *counts, int phase, cudaStream_t s) {
    constexpr int KC = D / 8 + 4;
    constexpr int NST = D == 128 ? 3 : 4;
    const int tiles = round_up(div_up(g.kp_cap, TC_M), 2);      // even: a CTA loads two adjacent A tiles
    if (phase == 0) {
        dim3 pgrid(tiles, 2 * n_pairs);
        l2_prep2_kernel<D><<<pgrid, TC_M, 0, s>>>(g, counts, b.fdesc, reinterpret_cast<uint4 *>(b.bf16desc), tiles);
    } else if (phase == 1) {
        const size_t smem = (size_t)(2 + NST) * KC * TC_M * 16;
        cudaFuncSetAttribute(l2_tc_pipe_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(tiles / 2, n_pairs, 2);
        l2_tc_pipe_kernel<D><<<grid, TP_THREADS, smem, s>>>(g, counts, reinterpret_cast<const uint4 *>(b.bf16desc), tiles, b.cand,
                                                             b.tc_error);
    } else {
        dim3 rgrid(div_up(g.kp_cap, 8), n_pairs, 2);
        l2_rerank_kernel<D><<<rgrid, 256, 0, s>>>(g, counts, b.fdesc, b.cand, b.best64, b.second64, b.allbest64, b.colbest64);
    }
    return 1;
}

template <int D>
static int launch_l2_tc_d(const Geom &g, int n_pairs, const Buffers &b, const uint32_t *counts, int phase, cudaStream_t s) {
    const int tiles = round_up(div_up(g.kp_cap, TC_M), 2);      // even: a CTA loads two adjacent A tiles
    if (phase == 0) {
        dim3 pgrid(tiles, 2 * n_pairs);
        l2_prep_kernel<D><<<pgrid, TC_M, 0, s>>>(g, counts, b.fdesc, reinterpret_cast<uint4 *>(b.bf16desc), b.fnorm, tiles);
    } else if (phase == 1) {
        const size_t smem = (size_t)3 * D * TC_M * 2;
        cudaFuncSetAttribute(l2_tc_topk_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(tiles / 2, n_pairs, 2);
        l2_tc_topk_kernel<D><<<grid, TC_THREADS2, smem, s>>>(g, counts, reinterpret_cast<const uint4 *>(b.bf16desc), b.fnorm, tiles,
                                                      b.cand, b.tc_error);
    } else {
        dim3 rgrid(div_up(g.kp_cap, 8), n_pairs, 2);
        l2_rerank_kernel<D><<<rgrid, 256, 0, s>>>(g, counts, b.fdesc, b.cand, b.best64, b.second64, b.allbest64, b.colbest64);
    }
    return 1;
}

// Unmasked L2: best64 / second64 / allbest64 (row side) and colbest64 (column side) for every pair.
// phase 0: operand preparation (bf16 tiles + norms), 1: the tcgen05 GEMM + top-4, 2: exact FP32 re-rank
int launch_l2_tensor(const Geom &g, int n_pairs, int dim, bool need_second, const Buffers &b, const uint32_t *counts,
                     int phase, cudaStream_t s) {
    static const int variant = getenv("FE_L2TC_VARIANT") ? atoi(getenv("FE_L2TC_VARIANT")) : 0;   // 1 = round-1 synchronous kernel (A/B)
    // The pipelined kernel proposes one candidate per 16-column group (exact for the arg-min up to bf16 near-ties); a
    // caller that needs the exact SECOND neighbour (unmasked kNN-2) gets the per-element top-4 kernel.  Its packed
    // column index is 14 bits wide.
    if (variant != 1 && !need_second && round_up(div_up(g.kp_cap, TC_M), 2) * TC_M <= 16384)
        return dim == 64 ? launch_l2_tp_d<64>(g, n_pairs, b, counts, phase, s) : launch_l2_tp_d<128>(g, n_pairs, b, counts, phase, s);
    return dim == 64 ? launch_l2_tc_d<64>(g, n_pairs, b, counts, phase, s) : launch_l2_tc_d<128>(g, n_pairs, b, counts, phase, s);
}

}  // namespace fe
