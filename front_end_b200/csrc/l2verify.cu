// Exact float-L2 cross-check on the tensor cores: ONE tcgen05 GEMM per stereo pair + candidate verification.
//
// Replaces BFMatcher(NORM_L2, crossCheck = true)::match followed by the live nodes' |yL - yR| <= max_dy filter
// (/root/reference src/live_stereo.cpp:240,364-377; features.py:463-467,670,724-733; algorithm.py:1122) for float
// descriptors (SURF / SURF_EXTENDED).  The L2 twin of the Hamming candidate verification (match.cu):
//
//   1. band pass (l2_band_kernel, exact FP32): a pair that survives the |dy| filter is the arg-min of its row and of
//      its column INSIDE the band, so the mutual band arg-mins name every possible survivor (q, t*, d*).
//   2. verification: (q, t*) survives iff no train anywhere beats t* for q and no query anywhere beats q for t*.  An
//      element (q', t') can only interfere if d(q', t') <= max(d*(q'), d*(t')).  The tensor cores evaluate
//            s(q', t') = q~'.t~' - |q~'|^2 / 2 - |t~'|^2 / 2 = -|q~' - t~'|^2 / 2
//      for ALL elements in one GEMM (fp16 operands q~', t~' = the descriptors scaled by a power of two and rounded;
//      the two norm terms ride in one extra K = 16 step: A-aug (1, 1, 1, qh, qm, ql) x B-aug (th, tm, tl, 1, 1, 1)), and
//      the epilogue compares s with min(L(q'), L(t')), L = -(d* + eps)^2 / 2 - eta: two ALU operations per element, no
//      top-k, no index packing.  eps bounds |q~ - q| + |t~ - t| (fp16 rounding, 2^-11 relative, + the subnormal floor),
//      eta the accumulation error, so an element that is NOT flagged provably has d > both thresholds and can neither
//      win nor tie.  Flagged elements (about two per row) go to a list.
//   3. l2v_eval_kernel measures the flagged elements exactly in FP32 (WarpRow, the banded kernel's definition) and folds
//      them into allbest64 / colbest64 with 64-bit atomicMin; l2_finalize_cross_kernel then keeps the mutual pairs.
// The result is EXACT (identical to the all-pairs FP32 kernel's, first-minimum ties included) and the GEMM runs once
// instead of once per direction.  A list overflow (degenerate descriptors) makes the evaluation kernel sweep that
// pair's whole matrix instead -- slow, still exact.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>

#include "fe_internal.cuh"
#include "fe_tc.cuh"

namespace fe {

using namespace tc;

constexpr unsigned long long KEY64_NONE_V = 0xFFFFFFFFFFFFFFFFull;
// A dead row / column (its candidate is proven not to be the minimum): a key no measured element can replace by atomicMin
// unless its distance is exactly 0 (then that element IS the true minimum), and whose index matches no keypoint.
constexpr unsigned long long KEY64_DEAD = 0x00000000FFFFFFFFull;
constexpr int VF_LIST_PER_KP = 32;           // flagged-element list capacity per pair = 32 x kp_cap

// ---- norms (error bound) and the per-image maximum ----------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
l2v_norm_kernel(Geom g, const uint32_t *__restrict__ counts, const float *__restrict__ fdesc, float *__restrict__ fnorm,
                uint32_t *__restrict__ maxnorm_bits) {
    // 64 rows per block (8 per warp, a quarter-warp of float4 lanes per row for D = 128); one atomicMax per block
    __shared__ float s_max[8];
    const int image = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = min((int)counts[image], g.kp_cap);
    if (blockIdx.x * 64 >= n) return;
    constexpr int LPR = D / 4 < 32 ? D / 4 : 32;          // lanes per row (float4 each): 32 (D = 128) / 16 (D = 64)
    constexpr int RPW = 32 / LPR;                         // rows per warp pass
    float wmax = 0.f;
    for (int it = 0; it < 8 / RPW; ++it) {
        const int row = blockIdx.x * 64 + warp * 8 + it * RPW + lane / LPR;
        float acc = 0.f;
        if (row < n) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(fdesc + ((size_t)image * g.kp_cap + row) * 128) + (lane % LPR));
            acc = __fmaf_rn(v.x, v.x, __fmaf_rn(v.y, v.y, __fmaf_rn(v.z, v.z, __fmul_rn(v.w, v.w))));
        }
#pragma unroll
        for (int off = LPR / 2; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        const float nrm = __fmul_rn(__fsqrt_ru(acc), 1.00001f);          // an upper bound of |x|
        if (row < n && (lane % LPR) == 0) fnorm[(size_t)image * g.kp_cap + row] = nrm;
        wmax = fmaxf(wmax, row < n ? nrm : 0.f);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_max[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_max[w]);
        atomicMax(&maxnorm_bits[image], __float_as_uint(m));              // non-negative floats order like their bits
    }
}

// power-of-two scale that brings the larger of the pair's maximum norms into [0.5, 1)
__device__ __forceinline__ float pair_scale(const uint32_t *maxnorm_bits, int pair) {
    const float m = fmaxf(__uint_as_float(maxnorm_bits[2 * pair]), __uint_as_float(maxnorm_bits[2 * pair + 1]));
    if (!(m > 0.f) || !(m < 3.0e38f)) return 1.f;
    int e;
    frexpf(m, &e);
    return ldexpf(1.f, -e);
}

// ---- operands: fp16 tiles in the UMMA K-major no-swizzle core-matrix layout + the augmentation columns ------------------
// per 128-row tile: [D/8 data chunks | B-aug | 0 | A-aug | 0], a chunk = 128 rows x 8 halfs (2 KB)
template <int D>
__global__ void __launch_bounds__(M)
l2v_prep_kernel(Geom g, const uint32_t *__restrict__ counts, const float *__restrict__ fdesc,
                const uint32_t *__restrict__ maxnorm_bits, uint4 *__restrict__ tiles, int tiles_per_image) {
    constexpr int KC = D / 8 + 4;
    const int image = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
    const int n = min((int)counts[image], g.kp_cap);
    // the GEMM reads the tiles that hold a keypoint, rounded up to a PAIR of tiles (a CTA loads two adjacent A tiles)
    if (tile >= round_up(div_up(max(n, 1), M), 2)) return;
    const int row = tile * M + r;
    const float sc = pair_scale(maxnorm_bits, image >> 1);
    uint4 *dst = tiles + ((size_t)image * tiles_per_image + tile) * KC * M;
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
                const __half lo = __float2half_rn(__fmul_rn(v[2 * e], sc)), hi = __float2half_rn(__fmul_rn(v[2 * e + 1], sc));
                w[e] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
                const float fl = __half2float(lo), fh = __half2float(hi);       // norm of the ROUNDED operand
                nrm = __fmaf_rn(fl, fl, nrm);
                nrm = __fmaf_rn(fh, fh, nrm);
            }
            dst[kc * M + r] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        for (int kc = 0; kc < D / 8; ++kc) dst[kc * M + r] = make_uint4(0, 0, 0, 0);
    }
    // -|x~|^2 / 2 as hi + mid + lo fp16 (33 bits)
    const float sv = __fmul_rn(-0.5f, nrm);
    const __half h = __float2half_rn(sv);
    const float r1 = __fsub_rn(sv, __half2float(h));
    const __half m = __float2half_rn(r1);
    const __half l = __float2half_rn(__fsub_rn(r1, __half2float(m)));
    const uint32_t one = 0x3C00u, neg = 0xF753u;                   // fp16 1.0, -30000 (rows / columns past the last keypoint)
    const uint32_t hm = (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(m) << 16);
    const uint32_t l1 = (uint32_t)__half_as_ushort(l) | (one << 16);
    if (row < n) {
        dst[(D / 8) * M + r] = make_uint4(hm, l1, one | (one << 16), 0u);                 // B-aug: th, tm, tl, 1, 1, 1, 0, 0
        dst[(D / 8 + 2) * M + r] = make_uint4(one | (one << 16), one | ((uint32_t)__half_as_ushort(h) << 16),
                                              (uint32_t)__half_as_ushort(m) | ((uint32_t)__half_as_ushort(l) << 16), 0u);   // A-aug: 1, 1, 1, qh, qm, ql
    } else {
        dst[(D / 8) * M + r] = make_uint4(neg, 0u, 0u, 0u);                                // B-aug: -30000, 0, ...
        dst[(D / 8 + 2) * M + r] = make_uint4(0u, neg << 16, 0u, 0u);                      // A-aug: 0, 0, 0, -30000, 0, ...
    }
    dst[(D / 8 + 1) * M + r] = make_uint4(0, 0, 0, 0);
    dst[(D / 8 + 3) * M + r] = make_uint4(0, 0, 0, 0);
}

// ---- seeds and thresholds from the band candidates -------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(1024)
l2v_classify_kernel(Geom g, const uint32_t *__restrict__ counts, const unsigned long long *__restrict__ candL,
                    const unsigned long long *__restrict__ candR, const float *__restrict__ fnorm,
                    const uint32_t *__restrict__ maxnorm_bits, unsigned long long *__restrict__ allbest,
                    unsigned long long *__restrict__ colbest, float *__restrict__ limq, float *__restrict__ limt,
                    float *__restrict__ limq_def, float *__restrict__ limt_def, uint32_t *__restrict__ npush, int cpad) {
    const int pair = blockIdx.x;
    const size_t o = (size_t)pair * g.kp_cap, ol = (size_t)pair * cpad;
    const float sc = pair_scale(maxnorm_bits, pair);
    const float inf = __int_as_float(0x7f800000);
    for (int side = 0; side < 2; ++side) {
        const int n = min((int)counts[2 * pair + side], g.kp_cap), n_other = min((int)counts[2 * pair + 1 - side], g.kp_cap);
        const unsigned long long *mine = (side ? candR : candL) + o, *other = (side ? candL : candR) + o;
        unsigned long long *seed = (side ? colbest : allbest) + o;
        float *lim = (side ? limt : limq) + ol;          // rows of cpad = round_up(kp_cap, 128) floats: the GEMM reads whole tiles
        const float other_max = __uint_as_float(maxnorm_bits[2 * pair + 1 - side]);
        const float *nrm = fnorm + (size_t)(2 * pair + side) * g.kp_cap;
        const int n_pad = round_up(max(n, 1), M);              // the GEMM reads whole tiles: +inf past the last entry
        for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
            float L = inf, Ldef = inf;
            unsigned long long s = KEY64_NONE_V;
            if (i < n) {
                const unsigned long long key = mine[i];
                const uint32_t j = (uint32_t)(key & 0xFFFFFFFFu);
                if (key != KEY64_NONE_V && (int)j < n_other && (uint32_t)(other[j] & 0xFFFFFFFFu) == (uint32_t)i) {
                    s = key;
                    // d* (an upper bound of it) in the scaled domain, plus the operand-rounding bound eps:
                    // |q~ - q| <= 2^-11 |q| + 2^-25 sqrt(D) per vector (fp16 normal / subnormal rounding)
                    // (the FP32 key itself is within D * 2^-24 relative of the real d^2: the 1.00002 covers it)
                    const float thr = __fmul_rn(__fmul_rn(__fsqrt_ru(__uint_as_float((uint32_t)(key >> 32))), sc), 1.00002f);
                    const float eps = __fmaf_rn(__fmul_rn(__fadd_rn(nrm[i], other_max), sc), 4.8829e-4f, 2.0f * 2.98e-8f * 11.32f);
                    const float T = __fadd_rn(thr, eps);
                    L = __fsub_rn(__fmul_rn(-0.5f, __fmul_rn(__fmul_rn(T, T), 1.000001f)), 1.0e-4f);   // eta: accumulation error
                    // "definitely closer than d*": s >= Ldef  =>  d~ <= Tm  =>  d <= Tm + eps < (a lower bound of) d*
                    const float thr_lo = __fmul_rn(__fmul_rn(__fsqrt_rd(__uint_as_float((uint32_t)(key >> 32))), sc), 0.99998f);
                    const float Tm = __fsub_rn(__fsub_rn(thr_lo, eps), 1.0e-6f);
                    if (Tm > 0.f) Ldef = __fadd_rn(__fmul_rn(-0.5f, __fmul_rn(__fmul_rn(Tm, Tm), 0.999998f)), 1.0e-4f);
                }
            }
            lim[i] = L;
            (side ? limt_def : limq_def)[ol + i] = Ldef;
            if (i < n) seed[i] = s;
        }
    }
    if (threadIdx.x == 0) npush[pair] = 0;
}

// Slow path of the GEMM epilogue (out of line and not unrolled into the hot loop, which stays small), entered by a whole warp
// when some lane's row (q = lane) passes the row or the column test somewhere in a group of 16 columns.
//   COLUMN KILL: an element definitely closer than column c's candidate (s >= Ldef(c)) proves that the candidate is not the
//   column minimum: the column stops flagging on its own account everywhere (L(t) := +inf in global memory; its far, false
//   candidate would otherwise let a few hundred elements through) and its seed is withdrawn.
//   PUSH (warp-aggregated, one atomicAdd): row-relevant elements always, column-relevant ones unless the column just died.
constexpr uint32_t VF_CHUNK = 64;            // list slots a warp reserves with one atomicAdd (the returning atomic is slow)

// (scalar arguments only: an aggregate would travel through local memory, whose cold stack lines cost thousands of cycles)
// The warp's list cursor (wbase, wused) goes in and comes back BY VALUE (a reference would pin both to the local-memory
// stack: an LDL / STL pair per tile step in the hot loop and cold stack lines in here -- ~5000 cycles per call measured).
__device__ __noinline__ uint2 vf_elem_slow(bool fr, bool fc, float sv, float ldef, int c, int q, int pair, int lane, int kp_cap,
                                           float *ltp, unsigned long long *colbest, uint32_t *npush, uint32_t *plist,
                                           uint32_t list_cap, uint32_t wbase, uint32_t wused) {
    const uint32_t bcd = __ballot_sync(0xffffffffu, fc && sv >= ldef);
    if (bcd && lane == __ffs(bcd) - 1) { __stcg(ltp + c, __int_as_float(0x7f800000)); colbest[(size_t)pair * kp_cap + c] = KEY64_DEAD; }
    const bool push = fr || (fc && !bcd);
    const uint32_t bp = __ballot_sync(0xffffffffu, push);
    if (bp) {
        const uint32_t n = (uint32_t)__popc(bp);
        if (wused + n > VF_CHUNK) {                                 // (warp-uniform) the chunk is full: pad it, take a new one
            if (wbase != 0xFFFFFFFFu)
                for (uint32_t k = wused + lane; k < VF_CHUNK; k += 32) if (wbase + k < list_cap) plist[wbase + k] = 0xFFFFFFFFu;
            uint32_t nb = 0;
            if (lane == 0) nb = atomicAdd(&npush[pair], VF_CHUNK);
            wbase = __shfl_sync(0xffffffffu, nb, 0);
            wused = 0;
        }
        const uint32_t k = wbase + wused + (uint32_t)__popc(bp & ((1u << lane) - 1u));
        if (push && k < list_cap) plist[k] = ((uint32_t)q << 16) | (uint32_t)c;
        wused += n;
    }
    return make_uint2(wbase, wused);
}

// ---- the GEMM: warp-specialised (producer / MMA issuer / 16 epilogue warps), 3-stage smem ring, 2 TMEM stages -----------
constexpr int VF_THREADS = 640;
constexpr int VF_EPI_WARPS = 16;

template <int D>
__global__ void __launch_bounds__(VF_THREADS, 1)
l2v_gemm_kernel(Geom g, const uint32_t *__restrict__ counts, const uint4 *__restrict__ tiles, int tiles_per_image,
                const float *__restrict__ limq, float *limt, const float *__restrict__ limq_def,
                const float *__restrict__ limt_def, int cpad, uint32_t *__restrict__ list, uint32_t *__restrict__ npush,
                unsigned long long *__restrict__ allbest, unsigned long long *__restrict__ colbest, int *__restrict__ error_flag) {
    constexpr int KC = D / 8 + 4;
    constexpr uint32_t TILE_BYTES = KC * M * 16;             // 40 KB (D = 128) / 24 KB (D = 64)
    constexpr int NST = D == 128 ? 3 : 4;
    constexpr uint32_t IDESC = idesc(0);                       // fp16 operands, fp32 accumulate
    extern __shared__ __align__(1024) uint8_t vf_smem[];
    __shared__ __align__(8) uint64_t s_full[NST], s_empty[NST], s_tfull[2], s_tempty[2], s_afull;
    __shared__ uint32_t s_tmem;

    const int pair = blockIdx.y;
    const int qi = 2 * pair, ti = 2 * pair + 1;
    const int nq = min((int)counts[qi], g.kp_cap), nt = min((int)counts[ti], g.kp_cap);
    const int q0 = blockIdx.x * (2 * M);
    if (q0 >= nq) return;                                    // uniform: before any allocation
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = div_up(nt, M);
    // every CTA sweeps the train tiles in the same cyclic order but starts somewhere else: a column killed by one row block
    // (below) is already dead when the others get to it
    const int rot = (int)((blockIdx.x * 7u) % (unsigned)max(n_tiles, 1));
    const uint32_t sA = smem_u32(vf_smem), sB = sA + 2 * TILE_BYTES;

    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&s_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_tfull[i]), 1); mbar_init(smem_u32(&s_tempty[i]), VF_EPI_WARPS); }
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
            // ===== producer: whole tiles by cp.async.bulk (a tile in HBM is the shared-memory image) =====
            const uint8_t *gA = reinterpret_cast<const uint8_t *>(tiles) + ((size_t)qi * tiles_per_image + 2 * blockIdx.x) * TILE_BYTES;
            const uint8_t *gB = reinterpret_cast<const uint8_t *>(tiles) + (size_t)ti * tiles_per_image * TILE_BYTES;
            mbar_expect_tx(smem_u32(&s_afull), 2 * TILE_BYTES);
            bulk_g2s(sA, gA, TILE_BYTES, smem_u32(&s_afull));
            bulk_g2s(sA + TILE_BYTES, gA + TILE_BYTES, TILE_BYTES, smem_u32(&s_afull));
            for (int j = 0; j < n_tiles && ok; ++j) {
                const int st = j % NST;
                ok = mbar_wait_bounded(smem_u32(&s_empty[st]), ((uint32_t)(j / NST) & 1u) ^ 1u);
                if (!ok) break;
                mbar_expect_tx(smem_u32(&s_full[st]), TILE_BYTES);
                const int jj = j + rot < n_tiles ? j + rot : j + rot - n_tiles;
                bulk_g2s(sB + st * TILE_BYTES, gB + (size_t)jj * TILE_BYTES, TILE_BYTES, smem_u32(&s_full[st]));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: (D / 16 + 1) x 2 tcgen05.mma per tile step =====
            ok = mbar_wait_bounded(smem_u32(&s_afull), 0);
            const uint64_t adesc0 = umma_desc(sA), adesc1 = umma_desc(sA + TILE_BYTES);
            constexpr uint64_t KSTEP = (uint64_t)((2 * LBO) >> 4);            // one K = 16 step = two core-matrix columns
            constexpr uint64_t AUG_B = (uint64_t)(((D / 8) * LBO) >> 4), AUG_A = (uint64_t)(((D / 8 + 2) * LBO) >> 4);
#ifdef FE_VF_TIMING
            long long twf = 0, twe = 0, t00 = clock64();
#endif
            for (int j = 0; j < n_tiles && ok; ++j) {
                const int st = j % NST, acc = j & 1;
#ifdef FE_VF_TIMING
                const long long ta = clock64();
#endif
                ok = mbar_wait_bounded(smem_u32(&s_full[st]), (uint32_t)(j / NST) & 1u);
#ifdef FE_VF_TIMING
                const long long tb = clock64();
                twf += tb - ta;
#endif
                if (ok) ok = mbar_wait_bounded(smem_u32(&s_tempty[acc]), ((uint32_t)(j >> 1) & 1u) ^ 1u);
#ifdef FE_VF_TIMING
                twe += clock64() - tb;
#endif
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint64_t bdesc = umma_desc(sB + st * TILE_BYTES);
                const uint32_t d0 = tmem + (uint32_t)(acc * 256), d1 = d0 + M;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    umma(d0, adesc0 + k * KSTEP, bdesc + k * KSTEP, IDESC, k > 0 ? 1u : 0u);
                    umma(d1, adesc1 + k * KSTEP, bdesc + k * KSTEP, IDESC, k > 0 ? 1u : 0u);
                }
                umma(d0, adesc0 + AUG_A, bdesc + AUG_B, IDESC, 1u);           // - |q~|^2 / 2 - |t~|^2 / 2
                umma(d1, adesc1 + AUG_A, bdesc + AUG_B, IDESC, 1u);
                umma_commit(smem_u32(&s_empty[st]));                          // shared stage free once these MMAs retire
                umma_commit(smem_u32(&s_tfull[acc]));                         // accumulator stage ready
            }
#ifdef FE_VF_TIMING
            if ((blockIdx.x == 3 || blockIdx.x == 11) && blockIdx.y == 5)
                printf("vf mma cta %d: tiles %d total %lld wait_full %lld wait_tempty %lld cycles/tile\n", blockIdx.x, n_tiles, (clock64() - t00) / n_tiles, twf / n_tiles, twe / n_tiles);
#endif
        }
    } else if (warp >= 4) {
        // ===== epilogue: warp w owns TMEM lanes 32 (w % 4).., A tile (w - 4) / 4 % 2, column half (w - 4) / 8 =====
        const int ew = warp - 4, a_tile = (ew >> 2) & 1, chalf = ew >> 3;
        const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a_tile * M + chalf * 64);
        const int q = q0 + a_tile * M + (warp & 3) * 32 + lane;
        const size_t po = (size_t)pair * cpad;
        float Lq = q < nq ? limq[po + q] : __int_as_float(0x7f800000);
        // a row is DEAD once some element is definitely closer than its candidate (s >= Ldef): the candidate cannot be the
        // row minimum, so the row stops flagging on its own account (its far, false candidate would otherwise let a few
        // hundred elements through) and its seed is withdrawn at the end
        const float Ldef = q < nq ? limq_def[po + q] : __int_as_float(0x7f800000);
        bool dead = false;
        float *ltp = limt + po;                          // column thresholds: read at L2 (__ldcg), +inf once the column is dead
        const float *ltdp = limt_def + po;
        const uint32_t list_cap = (uint32_t)VF_LIST_PER_KP * (uint32_t)g.kp_cap;
        uint32_t *plist = list + (size_t)pair * list_cap;
        const float inf = __int_as_float(0x7f800000);
        // this warp's 64 column thresholds of a tile step, double-buffered in shared memory: read at L2 (__ldcg: a column
        // another CTA killed is +inf) one tile step ahead, then served to all lanes as broadcast LDS.128
        float *s_lt = reinterpret_cast<float *>(vf_smem + (size_t)(2 + NST) * TILE_BYTES) + ew * 256;     // [2][lt 64 | ltdef 64]
        {
            const int jj0 = rot < n_tiles ? rot : 0;
            reinterpret_cast<float2 *>(s_lt)[lane] = __ldcg(reinterpret_cast<const float2 *>(ltp + jj0 * M + chalf * 64) + lane);
            reinterpret_cast<float2 *>(s_lt + 64)[lane] = __ldg(reinterpret_cast<const float2 *>(ltdp + jj0 * M + chalf * 64) + lane);
        }
        __syncwarp();
        uint32_t wbase = 0xFFFFFFFFu, wused = VF_CHUNK;        // this warp's reserved list chunk (none yet)
#ifdef FE_VF_TIMING
        long long e_wait = 0, e_ld = 0, e_main = 0, e_slow = 0, e_calls = 0, e_ncall = 0, e_t0 = clock64();
#endif
        for (int j = 0; j < n_tiles && ok; ++j) {
            const int acc = j & 1;
            float2 nxt = make_float2(inf, inf), nxd = make_float2(inf, inf);
            if (j + 1 < n_tiles) {
                const int jn = j + 1 + rot < n_tiles ? j + 1 + rot : j + 1 + rot - n_tiles;
                nxt = __ldcg(reinterpret_cast<const float2 *>(ltp + jn * M + chalf * 64) + lane);
                nxd = __ldg(reinterpret_cast<const float2 *>(ltdp + jn * M + chalf * 64) + lane);
            }
#ifdef FE_VF_TIMING
            const long long e_a = clock64();
#endif
            ok = mbar_wait_bounded(smem_u32(&s_tfull[acc]), (uint32_t)(j >> 1) & 1u);
            if (!ok) break;
#ifdef FE_VF_TIMING
            e_wait += clock64() - e_a;
#endif
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int jj = j + rot < n_tiles ? j + rot : j + rot - n_tiles;
            const int colbase = jj * M + chalf * 64;
            const float4 *lt4 = reinterpret_cast<const float4 *>(s_lt + (j & 1) * 128);
            // Per group of 16 columns: the row test is max(s) >= L(q), the column test max(s - L(t)) >= 0 -- one FADD (FMA
            // pipe) and half a 3-input max (ALU pipe) per element.  Only a group in which some lane of the warp passes one of
            // them is looked at element by element.  The 64 columns come out of TMEM in two halves (32 registers at a time).
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {          // (not unrolled: two copies of the group code, not four)
                uint32_t r[32];
#ifdef FE_VF_TIMING
                const long long e_b = clock64();
#endif
                tmem_ld32_nowait(tbase + (uint32_t)(acc * 256 + half * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#ifdef FE_VF_TIMING
                e_ld += clock64() - e_b;
#endif
                if (half == 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&s_tempty[acc]));      // TMEM stage released: the values are in registers
                }
#pragma unroll
                for (int h16 = 0; h16 < 32; h16 += 16) {
                    const int g16 = half * 32 + h16;
                    float lt[16];
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        const float4 t4 = lt4[(g16 >> 2) + v4];
                        lt[4 * v4] = t4.x; lt[4 * v4 + 1] = t4.y; lt[4 * v4 + 2] = t4.z; lt[4 * v4 + 3] = t4.w;
                    }
#define VF_S(e) __uint_as_float(r[h16 + (e)])
#define VF_T(e) __fsub_rn(VF_S(e), lt[e])
                    float rmax = fmaxf(fmaxf(VF_S(0), VF_S(1)), VF_S(2)), cmax = fmaxf(fmaxf(VF_T(0), VF_T(1)), VF_T(2));
#pragma unroll
                    for (int e = 3; e < 15; e += 2) { rmax = fmaxf(fmaxf(rmax, VF_S(e)), VF_S(e + 1)); cmax = fmaxf(fmaxf(cmax, VF_T(e)), VF_T(e + 1)); }
                    rmax = fmaxf(rmax, VF_S(15)); cmax = fmaxf(cmax, VF_T(15));
                    dead |= rmax >= Ldef;
                    if (dead) Lq = inf;              // at once: a dead row flags nothing on its own account, this group included
                    if (__any_sync(0xffffffffu, rmax >= Lq || cmax >= 0.f)) {
#ifdef FE_VF_TIMING
                        const long long e_c = clock64();
#endif
                        // which of the 16 columns have a qualifying element in SOME lane: one mask per lane, one warp reduction
                        uint32_t em = 0;
#pragma unroll
                        for (int e = 0; e < 16; ++e) em |= (VF_S(e) >= Lq || VF_S(e) >= lt[e]) ? (1u << e) : 0u;   // (L(t) of a dead column is +inf)
                        const uint32_t wm = __reduce_or_sync(0xffffffffu, em);
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            if (wm & (1u << e))                                       // rare: a few elements of a row qualify in total
                            {
                                const bool fr = VF_S(e) >= Lq, fc = VF_S(e) >= lt[e];
#ifdef FE_VF_TIMING
                                const long long e_d = clock64();
#endif
                                const uint2 wc = vf_elem_slow(fr, fc, VF_S(e), s_lt[(j & 1) * 128 + 64 + g16 + e], colbase + g16 + e, q, pair,
                                                              lane, g.kp_cap, ltp, colbest, npush, plist, list_cap, wbase, wused);
                                wbase = wc.x; wused = wc.y;
#ifdef FE_VF_TIMING
                                e_main += clock64() - e_d; ++e_ncall;
#endif
                            }
                        }
#ifdef FE_VF_TIMING
                        e_slow += clock64() - e_c; ++e_calls;
#endif
                    }
                }
            }
            if (dead) Lq = inf;
            reinterpret_cast<float2 *>(s_lt + ((j + 1) & 1) * 128)[lane] = nxt;      // next tile step's thresholds
            reinterpret_cast<float2 *>(s_lt + ((j + 1) & 1) * 128 + 64)[lane] = nxd;
            __syncwarp();
        }
#ifdef FE_VF_TIMING
        if ((blockIdx.x == 3 || blockIdx.x == 11) && blockIdx.y == 5 && lane == 0 && (warp == 4 || warp == 13))
            printf("vf epi cta %d warp %d: total %lld wait_tfull %lld tmem_ld %lld slow %lld (groups %lld: %lld cycles each; element calls %lld: %lld cycles each) cycles/tile\n", blockIdx.x, warp,
                   (clock64() - e_t0) / n_tiles, e_wait / n_tiles, e_ld / n_tiles, e_slow / n_tiles, e_calls, e_slow / (e_calls ? e_calls : 1), e_ncall, e_main / (e_ncall ? e_ncall : 1));
#endif
        // pad the unused tail of this warp's last chunk (the evaluation kernel skips the filler)
        if (wbase != 0xFFFFFFFFu)
            for (uint32_t k = wused + lane; k < VF_CHUNK; k += 32) if (wbase + k < list_cap) plist[wbase + k] = 0xFFFFFFFFu;
#undef VF_S
#undef VF_T
        if (dead && q < nq) allbest[(size_t)pair * g.kp_cap + q] = KEY64_DEAD;        // seed withdrawn (both column halves may write it)
    }
    if (!ok) atomicExch(error_flag, 1);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

// ---- exact evaluation of the flagged elements ----------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256, 8)
l2v_eval_kernel(Geom g, const uint32_t *__restrict__ counts, const float *__restrict__ fdesc, const uint32_t *__restrict__ list,
                const uint32_t *__restrict__ npush, unsigned long long *__restrict__ allbest, unsigned long long *__restrict__ colbest,
                int force_sweep) {
    const int pair = blockIdx.y, lane = threadIdx.x & 31;
    const int nq = min((int)counts[2 * pair], g.kp_cap), nt = min((int)counts[2 * pair + 1], g.kp_cap);
    const uint32_t list_cap = (uint32_t)VF_LIST_PER_KP * (uint32_t)g.kp_cap;
    const uint32_t pushed = npush[pair];
    const bool sweep = (force_sweep & 1) || pushed > list_cap;      // overflow: evaluate the whole matrix of this pair (still exact)
    if ((force_sweep & 2) && blockIdx.x == 0 && threadIdx.x == 0 && pair < 4) printf("l2verify pair %d: nq %d nt %d flagged %u (cap %u)\n", pair, nq, nt, pushed, list_cap);
    const unsigned long long total = sweep ? (unsigned long long)nq * (unsigned long long)nt : (unsigned long long)pushed;
    const uint32_t *plist = list + (size_t)pair * list_cap;
    const size_t po = (size_t)pair * g.kp_cap;
    const float *qd = fdesc + (size_t)(2 * pair) * g.kp_cap * 128, *td = fdesc + (size_t)(2 * pair + 1) * g.kp_cap * 128;
    const unsigned long long stride = (unsigned long long)gridDim.x * 8ull;
    int q_loaded = -1;
    WarpRow<D> qr;
    for (unsigned long long e = (unsigned long long)blockIdx.x * 8ull + (threadIdx.x >> 5); e < total; e += stride) {
        int q, t;
        if (sweep) { q = (int)(e / (unsigned long long)nt); t = (int)(e - (unsigned long long)q * (unsigned long long)nt); }
        else { const uint32_t w = plist[e]; q = (int)(w >> 16); t = (int)(w & 0xFFFFu); }
        if (q >= nq || t >= nt) continue;
        if (q != q_loaded) { qr.load(qd + (size_t)q * 128, lane); q_loaded = q; }
        const float d2 = qr.dist2(td + (size_t)t * 128, lane);
        if (lane == 0) {
            // most flagged elements do not beat the running minima: look before the (contended) atomic
            const unsigned long long bits = (unsigned long long)__float_as_uint(d2) << 32;
            const unsigned long long kq = bits | (unsigned)t, kt = bits | (unsigned)q;
            if (kq < *reinterpret_cast<volatile unsigned long long *>(&allbest[po + q])) atomicMin(&allbest[po + q], kq);
            if (kt < *reinterpret_cast<volatile unsigned long long *>(&colbest[po + t])) atomicMin(&colbest[po + t], kt);
        }
    }
}

template <int D>
static int launch_l2_verify_d(const Geom &g, int n_pairs, const Buffers &b, const uint32_t *counts, int phase, cudaStream_t s) {
    constexpr int KC = D / 8 + 4;
    constexpr int NST = D == 128 ? 3 : 4;
    const int tiles = round_up(div_up(g.kp_cap, M), 2);        // even: a CTA loads two adjacent A tiles
    const int cpad = round_up(g.kp_cap, M);
    if (phase == 0) {
        cudaMemsetAsync(b.vf_maxnorm, 0, sizeof(uint32_t) * 2 * n_pairs, s);
        dim3 ngrid(div_up(g.kp_cap, 64), 2 * n_pairs);
        l2v_norm_kernel<D><<<ngrid, 256, 0, s>>>(g, counts, b.fdesc, b.fnorm, b.vf_maxnorm);
        dim3 pgrid(tiles, 2 * n_pairs);
        l2v_prep_kernel<D><<<pgrid, M, 0, s>>>(g, counts, b.fdesc, b.vf_maxnorm, reinterpret_cast<uint4 *>(b.bf16desc), tiles);
        l2v_classify_kernel<D><<<n_pairs, 1024, 0, s>>>(g, counts, b.vf_candL, b.vf_candR, b.fnorm, b.vf_maxnorm, b.allbest64, b.colbest64,
                                                        b.vf_limq, b.vf_limt, b.vf_limqd, b.vf_limtd, b.vf_npush, cpad);
        return 3;
    }
    // FE_L2_VERIFY_SWEEP=1 (tests): skip the GEMM and evaluate every element exactly -- the reference the tensor path must equal
    static const int force_sweep = getenv("FE_L2_VERIFY_SWEEP") ? atoi(getenv("FE_L2_VERIFY_SWEEP")) : 0;
    if (phase == 1) {
        if (force_sweep & 1) return 0;
        const size_t smem = (size_t)(2 + NST) * KC * M * 16 + VF_EPI_WARPS * 256 * sizeof(float);
        cudaFuncSetAttribute(l2v_gemm_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(tiles / 2, n_pairs);
        l2v_gemm_kernel<D><<<grid, VF_THREADS, smem, s>>>(g, counts, reinterpret_cast<const uint4 *>(b.bf16desc), tiles, b.vf_limq, b.vf_limt, b.vf_limqd, b.vf_limtd, cpad, b.vf_list,
                                                          b.vf_npush, b.allbest64, b.colbest64, b.tc_error);
        return 1;
    }
    dim3 egrid(256, n_pairs);
    l2v_eval_kernel<D><<<egrid, 256, 0, s>>>(g, counts, b.fdesc, b.vf_list, b.vf_npush, b.allbest64, b.colbest64, force_sweep);
    return 1;
}

// phase 0: norms, fp16 operand tiles, seeds + thresholds from the band candidates (b.vf_candL / b.vf_candR);
// phase 1: the tcgen05 GEMM + threshold epilogue; phase 2: exact evaluation of the flagged elements
int launch_l2_verify(const Geom &g, int n_pairs, int dim, const Buffers &b, const uint32_t *counts, int phase, cudaStream_t s) {
    return dim == 64 ? launch_l2_verify_d<64>(g, n_pairs, b, counts, phase, s) : launch_l2_verify_d<128>(g, n_pairs, b, counts, phase, s);
}

size_t l2_verify_list_entries(int kp_cap) { return (size_t)VF_LIST_PER_KP * (size_t)kp_cap; }

}  // namespace fe
