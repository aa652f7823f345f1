// Detector setpoint: border filter + top-N by response with ties kept, compacted in raster order.
//
// Replaces KeyPointsFilter::runByImageBorder(edgeThreshold) + KeyPointsFilter::retainBest(N) inside
// cv::ORB::detect -- reached from /root/reference src/front_end/features.py:378-387
// (ORB_create(nfeatures,...)), src/utils.cpp:90, src/StereoCamera.cpp:438.  Semantics SURVEY.md A.2:
// keep edge <= x < W-edge, edge <= y < H-edge; if more than N remain keep every keypoint whose
// response >= the N-th largest response.  Responses are integers in [0,254], so the cut is read
// off a 256-bin histogram (built by fast.cu) instead of an nth_element; survivors are compacted
// strip by strip, which preserves the canonical raster order without a sort.
#include "fe_internal.cuh"

namespace fe {

constexpr int SEL_THREADS = 256;

__device__ __forceinline__ uint32_t block_incl_scan_256(uint32_t v, uint32_t *s_warp, uint32_t &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < SEL_THREADS / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < wid) base += c;
        total += c;
    }
    __syncthreads();
    return incl + base;
}

// Response cut T: keep score >= T.  T = 0 keeps everything, 256 keeps nothing.
__device__ int response_cut(const uint32_t *hist, int n_features, uint32_t *s_warp, int *s_cut) {
    // suffix counts S[i] = #{score >= i}: inclusive scan over the reversed histogram
    const int i = 255 - threadIdx.x;
    uint32_t total;
    const uint32_t S = block_incl_scan_256(hist[i], s_warp, total);  // = #{score >= i}
    if (threadIdx.x == 0) *s_cut = (n_features == 0) ? 256 : 0;
    __syncthreads();
    if (n_features > 0 && total > (uint32_t)n_features) {
        // largest i with S[i] >= N; S is non-increasing in i, so exactly one thread sees the edge
        const uint32_t S_next = S - hist[i];                          // = #{score >= i+1}
        if (S >= (uint32_t)n_features && S_next < (uint32_t)n_features) *s_cut = i;
    }
    __syncthreads();
    return *s_cut;
}

__device__ __forceinline__ bool survives(uint32_t rec, int y0, int cut, int edge, int w, int h) {
    const int x = rec & 0xFFFF, y = y0 + ((rec >> 16) & 0xFF), s = rec >> 24;
    return s >= cut && x >= edge && x < w - edge && y >= edge && y < h - edge;
}

__global__ void __launch_bounds__(SEL_THREADS)
select_count_kernel(Geom g, DetectParams p, StripView sv, const uint32_t *__restrict__ slab,
                    const uint32_t *__restrict__ strip_raw, const uint32_t *__restrict__ hist,
                    uint32_t *__restrict__ strip_sel) {
    __shared__ uint32_t s_warp[SEL_THREADS / 32];
    __shared__ int s_cut;
    const int strip = blockIdx.x, image = blockIdx.y;
    const int cut = response_cut(hist + image * 256, p.n_features, s_warp, &s_cut);
    const uint32_t n = strip_raw[image * g.n_strips + strip];
    const uint32_t *in = slab + (size_t)image * g.slab_img + (size_t)strip * sv.cap;
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS)
        c += survives(in[i], strip * sv.rows, cut, p.edge, g.w, g.h);
    uint32_t total;
    block_incl_scan_256(c, s_warp, total);
    if (threadIdx.x == 0) strip_sel[image * g.n_strips + strip] = total;
}

__global__ void __launch_bounds__(SEL_THREADS)
select_emit_kernel(Geom g, DetectParams p, StripView sv, const uint32_t *__restrict__ slab,
                   const uint32_t *__restrict__ strip_raw, const uint32_t *__restrict__ hist,
                   const uint32_t *__restrict__ strip_sel, uint32_t *__restrict__ kp_key,
                   uint8_t *__restrict__ kp_score, uint32_t *__restrict__ n_kp) {
    __shared__ uint32_t s_warp[SEL_THREADS / 32];
    __shared__ int s_cut;
    const int strip = blockIdx.x, image = blockIdx.y;
    const int cut = response_cut(hist + image * 256, p.n_features, s_warp, &s_cut);
    // exclusive offset of this strip = survivors in all earlier strips
    uint32_t part = 0;
    for (int sidx = threadIdx.x; sidx < strip; sidx += SEL_THREADS)
        part += strip_sel[image * g.n_strips + sidx];
    uint32_t offset;
    block_incl_scan_256(part, s_warp, offset);

    const uint32_t n = strip_raw[image * g.n_strips + strip];
    const uint32_t *in = slab + (size_t)image * g.slab_img + (size_t)strip * sv.cap;
    uint32_t *okey = kp_key + (size_t)image * g.kp_cap;
    uint8_t *oscore = kp_score + (size_t)image * g.kp_cap;
    const int y0 = strip * sv.rows;
    // Every warp owns one contiguous run of the strip's records (raster order = record order): it counts its survivors with
    // ballots, ONE block scan orders the eight warp totals, then the warp walks its run again (L1 hits) and writes each
    // survivor at the warp's base + its rank inside the ballot.  (First form: one block scan -- two barriers -- per 256 records.)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t per_warp = (uint32_t)round_up(div_up((int)n, SEL_THREADS / 32), 32);
    const uint32_t w_begin = min(n, (uint32_t)wid * per_warp), w_end = min(n, w_begin + per_warp);
    uint32_t mine = 0;
    for (uint32_t i = w_begin + lane; i < w_end + ((32u - ((w_end - w_begin) & 31u)) & 31u); i += 32)
        mine += __popc(__ballot_sync(0xffffffffu, i < w_end && survives(in[i], y0, cut, p.edge, g.w, g.h)));
    __shared__ uint32_t s_wtot[SEL_THREADS / 32];
    if (lane == 0) s_wtot[wid] = mine;
    __syncthreads();
    uint32_t wbase = offset, total = 0;
#pragma unroll
    for (int w2 = 0; w2 < SEL_THREADS / 32; ++w2) {
        const uint32_t c = s_wtot[w2];
        if (w2 < wid) wbase += c;
        total += c;
    }
    for (uint32_t i = w_begin + lane; i < w_end + ((32u - ((w_end - w_begin) & 31u)) & 31u); i += 32) {
        const uint32_t rec = i < w_end ? in[i] : 0u;
        const bool k = i < w_end && survives(rec, y0, cut, p.edge, g.w, g.h);
        const uint32_t m = __ballot_sync(0xffffffffu, k);
        if (k) {
            const uint32_t pos = wbase + __popc(m & ((1u << lane) - 1u));
            if (pos < (uint32_t)g.kp_cap) {
                const uint32_t x = rec & 0xFFFF, y = y0 + ((rec >> 16) & 0xFF);
                okey[pos] = (y << 16) | x;
                oscore[pos] = (uint8_t)(rec >> 24);
            }
        }
        wbase += __popc(m);
    }
    offset += total;
    if (strip == sv.n - 1 && threadIdx.x == 0) n_kp[image] = offset;
}

int launch_select(const Geom &g, const DetectParams &p, const Buffers &b, cudaStream_t s) {
    const StripView sv = strip_view(g, p);
    dim3 grid(sv.n, g.n_images);
    select_count_kernel<<<grid, SEL_THREADS, 0, s>>>(g, p, sv, b.slab, b.strip_raw, b.hist, b.strip_sel);
    select_emit_kernel<<<grid, SEL_THREADS, 0, s>>>(g, p, sv, b.slab, b.strip_raw, b.hist, b.strip_sel,
                                                    b.kp_key, b.kp_score, b.n_kp);
    return 2;
}

}  // namespace fe
