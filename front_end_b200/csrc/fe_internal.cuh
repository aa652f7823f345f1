// Internal declarations shared by the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/fe_abi.h"

namespace fe {

constexpr int STRIP_ROWS = 8;          // FAST strip height: one CTA = full image width x 8 rows
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ inline int div_up(int v, int m) { return (v + m - 1) / m; }

// Geometry of one batch launch (all images of a batch share width/height).
struct Geom {
    int w, h, pitch;       // pitch = round_up(w, 16) bytes
    int n_images;
    int n_strips;          // div_up(h, STRIP_ROWS)
    int slab_cap;          // candidate capacity per strip (u32 records)
    int kp_cap;            // per-image keypoint capacity
    size_t img_stride;     // bytes between consecutive images (pitch * h)
};

struct DetectParams {
    int threshold;         // FAST t
    int ps;                // 16 / 12 / 8
    int nonmax;
    int n_features;        // < 0: keep all
    int edge;              // border filter
};

// ---- device buffers (one set per ctx, sized for cfg.max_*) ------------------------------------
struct Buffers {
    uint8_t *img = nullptr;        // [n_images][h][pitch]
    uint8_t *blur = nullptr;       // same layout
    uint8_t *respmap = nullptr;    // same layout: NMS-surviving FAST responses (s - t), 0 elsewhere
    uint32_t *slab = nullptr;      // [n_images][n_strips][slab_cap]  (score<<24 | ylocal<<16 | x)
    uint32_t *strip_raw = nullptr; // [n_images][n_strips] candidates emitted per strip
    uint32_t *strip_sel = nullptr; // [n_images][n_strips] survivors of the top-N cut per strip
    uint32_t *hist = nullptr;      // [n_images][256] response histogram (inside the border)
    uint32_t *n_kp = nullptr;      // [n_images] keypoints found (may exceed kp_cap)
    uint32_t *kp_key = nullptr;    // [n_images][kp_cap]  (y<<16 | x), ascending = raster order
    uint8_t *kp_score = nullptr;   // [n_images][kp_cap]
    fe_kpoint *kp = nullptr;       // [n_images][kp_cap]  wire layout
    float *kx = nullptr, *ky = nullptr;   // [n_images][kp_cap]
    float2 *kcs = nullptr;         // [n_images][kp_cap]  (cos, sin) of the steering angle
    uint8_t *desc = nullptr;       // [n_images][kp_cap][32]
    float *fdesc = nullptr;        // [n_images][kp_cap][128] float descriptors (SURF), lazily allocated
    // matching, per pair (image 2p = query/left, 2p+1 = train/right)
    uint32_t *best = nullptr, *second = nullptr, *allbest = nullptr, *colbest = nullptr; // [n_pairs][kp_cap]
    fe_match *match_a = nullptr, *match_b = nullptr;   // [n_pairs][kp_cap]
    uint32_t *n_a = nullptr, *n_b = nullptr;           // [n_pairs]
    uint32_t *n_override = nullptr;                    // [n_images] counts for externally supplied kps
};

// ---- kernel launchers (each returns the number of kernels it launched) -------------------------
int launch_fast(const Geom &g, const DetectParams &p, const Buffers &b, cudaStream_t s);
int launch_select(const Geom &g, const DetectParams &p, const Buffers &b, cudaStream_t s);
int launch_orient_pack(const Geom &g, const DetectParams &p, const Buffers &b, bool orientation,
                       float kp_size, cudaStream_t s);
int launch_blur(const Geom &g, const Buffers &b, cudaStream_t s);
int launch_brief(const Geom &g, const Buffers &b, const uint32_t *counts, cudaStream_t s);
int launch_unpack_kps(const Geom &g, const Buffers &b, const uint32_t *counts, cudaStream_t s);

struct MatchParams {
    int mask;                  // fe_mask_kind for the (best, second) pair
    float epi_threshold, q_off, t_off;
    float half_w, half_h;
};
// unmasked row arg-min + column arg-min (cross-check)
int launch_hamming_cross(const Geom &g, int n_pairs, const Buffers &b, const uint32_t *counts, cudaStream_t s);
// masked kNN-2; train_sorted = train keypoints are in raster order (enables the banded kernel)
int launch_hamming_knn2(const Geom &g, int n_pairs, const MatchParams &mp, bool train_sorted, const Buffers &b,
                        const uint32_t *counts, cudaStream_t s);
int launch_finalize_ratio(const Geom &g, int n_pairs, double ratio, const Buffers &b,
                          const uint32_t *counts, cudaStream_t s);
int launch_finalize_cross(const Geom &g, int n_pairs, float max_dy, const Buffers &b,
                          const uint32_t *counts, cudaStream_t s);

}  // namespace fe
