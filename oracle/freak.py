"""Oracle: cv::FREAK (opencv_contrib xfeatures2d/src/freak.cpp) with caller-supplied selected pairs (test infrastructure).

The reference lists FREAK among the descriptors of its comparison runs (/root/reference bin/result_ONE:25, result_TWO:29,
result_THREE:23: features = [..., "FREAK", ...]) and creates it as  cv2.xfeatures2d.FREAK_create()  (bin/detect_node:43-45):
orientationNormalized = scaleNormalized = True, patternScale = 22, nOctaves = 4, selectedPairs = [].

Restated algorithm (scalar path, 8-bit image):
  * buildPattern: 43 receptive fields (7 rings of 6, staggered, + the centre) x 64 scales x 256 orientations, computed in
    double and stored as float (x, y, sigma); patternSizes[scale] = ceil((radius + sigma) * scale * patternScale) + 1;
    45 orientation pairs with integer weights  int(d / |d|^2 * 4096 + 0.5)  from the scale-0 / orientation-0 points;
    the 903 pairs (i > j) in order i = 1..42, j = 0..i-1; the 512 description pairs = allPairs[selectedPairs[k]].
  * compute: scale index = max(int(log(size / 7) * 64 / (ln2 * nOctaves) + 0.5), 0) clipped to 63 (or the constant of
    size 21 without scale normalisation); keypoints whose pattern does not fit (pt <= patternSize or pt >= dim - patternSize)
    are erased; meanIntensity = rounded box mean on the CV_32S integral image, box [int(c - r + 0.5), int(c + r + 1.5));
    sigma < 0.5: 10-bit fixed-point bilinear sample; orientation = atan2 of the integer-weighted sums of the 45 differences,
    stored in kp.angle (degrees, may be negative), quantised to 256 steps; descriptor bit = value[i] >= value[j], stored
    in the order of OpenCV's SSE path: pair 128 b + 16 u + t -> byte 16 b + 15 - t, bit u.

PARITY UNPINNED: no FREAK binary exists here (cv2 4.13 without xfeatures2d) and the default selection FREAK_DEF_PAIRS is a
constant table of opencv_contrib that is neither in the reference repository nor in this image, so the 512 selected pairs
are an INPUT (cv::FREAK's own `selectedPairs` argument).  The integer arithmetic is exact for any selection; the tests use
a seeded random one.  The pattern table is built with the C library's cos / sin / pow (Python's math module), the same
functions the product's host code calls."""
import math

import numpy as np

from .surf import integral_i32

NB_SCALES, NB_ORIENTATION, NB_POINTS, NB_PAIRS, NB_ORIENPAIRS, SMALLEST_KP_SIZE = 64, 256, 43, 512, 45, 7
LOG2 = 0.693147180559945
_RING_N = (6, 6, 6, 6, 6, 6, 6, 1)

# orientation pairs: opposite and second-neighbour fields on the four outer rings, the three diameters of the next three
ORIENT_PAIRS = []
for _b in (0, 6, 12, 18):
    ORIENT_PAIRS += [(_b + i, _b + j) for i, j in ((0, 3), (1, 4), (2, 5), (0, 2), (1, 3), (2, 4), (3, 5), (4, 0), (5, 1))]
for _b in (24, 30, 36):
    ORIENT_PAIRS += [(_b, _b + 3), (_b + 1, _b + 4), (_b + 2, _b + 5)]
assert len(ORIENT_PAIRS) == NB_ORIENPAIRS

ALL_PAIRS = [(i, j) for i in range(1, NB_POINTS) for j in range(i)]       # 903


def random_selection(seed=0):
    """A seeded stand-in for FREAK_DEF_PAIRS: 512 distinct indices into the 903 pairs."""
    return np.random.default_rng(seed).permutation(len(ALL_PAIRS))[:NB_PAIRS].astype(np.int32)


class Pattern:
    def __init__(self, pattern_scale=22.0, n_octaves=4):
        pattern_scale = float(np.float32(pattern_scale))
        big_r, small_r = 2.0 / 3.0, 2.0 / 24.0
        unit = (big_r - small_r) / 21.0
        radius = (big_r, big_r - 6 * unit, big_r - 11 * unit, big_r - 15 * unit, big_r - 18 * unit, big_r - 20 * unit, small_r, 0.0)
        sigma = tuple(r / 2.0 for r in radius[:7]) + (radius[6] / 2.0,)
        scale_step = math.pow(2.0, float(n_octaves) / NB_SCALES)
        self.n_octaves = n_octaves
        self.lookup = np.zeros((NB_SCALES, NB_ORIENTATION, NB_POINTS, 3), np.float32)
        self.sizes = np.zeros(NB_SCALES, np.int32)
        for s in range(NB_SCALES):
            sf = math.pow(scale_step, float(s))
            for o in range(NB_ORIENTATION):
                theta = float(o) * 2 * math.pi / float(NB_ORIENTATION)
                p = 0
                for i in range(8):
                    for k in range(_RING_N[i]):
                        beta = math.pi / _RING_N[i] * (i % 2)
                        alpha = float(k) * 2 * math.pi / float(_RING_N[i]) + beta + theta
                        self.lookup[s, o, p] = (radius[i] * math.cos(alpha) * sf * pattern_scale,
                                                radius[i] * math.sin(alpha) * sf * pattern_scale,
                                                sigma[i] * sf * pattern_scale)
                        p += 1
            for i in range(8):
                self.sizes[s] = max(self.sizes[s], int(math.ceil((radius[i] + sigma[i]) * sf * pattern_scale)) + 1)
        self.weights = np.zeros((NB_ORIENPAIRS, 2), np.int32)
        for m, (i, j) in enumerate(ORIENT_PAIRS):
            dx = np.float32(self.lookup[0, 0, i, 0] - self.lookup[0, 0, j, 0])
            dy = np.float32(self.lookup[0, 0, i, 1] - self.lookup[0, 0, j, 1])
            nsq = np.float32(np.float32(dx * dx) + np.float32(dy * dy))
            self.weights[m] = (int(float(np.float32(dx / nsq)) * 4096.0 + 0.5), int(float(np.float32(dy / nsq)) * 4096.0 + 0.5))

    def scale_index(self, size, scale_normalized=True):
        size_cst = float(np.float32(NB_SCALES / (LOG2 * self.n_octaves)))
        if scale_normalized:
            v = int(math.log(float(np.float32(np.float32(size) / np.float32(SMALLEST_KP_SIZE)))) * size_cst + 0.5)
        else:
            v = int(1.0986122886681 * size_cst + 0.5)
        return min(max(v, 0), NB_SCALES - 1)


def mean_intensity(img, S, pat, kx, ky, scale, rot, point):
    """freak.cpp meanIntensity<uchar, int>."""
    px, py, radius = pat.lookup[scale, rot, point]
    xf, yf = np.float32(px + np.float32(kx)), np.float32(py + np.float32(ky))
    x, y = int(xf), int(yf)
    if radius < 0.5:
        r_x, r_y = int(np.float32(np.float32(xf - np.float32(x)) * np.float32(1024))), int(np.float32(np.float32(yf - np.float32(y)) * np.float32(1024)))
        r_x_1, r_y_1 = 1024 - r_x, 1024 - r_y
        v = (r_x_1 * r_y_1 * int(img[y, x]) + r_x * r_y_1 * int(img[y, x + 1]) + r_x_1 * r_y * int(img[y + 1, x])
             + r_x * r_y * int(img[y + 1, x + 1]))
        # (sic) the weights sum to 2^20 but the source divides by 4 * 2^20: this branch returns a quarter of the mean
        return (((v + 2 * 1024 * 1024) & 0xFFFFFFFF) // (4 * 1024 * 1024)) & 0xFF
    x_left, y_top = int(float(np.float32(xf - radius)) + 0.5), int(float(np.float32(yf - radius)) + 0.5)
    x_right, y_bottom = int(float(np.float32(xf + radius)) + 1.5), int(float(np.float32(yf + radius)) + 1.5)
    v = int(S[y_bottom, x_right]) - int(S[y_bottom, x_left]) + int(S[y_top, x_left]) - int(S[y_top, x_right])
    area = (x_right - x_left) * (y_bottom - y_top)
    return ((v + area // 2) // area) & 0xFF


def _tdiv(a, b):
    """C integer division (truncation toward zero)."""
    q = abs(a) // b
    return q if a >= 0 else -q


def freak_compute(img, xs, ys, sizes, selected, orientation_normalized=True, scale_normalized=True, pattern_scale=22.0,
                  n_octaves=4, pattern=None):
    """Returns (keep bool[N], angle f32[n_kept], desc u8[n_kept x 64], values u8[n_kept x 43] at the final orientation)."""
    pat = pattern or Pattern(pattern_scale, n_octaves)
    h, w = img.shape
    S = integral_i32(img).astype(np.int64)
    pairs = [ALL_PAIRS[int(k)] for k in selected]
    assert len(pairs) == NB_PAIRS
    xs, ys = np.asarray(xs, np.float32), np.asarray(ys, np.float32)
    keep = np.zeros(len(xs), bool)
    angles, descs, values = [], [], []
    for k in range(len(xs)):
        sc = pat.scale_index(sizes[k], scale_normalized)
        ps = np.float32(pat.sizes[sc])
        if xs[k] <= ps or ys[k] <= ps or xs[k] >= np.float32(w - pat.sizes[sc]) or ys[k] >= np.float32(h - pat.sizes[sc]):
            continue
        keep[k] = True
        theta, angle = 0, np.float32(0)
        if orientation_normalized:
            v = [mean_intensity(img, S, pat, xs[k], ys[k], sc, 0, i) for i in range(NB_POINTS)]
            d0 = d1 = 0
            for m, (i, j) in enumerate(ORIENT_PAIRS):
                delta = v[i] - v[j]
                d0 += _tdiv(delta * int(pat.weights[m, 0]), 2048)
                d1 += _tdiv(delta * int(pat.weights[m, 1]), 2048)
            a = np.float32(math.atan2(float(np.float32(d1)), float(np.float32(d0))))       # atan2f, taken as correctly rounded
            angle = np.float32(float(a) * (180.0 / math.pi))
            t = float(np.float32(np.float32(NB_ORIENTATION) * angle)) * (1 / 360.0)
            theta = int(t - 0.5) if angle < 0 else int(t + 0.5)
            if theta < 0:
                theta += NB_ORIENTATION
            if theta >= NB_ORIENTATION:
                theta -= NB_ORIENTATION
        v = [mean_intensity(img, S, pat, xs[k], ys[k], sc, theta, i) for i in range(NB_POINTS)]
        d = np.zeros(64, np.uint8)
        for cnt, (i, j) in enumerate(pairs):
            b, u, t = cnt // 128, (cnt % 128) // 16, cnt % 16
            if v[i] >= v[j]:
                d[16 * b + 15 - t] |= 1 << u
        angles.append(angle)
        descs.append(d)
        values.append(v)
    n = len(descs)
    return (keep, np.array(angles, np.float32), np.stack(descs) if n else np.zeros((0, 64), np.uint8),
            np.array(values, np.uint8).reshape(n, NB_POINTS))
