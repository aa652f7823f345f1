"""Oracle: ORB single-level detect + describe (test infrastructure).

Restates what ``cv2.ORB_create(N, 1.2, nlevels=1, edgeThreshold=31, 0, WTA_K=2,
ORB_FAST_SCORE, patchSize=31, fastThreshold=t).detectAndCompute(img, None)``
computes -- the detector/descriptor the reference selects at
/root/reference/src/front_end/features.py:378-387, src/utils.cpp:90-94,
src/StereoCamera.cpp:434-444 (detect :84, compute :89) and bin/detect_node:50-51:
FAST-9_16 + NMS -> border filter -> retainBest (ties kept) -> intensity-centroid
angle -> 7x7 sigma=2 float Gaussian -> rBRIEF-256.  The arithmetic is OpenCV's
(un-vendored); semantics follow SURVEY.md Appendix A.2-A.4 and are pinned
bit-exactly against cv2 4.13.0 in tests/test_oracle_pins.py.

Canonical keypoint order of this oracle (and of the CUDA path) is raster order
(y, then x); cv2's own order after retainBest is nth_element-dependent, so pins
compare as sets keyed by (x, y).
"""
import os

import numpy as np

from . import fast as _fast

HALF_PATCH = 15
EDGE = 31
PATTERN = np.load(os.path.join(os.path.dirname(__file__), "orb_pattern.npy")).astype(np.int32)

# umax[v] = cvRound(sqrt(15^2 - v^2)) with OpenCV's symmetric fill (A.3)
UMAX = np.array([15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3], np.int32)

_f32 = np.float32
_SCALE = _f32(180.0 / np.pi)
ATAN2_P1 = _f32(_f32(0.9997878412794807) * _SCALE)
ATAN2_P3 = _f32(_f32(-0.3258083974640975) * _SCALE)
ATAN2_P5 = _f32(_f32(0.1555786518463281) * _SCALE)
ATAN2_P7 = _f32(_f32(-0.04432655554792128) * _SCALE)
_EPS = _f32(2.220446049250313e-16)

# getGaussianKernel(7, 2, CV_32F)
GAUSS7 = np.array([0.07015932351350784, 0.13107487559318542, 0.1907128244638443,
                   0.21610593795776367, 0.1907128244638443, 0.13107487559318542,
                   0.07015932351350784], np.float32)


def _fma(a, b, c):
    """Single-rounding a*b+c for float32 arrays (exact product in float64, then one rounding)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def fast_atan2_deg(y, x, use_fma=False):
    """cv::fastAtan2 (degrees, [0,360)) on float32 arrays; polynomial of A.3."""
    y = np.asarray(y, np.float32)
    x = np.asarray(x, np.float32)
    ax, ay = np.abs(x), np.abs(y)
    mx = np.maximum(ax, ay)
    mn = np.minimum(ax, ay)
    c = (mn / (mx + _EPS)).astype(np.float32)
    c2 = (c * c).astype(np.float32)
    if use_fma:
        a = _fma(ATAN2_P7 * np.ones_like(c2), c2, ATAN2_P5 * np.ones_like(c2))
        a = _fma(a, c2, ATAN2_P3 * np.ones_like(c2))
        a = _fma(a, c2, ATAN2_P1 * np.ones_like(c2))
    else:
        a = ((ATAN2_P7 * c2).astype(np.float32) + ATAN2_P5).astype(np.float32)
        a = ((a * c2).astype(np.float32) + ATAN2_P3).astype(np.float32)
        a = ((a * c2).astype(np.float32) + ATAN2_P1).astype(np.float32)
    a = (a * c).astype(np.float32)
    a = np.where(ax >= ay, a, (_f32(90.0) - a).astype(np.float32))
    a = np.where(x < 0, (_f32(180.0) - a).astype(np.float32), a)
    a = np.where(y < 0, (_f32(360.0) - a).astype(np.float32), a)
    return a.astype(np.float32)


def border_and_retain_best(xs, ys, resp, shape, n_features, edge=EDGE):
    """runByImageBorder(edge) then retainBest(n) with ties kept (A.2); raster order preserved."""
    H, W = shape
    inb = (xs >= edge) & (xs < W - edge) & (ys >= edge) & (ys < H - edge)
    xs, ys, resp = xs[inb], ys[inb], resp[inb]
    if n_features >= 0 and len(xs) > n_features:
        if n_features == 0:
            return xs[:0], ys[:0], resp[:0]
        cut = np.sort(resp)[::-1][n_features - 1]
        keep = resp >= cut
        xs, ys, resp = xs[keep], ys[keep], resp[keep]
    return xs, ys, resp


def harris_responses(img, xs, ys, block=7, k=0.04):
    """OpenCV orb.cpp HarrisResponses(img, pts, blockSize 7, HARRIS_K 0.04f): 3x3 Sobel-like integer gradients summed over
    the block, then  ((float)a*b - (float)c*c - k*((float)a+b)*((float)a+b)) * scale^4  in float32, one rounding per
    operation (no FMA), scale = 1.f / (4 * block * 255.f).  Reached when the `score` field of front_end/setDetector is
    HARRIS_SCORE (src/StereoCamera.cpp:445,462; src/utils.cpp:86-90).  Pinned bit-exactly against cv2."""
    f = np.float32
    I = np.pad(np.asarray(img, np.int64), block // 2 + 1, mode="reflect")
    p = block // 2 + 1
    r = block // 2
    a = np.zeros(len(xs), np.int64)
    b = a.copy()
    c = a.copy()
    for dy in range(-r, block - r):
        for dx in range(-r, block - r):
            y, x = ys + dy + p, xs + dx + p
            ix = (I[y, x + 1] - I[y, x - 1]) * 2 + (I[y - 1, x + 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y + 1, x - 1])
            iy = (I[y + 1, x] - I[y - 1, x]) * 2 + (I[y + 1, x - 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y - 1, x + 1])
            a += ix * ix
            b += iy * iy
            c += ix * iy
    scale = f(1.0) / f(f(4 * block) * f(255.0))
    s4 = f(f(f(scale * scale) * scale) * scale)
    af, bf, cf = a.astype(f), b.astype(f), c.astype(f)
    t = (af + bf).astype(f)
    det = ((af * bf).astype(f) - (cf * cf).astype(f)).astype(f)
    return ((det - ((f(k) * t).astype(f) * t).astype(f)).astype(f) * s4).astype(f)


def harris_retain_best(img, xs, ys, resp, n_features):
    """computeKeyPoints' second cull: Harris score of the 2N FAST survivors, retainBest(N) with ties kept; order kept."""
    hr = harris_responses(img, xs, ys)
    if n_features >= 0 and len(xs) > n_features:
        if n_features == 0:
            return xs[:0], ys[:0], hr[:0]
        cut = np.sort(hr)[::-1][n_features - 1]
        keep = hr >= cut
        xs, ys, hr = xs[keep], ys[keep], hr[keep]
    return xs, ys, hr


def ic_angle(img, xs, ys, use_fma=False):
    """Intensity-centroid orientation in degrees (float32) over the radius-15 disc."""
    I = img.astype(np.int32)
    m10 = np.zeros(len(xs), np.int64)
    m01 = np.zeros(len(xs), np.int64)
    for v in range(-HALF_PATCH, HALF_PATCH + 1):
        d = int(UMAX[abs(v)])
        for u in range(-d, d + 1):
            val = I[ys + v, xs + u]
            m10 += u * val
            m01 += v * val
    return fast_atan2_deg(m01.astype(np.float32), m10.astype(np.float32), use_fma), m10, m01


def _reflect101(idx, n):
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * (n - 1) - idx, idx)


def gaussian_blur_7x7(img):
    """ORB's internal blur: float32 separable 7-tap, BORDER_REFLECT_101, exact FMA order of A.4."""
    H, W = img.shape
    g = GAUSS7
    F = img.astype(np.float32)
    xi = [_reflect101(np.arange(W) + k - 3, W) for k in range(7)]
    s = (F[:, xi[0]] * g[0]).astype(np.float32)
    for k in range(1, 7):
        s = _fma(F[:, xi[k]], np.full_like(s, g[k]), s)
    T = s
    yi = [_reflect101(np.arange(H) + k, H) for k in range(-3, 4)]
    out = (T[yi[3]] * g[3]).astype(np.float32)
    for k in range(1, 4):
        pair = (T[yi[3 + k]] + T[yi[3 - k]]).astype(np.float32)
        out = _fma(pair, np.full_like(out, g[3 + k]), out)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def make_random_pattern(patch_size, npoints=512, seed=0x34985739):
    """OpenCV orb.cpp makeRandomPattern: cv::RNG(seed), x then y uniform in [-patchSize/2, patchSize/2].  cv::RNG is a
    multiply-with-carry generator (state = (u32)state * 4164903690 + (state >> 32)); uniform(a, b) = a + next() % (b - a).
    Used by ORB when patchSize != 31 (bin/detect_node:50-51 sets 70; features.py:292-352 sweeps 10/30/50/70).
    Returns a (256, 4) int32 table laid out like bit_pattern_31_ (x0, y0, x1, y1)."""
    state = seed
    lo, span = -(patch_size // 2), patch_size // 2 + 1 + patch_size // 2
    out = np.zeros(2 * npoints, np.int32)
    for i in range(2 * npoints):
        state = ((state & 0xFFFFFFFF) * 4164903690 + (state >> 32)) & 0xFFFFFFFFFFFFFFFF
        out[i] = lo + (state & 0xFFFFFFFF) % span
    return out.reshape(npoints // 2, 4)


def rbrief256(blurred, xs, ys, angles_deg, pattern=None, raw=None):
    """N x 32 u8 steered BRIEF (WTA_K=2) sampled from the blurred image (A.4).
    pattern: (256, 4) table, default ORB's bit_pattern_31_.  raw: the unblurred image -- when given, samples outside the
    image read raw[reflect101(y), reflect101(x)] (cv2 blurs only the image area of its bordered pyramid buffer; pinned
    with patchSize 70)."""
    n = len(xs)
    pat = PATTERN if pattern is None else np.asarray(pattern, np.int32)
    theta = (np.asarray(angles_deg, np.float32) * _f32(np.pi / 180.0)).astype(np.float32)
    a = np.cos(theta.astype(np.float64)).astype(np.float32)[:, None]
    b = np.sin(theta.astype(np.float64)).astype(np.float32)[:, None]
    px = pat.reshape(512, 2)[:, 0].astype(np.float32)[None, :]
    py = pat.reshape(512, 2)[:, 1].astype(np.float32)[None, :]
    fx = ((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.float32)
    fy = ((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.float32)
    ix = np.rint(fx).astype(np.int32)
    iy = np.rint(fy).astype(np.int32)
    X = np.asarray(xs, np.int32)[:, None] + ix
    Y = np.asarray(ys, np.int32)[:, None] + iy
    if raw is None:
        vals = blurred[Y, X].astype(np.int32)  # n x 512
    else:
        H, W = blurred.shape
        inside = (X >= 0) & (X < W) & (Y >= 0) & (Y < H)
        vals = np.where(inside, blurred[np.clip(Y, 0, H - 1), np.clip(X, 0, W - 1)],
                        raw[_reflect101(Y, H), _reflect101(X, W)]).astype(np.int32)
    bits = (vals[:, 0::2] < vals[:, 1::2]).astype(np.uint8)  # n x 256
    return np.packbits(bits.reshape(n, 32, 8), axis=2, bitorder="little").reshape(n, 32)


def init_wta_pattern(pattern0, wta_k, ntuples=128, seed=0x12345678):
    """orb.cpp initializeOrbPattern: ntuples tuples of wta_k DISTINCT points drawn from the 512-point base pattern with
    cv::RNG(seed).uniform(0, 512).  Returns (ntuples * wta_k, 2) int32."""
    pool = np.asarray(pattern0, np.int32).reshape(-1, 2)
    state = seed
    out = np.zeros((ntuples * wta_k, 2), np.int32)
    for i in range(ntuples):
        for k in range(wta_k):
            while True:
                state = ((state & 0xFFFFFFFF) * 4164903690 + (state >> 32)) & 0xFFFFFFFFFFFFFFFF
                pt = pool[(state & 0xFFFFFFFF) % len(pool)]
                if all(not np.array_equal(out[wta_k * i + k1], pt) for k1 in range(k)):
                    out[wta_k * i + k] = pt
                    break
    return out


def rbrief_wta(blurred, xs, ys, angles_deg, wta_k, pattern0=None):
    """ORB descriptor with WTA_K = 3 / 4 (orb.cpp computeOrbDescriptors): 128 two-bit symbols, four per byte."""
    pat = init_wta_pattern(PATTERN if pattern0 is None else pattern0, wta_k)
    n = len(xs)
    theta = (np.asarray(angles_deg, np.float32) * _f32(np.pi / 180.0)).astype(np.float32)
    a = np.cos(theta.astype(np.float64)).astype(np.float32)[:, None]
    b = np.sin(theta.astype(np.float64)).astype(np.float32)[:, None]
    px, py = pat[:, 0].astype(np.float32)[None, :], pat[:, 1].astype(np.float32)[None, :]
    ix = np.rint(((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.float32)).astype(np.int32)
    iy = np.rint(((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.float32)).astype(np.int32)
    v = blurred[np.asarray(ys, np.int32)[:, None] + iy, np.asarray(xs, np.int32)[:, None] + ix].astype(np.int32)
    v = v.reshape(n, 128, wta_k)
    if wta_k == 3:
        t0, t1, t2 = v[:, :, 0], v[:, :, 1], v[:, :, 2]
        val = np.where(t2 > t1, np.where(t2 > t0, 2, 0), (t1 > t0).astype(np.int32))
    else:
        t0, t1, t2, t3 = v[:, :, 0], v[:, :, 1], v[:, :, 2], v[:, :, 3]
        val = np.where(np.maximum(t0, t1) > np.maximum(t2, t3), (t1 > t0).astype(np.int32), np.where(t3 > t2, 3, 2))
    val = val.reshape(n, 32, 4)
    return (val[:, :, 0] | (val[:, :, 1] << 2) | (val[:, :, 2] << 4) | (val[:, :, 3] << 6)).astype(np.uint8)


def orb_compute(img, xs, ys, angles_deg, patch_size=31, edge=EDGE):
    """cv2.ORB_create(); setPatchSize(patch_size); .compute(img, kps) on supplied keypoints (angle used literally):
    border filter [edge, W - edge) x [edge, H - edge), blur, rBRIEF.  Returns (kept indices, descriptors)."""
    H, W = img.shape
    xs, ys = np.asarray(xs), np.asarray(ys)
    keep = np.nonzero((xs >= edge) & (xs < W - edge) & (ys >= edge) & (ys < H - edge))[0]
    blurred = gaussian_blur_7x7(img)
    pat = None if patch_size == 31 else make_random_pattern(patch_size)
    desc = rbrief256(blurred, np.rint(xs[keep]).astype(np.int32), np.rint(ys[keep]).astype(np.int32),
                     np.asarray(angles_deg, np.float32)[keep], pat, raw=img)
    return keep, desc


def orb_detect_and_compute(img, n_features=5000, fast_threshold=15, edge=EDGE, use_fma=False, wta_k=2, harris=False):
    """Full single-level ORB.  Returns dict(x, y, response, angle, desc) in raster order.  harris: scoreType HARRIS_SCORE."""
    xs, ys, resp = _fast.fast_detect(img, fast_threshold, 16, True)
    xs, ys, resp = border_and_retain_best(xs, ys, resp, img.shape, 2 * n_features if harris and n_features > 0 else n_features, edge)
    if harris:
        xs, ys, resp = harris_retain_best(img, xs, ys, resp, n_features)
    ang, _, _ = ic_angle(img, xs, ys, use_fma)
    blurred = gaussian_blur_7x7(img)
    desc = rbrief256(blurred, xs, ys, ang) if wta_k == 2 else rbrief_wta(blurred, xs, ys, ang, wta_k)
    return dict(x=xs, y=ys, response=resp, angle=ang, desc=desc)


# ---- multi-level ORB ("next" row 2) ---------------------------------------------------------------
def _linear_exact_coeffs(src, dst):
    """OpenCV resize.cpp interpolationLinear for INTER_LINEAR_EXACT: source offset and 8.8 fixed-point weights."""
    scale = 1.0 / (float(dst) / float(src))
    ofs = np.zeros(dst, np.int64)
    a1 = np.zeros(dst, np.int64)
    for i in range(dst):
        sf = scale * (i + 0.5) - 0.5
        si = int(np.floor(sf))
        if si < 0:
            ofs[i], a1[i] = 0, 0
        elif si + 1 >= src:
            ofs[i], a1[i] = src - 1, 0
        else:
            ofs[i], a1[i] = si, int(np.rint((sf - si) * 256))
    return ofs, 256 - a1, a1


def resize_linear_exact(img, dw, dh):
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT) for u8: horizontal pass in 8.8 fixed point,
    vertical pass in 16.16, round-half-up to u8.  Pinned bit-exactly against cv2 4.13."""
    H, W = img.shape
    ox, ax0, ax1 = _linear_exact_coeffs(W, dw)
    oy, ay0, ay1 = _linear_exact_coeffs(H, dh)
    src = img.astype(np.int64)
    hor = src[:, ox] * ax0[None, :] + src[:, np.minimum(ox + 1, W - 1)] * ax1[None, :]
    ver = hor[oy, :] * ay0[:, None] + hor[np.minimum(oy + 1, H - 1), :] * ay1[:, None]
    return np.clip((ver + (1 << 15)) >> 16, 0, 255).astype(np.uint8)


def pyramid_quotas(n_features, nlevels, scale_factor=1.2):
    """orb.cpp computeKeyPoints: features per level, geometric in 1/scaleFactor, the last level takes the remainder."""
    factor = _f32(1.0 / np.float64(_f32(scale_factor)))
    ndes = _f32(_f32(_f32(n_features) * _f32(_f32(1) - factor)) / _f32(_f32(1) - _f32(np.power(np.float64(factor), nlevels))))
    out, total = [], 0
    for lvl in range(nlevels - 1):
        out.append(int(np.rint(ndes)))
        total += out[-1]
        ndes = _f32(ndes * factor)
    out.append(max(n_features - total, 0))
    return out


def orb_pyramid_detect_and_compute(img, n_features=5000, nlevels=4, scale_factor=1.2, fast_threshold=15, edge=EDGE,
                                   harris=False):
    """cv2.ORB_create(n_features, scale_factor, nlevels, 31, 0, 2, ORB_FAST_SCORE, 31, fast_threshold)
    .detectAndCompute(img, None): level l = INTER_LINEAR_EXACT resize of level l-1 to (cvRound(W / s^l), cvRound(H / s^l));
    per level FAST-9_16 -> border 31 -> retainBest(quota_l, ties kept) -> IC angle -> blur -> rBRIEF; output
    pt = pt_l * s^l (f32), size = 31 * s^l, octave = l.  Level-major, raster order inside a level."""
    H, W = img.shape
    quotas = pyramid_quotas(n_features, nlevels, scale_factor)
    level_img = img
    out = dict(x=[], y=[], response=[], angle=[], desc=[], octave=[], size=[], lx=[], ly=[])
    for lvl in range(nlevels):
        # ORB stores scaleFactor as a double initialised from the float argument (1.2f -> 1.2000000476837158)
        s = _f32(np.power(np.float64(_f32(scale_factor)), lvl))
        if lvl > 0:
            dw, dh = int(np.rint(W * (1.0 / float(s)))), int(np.rint(H * (1.0 / float(s))))
            level_img = resize_linear_exact(level_img, dw, dh)
        xs, ys, resp = _fast.fast_detect(level_img, fast_threshold, 16, True)
        xs, ys, resp = border_and_retain_best(xs, ys, resp, level_img.shape, 2 * quotas[lvl] if harris else quotas[lvl], edge)
        if harris:
            xs, ys, resp = harris_retain_best(level_img, xs, ys, resp, quotas[lvl])
        ang, _, _ = ic_angle(level_img, xs, ys)
        desc = rbrief256(gaussian_blur_7x7(level_img), xs, ys, ang)
        out["lx"].append(xs); out["ly"].append(ys)
        out["x"].append((xs.astype(np.float32) * s).astype(np.float32))
        out["y"].append((ys.astype(np.float32) * s).astype(np.float32))
        out["response"].append(resp); out["angle"].append(ang); out["desc"].append(desc)
        out["octave"].append(np.full(len(xs), lvl, np.int32))
        out["size"].append(np.full(len(xs), _f32(_f32(31.0) * s), np.float32))
    return {k: np.concatenate(v) for k, v in out.items()}
