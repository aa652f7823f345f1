"""Oracle: SURF / SURF_EXTENDED descriptor at provided keypoints (test infrastructure).

Restates the reference's vendored cv::SURF compute path line by line:
  driver       /root/reference/src/surf.cpp:896-980  (integral CV_32S :913, SURFInvoker, removal of
               keypoints marked size = -1 :953-978)
  worker       src/surf.cpp:563-851 (SURFInvoker::operator()):
     s, grad_wav_size :592-607; orientation :617-670 (113 Haar samples in radius 6s, weights
     getGaussianKernel(13, 2.5), 60-degree window stepped by 5 degrees, fastAtan2(-besty, bestx));
     window extraction :675-769 (rotated bilinear with cvRound :706-740, upright :743-768);
     resize(win, 21x21, INTER_AREA) :772; weighted gradients :775-783; 4x4 cells x {4, 8} sums
     :790-843; f64 square_mag and scale 1/(sqrt + DBL_EPSILON) :846-849.
  helpers      calcHaarPattern :128-134, resizeHaarPattern :136-152.

PARITY UNPINNED as a whole: no SURF binary exists in this image (cv2 4.13 is built without
nonfree/xfeatures2d), and the reference holds no golden vectors.  What IS pinned, in
tests/test_oracle_pins.py, against cv2 primitives the reference delegates to: cv2.integral,
cv2.resize(INTER_AREA) for both the up-scaling (win_size < 21, area-mode fixed-point bilinear) and
the general down-scaling case, cv2.fastAtan2, and getGaussianKernel(13, 2.5).  The 20-tap sigma 3.3
descriptor Gaussian follows the plain formula of OpenCV 2.4, not cv2 4.13's.
cv2.resize(INTER_AREA) has THREE code paths and all three are restated and pinned: win_size < 21
(area-mode fixed-point bilinear), win_size an exact multiple of 21 (OpenCV's ResizeAreaFast: integer
box sums, (a+b+c+d+2)>>2 for x2 and cvRound(sum * (1.f/area)) otherwise -- every size-15 keypoint of
the Fast-Hessian detector has win_size 42), and the general fractional area table.
Stated deviation: sin/cos of the orientation are the correctly rounded float values (the
reference calls libm's float sinf/cosf).
"""
import numpy as np

from .orb import fast_atan2_deg

f32 = np.float32
ORI_RADIUS, ORI_WIN, PATCH_SZ = 6, 60, 20
ORI_SEARCH_INC = 5
DBL_EPSILON = 2.220446049250313e-16


def gaussian_kernel_f32(n, sigma):
    """OpenCV 2.4 getGaussianKernel(n, sigma, CV_32F) for sizes without a fixed table."""
    scale2x = -0.5 / (sigma * sigma)
    x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    cf = np.exp(scale2x * x * x).astype(np.float32)
    s = 1.0 / float(cf.astype(np.float64).sum())
    return (cf.astype(np.float64) * s).astype(np.float32)


G_ORI = gaussian_kernel_f32(2 * ORI_RADIUS + 1, 2.5)
G_DESC = gaussian_kernel_f32(PATCH_SZ, 3.3)
DW = (G_DESC[:, None] * G_DESC[None, :]).astype(np.float32)

# sample offsets (x = outer loop i, y = inner loop j) and weights, src/surf.cpp:541-552
APT = [(i, j) for i in range(-ORI_RADIUS, ORI_RADIUS + 1) for j in range(-ORI_RADIUS, ORI_RADIUS + 1)
       if i * i + j * j <= ORI_RADIUS * ORI_RADIUS]
APTW = np.array([f32(G_ORI[i + ORI_RADIUS] * G_ORI[j + ORI_RADIUS]) for i, j in APT], np.float32)
assert len(APT) == 113

DX_S = [(0, 0, 2, 4, -1), (2, 0, 4, 4, 1)]
DY_S = [(0, 0, 4, 2, 1), (0, 2, 4, 4, -1)]


def cv_round(v):
    return int(np.rint(v))


def integral_i32(img):
    H, W = img.shape
    s = np.zeros((H + 1, W + 1), np.int64)
    s[1:, 1:] = img.astype(np.int64).cumsum(0).cumsum(1)
    return s.astype(np.int32)


def _resize_haar(src, new_size):
    """resizeHaarPattern: list of (dx1, dy1, dx2, dy2, w)."""
    ratio = f32(f32(new_size) / f32(4))
    out = []
    for (a, b, c, d, w) in src:
        dx1, dy1 = cv_round(f32(ratio * f32(a))), cv_round(f32(ratio * f32(b)))
        dx2, dy2 = cv_round(f32(ratio * f32(c))), cv_round(f32(ratio * f32(d)))
        out.append((dx1, dy1, dx2, dy2, f32(f32(w) / f32(f32(dx2 - dx1) * f32(dy2 - dy1)))))
    return out


def _haar(S, y, x, pat):
    d = 0.0
    for (dx1, dy1, dx2, dy2, w) in pat:
        v = int(S[y + dy1, x + dx1]) + int(S[y + dy2, x + dx2]) - int(S[y + dy2, x + dx1]) - int(S[y + dy1, x + dx2])
        d += float(f32(f32(v) * w))
    return f32(d)


def orientation(S, cx, cy, s, grad_wav_size):
    """Dominant orientation in degrees (float32), or None when no sample fits (keypoint dropped)."""
    dx_t, dy_t = _resize_haar(DX_S, grad_wav_size), _resize_haar(DY_S, grad_wav_size)
    half = f32(f32(grad_wav_size - 1) / f32(2))
    X, Y = [], []
    rows, cols = S.shape
    for kk, (ax, ay) in enumerate(APT):
        x = cv_round(f32(f32(cx + f32(f32(ax) * s)) - half))
        y = cv_round(f32(f32(cy + f32(f32(ay) * s)) - half))
        if y < 0 or y >= rows - grad_wav_size or x < 0 or x >= cols - grad_wav_size:
            continue
        vx, vy = _haar(S, y, x, dx_t), _haar(S, y, x, dy_t)
        X.append(f32(vx * APTW[kk]))
        Y.append(f32(vy * APTW[kk]))
    if not X:
        return None
    X, Y = np.array(X, np.float32), np.array(Y, np.float32)
    ang = np.rint(fast_atan2_deg(Y, X)).astype(np.int32)
    bestx = besty = f32(0)
    best_mod = f32(0)
    for i in range(0, 360, ORI_SEARCH_INC):
        d = np.abs(ang - i)
        sel = (d < ORI_WIN // 2) | (d > 360 - ORI_WIN // 2)
        sumx = sumy = f32(0)
        for j in np.nonzero(sel)[0]:
            sumx = f32(sumx + X[j])
            sumy = f32(sumy + Y[j])
        mod = f32(f32(sumx * sumx) + f32(sumy * sumy))
        if mod > best_mod:
            best_mod, bestx, besty = mod, sumx, sumy
    return f32(fast_atan2_deg(np.array([-besty], np.float32), np.array([bestx], np.float32))[0])


def window_upright(img, cx, cy, win_size):
    H, W = img.shape
    win_offset = f32(-f32(win_size - 1) / f32(2))
    start_x = cv_round(f32(cx + win_offset))
    start_y = cv_round(f32(cy - win_offset))
    xs = np.clip(start_x + np.arange(win_size), 0, W - 1)
    ys = np.clip(start_y - np.arange(win_size), 0, H - 1)
    return np.ascontiguousarray(img[ys[None, :], xs[:, None]])      # WIN[i][j] = img(y_j, x_i)


def window_rotated(img, cx, cy, win_size, dir_deg):
    H, W = img.shape
    d = f32(f32(dir_deg) * f32(np.pi / 180.0))
    sin_dir = f32(-f32(np.sin(np.float64(d))))
    cos_dir = f32(np.cos(np.float64(d)))
    win_offset = f32(-f32(win_size - 1) / f32(2))
    start_x = f32(f32(cx + f32(win_offset * cos_dir)) + f32(win_offset * sin_dir))
    start_y = f32(f32(cy - f32(win_offset * sin_dir)) + f32(win_offset * cos_dir))
    ncols1, nrows1 = W - 1, H - 1
    win = np.zeros((win_size, win_size), np.uint8)
    one = f32(1)
    for i in range(win_size):
        px, py = float(start_x), float(start_y)
        for j in range(win_size):
            ix, iy = int(np.floor(px)), int(np.floor(py))
            if 0 <= ix < ncols1 and 0 <= iy < nrows1:
                a, b = f32(px - ix), f32(py - iy)
                p00, p01, p10, p11 = (f32(img[iy, ix]), f32(img[iy, ix + 1]), f32(img[iy + 1, ix]),
                                      f32(img[iy + 1, ix + 1]))
                v = f32(f32(p00 * f32(one - a)) * f32(one - b))
                v = f32(v + f32(f32(p01 * a) * f32(one - b)))
                v = f32(v + f32(f32(p10 * f32(one - a)) * b))
                v = f32(v + f32(f32(p11 * a) * b))
                win[i, j] = np.uint8(cv_round(v))
            else:
                x = min(max(cv_round(px), 0), ncols1)
                y = min(max(cv_round(py), 0), nrows1)
                win[i, j] = img[y, x]
            px += float(cos_dir)
            py -= float(sin_dir)
        start_x = f32(start_x + sin_dir)
        start_y = f32(start_y + cos_dir)
    return win


# ---- cv::resize(win, 21 x 21, INTER_AREA) --------------------------------------------------------------
def _linear_area_tab(S, D):
    """INTER_AREA with scale < 1 falls back to bilinear with area-mode coefficients, fixed point 2^11."""
    scale, inv = S / D, D / S
    ofs, co = [], []
    for dx in range(D):
        sx = int(np.floor(dx * scale))
        fx = f32((dx + 1) - (sx + 1) * inv)
        fx = f32(0) if fx <= 0 else f32(fx - np.floor(fx))
        if sx < 0:
            fx, sx = f32(0), 0
        if sx >= S - 1:
            fx, sx = f32(0), S - 1
        ofs.append(sx)
        co.append((int(np.rint(f32(f32(1) - fx) * f32(2048))), int(np.rint(fx * f32(2048)))))
    return ofs, co


def area_tab(S, D):
    """computeResizeAreaTab: list of (dst index, src index, alpha f32)."""
    scale = S / D
    tab = []
    for dx in range(D):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cw = min(scale, S - fsx1)
        sx1, sx2 = int(np.ceil(fsx1)), int(np.floor(fsx2))
        sx2 = min(sx2, S - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, f32((sx1 - fsx1) / cw)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, f32(1.0 / cw)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, f32(min(min(fsx2 - sx2, 1.0), cw) / cw)))
    return tab


def resize_area_21(win):
    S, D = win.shape[0], PATCH_SZ + 1
    if S == D:
        return win.copy()
    if S % D == 0:
        # integer decimation (is_area_fast in cv::resize): ResizeAreaFastVec for x2 rounds half up in integers,
        # ResizeAreaFast_Invoker multiplies the int box sum by the float 1.f/area and cvRounds (half even)
        k = S // D
        box = win.astype(np.int64).reshape(D, k, D, k).sum(axis=(1, 3))
        if k == 2:
            return ((box + 2) >> 2).astype(np.uint8)
        scale = f32(f32(1) / f32(k * k))
        return np.clip(np.rint(box.astype(np.float32) * scale), 0, 255).astype(np.uint8)
    if S < D:
        ofs, co = _linear_area_tab(S, D)
        src = win.astype(np.int64)
        Hh = np.zeros((S, D), np.int64)
        for dx in range(D):
            sx, (a0, a1) = ofs[dx], co[dx]
            Hh[:, dx] = src[:, sx] * a0 + src[:, min(sx + 1, S - 1)] * a1
        out = np.zeros((D, D), np.uint8)
        for dy in range(D):
            sy, (b0, b1) = ofs[dy], co[dy]
            out[dy] = ((((b0 * (Hh[sy] >> 4)) >> 16) + ((b1 * (Hh[min(sy + 1, S - 1)] >> 4)) >> 16) + 2) >> 2).astype(np.uint8)
        return out
    tab = area_tab(S, D)
    F = win.astype(np.float32)
    res = np.zeros((D, D), np.float32)
    bufs = {}

    def rowbuf(sy):
        if sy not in bufs:
            buf = np.zeros(D, np.float32)
            for (dx, sx, a) in tab:
                buf[dx] = f32(buf[dx] + f32(F[sy, sx] * a))
            bufs[sy] = buf
        return bufs[sy]

    prev, sums = -1, None
    for (dy, sy, b) in tab:
        t = (b * rowbuf(sy)).astype(np.float32)
        if dy != prev:
            if prev >= 0:
                res[prev] = sums
            sums, prev = t, dy
        else:
            sums = (sums + t).astype(np.float32)
    res[prev] = sums
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def descriptor_from_patch(patch, extended):
    """patch: 21 x 21 u8 -> 64 / 128 float32 (src/surf.cpp:775-849)."""
    P = patch.astype(np.int32)
    DXm = ((P[:-1, 1:] - P[:-1, :-1] + P[1:, 1:] - P[1:, :-1]).astype(np.float32) * DW).astype(np.float32)
    DYm = ((P[1:, :-1] - P[:-1, :-1] + P[1:, 1:] - P[:-1, 1:]).astype(np.float32) * DW).astype(np.float32)
    n = 8 if extended else 4
    vec = np.zeros(16 * n, np.float32)
    sq = 0.0
    for i in range(4):
        for j in range(4):
            v = np.zeros(n, np.float32)
            for y in range(i * 5, i * 5 + 5):
                for x in range(j * 5, j * 5 + 5):
                    tx, ty = DXm[y, x], DYm[y, x]
                    if extended:
                        if ty >= 0:
                            v[0] = f32(v[0] + tx); v[1] = f32(v[1] + abs(tx))
                        else:
                            v[2] = f32(v[2] + tx); v[3] = f32(v[3] + abs(tx))
                        if tx >= 0:
                            v[4] = f32(v[4] + ty); v[5] = f32(v[5] + abs(ty))
                        else:
                            v[6] = f32(v[6] + ty); v[7] = f32(v[7] + abs(ty))
                    else:
                        v[0] = f32(v[0] + tx); v[1] = f32(v[1] + ty)
                        v[2] = f32(v[2] + abs(tx)); v[3] = f32(v[3] + abs(ty))
            for kk in range(n):
                sq += float(f32(v[kk] * v[kk]))
            vec[(i * 4 + j) * n:(i * 4 + j + 1) * n] = v
    scale = f32(1.0 / (np.sqrt(sq) + DBL_EPSILON))
    return (vec * scale).astype(np.float32)


def surf_compute(img, xs, ys, sizes, extended=True, upright=True):
    """cv::SURF::operator()(img, noArray(), keypoints, descriptors, useProvidedKeypoints=true).

    Returns (keep bool[N], angle f32[N], desc f32[n_kept x 64/128]); keep marks the keypoints that
    survive (size != -1)."""
    H, W = img.shape
    S = integral_i32(img) if not upright else None
    n = len(xs)
    keep = np.ones(n, bool)
    angles = np.full(n, 270.0, np.float32)
    descs = []
    for k in range(n):
        cx, cy, size = f32(xs[k]), f32(ys[k]), f32(sizes[k])
        s = f32(f32(size * f32(1.2)) / f32(9.0))
        grad_wav_size = 2 * cv_round(f32(f32(2) * s))
        if H + 1 < grad_wav_size or W + 1 < grad_wav_size:
            keep[k] = False
            continue
        dir_deg = f32(270.0)
        if not upright:
            o = orientation(S, cx, cy, s, grad_wav_size)
            if o is None:
                keep[k] = False
                continue
            dir_deg = o
        angles[k] = dir_deg
        win_size = int(f32(f32(PATCH_SZ + 1) * s))
        win = window_upright(img, cx, cy, win_size) if upright else window_rotated(img, cx, cy, win_size, dir_deg)
        descs.append(descriptor_from_patch(resize_area_21(win), extended))
    d = np.stack(descs) if descs else np.zeros((0, 128 if extended else 64), np.float32)
    return keep, angles, d
