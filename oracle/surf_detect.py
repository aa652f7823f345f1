"""CPU restatement of the vendored SURF Fast-Hessian detector (TEST INFRASTRUCTURE ONLY).

Follows /root/reference src/surf.cpp line by line:
  * resizeHaarPattern :136-152, calcHaarPattern :128-134 (int box sums * float weight, accumulated in double);
  * calcLayerDetAndTrace :167-206  (9x9 box-filter Hessian scaled to `size`, det = dx*dy - 0.81*dxy^2, trace = dx+dy);
  * findMaximaInLayer :346-443     (threshold, strict 3x3x3 non-maximum suppression, sub-sample interpolation);
  * interpolateKeypoint :228-258   (central differences, 3x3 solve -- OpenCV 2.4's Matx<float,3,3>::solve(DECOMP_LU) is
                                    the closed-form Cramer rule of Matx_FastSolveOp, restated here in float32);
  * fastHessianDetector :462-512   (octave / layer schedule, final std::sort with KeypointGreater :445-460).

PARITY UNPINNED: no SURF binary exists in this image (cv2 4.13 headless has no xfeatures2d, and src/surf.cpp needs
OpenCV-2.4 internal headers to compile), so this restatement is checked only for internal consistency (tests compare
its box responses with an independent float64 box-filter evaluation) and is the reference the CUDA path is compared with.
"""
import numpy as np

from .surf import cv_round, f32, integral_i32

HAAR_SIZE0, HAAR_SIZE_INC = 9, 6
DX_S = [(0, 2, 3, 7, 1), (3, 2, 6, 7, -2), (6, 2, 9, 7, 1)]
DY_S = [(2, 0, 7, 3, 1), (2, 3, 7, 6, -2), (2, 6, 7, 9, 1)]
DXY_S = [(1, 1, 4, 4, 1), (5, 1, 8, 4, -1), (1, 5, 4, 8, -1), (5, 5, 8, 8, 1)]


def resize_haar9(src, new_size):
    ratio = f32(f32(new_size) / f32(9))
    out = []
    for (a, b, c, d, w) in src:
        dx1, dy1 = cv_round(f32(ratio * f32(a))), cv_round(f32(ratio * f32(b)))
        dx2, dy2 = cv_round(f32(ratio * f32(c))), cv_round(f32(ratio * f32(d)))
        out.append((dx1, dy1, dx2, dy2, f32(f32(w) / f32(f32(dx2 - dx1) * f32(dy2 - dy1)))))
    return out


def _haar_map(S, pat, ys, xs):
    """calcHaarPattern at every (ys[i], xs[j]): float32 result of a double accumulation of float products."""
    acc = np.zeros((len(ys), len(xs)), np.float64)
    Y, X = ys[:, None], xs[None, :]
    for (dx1, dy1, dx2, dy2, w) in pat:
        v = (S[Y + dy1, X + dx1].astype(np.int64) + S[Y + dy2, X + dx2] - S[Y + dy2, X + dx1] - S[Y + dy1, X + dx2])
        acc += (v.astype(np.float32) * w).astype(np.float32).astype(np.float64)
    return acc.astype(np.float32)


def layer_det_trace(S, size, step):
    R, C = S.shape[0] - 1, S.shape[1] - 1
    det = np.zeros((R // step, C // step), np.float32)
    trace = np.zeros_like(det)
    if size > R or size > C:
        return det, trace
    si, sj = 1 + (R - size) // step, 1 + (C - size) // step
    margin = (size // 2) // step
    ys, xs = np.arange(si) * step, np.arange(sj) * step
    dx = _haar_map(S, resize_haar9(DX_S, size), ys, xs)
    dy = _haar_map(S, resize_haar9(DY_S, size), ys, xs)
    dxy = _haar_map(S, resize_haar9(DXY_S, size), ys, xs)
    det[margin:margin + si, margin:margin + sj] = ((dx * dy).astype(np.float32) -
                                                   ((f32(0.81) * dxy).astype(np.float32) * dxy).astype(np.float32)).astype(np.float32)
    trace[margin:margin + si, margin:margin + sj] = (dx + dy).astype(np.float32)
    return det, trace


def _solve3(A, b):
    """Matx<float,3,3>::solve(b, DECOMP_LU) == Matx_FastSolveOp<float,3,1>: Cramer's rule in float32.  None if det == 0."""
    a = A
    d = f32(f32(f32(a[0][0] * f32(f32(a[1][1] * a[2][2]) - f32(a[2][1] * a[1][2]))) -
                f32(a[0][1] * f32(f32(a[1][0] * a[2][2]) - f32(a[2][0] * a[1][2])))) +
            f32(a[0][2] * f32(f32(a[1][0] * a[2][1]) - f32(a[2][0] * a[1][1]))))
    if d == 0:
        return None
    d = f32(f32(1) / d)

    def m2(p, q, r, s):
        return f32(f32(p * q) - f32(r * s))
    x0 = f32(d * f32(f32(f32(b[0] * m2(a[1][1], a[2][2], a[1][2], a[2][1])) -
                         f32(a[0][1] * m2(b[1], a[2][2], a[1][2], b[2]))) +
                     f32(a[0][2] * m2(b[1], a[2][1], a[1][1], b[2]))))
    x1 = f32(d * f32(f32(f32(a[0][0] * m2(b[1], a[2][2], a[1][2], b[2])) -
                         f32(b[0] * m2(a[1][0], a[2][2], a[1][2], a[2][0]))) +
                     f32(a[0][2] * m2(a[1][0], b[2], b[1], a[2][0]))))
    x2 = f32(d * f32(f32(f32(a[0][0] * m2(a[1][1], b[2], b[1], a[2][1])) -
                         f32(a[0][1] * m2(a[1][0], b[2], b[1], a[2][0]))) +
                     f32(b[0] * m2(a[1][0], a[2][1], a[1][1], a[2][0]))))
    return x0, x1, x2


def interpolate_keypoint(N9, dx, dy, ds, x, y, size):
    """Returns (ok, x, y, size)."""
    two, four = f32(2), f32(4)
    b = (f32(-f32(f32(N9[1][5] - N9[1][3]) / two)), f32(-f32(f32(N9[1][7] - N9[1][1]) / two)),
         f32(-f32(f32(N9[2][4] - N9[0][4]) / two)))
    dxy_ = f32(f32(f32(f32(N9[1][8] - N9[1][6]) - N9[1][2]) + N9[1][0]) / four)
    dxs_ = f32(f32(f32(f32(N9[2][5] - N9[2][3]) - N9[0][5]) + N9[0][3]) / four)
    dys_ = f32(f32(f32(f32(N9[2][7] - N9[2][1]) - N9[0][7]) + N9[0][1]) / four)
    A = ((f32(f32(N9[1][3] - f32(two * N9[1][4])) + N9[1][5]), dxy_, dxs_),
         (dxy_, f32(f32(N9[1][1] - f32(two * N9[1][4])) + N9[1][7]), dys_),
         (dxs_, dys_, f32(f32(N9[0][4] - f32(two * N9[1][4])) + N9[2][4])))
    sol = _solve3(A, b)
    if sol is None:
        sol = (f32(0), f32(0), f32(0))
    x0, x1, x2 = sol
    ok = (x0 != 0 or x1 != 0 or x2 != 0) and abs(x0) <= 1 and abs(x1) <= 1 and abs(x2) <= 1
    if ok:
        x = f32(x + f32(x0 * f32(dx)))
        y = f32(y + f32(x1 * f32(dy)))
        size = f32(cv_round(f32(size + f32(x2 * f32(ds)))))
    return ok, x, y, size


def fast_hessian(img, hessian_threshold=100.0, n_octaves=4, n_octave_layers=2):
    """fastHessianDetector.  Returns a structured array (x, y, size, response, octave, laplacian) in the reference's
    final order (std::sort with KeypointGreater)."""
    S = integral_i32(img)
    R, C = S.shape[0] - 1, S.shape[1] - 1
    sizes, steps, middle = [], [], []
    step = 1
    for o in range(n_octaves):
        for l in range(n_octave_layers + 2):
            sizes.append((HAAR_SIZE0 + HAAR_SIZE_INC * l) << o)
            steps.append(step)
            if 0 < l <= n_octave_layers:
                middle.append(len(sizes) - 1)
        step *= 2
    dets, traces = zip(*[layer_det_trace(S, sizes[i], steps[i]) for i in range(len(sizes))])
    thr = f32(hessian_threshold)
    out = []
    for mi, layer in enumerate(middle):
        octave = mi // n_octave_layers
        size, st = sizes[layer], steps[layer]
        rows, cols = R // st, C // st
        margin = (sizes[layer + 1] // 2) // st + 1
        if rows - margin <= margin or cols - margin <= margin:
            continue
        d0, d1, d2 = dets[layer - 1], dets[layer], dets[layer + 1]
        sl = (slice(margin, rows - margin), slice(margin, cols - margin))
        c = d1[sl]
        keep = c > thr
        for dd in (d0, d1, d2):
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    if dd is d1 and di == 0 and dj == 0:
                        continue
                    keep &= c > dd[margin + di:rows - margin + di, margin + dj:cols - margin + dj]
        for (ii, jj) in zip(*np.nonzero(keep)):
            i, j = ii + margin, jj + margin
            sum_i, sum_j = st * (i - (size // 2) // st), st * (j - (size // 2) // st)
            N9 = [[f32(v) for v in dd[i - 1:i + 2, j - 1:j + 2].ravel()] for dd in (d0, d1, d2)]
            cy, cx = f32(f32(sum_i) + f32(f32(size - 1) * f32(0.5))), f32(f32(sum_j) + f32(f32(size - 1) * f32(0.5)))
            ok, x, y, sz = interpolate_keypoint(N9, st, st, size - sizes[layer - 1], cx, cy, f32(size))
            if ok:
                t = traces[layer][i, j]
                out.append((x, y, sz, d1[i, j], octave, 1 if t > 0 else (-1 if t < 0 else 0)))
    kp = np.array(out, dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("response", "<f4"), ("octave", "<i4"),
                              ("laplacian", "<i4")])
    # KeypointGreater: response desc, size desc, octave desc, y desc, x asc
    order = np.lexsort((kp["x"], -kp["y"], -kp["octave"], -kp["size"], -kp["response"]))
    return kp[order]
