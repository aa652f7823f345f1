"""CPU restatement of cv::cornerSubPix and the 2x3 grid FAST detector of the live nodes (TEST INFRASTRUCTURE
ONLY -- never imported by front_end_b200/).

Reference call sites (RyanEvanWolf/front_end):
  * src/live_stereo.cpp:235-237 (winSize 5x5, zeroZone -1, 40 iterations / eps 1e-3), :321-337 (one call per
    keypoint, on the CELL sub-image, before the cell / ROI offsets are added at :340-350);
  * src/front_end/features.py:600-601, 637-640 (same parameters, on the FULL rectified image after offsets).
  * grid: src/live_stereo.cpp:277-318 (2x3 cells of the ROI, FASTX TYPE_7_12 per cell with the cell's own
    threshold, controller step after each cell), features.py:609-636 (Python twin).

The arithmetic lives in OpenCV (un-vendored): imgproc/src/cornersubpix.cpp and imgproc/src/samplers.cpp
(getRectSubPix_8u32f).  Restated here from the published algorithm and pinned bit-for-bit against the cv2
4.13.0 of this image (tests/test_oracle_pins.py::test_cornersubpix_pinned_against_cv2 and the golden fixture
tests/golden/grid_subpix_*.npz).
"""
import math

import numpy as np

from . import fast as ofast

F32 = np.float32
DBL_EPS = np.finfo(np.float64).eps


def rect_subpix_8u32f(img, win_w, win_h, cx, cy):
    """cv::getRectSubPix(img u8, Size(win_w, win_h), center, CV_32F) -- samplers.cpp getRectSubPix_Cn_<uchar,
    float, float> + adjustRect: float bilinear weights, replicated border.  (OpenCV 2.4's getRectSubPix_8u32f
    used a running-sum form with a = max(a, 1e-4); cv2 4.13 -- the pin -- does not: integer centres return the
    pixels exactly.)"""
    H, W = img.shape
    cx0, cy0 = F32(cx), F32(cy)
    x = F32(cx0 - F32((win_w - 1) * 0.5))
    y = F32(cy0 - F32((win_h - 1) * 0.5))
    ipx, ipy = int(np.floor(x)), int(np.floor(y))
    out = np.empty((win_h, win_w), F32)
    # generic path: coefficients in float, border replicated through adjustRect
    a = F32(x - F32(ipx))
    b = F32(y - F32(ipy))
    a11 = F32(F32(F32(1) - a) * F32(F32(1) - b))
    a12 = F32(a * F32(F32(1) - b))
    a21 = F32(F32(F32(1) - a) * b)
    a22 = F32(a * b)
    b1 = F32(F32(1) - b)
    b2 = b
    interior = 0 <= ipx < W - win_w and 0 <= ipy < H - win_h
    # adjustRect: the source window [ip, ip + win) clipped to the image; r = (x0, y0, x1, y1) in window coords
    if ipx >= 0:
        sx, rx = ipx, 0
    else:
        sx, rx = 0, min(-ipx, win_w)
    if ipx < W - win_w:
        rw = win_w
    else:
        rw = W - ipx - 1
        if rw < 0:
            sx += rw
            rw = 0
    if ipy >= 0:
        sy, ry = ipy, 0
    else:
        sy, ry = 0, -ipy
    if ipy < H - win_h:
        rh = win_h
    else:
        rh = H - ipy - 1
        if rh < 0:
            sy += rh
            rh = 0
    base_x = sx - rx      # src pointer is moved back by r.x so that window column j reads src[j]
    row = sy
    imgf = img.astype(F32)
    for i in range(win_h):
        row2 = row + 1
        if i < ry or i >= rh:
            row2 = row
        r0 = imgf[min(max(row, 0), H - 1)]
        r1 = imgf[min(max(row2, 0), H - 1)]

        def px(rr, j):
            return rr[min(max(base_x + j, 0), W - 1)]
        s0 = F32(F32(px(r0, rx) * b1) + F32(px(r1, rx) * b2))
        out[i, :rx] = s0
        # cv2 4.13 quirk (pinned empirically): in rows replicated ABOVE the image the right-hand fill reads the
        # pixel before the clamp column (W - 2 instead of W - 1); every other edge replicates the edge pixel
        jr = rw - 1 if i < ry else rw
        s0 = F32(F32(px(r0, jr) * b1) + F32(px(r1, jr) * b2))
        out[i, rw:] = s0
        for j in range(rx, rw):
            # association pinned against cv2 4.13 (no FMA): (p00*a11 + p01*a12) + (p10*a21 + p11*a22)
            if row2 != row:
                out[i, j] = F32(F32(F32(px(r0, j) * a11) + F32(px(r0, j + 1) * a12)) +
                                F32(F32(px(r1, j) * a21) + F32(px(r1, j + 1) * a22)))
            else:
                # replicated row (above / below the image): cv2 4.13 returns fma(p01, a, p00 * (1 - a)) -- pinned
                # empirically on 149k samples; the double expression below is an exact FMA (8-bit x 24-bit
                # product + f32 addend fits in 53 bits)
                out[i, j] = F32(np.float64(px(r0, j + 1)) * np.float64(a) + np.float64(F32(px(r0, j) * F32(F32(1) - a))))
        if i < rh:
            row = row2
    return out


def _mask(win):
    n = 2 * win + 1
    idx = (np.arange(n, dtype=F32) - F32(win)) / F32(win)
    # std::exp(float) -> glibc expf, correctly rounded: exp in double, rounded once to float.  (numpy's SIMD
    # float32 exp is 1 ulp off for 8 of the 11 taps, which showed up as 6 % of the refined points differing.)
    e = np.array([F32(math.exp(float(F32(-(v * v))))) for v in idx], F32)
    return (e[:, None] * e[None, :]).astype(F32)


def corner_subpix(img, pts, win=5, max_iters=40, epsilon=0.001):
    """cv::cornerSubPix(img, pts, Size(win, win), Size(-1, -1), TermCriteria(EPS + ITER, max_iters, epsilon)).
    pts: (n, 2) float32 (x, y); returns refined (n, 2) float32."""
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    n = 2 * win + 1
    mask = _mask(win).astype(np.float64)
    eps2 = float(epsilon) * float(epsilon)
    px = (np.arange(n) - win).astype(np.float64)[None, :]
    py = (np.arange(n) - win).astype(np.float64)[:, None]
    out = np.array(pts, F32).reshape(-1, 2).copy()
    for k in range(len(out)):
        cTx, cTy = F32(out[k, 0]), F32(out[k, 1])
        cIx, cIy = cTx, cTy
        it = 0
        while True:
            sp = rect_subpix_8u32f(img, n + 2, n + 2, cIx, cIy)
            tgx = (sp[1:-1, 2:] - sp[1:-1, :-2]).astype(F32).astype(np.float64)
            tgy = (sp[2:, 1:-1] - sp[:-2, 1:-1]).astype(F32).astype(np.float64)
            gxx = tgx * tgx * mask
            gxy = tgx * tgy * mask
            gyy = tgy * tgy * mask
            # sequential double accumulation in raster order (np.sum would pair-sum: different rounding)
            a = b = c = bb1 = bb2 = 0.0
            t1 = gxx * px + gxy * py
            t2 = gxy * px + gyy * py
            for v in gxx.ravel():
                a += v
            for v in gxy.ravel():
                b += v
            for v in gyy.ravel():
                c += v
            for v in t1.ravel():
                bb1 += v
            for v in t2.ravel():
                bb2 += v
            det = a * c - b * b
            if abs(det) <= DBL_EPS * DBL_EPS:
                break
            scale = 1.0 / det
            nx = F32(float(cIx) + c * scale * bb1 - b * scale * bb2)
            ny = F32(float(cIy) - b * scale * bb1 + a * scale * bb2)
            dx, dy = F32(nx - cIx), F32(ny - cIy)
            err = float(F32(F32(dx * dx) + F32(dy * dy)))
            # cv2 4.13 (the pin): a step that leaves the image is discarded -- the estimate stays at the last
            # in-bounds value (OpenCV 2.4 assigned first and broke afterwards)
            if nx < 0 or nx >= W or ny < 0 or ny >= H:
                break
            cIx, cIy = nx, ny
            it += 1
            if not (it < max_iters and err > eps2):
                break
        if abs(float(cIx) - float(cTx)) > win or abs(float(cIy) - float(cTy)) > win:
            cIx, cIy = cTx, cTy
        out[k] = (cIx, cIy)
    return out


def grid_cells(roi, rows=2, cols=3, python_variant=False):
    """Cell rectangles (x, y, w, h) in full-image coordinates, row-major.
    C++ (live_stereo.cpp:150-153,282-287): gridWidth = roi.width / cols, gridHeight = roi.height / rows (integer
    division) inside the ROI.  Python (features.py:610-620): the ROI slice is [y : h + 1, x : w + 1] -- width and
    height are used as END coordinates -- and the cell size is int(w / cols) x int(h / rows), clipped by the slice."""
    x, y, w, h = roi
    cells = []
    if python_variant:
        cw, ch = int(w / cols), int(h / rows)
        x_end, y_end = w + 1, h + 1
        for r in range(rows):
            for c in range(cols):
                x0, y0 = x + c * cw, y + r * ch
                x1, y1 = min(x0 + cw, x_end), min(y0 + ch, y_end)
                cells.append((x0, y0, max(x1 - x0, 0), max(y1 - y0, 0)))
    else:
        cw, ch = w // cols, h // rows
        for r in range(rows):
            for c in range(cols):
                cells.append((x + c * cw, y + r * ch, cw, ch))
    return cells


def grid_detect(img, roi, thresholds, set_point, rows=2, cols=3, ps=12, python_variant=False, subpix=True,
                update=True, subpix_step=1):
    """One frame of the live nodes' detector for one eye.  Returns (pts (n, 2) f32 full-image coordinates,
    responses (n,), per-cell counts (rows, cols), new thresholds (rows, cols)).
    Keypoints are concatenated cell by cell (row-major), raster order inside a cell.  subpix_step > 1 refines
    only every subpix_step-th keypoint (global index) -- the pure-Python refinement is slow; the others keep their
    integer position."""
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    thr = np.array(thresholds, np.int64).reshape(rows, cols)
    counts = np.zeros((rows, cols), np.int64)
    pts, resp = [], []
    n_before = 0
    cells = grid_cells(roi, rows, cols, python_variant)
    for idx, (x0, y0, cw, ch) in enumerate(cells):
        r, c = divmod(idx, cols)
        x1, y1 = min(x0 + cw, W), min(y0 + ch, H)
        cell = img[y0:y1, x0:x1]
        if cell.shape[0] < 7 or cell.shape[1] < 7:
            xs = ys = sc = np.zeros(0, np.int64)
        else:
            xs, ys, sc = ofast.fast_detect(cell, int(thr[r, c]), ps, True)
        counts[r, c] = len(xs)
        p = np.stack([xs.astype(F32), ys.astype(F32)], 1) if len(xs) else np.zeros((0, 2), F32)
        if subpix and not python_variant and len(p):
            sel = np.nonzero((np.arange(len(p)) + n_before) % subpix_step == 0)[0]
            if len(sel):
                p[sel] = corner_subpix(cell, p[sel])   # C++: on the cell sub-image, cell coordinates
        n_before += len(p)
        if len(p):
            p = (p + np.array([x0, y0], F32)).astype(F32)
        pts.append(p)
        resp.append(sc.astype(F32))
    pts = np.concatenate(pts) if pts else np.zeros((0, 2), F32)
    resp = np.concatenate(resp) if resp else np.zeros(0, F32)
    if subpix and python_variant and len(pts):
        sel = np.arange(0, len(pts), subpix_step)
        pts[sel] = corner_subpix(img, pts[sel])        # Python: on the full image, after offsets
    new_thr = ofast.setpoint_step(thr, counts, set_point, rows, cols, lo=6 if python_variant else 4, hi=80,
                                  python_variant=python_variant) if update else thr
    return pts, resp, counts, np.asarray(new_thr)
