"""Oracle: brute-force descriptor matching + the reference's mask / ratio / cross-check glue
(test infrastructure).

Restates cv::BFMatcher::knnMatch(q, t, k=2, mask) and BFMatcher(crossCheck=true)::match as the
reference calls them, and the host loops around them:
  * epipolar-band mask            /root/reference/src/StereoCamera.cpp:182-196 (|dy| <= 1 after ROI
                                  offsets), src/front_end/algorithm.py:825-836 (|dy| <= 2.0)
  * masked kNN-2                  StereoCamera.cpp:199-201, WindowMatcher.cpp:150-153, algorithm.py:848-853
  * Lowe ratio 0.8 + singleton    StereoCamera.cpp:206-264, WindowMatcher.cpp:161-224, algorithm.py:838-846
  * cross-check + |dy| <= 0.7     src/live_stereo.cpp:240,364-377, features.py:670,724-733
  * search-window box mask        WindowMatcher.cpp:104-128 (|dx| < w/2 and |dy| < h/2, Rect(0,0,100,100))
Semantics of the OpenCV calls per SURVEY.md Appendix A.5, pinned against cv2 4.13.0 in
tests/test_oracle_pins.py.
"""
import numpy as np

NO_MATCH = -1


def hamming_matrix(qd, td, chunk=512):
    """N x M uint16 Hamming distances between rows of two u8 descriptor matrices (D % 8 == 0)."""
    q = np.ascontiguousarray(qd).view(np.uint64)
    t = np.ascontiguousarray(td).view(np.uint64)
    out = np.empty((q.shape[0], t.shape[0]), np.uint16)
    for s in range(0, q.shape[0], chunk):
        x = q[s:s + chunk, None, :] ^ t[None, :, :]
        out[s:s + chunk] = np.bitwise_count(x).sum(axis=2, dtype=np.uint16)
    return out


def hamming2_matrix(qd, td, chunk=512):
    """cv::NORM_HAMMING2: number of differing two-bit symbols (ORB WTA_K 3 / 4; src/StereoCamera.cpp:504-511)."""
    q = np.ascontiguousarray(qd).view(np.uint64)
    t = np.ascontiguousarray(td).view(np.uint64)
    out = np.empty((q.shape[0], t.shape[0]), np.uint16)
    m = np.uint64(0x5555555555555555)
    for s in range(0, q.shape[0], chunk):
        x = q[s:s + chunk, None, :] ^ t[None, :, :]
        out[s:s + chunk] = np.bitwise_count((x | (x >> np.uint64(1))) & m).sum(axis=2, dtype=np.uint16)
    return out


def _dist_matrix(qd, td, norm):
    return {"hamming": hamming_matrix, "hamming2": hamming2_matrix, "l2": l2_matrix}[norm](qd, td)


def l2_matrix(qd, td, chunk=512):
    """N x M float32 L2 distances (sqrt, not squared), accumulated in float64."""
    q = np.asarray(qd, np.float64)
    t = np.asarray(td, np.float64)
    out = np.empty((q.shape[0], t.shape[0]), np.float32)
    tn = (t * t).sum(axis=1)
    for s in range(0, q.shape[0], chunk):
        qq = q[s:s + chunk]
        d2 = (qq * qq).sum(axis=1)[:, None] + tn[None, :] - 2.0 * (qq @ t.T)
        # recompute exactly where cancellation matters is unnecessary at 1e-4 tolerance,
        # but keep non-negativity
        out[s:s + chunk] = np.sqrt(np.maximum(d2, 0.0)).astype(np.float32)
    return out


def epipolar_mask(ly, ry, threshold, l_off=0.0, r_off=0.0):
    """mask[i,j] = |(ly_i + l_off) - (ry_j + r_off)| <= threshold   (float32 arithmetic)."""
    a = np.asarray(ly, np.float32) + np.float32(l_off)
    b = np.asarray(ry, np.float32) + np.float32(r_off)
    return np.abs(a[:, None] - b[None, :]) <= np.float32(threshold)


def window_mask(cx, cy, px, py, width=100, height=100):
    """mask[i,j] = |cx_i - px_j| < width/2 and |cy_i - py_j| < height/2 (integer halves, strict)."""
    cx, cy = np.asarray(cx, np.float32), np.asarray(cy, np.float32)
    px, py = np.asarray(px, np.float32), np.asarray(py, np.float32)
    hw, hh = np.float32(int(width) // 2), np.float32(int(height) // 2)
    return (np.abs(cx[:, None] - px[None, :]) < hw) & (np.abs(cy[:, None] - py[None, :]) < hh)


def knn2(dist, mask=None):
    """Per query: the <=2 smallest allowed distances, ties -> lower train index.

    Returns (idx N x 2 int32 with -1 for absent, d N x 2 float32 (inf for absent), count N)."""
    D = dist.astype(np.float32)
    if mask is not None:
        D = np.where(mask, D, np.float32(np.inf))
    n, m = D.shape
    idx = np.full((n, 2), NO_MATCH, np.int32)
    dd = np.full((n, 2), np.inf, np.float32)
    if m == 0 or n == 0:
        return idx, dd, np.zeros(n, np.int32)
    rows = np.arange(n)
    i0 = np.argmin(D, axis=1)           # first occurrence of the minimum
    d0 = D[rows, i0]
    ok0 = np.isfinite(d0)
    idx[ok0, 0] = i0[ok0]
    dd[ok0, 0] = d0[ok0]
    if m > 1:
        D2 = D.copy()
        D2[rows, i0] = np.inf
        i1 = np.argmin(D2, axis=1)
        d1 = D2[rows, i1]
        ok1 = np.isfinite(d1)
        idx[ok1, 1] = i1[ok1]
        dd[ok1, 1] = d1[ok1]
    return idx, dd, (idx >= 0).sum(axis=1).astype(np.int32)


def lowe_ratio(idx, dd, ratio=0.8):
    """Accept singleton rows, and rows with d0 < ratio*d1 (double arithmetic, strict).
    Returns (queryIdx, trainIdx, distance) ordered by query."""
    cnt = (idx >= 0).sum(axis=1)
    d0 = dd[:, 0].astype(np.float64)
    d1 = dd[:, 1].astype(np.float64)
    with np.errstate(invalid="ignore"):
        good = (cnt == 1) | ((cnt >= 2) & (d0 < float(ratio) * d1))
    q = np.nonzero(good)[0].astype(np.int32)
    return q, idx[q, 0], dd[q, 0]


def cross_check(dist):
    """BFMatcher(crossCheck=True).match: mutual first-argmin, ordered by queryIdx."""
    n, m = dist.shape
    if n == 0 or m == 0:
        z = np.zeros(0, np.int32)
        return z, z, np.zeros(0, np.float32)
    s = np.argmin(dist, axis=1)       # nearest train of each query (first min)
    tq = np.argmin(dist, axis=0)      # nearest query of each train (first min)
    q = np.nonzero(tq[s] == np.arange(n))[0].astype(np.int32)
    return q, s[q].astype(np.int32), dist[q, s[q]].astype(np.float32)


def stereo_match_ratio(lkp_y, rkp_y, ld, rd, epi_threshold=2.0, ratio=0.8, norm="hamming",
                       l_off=0.0, r_off=0.0):
    """Path A (algorithm_one / StereoCamera::processStereo): band mask -> kNN-2 -> ratio."""
    D = _dist_matrix(ld, rd, norm)
    mask = epipolar_mask(lkp_y, rkp_y, epi_threshold, l_off, r_off)
    idx, dd, _ = knn2(D, mask)
    return lowe_ratio(idx, dd, ratio)


def stereo_match_crosscheck(lkp_y, rkp_y, ld, rd, max_dy=0.7, norm="hamming"):
    """Path B (live nodes): cross-check match, then keep |yL - yR| <= max_dy."""
    D = _dist_matrix(ld, rd, norm)
    q, t, d = cross_check(D)
    ly = np.asarray(lkp_y, np.float32)
    ry = np.asarray(rkp_y, np.float32)
    keep = np.abs(ly[q] - ry[t]) <= np.float32(max_dy)
    return q[keep], t[keep], d[keep]


def window_match(cur_xy, prev_xy, cur_desc, prev_desc, width=100, height=100, ratio=0.8,
                 norm="hamming"):
    """WindowMatcher::newStereo matching stage (WindowMatcher.cpp:104-231)."""
    D = _dist_matrix(cur_desc, prev_desc, norm)
    mask = window_mask(cur_xy[:, 0], cur_xy[:, 1], prev_xy[:, 0], prev_xy[:, 1], width, height)
    idx, dd, _ = knn2(D, mask)
    return lowe_ratio(idx, dd, ratio)


def live_graph_tracks(cur_ldesc, prev_ldesc, cur_rdesc, prev_rdesc, norm="hamming"):
    """liveGraph.updateMatches (src/front_end/algorithm.py:1122,1160-1190): bf.match(crossCheck) of the current vs previous
    LEFT descriptors and of the RIGHT descriptors; current landmark c continues previous landmark t iff (c, t) is a
    mutual match on both sides.  Returns (cur index, prev index, left distance) ordered by cur index."""
    ql, tl, dl = cross_check(_dist_matrix(cur_ldesc, prev_ldesc, norm))
    qr, tr, _ = cross_check(_dist_matrix(cur_rdesc, prev_rdesc, norm))
    right = dict(zip(qr.tolist(), tr.tolist()))
    keep = np.array([right.get(int(q), -1) == int(t) for q, t in zip(ql, tl)], bool)
    return ql[keep], tl[keep], dl[keep]
