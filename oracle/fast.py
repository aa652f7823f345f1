"""Oracle: FAST corner detector with score + 3x3 non-max suppression (test infrastructure).

Restates cv::FASTX / cv2.FastFeatureDetector as the reference calls it
(/root/reference/src/live_stereo.cpp:293,306 TYPE_7_12; src/utils.cpp:30;
src/front_end/features.py:62-67,595-597,621; ORB's internal FAST-9_16 via
features.py:378-387).  The arithmetic lives in OpenCV (un-vendored); semantics
follow SURVEY.md Appendix A.1 and are pinned bit-exactly against cv2 4.13.0 in
tests/test_oracle_pins.py.
"""
import numpy as np

RING = {
    16: [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3),
         (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)],
    12: [(0, 2), (1, 2), (2, 1), (2, 0), (2, -1), (1, -2), (0, -2), (-1, -2), (-2, -1), (-2, 0),
         (-2, 1), (-1, 2)],
    8: [(0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1)],
}


def _ring_diffs(img, ps):
    """d[k] = v - p_k (int16) for the interior [3,H-3) x [3,W-3)."""
    H, W = img.shape
    I = img.astype(np.int16)
    v = I[3:H - 3, 3:W - 3]
    d = np.empty((ps,) + v.shape, np.int16)
    for k, (dx, dy) in enumerate(RING[ps]):
        d[k] = v - I[3 + dy:H - 3 + dy, 3 + dx:W - 3 + dx]
    return d


def fast_score_map(img, threshold, ps=16, nonmax=True):
    """Return (corner_mask, score) both H x W; score is the FAST response (0 off-corner).

    corner: exists a run of K+1 (K = ps/2) contiguous ring pixels all darker than
    v - t or all brighter than v + t, gated by OpenCV's literal-index quick test
    (a no-op for ps=16, an extra filter for 12 and 8).
    score (only meaningful with nonmax): max(t, max_arc min d, max_arc min -d) - 1.
    """
    H, W = img.shape
    corner = np.zeros((H, W), bool)
    score = np.zeros((H, W), np.int32)
    if H < 7 or W < 7:
        return corner, score
    K = ps // 2
    arc = K + 1
    t = int(threshold)
    d = _ring_diffs(img, ps)
    dark = d > t       # class bit 1: ring pixel darker than v - t
    bright = d < -t    # class bit 2
    qd = np.ones(d.shape[1:], bool)
    qb = np.ones(d.shape[1:], bool)
    for a, b in ((0, 8), (2, 10), (4, 12), (6, 14), (1, 9), (3, 11), (5, 13), (7, 15)):
        qd &= dark[a % ps] | dark[b % ps]
        qb &= bright[a % ps] | bright[b % ps]
        # OpenCV ORs the two class bits before AND-ing; since a pixel cannot be both
        # darker and brighter the per-bit accumulation is equivalent.

    def has_arc(flag):
        # linear scan over ring indices 0..ps+K (wrapped): any run of `arc` set flags
        ext = np.concatenate([flag, flag[:K + 1]], axis=0)  # N = ps+K+1 entries
        run = np.zeros(flag.shape[1:], np.int16)
        hit = np.zeros(flag.shape[1:], bool)
        for k in range(ext.shape[0]):
            run = np.where(ext[k], run + 1, 0).astype(np.int16)
            hit |= run >= arc
        return hit

    is_corner = (qd & has_arc(dark)) | (qb & has_arc(bright))
    corner[3:H - 3, 3:W - 3] = is_corner
    if nonmax:
        # sliding min over all ps cyclic arcs of length `arc`
        ext = np.concatenate([d, d[:arc - 1]], axis=0)
        best_pos = np.full(d.shape[1:], -32768, np.int16)   # max_arc min d
        best_neg = np.full(d.shape[1:], -32768, np.int16)   # max_arc min (-d)
        for s in range(ps):
            win = ext[s:s + arc]
            best_pos = np.maximum(best_pos, win.min(axis=0))
            best_neg = np.maximum(best_neg, (-win).min(axis=0))
        sc = np.maximum(np.maximum(best_pos, best_neg), t).astype(np.int32) - 1
        score[3:H - 3, 3:W - 3] = np.where(is_corner, sc, 0)
    return corner, score


def nms3x3(corner, score):
    """Keep a corner iff its score is strictly greater than all 8 neighbours (non-corners = 0)."""
    H, W = score.shape
    p = np.zeros((H + 2, W + 2), score.dtype)
    p[1:-1, 1:-1] = score
    keep = corner.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            keep &= score > p[1 + dy:H + 1 + dy, 1 + dx:W + 1 + dx]
    return keep


def fast_detect(img, threshold, ps=16, nonmax=True):
    """Keypoints in raster order: (xs int32, ys int32, response int32).

    Matches cv2.FastFeatureDetector_create(threshold, nonmax, type).detect(img):
    pt=(x,y), size=7, angle=-1, response=score (0 without nonmax), octave=0, class_id=-1.
    """
    corner, score = fast_score_map(img, threshold, ps, nonmax)
    keep = nms3x3(corner, score) if nonmax else corner
    ys, xs = np.nonzero(keep)  # row-major == raster order
    return xs.astype(np.int32), ys.astype(np.int32), score[ys, xs].astype(np.int32)


def setpoint_step(thresholds, counts, set_point, rows=2, cols=3, lo=4, hi=80, python_variant=False):
    """One update of the per-cell threshold controller.

    C++ (/root/reference/src/live_stereo.cpp:84-102,294-318): target = int(setPoint/(rows*cols));
    thr += sign(n - target) when |n - target| > 0.2*target; clip [4,80].
    Python (/root/reference/src/front_end/features.py:604-608,626-636): bottom row (row==1)
    targets 2x the bucket, other rows 0.5x; clip [6,80].
    """
    thr = np.array(thresholds, dtype=np.int64).reshape(rows, cols).copy()
    cnt = np.asarray(counts).reshape(rows, cols)
    bucket = int(float(set_point) / float(rows * cols))
    for r in range(rows):
        for c in range(cols):
            if python_variant:
                target = 2 * bucket if r == 1 else 0.5 * bucket
                clo = 6
            else:
                target = bucket
                clo = lo
            err = cnt[r, c] - target
            if abs(err) > 0.2 * target:
                thr[r, c] = min(max(thr[r, c] + (1 if err > 0 else -1), clo), hi)
    return thr
