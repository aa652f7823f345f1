#!/usr/bin/env python
"""Golden vectors for the live nodes' grid detector: the reference's own loops executed through cv2 4.13.

C++ variant  = src/live_stereo.cpp:277-352 (FASTX TYPE_7_12 per cell, controller, cornerSubPix on the cell, offsets);
Python variant = src/front_end/features.py:609-641 (gridDetector.detect).  Three consecutive frames per variant so the
threshold trajectory of the setpoint controller is part of the fixture.  Run: python tests/golden/make_golden_grid.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CRIT = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 40, 0.001)


def clip(v, lo, hi):
    return max(lo, min(hi, v))


def cpp_frame(img, roi, thr, set_point, rows=2, cols=3):
    """live_stereo.cpp:272-352 for one eye."""
    x, y, w, h = roi
    roi_img = img[y:y + h, x:x + w]
    gw, gh = w // cols, h // rows
    grid_set = int(float(set_point) / float(rows * cols))
    pts, resp, counts = [], [], np.zeros((rows, cols), np.int32)
    for r in range(rows):
        for c in range(cols):
            cell = roi_img[r * gh:(r + 1) * gh, c * gw:(c + 1) * gw]
            det = cv2.FastFeatureDetector_create(int(thr[r, c]), True, cv2.FAST_FEATURE_DETECTOR_TYPE_7_12)
            kps = det.detect(np.ascontiguousarray(cell))
            counts[r, c] = len(kps)
            err = len(kps) - grid_set
            if abs(err) > 0.2 * grid_set:
                thr[r, c] = clip(thr[r, c] + (1 if err > 0 else -1), 4, 80)
            for k in kps:
                p = np.array([[k.pt]], np.float32)
                cv2.cornerSubPix(np.ascontiguousarray(cell), p, (5, 5), (-1, -1), CRIT)
                px = np.float32(np.float32(p[0, 0, 0] + np.float32(c * gw)) + np.float32(x))
                py = np.float32(np.float32(p[0, 0, 1] + np.float32(r * gh)) + np.float32(y))
                pts.append((px, py))
                resp.append(k.response)
    return np.array(pts, np.float32).reshape(-1, 2), np.array(resp, np.float32), counts


def py_frame(img, roi, thr, set_point, rows=2, cols=3):
    """features.py:609-641 (gridDetector.detect) for one eye."""
    x, y, w, h = roi
    roi_img = img[y:h + 1, x:w + 1]
    iw, ih = int(w / cols), int(h / rows)
    bucket = int(set_point / float(rows * cols))
    det = cv2.FastFeatureDetector_create()
    det.setType(cv2.FAST_FEATURE_DETECTOR_TYPE_7_12)
    det.setNonmaxSuppression(True)
    allk, counts = [], np.zeros((rows, cols), np.int32)
    for r in range(rows):
        for c in range(cols):
            xo, yo = c * iw, r * ih
            mini = roi_img[yo:yo + ih, xo:xo + iw]
            det.setThreshold(int(thr[r, c]))
            kps = det.detect(np.ascontiguousarray(mini))
            counts[r, c] = len(kps)
            for d in kps:
                allk.append((d.pt[0] + xo + x, d.pt[1] + yo + y, d.response))
            if r == 1:
                err = len(kps) - 2 * bucket
                hyst = 0.2 * 2 * bucket
            else:
                err = len(kps) - 0.5 * bucket
                hyst = 0.2 * 0.5 * bucket
            if abs(err) > hyst:
                thr[r, c] = np.clip(thr[r, c] + (1 if err > 0 else -1), 6, 80)
    pts = []
    for (px, py, _) in allk:
        ref = np.float32([[px, py]])
        cv2.cornerSubPix(img, ref, (5, 5), (-1, -1), CRIT)
        pts.append((ref[0, 0], ref[0, 1]))
    return (np.array(pts, np.float32).reshape(-1, 2), np.array([k[2] for k in allk], np.float32), counts)


def main():
    h, w = 360, 480
    frames = synth.stereo_sequence(h, w, 21, 3)
    d = {}
    for variant, fn, roi, sp, t0 in (("cpp", cpp_frame, (16, 8, 450, 340), 1500, 15), ("py", py_frame, (0, 0, 470, 350), 1500, 10)):
        for eye in (0, 1):
            thr = np.full((2, 3), t0, np.int32)
            for f, pair in enumerate(frames):
                d["%s_e%d_f%d_thr_in" % (variant, eye, f)] = thr.copy()
                pts, resp, counts = fn(pair[eye], roi, thr, sp)
                d["%s_e%d_f%d_pts" % (variant, eye, f)] = pts
                d["%s_e%d_f%d_resp" % (variant, eye, f)] = resp
                d["%s_e%d_f%d_counts" % (variant, eye, f)] = counts
                d["%s_e%d_f%d_thr_out" % (variant, eye, f)] = thr.copy()
        d[variant + "_roi"] = np.array(roi, np.int32)
        d[variant + "_set_point"] = np.int32(sp)
    for f, pair in enumerate(frames):
        d["img_f%d_l" % f], d["img_f%d_r" % f] = pair[0], pair[1]
    # plain cornerSubPix on arbitrary float points, incl. points near every border
    rng = np.random.default_rng(3)
    img = frames[0][0]
    p = np.stack([rng.uniform(0, w - 1e-3, 400), rng.uniform(0, h - 1e-3, 400)], 1).astype(np.float32)
    p[:40, 0] = rng.uniform(0, 4, 40)
    p[40:80, 1] = rng.uniform(0, 4, 40)
    p[80:120, 0] = rng.uniform(w - 5, w - 0.01, 40)
    p[120:160, 1] = rng.uniform(h - 5, h - 0.01, 40)
    ref = p.copy().reshape(-1, 1, 2)
    cv2.cornerSubPix(img, ref, (5, 5), (-1, -1), CRIT)
    d["subpix_in"], d["subpix_out"] = p, ref.reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "grid_subpix_480x360.npz"), **d)
    print("grid_subpix_480x360.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if "f0" in k and "e0" in k})


if __name__ == "__main__":
    main()
