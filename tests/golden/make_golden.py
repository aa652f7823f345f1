#!/usr/bin/env python
"""Generate golden vectors by running the reference's own call sequence through cv2 (4.13.0 here).

The reference (RyanEvanWolf/front_end) holds no tests or fixtures; its hot
path is a sequence of OpenCV calls.  This script executes exactly those calls on seeded synthetic
inputs and stores the outputs, so that the oracle and the CUDA path can be checked on machines
without cv2.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FAST_TYPES = {16: cv2.FAST_FEATURE_DETECTOR_TYPE_9_16, 12: cv2.FAST_FEATURE_DETECTOR_TYPE_7_12,
              8: cv2.FAST_FEATURE_DETECTOR_TYPE_5_8}


def kp_arrays(kps):
    x = np.array([k.pt[0] for k in kps], np.float32)
    y = np.array([k.pt[1] for k in kps], np.float32)
    r = np.array([k.response for k in kps], np.float32)
    a = np.array([k.angle for k in kps], np.float32)
    return x, y, r, a


def orb(img, n, thr):
    o = cv2.ORB_create(nfeatures=n, scaleFactor=1.2, nlevels=1, edgeThreshold=31, firstLevel=0,
                       WTA_K=2, scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=thr)
    kps, desc = o.detectAndCompute(img, None)
    x, y, r, a = kp_arrays(kps)
    order = np.lexsort((x, y))  # canonical raster order
    return x[order], y[order], r[order], a[order], desc[order]


def knn_arrays(res):
    idx = np.full((len(res), 2), -1, np.int32)
    dd = np.full((len(res), 2), np.inf, np.float32)
    for i, r in enumerate(res):
        for j, m in enumerate(r[:2]):
            idx[i, j] = m.trainIdx
            dd[i, j] = m.distance
    return idx, dd


def match_arrays(ms):
    return (np.array([m.queryIdx for m in ms], np.int32), np.array([m.trainIdx for m in ms], np.int32),
            np.array([m.distance for m in ms], np.float32))


def main():
    # --- FAST on a small image, all three ring sizes, with/without NMS -------------------------
    img, _ = synth.stereo_pair(120, 160, 5)
    d = {"img": img}
    for ps, typ in FAST_TYPES.items():
        for thr in (15, 40):
            for nms in (1, 0):
                kps = cv2.FastFeatureDetector_create(thr, bool(nms), typ).detect(img)
                x, y, r, _ = kp_arrays(kps)
                d["x_%d_%d_%d" % (ps, thr, nms)] = x.astype(np.int16)
                d["y_%d_%d_%d" % (ps, thr, nms)] = y.astype(np.int16)
                d["r_%d_%d_%d" % (ps, thr, nms)] = r.astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "fast_160x120.npz"), **d)

    # --- config C1: 640x480 pair, ORB-5000, both matching paths ---------------------------------
    for name, (h, w, seed, n, store_img) in {"c1_640x480": (480, 640, 1, 5000, True),
                                             "c2_1280x720": (720, 1280, 0, 5000, False),
                                             "small_320x240": (240, 320, 3, 500, True)}.items():
        L, R = synth.stereo_pair(h, w, seed)
        d = {"h": h, "w": w, "seed": seed, "n_features": n, "fast_threshold": 15}
        if store_img:
            d["L"], d["R"] = L, R
        else:
            d["L_sum"], d["R_sum"] = np.int64(L.astype(np.int64).sum()), np.int64(R.astype(np.int64).sum())
        feats = {}
        for eye, im in (("l", L), ("r", R)):
            x, y, r, a, desc = orb(im, n, 15)
            feats[eye] = (x, y, desc)
            d[eye + "x"], d[eye + "y"] = x.astype(np.int16), y.astype(np.int16)
            d[eye + "resp"], d[eye + "angle"], d[eye + "desc"] = r.astype(np.int16), a, desc
        (lx, ly, ld), (rx, ry, rd) = feats["l"], feats["r"]
        for thr in (1.0, 2.0):
            mask = (np.abs(ly[:, None] - ry[None, :]) <= np.float32(thr)).astype(np.uint8)
            res = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(ld, rd, 2, mask)
            idx, dd = knn_arrays(res)
            d["knn_idx_%d" % int(thr)], d["knn_dist_%d" % int(thr)] = idx, dd
        q, t, dist = match_arrays(cv2.BFMatcher(cv2.NORM_HAMMING, True).match(ld, rd))
        d["cc_q"], d["cc_t"], d["cc_d"] = q, t, dist
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, len(lx), len(rx), "cc", len(q))

    # --- window matching: two consecutive frames of a sequence, box mask 100x100 ---------------
    seq = synth.stereo_sequence(240, 320, 11, 2)
    d = {}
    fr = []
    for f, (L, R) in enumerate(seq):
        x, y, r, a, desc = orb(L, 500, 15)
        fr.append((x, y, desc))
        d["x%d" % f], d["y%d" % f], d["desc%d" % f] = x, y, desc
    (px, py, pd), (cx, cy, cd) = fr
    mask = ((np.abs(cx[:, None] - px[None, :]) < 50) & (np.abs(cy[:, None] - py[None, :]) < 50)).astype(np.uint8)
    idx, dd = knn_arrays(cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(cd, pd, 2, mask))
    d["knn_idx"], d["knn_dist"] = idx, dd
    np.savez_compressed(os.path.join(OUT, "window_320x240.npz"), **d)

    # --- L2 matcher on random 128-d unit vectors (SURF_EXTENDED-shaped) --------------------------
    rng = np.random.default_rng(7)
    a = rng.random((300, 128), dtype=np.float32)
    b = np.vstack([a[:200] + 0.05 * rng.standard_normal((200, 128)).astype(np.float32),
                   rng.random((150, 128), dtype=np.float32)]).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    idx, dd = knn_arrays(cv2.BFMatcher(cv2.NORM_L2, False).knnMatch(a, b, 2))
    q, t, dist = match_arrays(cv2.BFMatcher(cv2.NORM_L2, True).match(a, b))
    np.savez_compressed(os.path.join(OUT, "l2_300x350.npz"), a=a, b=b, knn_idx=idx, knn_dist=dd,
                        cc_q=q, cc_t=t, cc_d=dist)


if __name__ == "__main__":
    main()
