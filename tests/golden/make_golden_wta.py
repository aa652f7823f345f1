#!/usr/bin/env python
"""Golden vectors for ORB with WTA_K = 3 / 4 and NORM_HAMMING2 matching (features.py:378-387 sweeps wta 2 / 3 / 4;
src/StereoCamera.cpp:504-511 picks NORM_HAMMING2 when WTA_K > 2): cv2.ORB_create(..., WTA_K=k).detectAndCompute on a
seeded stereo pair, then BFMatcher(NORM_HAMMING2).knnMatch with the epipolar band mask and the cross-check match.
Run: python tests/golden/make_golden_wta.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    L, R = synth.stereo_pair(240, 320, 13)
    d = {"l_img": L, "r_img": R}
    for k in (3, 4):
        o = cv2.ORB_create(nfeatures=600, scaleFactor=1.2, nlevels=1, edgeThreshold=31, firstLevel=0, WTA_K=k,
                           scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=15)
        feats = {}
        for eye, im in (("l", L), ("r", R)):
            kps, desc = o.detectAndCompute(im, None)
            x = np.array([p.pt[0] for p in kps], np.float32)
            y = np.array([p.pt[1] for p in kps], np.float32)
            order = np.lexsort((x, y))
            feats[eye] = (x[order], y[order], desc[order])
            d["k%d_%s_x" % (k, eye)], d["k%d_%s_y" % (k, eye)], d["k%d_%s_desc" % (k, eye)] = feats[eye]
            d["k%d_%s_angle" % (k, eye)] = np.array([p.angle for p in kps], np.float32)[order]
        (lx, ly, ld), (rx, ry, rd) = feats["l"], feats["r"]
        mask = (np.abs(ly[:, None] - ry[None, :]) <= np.float32(2.0)).astype(np.uint8)
        knn = cv2.BFMatcher(cv2.NORM_HAMMING2, False).knnMatch(ld, rd, 2, mask)
        idx = np.full((len(knn), 2), -1, np.int32)
        dist = np.full((len(knn), 2), np.inf, np.float32)
        for i, row in enumerate(knn):
            for j, m in enumerate(row[:2]):
                idx[i, j], dist[i, j] = m.trainIdx, m.distance
        cc = cv2.BFMatcher(cv2.NORM_HAMMING2, True).match(ld, rd)
        d["k%d_knn_idx" % k], d["k%d_knn_dist" % k] = idx, dist
        d["k%d_cc_q" % k] = np.array([m.queryIdx for m in cc], np.int32)
        d["k%d_cc_t" % k] = np.array([m.trainIdx for m in cc], np.int32)
        d["k%d_cc_d" % k] = np.array([m.distance for m in cc], np.float32)
    np.savez_compressed(os.path.join(OUT, "orb_wta_320x240.npz"), **d)
    print({k: v.shape for k, v in d.items() if "desc" in k or "cc_q" in k})


if __name__ == "__main__":
    main()
