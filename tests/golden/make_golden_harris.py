#!/usr/bin/env python
"""Golden vectors for cv::ORB with scoreType = HARRIS_SCORE (the `score` field of front_end/setDetector,
src/StereoCamera.cpp:445,462; src/utils.cpp:86-90; cv2.ORB_create()'s own default as bin/detect_node:50 constructs it):
cv2.ORB_create(N, 1.2, nlevels, 31, 0, 2, ORB_HARRIS_SCORE, 31, fastThreshold).detectAndCompute on seeded synthetic pairs
(regenerated from oracle.synth by the tests: only h, w, seed are stored).  Stored level-major, raster order inside a level.
Run: python tests/golden/make_golden_harris.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {"a": (240, 320, 300, 1, 15, 11), "b": (480, 640, 1500, 4, 15, 12), "c": (360, 480, 500, 8, 20, 13)}


def main():
    d = {}
    for tag, (h, w, n, lv, thr, seed) in CASES.items():
        pair = synth.stereo_pair(h, w, seed)
        if tag == "c":
            o = cv2.ORB_create()             # 500 features, 8 levels, HARRIS_SCORE, fastThreshold 20
            assert o.getScoreType() == cv2.ORB_HARRIS_SCORE and o.getNLevels() == 8 and o.getFastThreshold() == 20
        else:
            o = cv2.ORB_create(nfeatures=n, scaleFactor=1.2, nlevels=lv, edgeThreshold=31, firstLevel=0, WTA_K=2,
                               scoreType=cv2.ORB_HARRIS_SCORE, patchSize=31, fastThreshold=thr)
        for eye, im in zip("lr", pair):
            kps, desc = o.detectAndCompute(im, None)
            x = np.array([k.pt[0] for k in kps], np.float32)
            y = np.array([k.pt[1] for k in kps], np.float32)
            oc = np.array([k.octave for k in kps], np.int32)
            order = np.lexsort((x, y, oc))
            p = "%s_%s_" % (tag, eye)
            d[p + "x"], d[p + "y"], d[p + "octave"] = x[order], y[order], oc[order]
            d[p + "size"] = np.array([k.size for k in kps], np.float32)[order]
            d[p + "angle"] = np.array([k.angle for k in kps], np.float32)[order]
            d[p + "response"] = np.array([k.response for k in kps], np.float32)[order]
            d[p + "desc"] = desc[order]
        d[tag + "_params"] = np.array([h, w, n, lv, thr, seed], np.int32)
    np.savez_compressed(os.path.join(OUT, "orb_harris.npz"), **d)
    print({k: v.shape for k, v in d.items() if k.endswith("_x")})


if __name__ == "__main__":
    main()
