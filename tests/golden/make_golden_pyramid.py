#!/usr/bin/env python
"""Golden vectors for multi-level ORB: cv2.ORB_create(N, 1.2, nlevels, 31, 0, 2, ORB_FAST_SCORE, 31, 15).detectAndCompute
on seeded synthetic images (features.py:378-387 with the nLevels sweep of :292-352; bin/detect_node:50 default 8 levels).
Stored level-major, raster order inside a level (cv2's own order after retainBest is nth_element-dependent).
Run: python tests/golden/make_golden_pyramid.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    d = {}
    for tag, (h, w, n, lv, seed) in {"a": (240, 320, 500, 3, 7), "b": (480, 640, 3000, 4, 8), "c": (360, 480, 2000, 8, 9)}.items():
        img, right = synth.stereo_pair(h, w, seed)
        o = cv2.ORB_create(nfeatures=n, scaleFactor=1.2, nlevels=lv, edgeThreshold=31, firstLevel=0, WTA_K=2,
                           scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=15)
        for eye, im in (("l", img), ("r", right)):
            kps, desc = o.detectAndCompute(im, None)
            x = np.array([k.pt[0] for k in kps], np.float32)
            y = np.array([k.pt[1] for k in kps], np.float32)
            oc = np.array([k.octave for k in kps], np.int32)
            order = np.lexsort((x, y, oc))
            d["%s_%s_img" % (tag, eye)] = im
            d["%s_%s_x" % (tag, eye)], d["%s_%s_y" % (tag, eye)], d["%s_%s_octave" % (tag, eye)] = x[order], y[order], oc[order]
            d["%s_%s_size" % (tag, eye)] = np.array([k.size for k in kps], np.float32)[order]
            d["%s_%s_angle" % (tag, eye)] = np.array([k.angle for k in kps], np.float32)[order]
            d["%s_%s_response" % (tag, eye)] = np.array([k.response for k in kps], np.float32)[order]
            d["%s_%s_desc" % (tag, eye)] = desc[order]
        d[tag + "_params"] = np.array([n, lv], np.int32)
    np.savez_compressed(os.path.join(OUT, "orb_pyramid.npz"), **d)
    print({k: v.shape for k, v in d.items() if k.endswith("_x")})


if __name__ == "__main__":
    main()
