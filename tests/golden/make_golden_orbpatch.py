#!/usr/bin/env python
"""Golden vectors for rBRIEF with ORB::setPatchSize != 31: the live Python node's descriptor (bin/detect_node:50-51,
cv2.ORB_create(); setPatchSize(70)) computed on FAST-7_12 keypoints, and two sizes of the features.py:292-352 sweep.
Run: python tests/golden/make_golden_orbpatch.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    img, _ = synth.stereo_pair(240, 320, 5)
    kps = cv2.FastFeatureDetector_create(15, True, cv2.FAST_FEATURE_DETECTOR_TYPE_7_12).detect(img)
    d = {"img": img, "x": np.array([k.pt[0] for k in kps], np.float32), "y": np.array([k.pt[1] for k in kps], np.float32),
         "response": np.array([k.response for k in kps], np.float32)}
    for ps in (70, 50, 10):
        o = cv2.ORB_create()
        o.setPatchSize(ps)
        k2, desc = o.compute(img, kps)
        d["p%d_x" % ps] = np.array([k.pt[0] for k in k2], np.float32)
        d["p%d_y" % ps] = np.array([k.pt[1] for k in k2], np.float32)
        d["p%d_desc" % ps] = desc
    np.savez_compressed(os.path.join(OUT, "orbpatch_320x240.npz"), **d)
    print({k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
