#!/usr/bin/env python
"""Golden vectors for the reference's ORB parameter table (src/front_end/features.py:292-352: edgeThreshold 5..45,
patchSize 10/30/50, nLevels 2/4, wta 3/4, scaleFactor 1.1..2.0): cv2.ORB_create(N, scale, nlevels, edge, 0, wta,
ORB_FAST_SCORE, patch, 15).detectAndCompute on a seeded image.  Stored level-major, raster order inside a level.
Run: python tests/golden/make_golden_orbparams.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {  # tag: (nfeatures, scale, nlevels, edge, wta, patch)
    "e5p31": (400, 1.2, 1, 5, 2, 31), "e15p30": (400, 1.2, 1, 15, 2, 30), "e5p10": (400, 1.2, 1, 5, 2, 10),
    "e25p50": (400, 1.2, 1, 25, 2, 50), "e15p30l2s14w3": (500, 1.4, 2, 15, 3, 30), "e35p10l4s11w4": (600, 1.1, 4, 35, 4, 10),
}


def main():
    img, _ = synth.stereo_pair(200, 280, 17)
    d = {"img": img}
    for tag, (n, sc, lv, edge, wta, patch) in CASES.items():
        o = cv2.ORB_create(nfeatures=n, scaleFactor=sc, nlevels=lv, edgeThreshold=edge, firstLevel=0, WTA_K=wta,
                           scoreType=cv2.ORB_FAST_SCORE, patchSize=patch, fastThreshold=15)
        kps, desc = o.detectAndCompute(img, None)
        x = np.array([k.pt[0] for k in kps], np.float32)
        y = np.array([k.pt[1] for k in kps], np.float32)
        oc = np.array([k.octave for k in kps], np.int32)
        order = np.lexsort((x, y, oc))
        d[tag + "_x"], d[tag + "_y"], d[tag + "_octave"] = x[order], y[order], oc[order]
        d[tag + "_size"] = np.array([k.size for k in kps], np.float32)[order]
        d[tag + "_angle"] = np.array([k.angle for k in kps], np.float32)[order]
        d[tag + "_response"] = np.array([k.response for k in kps], np.float32)[order]
        d[tag + "_desc"] = desc[order]
        d[tag + "_params"] = np.array([n, lv, edge, wta, patch], np.int32)
        d[tag + "_scale"] = np.float32(sc)
    np.savez_compressed(os.path.join(OUT, "orb_params.npz"), **d)
    print({k: v.shape for k, v in d.items() if k.endswith("_x")})


if __name__ == "__main__":
    main()
