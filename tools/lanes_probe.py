#!/usr/bin/env python
"""Would two concurrent compute lanes (half batches on two streams) beat one stream for a RESIDENT batch?  A/B without touching
the library: one fe_ctx with 96 pairs vs two fe_ctx with 48 pairs each, their asynchronous fe_batch_run calls enqueued
alternately (lane B one call behind lane A, so different stages overlap), wall clock over the same number of pairs."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import front_end_b200 as fe
from front_end_b200 import synth

h, w, N, P = 720, 1280, 5000, 96
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
uniq = 24
Lu, Ru = synth.stereo_batch(h, w, uniq, seed0=0, n_scenes=4)
Ls, Rs = np.concatenate([Lu] * (P // uniq)), np.concatenate([Ru] * (P // uniq))
ca = fe.match_cfg(mask=fe.MASK_EPIPOLAR, epi_threshold=2.0)
cb = fe.match_cfg(mode=fe.MATCH_CROSSCHECK, mask=fe.MASK_NONE, max_dy=0.7)


def ctx(pairs):
    return fe.FrontEnd(max_width=w, max_height=h, max_pairs=pairs, max_keypoints=8192, n_features=N, fast_threshold=15)


def timed(ctxs, n):
    for f in ctxs:
        f.batch_run(ca, cb, sync=True)
    t0 = time.perf_counter()
    for _ in range(n):
        for f in ctxs:
            f.batch_run(ca, cb, sync=False)
    for f in ctxs:
        f.batch_run(ca, cb, sync=True)          # one more each: drains the lane
    return (time.perf_counter() - t0) / (n + 1) * 1e3


one = ctx(P)
one.batch_upload(Ls, Rs)
for _ in range(3):
    one.batch_run(ca, cb, sync=True)
t1 = timed([one], steps)
nb1 = int(one.batch_download(want=("b",))["n_b"].sum())
del one
for lanes in (2, 3, 4):
    fs = [ctx(P // lanes) for _ in range(lanes)]
    for i, f in enumerate(fs):
        f.batch_upload(Ls[i * (P // lanes):(i + 1) * (P // lanes)], Rs[i * (P // lanes):(i + 1) * (P // lanes)])
        for _ in range(3):
            f.batch_run(ca, cb, sync=True)
    t = timed(fs, steps)
    nb = sum(int(f.batch_download(want=("b",))["n_b"].sum()) for f in fs)
    print("lanes=%d: %.3f ms per 96 pairs (one stream: %.3f ms)  matches %d / %d" % (lanes, t, t1, nb, nb1))
    del fs
