#!/usr/bin/env python
"""Per-stage device times of one resident c2 / c3 batch (CUDA events inside the library): the quick A/B harness for kernel work.

usage: python tools/stage_time.py [c2|c3|c5] [steps]     (env knobs such as FE_FAST_FMA are read by the library)
Prints one line: total ms per step, then `stage=ms` for every stage.  24 distinct pairs are generated and tiled to the
batch size (the images still occupy distinct HBM addresses, 177 MB per step: larger than L2)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import front_end_b200 as fe
from front_end_b200 import synth

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
h, w, N, P = {"c2": (720, 1280, 5000, 96), "c3": (720, 1280, 5000, 96), "c5": (1200, 1920, 10000, 48)}[wl]
uniq = min(P, 24)
Lu, Ru = synth.stereo_batch(h, w, uniq, seed0=0, n_scenes=4)
Ls, Rs = np.concatenate([Lu] * (P // uniq)), np.concatenate([Ru] * (P // uniq))
surf = wl == "c3"
f = fe.FrontEnd(max_width=w, max_height=h, max_pairs=P, max_keypoints=8192 if N <= 5000 else 16384, n_features=N, fast_threshold=15,
                orientation=not surf, surf_upright=True)
if surf:
    f.set_batch_descriptor(fe.DESC_SURF128)
norm = fe.NORM_L2 if surf else fe.NORM_HAMMING
ca = fe.match_cfg(mask=fe.MASK_EPIPOLAR, epi_threshold=2.0, norm=norm)
cb = fe.match_cfg(mode=fe.MATCH_CROSSCHECK, mask=fe.MASK_NONE, norm=norm, max_dy=0.7)
f.batch_upload(Ls, Rs)
for _ in range(3):
    f.batch_run(ca, cb, sync=True)
f.profile(True)
f.profile_reset()
for _ in range(steps):
    f.batch_run(ca, cb, sync=True)
st = f.stage_times()
tot = sum(v[0] for v in st.values()) / steps
out = f.batch_download(want=("a", "b"))
print("%s total=%.3f  " % (wl, tot) + "  ".join("%s=%.3f" % (k, v[0] / steps) for k, v in sorted(st.items(), key=lambda kv: -kv[1][0]) if v[1])
      + "  | kps=%d a=%d b=%d" % (int(out["n_kps"].sum()), int(out["n_a"].sum()), int(out["n_b"].sum())))
