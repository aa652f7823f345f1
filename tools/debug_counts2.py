import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import front_end_b200 as fe
from oracle import synth, orb as oorb
L, R = synth.stereo_pair(240, 320, 3)
print("flags", L.flags.c_contiguous, L.strides, L.ctypes.data % 256, R.ctypes.data % 256)
ref = oorb.orb_detect_and_compute(L, 500, 15)
ca, cb = fe.match_cfg(), fe.match_cfg(mode=fe.MATCH_CROSSCHECK, mask=fe.MASK_NONE)
def show(tag, out):
    n = out["n_kps"][0]; k = out["kps"][0][:n]
    print(tag, out["n_kps"].tolist())
    if n != len(ref["x"]):
        have = set(zip(k["x"].astype(int).tolist(), k["y"].astype(int).tolist()))
        want = set(zip(ref["x"].tolist(), ref["y"].tolist()))
        extra = sorted(have - want, key=lambda p: (p[1], p[0]))
        print("  extra", len(extra), "missing", len(want - have))
        sc = {(int(a), int(b)): float(c) for a, b, c in zip(k["x"], k["y"], k["response"])}
        print("  extra (x,y,score):", [(p[0], p[1], sc[p]) for p in extra][:40])
        print("  min ref score", ref["response"].min())
for trial in range(3):
    with fe.FrontEnd(device=0, max_width=320, max_height=240, max_pairs=1, max_keypoints=2048, n_features=500) as f:
        out = f.pipeline_batch(L[None], R[None], ca, cb)
        show("L[None] inside with, trial %d" % trial, out)
    show("after with", out)
with fe.FrontEnd(device=0, max_width=320, max_height=240, max_pairs=1, max_keypoints=2048, n_features=500) as f:
    out = f.pipeline_batch(L[None].copy(), R[None].copy(), ca, cb)
    show("copies", out)
import __graft_entry__ as g
try:
    g.smoke()
except AssertionError as e:
    print("smoke failed", e)
