import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import front_end_b200 as fe
from oracle import synth, orb as oorb
L, R = synth.stereo_pair(240, 320, 3)
exp = [len(oorb.orb_detect_and_compute(im, 500, 15)["x"]) for im in (L, R)]
print("expected", exp)
ca, cb = fe.match_cfg(), fe.match_cfg(mode=fe.MATCH_CROSSCHECK, mask=fe.MASK_NONE)
def run(tag, P, cap, a, b, prof=False, first_detect=False):
    with fe.FrontEnd(max_width=320, max_height=240, max_pairs=max(P,1), max_keypoints=cap, n_features=500) as f:
        if prof: f.profile(True)
        if first_detect: print(tag, "detect first", len(f.detect(L)))
        Ls = np.repeat(L[None], P, 0); Rs = np.repeat(R[None], P, 0)
        out = f.pipeline_batch(Ls, Rs, a, b)
        print(tag, "P", P, "cap", cap, out["n_kps"].tolist(), out["n_a"].tolist(), out["n_b"].tolist())
        out = f.pipeline_batch(Ls, Rs, a, b)
        print(tag, "again", out["n_kps"].tolist())
        lk, ld, rk, rd, _ = f.stereo_features(L, R)
        print(tag, "stereo_features", len(lk), len(rk))
run("both", 1, 2048, ca, cb)
run("both-prof", 1, 2048, ca, cb, prof=True)
run("a-only", 1, 2048, ca, None)
run("b-only", 1, 2048, None, cb)
run("none", 1, 2048, None, None)
run("both", 2, 2048, ca, cb)
run("both", 9, 2048, ca, cb)
run("both-cap16k", 1, 16384, ca, cb)
run("both-detect-first", 1, 2048, ca, cb, first_detect=True)
