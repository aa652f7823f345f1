"""The C++ host side above the C-ABI (include/fe_host.hpp): StereoCamera (src/StereoCamera.cpp:5-381), WindowMatcher
(src/WindowMatcher.cpp:75-231) and the live node's stereoMatch loop (src/live_stereo.cpp:227-404), compiled with g++
against libfe_b200.so only and driven by tests/host_cpp/host_check.cpp.  What the classes publish is compared with the
oracle's restatement of the same reference code on the same frames."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import fast as ofast  # noqa: F401  (oracle = checker only)
from oracle import match as omatch
from oracle import orb as oorb
from oracle import subpix as osub
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tests", "host_cpp")
BIN = os.path.join(HOST, "host_check")


def _build():
    subprocess.check_call(["make", "-C", HOST, "-s"])
    assert os.path.exists(BIN)


def _write_input(path, Ls, Rs, n_features, thr, roi, set_point, live_thr, Q):
    F, H, W = Ls.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<11i", W, H, F, n_features, thr, *roi, set_point, live_thr))
        for l, r in zip(Ls, Rs):
            f.write(np.ascontiguousarray(l, np.uint8).tobytes())
            f.write(np.ascontiguousarray(r, np.uint8).tobytes())
        f.write(np.ascontiguousarray(Q, np.float64).tobytes())


def test_host_cpp_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """The host classes compile against the C-ABI header alone; with no CUDA device fe_create reports
    FE_ERR_NO_DEVICE and the program stops (exit 3) -- there is no CPU path to fall back to."""
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    Ls = np.zeros((1, 64, 64), np.uint8)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_input(inp, Ls, Ls, 100, 15, (0, 0, 64, 64), 100, 15, np.eye(4))
    p = subprocess.run([BIN, inp, out], capture_output=True, text=True, timeout=120)
    assert p.returncode == 3, (p.returncode, p.stderr)
    assert "fe_create" in p.stderr


class _Reader:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        self.o = 0

    def take(self, dtype, n):
        dt = np.dtype(dtype)
        a = np.frombuffer(self.b, dt, n, self.o)
        self.o += dt.itemsize * n
        return a

    def done(self):
        return self.o == len(self.b)


@pytest.mark.gpu
def test_host_cpp_classes_vs_oracle(fe, tmp_path):
    _build()
    h, w, F, N, thr = 240, 320, 5, 400, 15
    roi, set_point, live_thr = (16, 8, 288, 224), 900, 20
    frames = synth.stereo_sequence(h, w, 23, F)
    Ls, Rs = np.stack([p[0] for p in frames]), np.stack([p[1] for p in frames])
    Q = np.array([[1, 0, 0, -160.5], [0, 1, 0, -120.25], [0, 0, 0, 420.0], [0, 0, 1.0 / 0.12, 0]], np.float64)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_input(inp, Ls, Rs, N, thr, roi, set_point, live_thr, Q)
    p = subprocess.run([BIN, inp, out], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    rd = _Reader(out)

    # StereoCamera: ORB detect + compute per eye, mask |2 dy| <= 2, kNN-2, Lowe 0.8, StereoFrame packing
    lm = []
    for f in range(F):
        n = int(rd.take(np.int32, 1)[0])
        v = rd.take(np.float32, 5 * n).reshape(n, 5)
        ld = rd.take(np.uint8, 32 * n).reshape(n, 32)
        rdesc = rd.take(np.uint8, 32 * n).reshape(n, 32)
        ref = [oorb.orb_detect_and_compute(im, N, thr) for im in (Ls[f], Rs[f])]
        q, t, d = omatch.stereo_match_ratio(ref[0]["y"], ref[1]["y"], ref[0]["desc"], ref[1]["desc"], 1.0, 0.8)
        assert n == len(q) and n > 100
        assert np.array_equal(v[:, 0], ref[0]["x"][q].astype(np.float32)) and np.array_equal(v[:, 1], ref[0]["y"][q].astype(np.float32))
        assert np.array_equal(v[:, 2], ref[1]["x"][t].astype(np.float32)) and np.array_equal(v[:, 3], ref[1]["y"][t].astype(np.float32))
        assert np.array_equal(v[:, 4], d.astype(np.float32))
        assert np.array_equal(ld, ref[0]["desc"][q]) and np.array_equal(rdesc, ref[1]["desc"][t])
        lm.append((v[:, :2].copy(), ld.copy()))

    # WindowMatcher: consecutive-frame box mask + kNN-2 + Lowe on the left features of the published frames
    for f in range(1, F):
        n = int(rd.take(np.int32, 1)[0])
        cur, prev = rd.take(np.int32, n), rd.take(np.int32, n)
        q, t, _ = omatch.window_match(lm[f][0], lm[f - 1][0], lm[f][1], lm[f - 1][1])
        assert np.array_equal(cur, q) and np.array_equal(prev, t) and n > 50

    # LiveDetector: 2x3 grid FAST-7_12 + controller (exact vs the oracle) ; refined points, descriptors and the
    # cross-check stage against the Python host mirror on the same library
    lthr = np.full((2, 3), live_thr)
    rthr = lthr.copy()
    with fe.FrontEnd(max_width=w, max_height=h, max_pairs=3, max_keypoints=8192, n_features=N, fast_threshold=thr) as g:
        for f in range(F):
            got_thr = rd.take(np.int32, 6).reshape(2, 3)
            nl, nr, ng = (int(x) for x in rd.take(np.int32, 3))
            lk, rk = rd.take(fe.lib.KPOINT, nl), rd.take(fe.lib.KPOINT, nr)
            mm = rd.take(fe.lib.MATCH, ng)
            _, _, counts, want_thr = osub.grid_detect(Ls[f], roi, lthr, set_point, subpix=False)
            pk_l, c_l, lthr2 = g.grid_detect(Ls[f], lthr, set_point, roi=roi)
            pk_r, _, rthr = g.grid_detect(Rs[f], rthr, set_point, roi=roi)
            assert np.array_equal(c_l, counts) and np.array_equal(lthr2, want_thr) and np.array_equal(got_thr, want_thr)
            lthr = lthr2
            kl, dl = g.compute(Ls[f], pk_l)
            kr, dr = g.compute(Rs[f], pk_r)
            assert np.array_equal(lk, kl) and np.array_equal(rk, kr) and nl > 100
            want = g.stereo_match(kl, dl, kr, dr, fe.match_cfg(mode=fe.MATCH_CROSSCHECK, mask=fe.MASK_NONE))
            assert np.array_equal(mm, want)
            q, t, _ = omatch.stereo_match_crosscheck(kl["y"], kr["y"], dl, dr, 0.7)
            assert np.array_equal(mm["queryIdx"], q) and np.array_equal(mm["trainIdx"], t)
    assert rd.done()
