"""CPU oracle for the stereo feature front-end hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in numpy, the arithmetic that the reference
(RyanEvanWolf/front_end) delegates to OpenCV on its hot path: FAST + NMS,
ORB top-N / IC-angle / rBRIEF-256, SURF descriptors, brute-force Hamming / L2
matching with epipolar-band, search-window, Lowe-ratio and cross-check glue.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker.  The product
(``front_end_b200``) never imports this package and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors, and
its arithmetic lives in an un-vendored, un-pinned OpenCV.  The oracle is therefore
pinned against cv2 4.13.0 as installed in this image (``tests/test_oracle_pins.py``
runs the reference's own call sequence through cv2 and demands bit-equality) and
against fixtures under ``tests/golden/`` produced by ``tests/golden/make_golden.py``.
SURF has no binary available here: its restatement of ``src/surf.cpp`` is pinned
only primitive-by-primitive -- "parity unpinned" for the SURF descriptor as a whole.
"""
