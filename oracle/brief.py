"""Oracle: cv::BriefDescriptorExtractor (BRIEF-16 / 32 / 64) with a supplied test table (test infrastructure).

Restates OpenCV's features2d/src/brief.cpp (2.4, the reference's version; identical in opencv_contrib's xfeatures2d)
as the reference calls it:
  /root/reference src/live_stereo.cpp:238,359-360   cv::BriefDescriptorExtractor extractor(16); extractor.compute(img, kps, desc)
  src/front_end/features.py:93-96                    cv2.xfeatures2d.BriefDescriptorExtractor_create(bytes, use_orientation)
  bin/detect_node:28-29                              BriefDescriptorExtractor_create(16 / 64, False)

  * PATCH_SIZE = 48, KERNEL_SIZE = 9; KeyPointsFilter::runByImageBorder(kps, size, 48/2 + 9/2 = 28);
  * sum = integral(img) (CV_32S); smoothedSum(y, x) = box sum of the 9 x 9 window centred on
    ((int)(pt.y + 0.5) + y, (int)(pt.x + 0.5) + x);
  * pixelTests{16,32,64}: desc[b] = sum_j (smoothedSum(y1, x1) < smoothedSum(y2, x2)) << (7 - j) over the byte's 8 tests
    (the first test of a byte is its most significant bit);
  * use_orientation (contrib only): offsets rotated by kp.angle, truncated to int and clamped to [-24, 24].

PARITY UNPINNED: the test tables (generated_16.i / _32.i / _64.i) are part of OpenCV's sources, which are neither in the
reference repository nor in this image (cv2 4.13 here is built without xfeatures2d), and no BRIEF binary exists to run.
The arithmetic is integer-exact, so any table the caller supplies gives identical results on both sides; the tests use
a seeded random table."""
import numpy as np

from .surf import integral_i32

BORDER = 28


def random_tests(n_bytes, seed=0):
    """A seeded stand-in for OpenCV's table: (n_bytes * 8, 4) int8 rows (y1, x1, y2, x2), Gaussian offsets clipped to +-24."""
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.normal(0.0, 48 / 5.0, size=(n_bytes * 8, 4))), -24, 24).astype(np.int8)


def keep_mask(xs, ys, w, h):
    xs, ys = np.asarray(xs, np.float32), np.asarray(ys, np.float32)
    return (xs >= BORDER) & (xs < w - BORDER) & (ys >= BORDER) & (ys < h - BORDER)


def brief_compute(img, xs, ys, tests, angles=None, use_orientation=False):
    """Returns (keep bool[N], desc u8[n_kept x bytes])."""
    h, w = img.shape
    S = integral_i32(img).astype(np.int64)
    keep = keep_mask(xs, ys, w, h)
    tests = np.asarray(tests, np.int64)
    n_bytes = len(tests) // 8
    out = []
    for k in np.nonzero(keep)[0]:
        cy, cx = int(np.float64(np.float32(ys[k])) + 0.5), int(np.float64(np.float32(xs[k])) + 0.5)
        t = tests.copy()
        if use_orientation:
            a = np.float32(np.float32(angles[k]) * np.float32(np.pi / 180.0))
            r0, r1 = np.float32(np.sin(np.float64(a))), np.float32(np.cos(np.float64(a)))
            for cy_, cx_ in ((0, 1), (2, 3)):
                y, x = tests[:, cy_].astype(np.float32), tests[:, cx_].astype(np.float32)
                rx = np.trunc((x * r1).astype(np.float32) - (y * r0).astype(np.float32)).astype(np.int64)
                ry = np.trunc((x * r0).astype(np.float32) + (y * r1).astype(np.float32)).astype(np.int64)
                t[:, cx_], t[:, cy_] = np.clip(rx, -24, 24), np.clip(ry, -24, 24)

        def smoothed(y, x):
            # a keypoint whose coordinate rounds UP onto the border (y in [h - 28.5, h - 28)) with an offset of +24 makes
            # OpenCV read one row / column past the integral image (undefined there); both sides clamp that index
            Y, X = cy + y, cx + x
            Y5, X5 = np.minimum(Y + 5, h), np.minimum(X + 5, w)
            return S[Y5, X5] - S[Y5, X - 4] - S[Y - 4, X5] + S[Y - 4, X - 4]

        bits = (smoothed(t[:, 0], t[:, 1]) < smoothed(t[:, 2], t[:, 3])).astype(np.uint8).reshape(n_bytes, 8)
        out.append(np.packbits(bits, axis=1, bitorder="big").reshape(n_bytes))
    d = np.stack(out) if out else np.zeros((0, n_bytes), np.uint8)
    return keep, d
