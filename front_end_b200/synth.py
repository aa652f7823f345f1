"""Deterministic synthetic rectified stereo generator (numpy only; workload generator of bench.py and the tests).

Recipe from SURVEY.md section 8(d): a scene canvas (smooth random field blended with
white noise, plus small random rectangles for corner structure) from which views
are integer-offset crops with independent Gaussian sensor noise.  A stereo pair
is two crops of the same rows displaced by the disparity, so the pair is
rectified by construction (true matches share the image row, xL - xR = +d).
"""
import numpy as np

MARGIN = 96


def _upsample4(coarse, H, W):
    """Bilinear x4 up-sampling of a coarse float field to H x W (separable)."""
    ch, cw = coarse.shape
    ys = np.arange(H, dtype=np.float32) / 4.0
    xs = np.arange(W, dtype=np.float32) / 4.0
    y0 = np.minimum(ys.astype(np.int32), ch - 2)
    x0 = np.minimum(xs.astype(np.int32), cw - 2)
    fy = (ys - y0)[:, None].astype(np.float32)
    fx = (xs - x0)[None, :].astype(np.float32)
    rows = coarse[y0] * (1 - fy) + coarse[y0 + 1] * fy
    return rows[:, x0] * (1 - fx) + rows[:, x0 + 1] * fx


def scene_canvas(h, w, seed):
    """(h+192) x (w+192) u8 canvas for one sequence."""
    rng = np.random.default_rng(seed)
    H, W = h + 2 * MARGIN, w + 2 * MARGIN
    coarse = rng.integers(0, 256, size=(H // 4 + 2, W // 4 + 2)).astype(np.float32)
    smooth = _upsample4(coarse, H, W)
    white = rng.integers(0, 256, size=(H, W)).astype(np.float32)
    img = 0.6 * smooth + 0.4 * white
    n_rect = int(400 * (H * W) / float(704 * 480))
    xs = rng.integers(0, W - 20, size=n_rect)
    ys = rng.integers(0, H - 20, size=n_rect)
    ws = rng.integers(4, 20, size=n_rect)
    hs = rng.integers(4, 20, size=n_rect)
    gs = rng.integers(0, 256, size=n_rect)
    for x, y, rw, rh, g in zip(xs, ys, ws, hs, gs):
        img[y:y + rh, x:x + rw] = g
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def view(canvas, h, w, ox, oy, sigma, seed):
    """w x h crop at integer offset (ox, oy) plus Gaussian sensor noise (own seed)."""
    crop = canvas[oy:oy + h, ox:ox + w]
    if sigma <= 0:
        return np.ascontiguousarray(crop)
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal(size=(h, w), dtype=np.float32) * np.float32(sigma)
    return np.ascontiguousarray(np.clip(np.rint(crop.astype(np.float32) + noise), 0, 255).astype(np.uint8))


def stereo_pair(h, w, seed, disparity=12, sigma=2.0, shift=(0, 0), canvas=None):
    """Rectified (left, right) u8 pair.  ``shift`` moves both views (inter-frame motion)."""
    if canvas is None:
        canvas = scene_canvas(h, w, seed)
    sx, sy = shift
    left = view(canvas, h, w, MARGIN + sx, MARGIN + sy, sigma, seed * 2 + 1_000_003)
    right = view(canvas, h, w, MARGIN + sx + disparity, MARGIN + sy, sigma, seed * 2 + 1_000_004)
    return left, right


def stereo_sequence(h, w, seed, n_frames, step=(3, 1), disparity=12, sigma=2.0):
    """n_frames pairs of one scene, the camera translating by ``step`` px per frame."""
    canvas = scene_canvas(h, w, seed)
    out = []
    for f in range(n_frames):
        sx, sy = step[0] * f, step[1] * f
        left = view(canvas, h, w, MARGIN + sx, MARGIN + sy, sigma, (seed * 131 + f) * 2 + 2_000_003)
        right = view(canvas, h, w, MARGIN + sx + disparity, MARGIN + sy, sigma,
                     (seed * 131 + f) * 2 + 2_000_004)
        out.append((left, right))
    return out


def stereo_batch(h, w, n_pairs, seed0=0, disparity=12, sigma=2.0, n_scenes=None, first=0):
    """(n_pairs, h, w) left and right stacks.  Scenes are reused round-robin (``n_scenes``
    distinct canvases) with fresh sensor noise per pair so that every pair is distinct.
    ``first``: index of the first pair inside a larger (global) batch -- pair i of the result is
    pair first + i of stereo_batch(..., first=0), so a batch can be generated shard by shard."""
    if n_scenes is None:
        n_scenes = min(n_pairs, 8)
    canvases = [scene_canvas(h, w, seed0 + s) for s in range(n_scenes)]
    L = np.empty((n_pairs, h, w), np.uint8)
    R = np.empty((n_pairs, h, w), np.uint8)
    for k in range(n_pairs):
        i = first + k
        c = canvases[i % n_scenes]
        j = i // n_scenes
        L[k], R[k] = stereo_pair(h, w, seed0 + 7919 * i + 17, disparity, sigma,
                                 shift=((5 * j) % 64, (3 * j) % 48), canvas=c)
    return L, R
