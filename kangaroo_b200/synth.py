"""Deterministic synthetic stereo pairs: textured left image, piecewise-planar
ground-truth disparity, right image warped from the left (left x <-> right x-d, i.e. sd = -1).

Harness code shared by the tests and bench.py; nothing here is on the measured path.
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0x5EED0000


def _value_noise(rng: np.random.Generator, h: int, w: int, cell: int) -> np.ndarray:
    """Bilinearly interpolated lattice noise with `cell`-pixel cells, in [0, 1)."""
    if cell == 1:
        return rng.random((h, w), dtype=np.float32)
    gh, gw = h // cell + 2, w // cell + 2
    g = rng.random((gh, gw), dtype=np.float32)
    ys = np.arange(h, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0 = ys.astype(np.int64)
    x0 = xs.astype(np.int64)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def stereo_pair(w: int, h: int, max_disp: int, config: int = 0, index: int = 0, stress: bool = False):
    """Returns (left u8 HxW, right u8 HxW, d_gt int32 HxW)."""
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + 1000 * config + index))
    if stress:  # iid uniform images: worst-case tie rate
        left = rng.integers(0, 256, (h, w), dtype=np.uint8)
        right = rng.integers(0, 256, (h, w), dtype=np.uint8)
        return left, right, np.zeros((h, w), np.int32)
    tex = np.zeros((h, w), np.float32)
    for cell, weight in ((64, 8.0), (16, 4.0), (4, 2.0), (1, 1.0)):
        tex += weight * _value_noise(rng, h, w, cell)
    tex -= tex.min()
    tex /= max(float(tex.max()), 1e-6)
    left = np.clip(np.rint(tex * 255.0), 0, 255).astype(np.uint8)

    # piecewise-planar disparity: 6 regions (a background plane + 5 rectangles), fronto-parallel or slanted
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    dmax = 0.75 * max_disp
    d_gt = np.full((h, w), 0.15 * dmax, np.float32) + (yy / max(h - 1, 1)) * 0.25 * dmax
    for _ in range(5):
        x0, x1 = sorted(rng.integers(0, w, 2).tolist())
        y0, y1 = sorted(rng.integers(0, h, 2).tolist())
        base = float(rng.uniform(0.1, 1.0)) * dmax
        sx = float(rng.uniform(-0.05, 0.05))
        sy = float(rng.uniform(-0.05, 0.05))
        plane = base + sx * (xx - x0) + sy * (yy - y0)
        d_gt[y0:y1 + 1, x0:x1 + 1] = plane[y0:y1 + 1, x0:x1 + 1]
    d_gt = np.clip(np.rint(d_gt), 0, int(dmax)).astype(np.int32)

    # right(x', y) = left(x' + d, y): forward-warp far-to-near so that nearer surfaces win, then fill
    # holes from the left neighbour (clamp-to-edge at the image border)
    right = np.zeros((h, w), np.int32) - 1
    xs = np.arange(w)
    for y in range(h):
        order = np.argsort(d_gt[y], kind="stable")
        xr = xs[order] - d_gt[y][order]
        ok = xr >= 0
        right[y][xr[ok]] = left[y][order][ok]
    for y in range(h):
        row = right[y]
        holes = row < 0
        if holes.any():
            idx = np.where(~holes, xs, -1)
            np.maximum.accumulate(idx, out=idx)
            first = int(np.argmax(~holes)) if (~holes).any() else 0
            idx[idx < 0] = first
            row[:] = np.where((~holes).any(), row[idx], left[y])
    noise = rng.integers(-2, 3, (h, w))
    right = np.clip(right + noise, 0, 255).astype(np.uint8)
    return left, right, d_gt
