"""Generates tests/golden/warp.npz: the UNMODIFIED reference kernel roo::Warp (oracle/_ref) on a B200:

    gpurun -- 'python tests/golden/make_golden_warp.py gpurun_out/golden'

Lookup tables: a radial-distortion map of the kind CreateMatlabLookupTable produces (clamped to [1, w-2] x [1, h-2]
like cu_lookup_warp.cu:69-73), an identity map on the integer grid, and a half-pixel shift.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_gpu as ref  # noqa: E402


def main(out_dir: str) -> None:
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.Generator(np.random.PCG64(20261020))
    h, w = 48, 64
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    fu, fv, u0, v0, k1, k2 = np.float32(60), np.float32(58), np.float32(31.5), np.float32(23.2), np.float32(-0.21), np.float32(0.05)
    pnu, pnv = (xx - u0) / fu, (yy - v0) / fv
    rr = pnu * pnu + pnv * pnv
    rf = 1 + k1 * rr + k2 * rr * rr
    radial = np.stack([np.clip(pnu * rf * fu + u0, 1, w - 2), np.clip(pnv * rf * fv + v0, 1, h - 2)], -1).astype(np.float32)
    ident = np.stack([np.clip(xx, 0, w - 2), np.clip(yy, 0, h - 2)], -1).astype(np.float32)
    half = np.stack([np.clip(xx + 0.5, 0, w - 2), np.clip(yy + 0.25, 0, h - 2)], -1).astype(np.float32)
    g = {"img": img, "radial": radial, "ident": ident, "half": half}
    for nm in ("radial", "ident", "half"):
        g["out_" + nm] = ref.warp(img, g[nm])
    np.savez_compressed(os.path.join(out_dir, "warp.npz"), **g)
    print("wrote", os.path.join(out_dir, "warp.npz"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
