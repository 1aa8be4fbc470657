"""Generates tests/golden/median.npz: outputs of the UNMODIFIED reference kernels MedianFilterRejectNegative{5x5,7x7,9x9}
(oracle/_ref), called OUT OF PLACE, on a B200:

    gpurun -- 'python tests/golden/make_golden_median.py gpurun_out/golden'

Two inputs: a NaN-free subpixel disparity image (what applications/stereo2/main.cpp:438-444 feeds the filter) and the
same image with invalid pixels (NaN, +-inf) sprinkled in, which exercises the bad-pixel counting.
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
    rng = np.random.Generator(np.random.PCG64(20261019))
    h, w = 48, 64   # divides by the reference's 16x16 blocks
    yy, xx = np.mgrid[0:h, 0:w]
    clean = (20 + 0.2 * xx + 5 * (yy > 20) + rng.normal(0, 1.5, (h, w))).astype(np.float32)
    clean[rng.random((h, w)) < 0.05] += 30          # outliers the filter is there to remove
    clean[10:14, 30:36] = 7.25                      # a constant patch: ties
    dirty = clean.copy()
    m = rng.random((h, w))
    dirty[m < 0.06] = np.nan
    dirty[(m >= 0.06) & (m < 0.07)] = np.inf
    dirty[(m >= 0.07) & (m < 0.075)] = -np.inf
    dirty[30:40, 5:15] = np.nan                     # a hole larger than the windows
    g = {"clean": clean, "dirty": dirty}
    for size in (5, 7, 9):
        g[f"clean_{size}_mb100"] = ref.median_filter_reject_negative(clean, size, 100)
        g[f"clean_{size}_mb0"] = ref.median_filter_reject_negative(clean, size, 0)      # bad < maxbad never holds
        for mb in (1, 4, 100):
            g[f"dirty_{size}_mb{mb}"] = ref.median_filter_reject_negative(dirty, size, mb)
    np.savez_compressed(os.path.join(out_dir, "median.npz"), **g)
    print("wrote", os.path.join(out_dir, "median.npz"), len(g), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
