"""Generates tests/golden/guided.npz: outputs of the UNMODIFIED reference operators behind the applications' guided
filtering of a cost volume (applications/stereo2/main.cpp:392-405), compiled for sm_100a (oracle/_ref) and run on a B200:
BoxFilter<float,float,float> (cu_integral_image.h:26-38: two work-efficient prefix sums + a four-tap lookup),
the float elementwise operators it is composed with (cu_operations.cu:85-190) and the per-slice
ComputeMeanVarience / ComputeCovariance / GuidedFilter sequence (cu_integral_image.h:42-93):

    gpurun -- 'python tests/golden/make_golden_guided.py gpurun_out/golden'
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
    g = {}
    # box filter: sizes around the scan's power-of-two padding, radii below / above the image size
    for name, (h, w, rad) in {"a": (37, 70, 3), "b": (64, 128, 9), "c": (33, 257, 1), "d": (130, 45, 40), "e": (300, 520, 14)}.items():
        img = rng.random((h, w), dtype=np.float32)
        if name == "c":
            img = (img * 255).astype(np.float32)           # 8-bit-like magnitudes: larger prefix sums
        g[f"box_in_{name}"] = img
        g[f"box_out_{name}"] = ref.box_filter(img, rad)
        g[f"box_rad_{name}"] = np.int32(rad)
    # elementwise operators with non-trivial scalars
    a = (rng.random((29, 53), dtype=np.float32) - 0.5).astype(np.float32)
    b = (rng.random((29, 53), dtype=np.float32) + 0.05).astype(np.float32)
    c = (rng.random((29, 53), dtype=np.float32) * 3).astype(np.float32)
    g["ew_a"], g["ew_b"], g["ew_c"] = a, b, c
    g["ew_mul"] = ref.elementwise(0, a, b, None, 1.7, -0.3)
    g["ew_div"] = ref.elementwise(1, a, b, None, 0.25, 0.01, 1.3, 0.5)
    g["ew_sq"] = ref.elementwise(2, a, None, None, 0.9, 0.1)
    g["ew_mad"] = ref.elementwise(3, a, b, c, -1.0, 1.0, 0.0)
    g["ew_mad2"] = ref.elementwise(3, a, b, c, 0.7, -1.1, 0.2)
    # the applications' loop on a Hamming-like cost volume with a piecewise-smooth guide image
    h, w, D = 61, 90, 6
    vol = (rng.integers(0, 64, (D, h, w)) / np.float32(64)).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    guide = (np.clip(40 + 2 * xx + 60 * (yy > 25) + rng.normal(0, 6, (h, w)), 0, 255) / 255).astype(np.float32)
    g["gf_vol"], g["gf_guide"] = vol, guide
    for name, (rad, eps) in {"r4": (4, 1e-4), "r9": (9, 1e-2), "r1": (1, 1e-3)}.items():
        g[f"gf_out_{name}"] = ref.guided_filter_volume(vol, guide, rad, eps)
        g[f"gf_par_{name}"] = np.array([rad, eps], np.float32)
    np.savez_compressed(os.path.join(out_dir, "guided.npz"), **g)
    print("wrote guided.npz")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
