"""Generates tests/golden/bilateral.npz: outputs of the UNMODIFIED reference kernel
BilateralFilter<float,float,{unsigned char,float}>(dOut, dIn, dImg, gs, gr, gc, size) (cu_bilateral.cu:110-155), compiled for
sm_100a (oracle/_ref) and run on a B200 -- the filter the applications apply to every cost-volume slice
(applications/stereo2/main.cpp:407-421):

    gpurun -- 'python tests/golden/make_golden_bilateral.py gpurun_out/golden'
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
    rng = np.random.Generator(np.random.PCG64(20261018))
    h, w, D = 37, 70, 5                      # not multiples of the 32x32 blocks; a few slices of a Hamming-like cost volume
    vol = (rng.integers(0, 64, (D, h, w)) / np.float32(64)).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    g8 = np.clip(40 + 2 * xx + 60 * (yy > 15) + rng.normal(0, 6, (h, w)), 0, 255).astype(np.uint8)
    gf = (g8 / np.float32(255)).astype(np.float32)
    g = {"vol": vol, "guide_u8": g8, "guide_f32": gf}
    for name, (guide, gs, gr, gc, size) in {"u8_s2": (g8, 2.0, 0.2, 10.0, 2), "u8_s5": (g8, 3.0, 0.1, 25.0, 5),
                                            "f32_s3": (gf, 1.5, 0.3, 0.05, 3), "f32_s0": (gf, 1.0, 0.2, 0.1, 0)}.items():
        g[f"out_{name}"] = np.stack([ref.bilateral_filter_joint(vol[d], guide, gs, gr, gc, size) for d in range(D)])
        g[f"par_{name}"] = np.array([gs, gr, gc, size], np.float32)
    np.savez_compressed(os.path.join(out_dir, "bilateral.npz"), **g)
    print("wrote bilateral.npz")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
