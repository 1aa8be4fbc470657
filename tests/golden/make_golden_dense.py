"""Generates tests/golden/dense.npz: outputs of the UNMODIFIED reference block matcher
DenseStereo<{unsigned char, char}, unsigned char> (cu_dense_stereo.cu:209-253,376-406), compiled for sm_100a (oracle/_ref)
and run on a B200, for every score radius (0 = single-pixel squared difference, 1..7 = SANDPatchScore), positive and
negative disparity ranges and several acceptance thresholds, on tightly packed images:

    gpurun -- 'python tests/golden/make_golden_dense.py gpurun_out/golden'
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_gpu as ref  # noqa: E402
from kangaroo_b200.synth import stereo_pair  # noqa: E402


def main(out_dir: str) -> None:
    os.makedirs(out_dir, exist_ok=True)
    g = {}
    L, R, _ = stereo_pair(200, 64, 40, config=3)
    g["left"], g["right"] = L, R
    cases = {}
    for rad in range(8):
        cases[f"u8_r{rad}"] = (L, R, 40, 0.05 if rad else 0.5, rad, False)
    cases["u8_r2_t0"] = (L, R, 40, 0.0, 2, False)
    cases["u8_r3_big"] = (L, R, 120, 0.2, 3, False)
    cases["i8_r2_pos"] = (L, R, 40, 0.05, 2, True)
    cases["i8_r1_neg"] = (R, L, -40, 0.05, 1, True)          # right-to-left: negative disparities
    cases["i8_r4_neg"] = (R, L, -100, 0.1, 4, True)
    for name, (a, b, md, th, rad, signed) in cases.items():
        g[f"out_{name}"] = ref.dense_stereo(a, b, md, th, rad, signed)
        g[f"par_{name}"] = np.array([md, th, rad, int(signed), int(a is R)], np.float32)
    np.savez_compressed(os.path.join(out_dir, "dense.npz"), **g)
    print("wrote dense.npz")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
