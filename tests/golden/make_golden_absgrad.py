"""Generates tests/golden/absgrad.npz: the UNMODIFIED reference kernel CostVolumeFromStereoTruncatedAbsAndGrad
(oracle/_ref) on a B200:

    gpurun -- 'python tests/golden/make_golden_absgrad.py gpurun_out/golden'
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
    L, R, _ = stereo_pair(64, 24, 16, config=91)
    lf, rf = (L / np.float32(255)).astype(np.float32), (R / np.float32(255)).astype(np.float32)
    g = {"left": lf, "right": rf}
    g["vol_sdm1"] = ref.costvol_abs_and_grad(lf, rf, 16, -1.0, 0.9, 0.03, 0.008)
    g["vol_sdp1"] = ref.costvol_abs_and_grad(rf, lf, 16, +1.0, 0.5, 0.1, 0.02)
    np.savez_compressed(os.path.join(out_dir, "absgrad.npz"), **g)
    print("wrote", os.path.join(out_dir, "absgrad.npz"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
