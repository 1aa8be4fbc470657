"""Generates tests/golden/lookup.npz: the UNMODIFIED reference kernel CreateMatlabLookupTable (no homography) on a B200:

    gpurun -- 'python tests/golden/make_golden_lookup.py gpurun_out/golden'
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
    g = {"params_a": np.array([64, 48, 60.0, 58.0, 31.5, 23.2, -0.21, 0.05], np.float32),
         "params_b": np.array([96, 32, 410.3, 409.1, 47.7, 15.1, 0.12, -0.3], np.float32)}
    for nm in ("a", "b"):
        p = g["params_" + nm]
        g["lut_" + nm] = ref.create_matlab_lookup_table(int(p[0]), int(p[1]), *[float(x) for x in p[2:]])
    # the overload with a homography (a small rotation + shear + perspective term), clamped to [1, w-2] x [1, h-2]
    g["H"] = np.array([0.998, -0.021, 1.7, 0.019, 1.003, -0.9, 1.2e-5, -0.8e-5, 1.0], np.float32)
    for nm in ("a", "b"):
        p = g["params_" + nm]
        g["lut_h_" + nm] = ref.create_matlab_lookup_table_h(int(p[0]), int(p[1]), *[float(x) for x in p[2:]], g["H"])
    np.savez_compressed(os.path.join(out_dir, "lookup.npz"), **g)
    print("wrote", os.path.join(out_dir, "lookup.npz"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
