"""Generates tests/golden/sqpen.npz and tests/golden/filtgrad.npz: outputs of the UNMODIFIED reference kernels
CostVolMinimumSquarePenaltySubpix (cu_dense_stereo.cu:122-174) and FilterDispGrad (cu_dense_stereo.cu:793-812),
compiled for sm_100a (oracle/_ref) and run on a B200:

    gpurun -- 'python tests/golden/make_golden_n4b.py gpurun_out/golden'

FilterDispGrad is called OUT OF PLACE with the output image pre-filled (the kernel differentiates the image it writes);
the applications call it in place, where the reference races with itself.
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
    rng = np.random.Generator(np.random.PCG64(20261017))
    # ---- CostVolMinimumSquarePenaltySubpix
    h, w, D = 40, 72, 32
    vol = (rng.integers(0, 64, (D, h, w)) / np.float32(64)).astype(np.float32)       # Hamming-like costs, many ties
    yy, xx = np.mgrid[0:h, 0:w]
    lastd = np.clip(8 + 0.2 * xx + rng.normal(0, 2.0, (h, w)), 0, D - 1).astype(np.float32)
    g = {"vol": vol, "lastd": lastd}
    for name, (sd, lam, theta) in {"a": (-1.0, 1.0, 0.5), "b": (-1.0, 0.25, 8.0), "c": (1.0, 4.0, 2.0)}.items():
        g[f"out_{name}"] = ref.costvol_minimum_square_penalty_subpix(vol, lastd, D, sd, lam, theta)
        g[f"par_{name}"] = np.array([sd, lam, theta], np.float32)
    np.savez_compressed(os.path.join(out_dir, "sqpen.npz"), **g)
    # ---- FilterDispGrad
    h, w = 48, 64
    yy, xx = np.mgrid[0:h, 0:w]
    grad_src = (20 + 0.3 * xx + 0.1 * yy + 6 * (xx > 30) + rng.normal(0, 0.3, (h, w))).astype(np.float32)
    grad_src[rng.random((h, w)) < 0.03] = np.nan                       # invalid disparities (left-right check)
    img_in = rng.random((h, w), dtype=np.float32) * 50
    f = {"grad_src": grad_src, "img_in": img_in}
    for thr in (0.05, 0.5, 4.0):
        f[f"out_{thr}"] = ref.filter_disp_grad(grad_src, img_in, thr)
    f["inplace_0.5"] = ref.filter_disp_grad(grad_src, grad_src, 0.5)   # values = the snapshot itself: what an in-place call means
    np.savez_compressed(os.path.join(out_dir, "filtgrad.npz"), **f)
    print("wrote sqpen.npz, filtgrad.npz")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
