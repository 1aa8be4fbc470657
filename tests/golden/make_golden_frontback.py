"""Generates tests/golden/frontback.npz: outputs of the UNMODIFIED reference kernels (oracle/_ref) for the
operators either side of the stereo path -- ElementwiseScaleBias, BoxHalf, Disp2Depth, DisparityImageToVbo --
run on a B200:

    gpurun -- 'python tests/golden/make_golden_frontback.py gpurun_out/golden'

then gpurun_out/golden/frontback.npz is copied into tests/golden/ and committed (inputs + reference outputs, so
the CPU tests need neither a GPU nor /root/reference).
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
    g = {}
    # ---- ElementwiseScaleBias: u8 / u16 / f32 inputs, the app's call (s = 1/255, offset 0) and a biased one
    g["sb_u8"] = rng.integers(0, 256, (36, 52), dtype=np.uint8)
    g["sb_u16"] = rng.integers(0, 65536, (36, 52), dtype=np.uint16)
    g["sb_f32"] = (rng.random((36, 52), dtype=np.float32) * 200 - 100).astype(np.float32)
    for nm in ("u8", "u16", "f32"):
        g[f"sb_{nm}_app"] = ref.elementwise_scale_bias(g["sb_" + nm], 1.0 / 255.0, 0.0)
        g[f"sb_{nm}_bias"] = ref.elementwise_scale_bias(g["sb_" + nm], 0.37, -1.25)
    # ---- BoxHalf: two pyramid levels, u8 and f32 (48x64 divides by the reference's gcd-sized blocks)
    g["bh_u8"] = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    g["bh_f32"] = rng.random((48, 64), dtype=np.float32)
    for nm in ("u8", "f32"):
        g[f"bh_{nm}_l1"] = ref.box_half(g["bh_" + nm])
        g[f"bh_{nm}_l2"] = ref.box_half(g[f"bh_{nm}_l1"])
    # ---- Disp2Depth / DisparityImageToVbo: subpixel disparities with zeros, negatives, NaN and a denormal
    d = (rng.random((32, 48), dtype=np.float32) * 64).astype(np.float32)
    d[0, :6] = [0.0, -0.0, -1.5, np.nan, 1e-40, np.inf]
    d[5:9, 7:11] = 0.0
    g["disp"] = d
    g["depth_min0"] = ref.disp2depth(d, 570.3, 0.12, 0.0)
    g["depth_min2"] = ref.disp2depth(d, 570.3, 0.12, 2.0)
    g["vbo"] = ref.disparity_image_to_vbo(d, 0.12, 570.3, 568.9, 23.4, 15.7)
    np.savez_compressed(os.path.join(out_dir, "frontback.npz"), **g)
    print("wrote", os.path.join(out_dir, "frontback.npz"), {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
