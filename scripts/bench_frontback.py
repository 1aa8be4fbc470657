#!/usr/bin/env python
"""Device-time of the front-end / back-end operators against the HBM roofline.

One-touch elementwise kernels: algorithmic bytes = bytes read + bytes written once.  Images larger than L2 are
not realistic for camera frames, so L2 is flushed between timed launches (a 256 MB write) and each launch is
timed alone with CUDA events on the launching (current) stream.  Writes gpurun_out/frontback_bench.json.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import roo  # noqa: E402

PEAK = 6454.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, flush, reps=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        fn()
        t1.record()
        torch.cuda.synchronize()
        ms.append(t0.elapsed_time(t1))
    return float(np.median(ms))


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"peak_gbs": PEAK, "note": "median of 20 launches, L2 flushed before each, CUDA events", "results": []}
    for w, h in ((1280, 720), (3840, 2160)):
        raw = roo.Image(w, h, np.uint8)
        imgf = roo.Image(w, h, np.float32)
        half = roo.Image(w // 2, h // 2, np.float32)
        half8 = roo.Image(w // 2, h // 2, np.uint8)
        disp = roo.Image.from_numpy((np.random.default_rng(0).random((h, w), dtype=np.float32) * 100).astype(np.float32))
        depth = roo.Image(w, h, np.float32)
        vbo = roo.Image(w, h, roo.FLOAT4)
        px = w * h
        cases = [
            ("ElementwiseScaleBias<float,uchar,float>", lambda: roo.ElementwiseScaleBias(imgf, raw, 1 / 255.0), px * 5),
            ("BoxHalf<float,float,float>", lambda: roo.BoxHalf(half, imgf), px * 4 + px),
            ("BoxHalf<uchar,uint,uchar>", lambda: roo.BoxHalf(half8, raw), px + px // 4),
            ("Disp2Depth", lambda: roo.Disp2Depth(disp, depth, 500.0, 0.1), px * 8),
            ("DisparityImageToVbo", lambda: roo.DisparityImageToVbo(vbo, disp, 0.1, 500.0, 500.0, w / 2, h / 2), px * 20),
        ]
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        lut = roo.Image.from_numpy(np.ascontiguousarray(np.stack([np.clip(xx * 0.98 + 3.3, 1, w - 2),
                                                                    np.clip(yy * 0.98 + 2.7, 1, h - 2)], -1).astype(np.float32)))
        rect = roo.Image(w, h, np.uint8)
        cases.append(("Warp (rectification lookup)", lambda: roo.Warp(rect, raw, lut), px * (8 + 1 + 1)))
        med = roo.Image(w, h, np.float32)
        for size in (5, 7, 9):
            cases.append((f"MedianFilterRejectNegative{size}x{size}",
                          (lambda f: (lambda: f(med, disp, 50)))(getattr(roo, f"MedianFilterRejectNegative{size}x{size}")), px * 8))
        for name, fn, nbytes in cases:
            ms = timed(fn, flush)
            gbs = nbytes / (ms * 1e-3) / 1e9
            out["results"].append({"op": name, "w": w, "h": h, "ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs,
                                   "frac_of_peak": gbs / PEAK})
            print(f"{name:42s} {w}x{h}: {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  {gbs / PEAK:5.1%}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/frontback_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
