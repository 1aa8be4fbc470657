#!/usr/bin/env python
"""Device time of the guided cost-volume filter at the c2 working size, against the reference's own
per-slice sequence on the same GPU.

ours       roo.GuidedFilterVolume on a 1024 x 720 x 128 fp32 volume (in place), CUDA events on the launching stream,
           median of 5 runs, the volume (377 MB) is larger than L2.
reference  oracle/_ref (the unmodified kernels compiled for sm_100a): ComputeCovariance + GuidedFilter per slice as
           applications/stereo2/main.cpp:392-405 calls them, 16 slices timed with the host clock around the shim call
           (it synchronises after each slice and copies the slice in and out; both are small against its 37 launches
           per slice) and scaled to 128.
Algorithmic bytes per pixel and slice: read P, write q = 8 B.  Bytes the five passes really move: row scans 4 + 8 (P read once
for P and I*P) and 8 + 8 (a, b computed from the integral images of P and I*P, only their row sums written), column scans
2 x 16 in place, the last lookup 8 read + 4 written -> 72 B.
Writes gpurun_out/gfilter_bench.json.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import roo  # noqa: E402

PEAK = 6454.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def main():
    rng = np.random.default_rng(0)
    w, h, D, rad, eps = 1024, 720, 128, 9, 1e-2
    vol = (rng.integers(0, 64, (D, h, w)) / np.float32(64)).astype(np.float32)
    guide = rng.random((h, w), dtype=np.float32)
    v, g = roo.Volume.from_numpy(vol), roo.Image.from_numpy(guide)
    ms = []
    for i in range(8):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        roo.GuidedFilterVolume(v, g, rad, eps, D)
        t1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(t0.elapsed_time(t1))
    ours = float(np.median(ms))
    px = w * h * D
    res = {"workload": f"{w}x{h}x{D} fp32 cost volume, rad {rad}", "ms": round(ours, 3), "launches": 3 + 5,
           "algorithmic_GBps": round(px * 8 / ours / 1e6, 1), "moved_bytes_per_px": 72,
           "moved_GBps": round(px * 72 / ours / 1e6, 1), "peak_gbs": PEAK, "frac_moved": round(px * 72 / ours / 1e6 / PEAK, 3)}
    try:
        from oracle import ref_gpu as ref
        n = 16
        ref.guided_filter_volume(vol[:2], guide, rad, eps)          # warm-up
        t = time.perf_counter()
        ref.guided_filter_volume(vol[:n], guide, rad, eps)
        dt = (time.perf_counter() - t) * 1e3
        res["reference_ms_scaled_to_128_slices"] = round(dt * D / n, 1)
        res["reference_launches"] = 21 + 37 * D
        res["speedup_vs_reference_kernels"] = round(dt * D / n / ours, 1)
    except Exception as e:  # oracle/_ref not built
        res["reference"] = f"unavailable: {e}"
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gfilter_bench.json", "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
