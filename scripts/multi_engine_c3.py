#!/usr/bin/env python
"""BASELINE config 3: KITTI-shaped 1242x375 pairs, 128 disparities, a batch of 512 pairs sharded across the GPUs of the box
by roo_multi_engine_run_host (one engine + one host thread per device inside the C++ library, no collective).
Writes pairs/s for 1 .. all visible GPUs and whether the disparities are identical whatever the device count.

    python scripts/multi_engine_c3.py [--pairs 512] [--out gpurun_out/r2_multi_engine.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kangaroo_b200 import roo  # noqa: E402
from kangaroo_b200.synth import stereo_pair  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=512)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    w, h, D, n = 1242, 375, 128, args.pairs
    base = [stereo_pair(w, h, D, config=3, index=i) for i in range(8)]   # 8 distinct pairs, cycled and row-rolled
    L = torch.from_numpy(np.stack([np.roll(base[i % 8][0], i // 8, axis=0) for i in range(n)])).pin_memory()
    R = torch.from_numpy(np.stack([np.roll(base[i % 8][1], i // 8, axis=0) for i in range(n)])).pin_memory()
    ndev = torch.cuda.device_count()
    rec = {"workload": f"c3 {w}x{h}x{D} 4-path + WTA, batch of {n} pairs from and to pinned host memory",
           "api": "roo_multi_engine_run_host", "gpu_count_visible": ndev, "runs": []}
    ref = None
    for g in [c for c in (1, 2, 4, 8) if c <= ndev]:
        m = roo.MultiGpuStereoEngine(w, h, D, devices=list(range(g)), max_batch=16)
        out = torch.empty((n, h, w), dtype=torch.float32).pin_memory()
        m.run_host(L[:32 * g], R[:32 * g], out[:32 * g])   # warm-up: scratch, staging buffers, streams
        t0 = time.perf_counter()
        m.run_host(L, R, out)
        dt = time.perf_counter() - t0
        m.close()
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(out, ref))
        rec["runs"].append({"gpus": g, "pairs_per_s": n / dt, "seconds": dt, "identical_disparities": same})
        print(json.dumps(rec["runs"][-1]), flush=True)
    print(json.dumps(rec))
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(rec, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
