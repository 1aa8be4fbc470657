#!/usr/bin/env python
"""BASELINE config 5: ONE 3840x2160 pair, 256 disparities, 8 paths + subpixel + left-right check, split into row strips
across the visible GPUs (roo_split_engine_*).  Prints / writes one JSON record: milliseconds per pair at N strips, the
bytes handed between strips over NVLink, and whether the disparities equal the single-GPU engine's bit for bit.

    python scripts/c5_split.py [--gpus N] [--reps K] [--out gpurun_out/r2_c5_split.json] [--small]
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
    ap.add_argument("--gpus", type=int, default=0, help="strips = GPUs used (0: every power of two up to the visible count)")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default="")
    ap.add_argument("--small", action="store_true", help="1920x1080 instead of 3840x2160 (quick check)")
    ap.add_argument("--strip-ctas", type=int, default=-1, help="roo_set_tuning(ROO_TUNE_STRIP_CTAS_PER_SM): CTAs per SM of a strip-crossing sweep")
    args = ap.parse_args()
    w, h, D = (1920, 1080, 256) if args.small else (3840, 2160, 256)
    opts = dict(dodiag=True, subpix=True, lrcheck=True, lr_maxdiff=1.0)
    L, R, _ = stereo_pair(w, h, D, config=5)
    lp, rp = torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()
    ndev = torch.cuda.device_count()
    # single-GPU engine: the reference result and its time (fused vertical groups)
    torch.cuda.set_device(0)
    eng = roo.StereoEngine(w, h, D, max_batch=1, **opts)
    l1, r1 = lp[None].cuda(), rp[None].cuda()
    ref = eng.run_device(l1, r1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        eng.run_device(l1, r1, ref)
    e1.record()
    torch.cuda.synchronize()
    single_ms = e0.elapsed_time(e1) / args.reps
    ref = ref[0].cpu()
    eng.close()
    rec = {"workload": f"c5 {w}x{h}x{D} 8-path + subpix + LR check, one pair", "gpu_count_visible": ndev,
           "single_gpu_engine_ms_per_pair_device": single_ms, "splits": []}
    if args.strip_ctas >= 0:
        roo.set_tuning(roo.capi.TUNE_STRIP_CTAS_PER_SM, args.strip_ctas)
        rec["strip_ctas_per_sm"] = args.strip_ctas
    counts = [args.gpus] if args.gpus else [n for n in (1, 2, 4, 8) if n <= ndev]
    for n in counts:
        se = roo.SplitStereoEngine(w, h, D, devices=list(range(n)), **opts)
        out = torch.empty((h, w), dtype=torch.float32).pin_memory()
        for _ in range(2):
            se.run_host(lp, rp, out)
        t0 = time.perf_counter()
        dev_ms = []
        for _ in range(args.reps):
            se.run_host(lp, rp, out)
            dev_ms.append(se.last_stats()[0])
        wall = (time.perf_counter() - t0) / args.reps * 1e3
        _, nbytes = se.last_stats()
        se.close()
        same = bool(torch.equal(torch.nan_to_num(out, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)))
        rec["splits"].append({"strips": n, "device_ms_per_pair": float(np.median(dev_ms)), "host_to_host_ms_per_pair": wall,
                              "nvlink_bytes_per_pair": nbytes, "identical_to_single_gpu_engine": same})
        print(json.dumps(rec["splits"][-1]), flush=True)
    print(json.dumps(rec))
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(rec, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
