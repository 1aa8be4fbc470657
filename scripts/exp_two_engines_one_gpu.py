#!/usr/bin/env python
"""Experiment: roo_multi_engine with the same device listed twice = two engines (and host threads) on one GPU, so the
tail of one engine's launches overlaps the other's (development aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import roo  # noqa: E402
from kangaroo_b200.synth import stereo_pair  # noqa: E402

w, h, D, B, N = 1280, 720, 128, 16, 128
L, R, _ = stereo_pair(w, h, D, config=2)
lp = torch.from_numpy(np.stack([L] * N)).pin_memory()
rp = torch.from_numpy(np.stack([R] * N)).pin_memory()
dp = torch.empty((N, h, w), dtype=torch.float32).pin_memory()
ref = None
for devs in ([0], [0, 0], [0, 0, 0]):
    m = roo.MultiGpuStereoEngine(w, h, D, devices=devs, dodiag=True, max_batch=B)
    m.run_host(lp, rp, dp)
    t0 = time.perf_counter()
    for _ in range(3):
        m.run_host(lp, rp, dp)
    dt = (time.perf_counter() - t0) / 3
    if ref is None:
        ref = dp.numpy().copy()
    print(devs, round(N / dt, 1), "pairs/s from/to pinned host memory; identical:", np.array_equal(ref, dp.numpy()))
    m.close()
