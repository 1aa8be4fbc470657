#!/usr/bin/env python
"""Experiment: fused vertical groups vs one pass per path as a function of the batch (development aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import roo  # noqa: E402
from kangaroo_b200.synth import stereo_pair  # noqa: E402

for (w, h, D, batches) in ((640, 480, 64, (1, 2, 4, 8)), (1280, 720, 128, (1, 2, 3)), (1920, 1080, 256, (1, 2)), (3840, 2160, 256, (1,))):
    L, R, _ = stereo_pair(w, h, D, config=2)
    for B in batches:
        res = []
        for fuse in (True, False):
            eng = roo.StereoEngine(w, h, D, dodiag=True, max_batch=B, fuse_vertical=fuse)
            l = torch.from_numpy(np.stack([L] * B)).cuda()
            r = torch.from_numpy(np.stack([R] * B)).cuda()
            out = torch.empty((B, h, w), dtype=torch.float32, device="cuda")
            for _ in range(3):
                eng.run_device(l, r, out)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                eng.run_device(l, r, out)
            torch.cuda.synchronize()
            res.append((time.perf_counter() - t0) / 10 * 1e3)
            eng.close()
        print(f"{w}x{h}x{D} batch {B}: fused {res[0]:.2f} ms, separate {res[1]:.2f} ms")
