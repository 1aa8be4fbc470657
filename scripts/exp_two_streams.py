#!/usr/bin/env python
"""Experiment: do two engines on two streams (half the batch each) beat one engine with the full batch?"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import roo  # noqa: E402
from kangaroo_b200.synth import stereo_pair  # noqa: E402

w, h, D = 1280, 720, 128
L, R, _ = stereo_pair(w, h, D, config=2)


def bench(n_eng, B, steps=20):
    engs = [roo.StereoEngine(w, h, D, dodiag=True, max_batch=B) for _ in range(n_eng)]
    streams = [torch.cuda.Stream() for _ in range(n_eng)]
    l = torch.from_numpy(np.stack([L] * B)).cuda()
    r = torch.from_numpy(np.stack([R] * B)).cuda()
    outs = [torch.empty((B, h, w), dtype=torch.float32, device="cuda") for _ in range(n_eng)]
    for _ in range(3):
        for e, s, o in zip(engs, streams, outs):
            e.run_device(l, r, o, stream=s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for e, s, o in zip(engs, streams, outs):
            e.run_device(l, r, o, stream=s)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for e in engs:
        e.close()
    return n_eng * B * steps / dt


for n_eng, B in ((1, 16), (2, 8), (2, 16), (4, 4), (3, 8)):
    print(n_eng, "engines x", B, "pairs:", round(bench(n_eng, B), 1), "pairs/s")
