#!/usr/bin/env python
"""Small end-to-end runs for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py` (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import roo  # noqa: E402
from kangaroo_b200.synth import stereo_pair  # noqa: E402

for (w, h, D, kw) in [(200, 70, 64, dict(dodiag=True, subpix=True, lrcheck=True)),
                      (77, 130, 128, dict(dodiag=True)),
                      (130, 33, 256, dict(dodiag=True, subpix=True, lrcheck=True)),
                      (64, 48, 32, dict()),
                      (33, 17, 40, dict(dodiag=True, window=roo.WIN_16x16))]:
    L, R, _ = stereo_pair(w, h, D, config=7)
    eng = roo.StereoEngine(w, h, D, max_batch=3, **kw)
    l = torch.from_numpy(np.stack([L] * 3)).cuda()
    r = torch.from_numpy(np.stack([R] * 3)).cuda()
    d = eng.run_device(l, r)
    torch.cuda.synchronize()
    print(w, h, D, kw, float(torch.nan_to_num(d).sum()))
    eng.close()
img = roo.Image.from_numpy(np.random.default_rng(0).random((37, 53), dtype=np.float32))
out = roo.Image(53, 37, np.float32)
for f in (roo.MedianFilterRejectNegative5x5, roo.MedianFilterRejectNegative7x7, roo.MedianFilterRejectNegative9x9):
    f(out, img, 10)
half = roo.Image(26, 18, np.float32)
roo.BoxHalf(half, img)
vbo = roo.Image(53, 37, roo.FLOAT4)
roo.DisparityImageToVbo(vbo, img, 0.1, 500.0, 500.0, 26.0, 18.0)
roo.Disp2Depth(img, out, 500.0, 0.1)
lut = roo.Image(53, 37, roo.FLOAT2)
roo.CreateMatlabLookupTable(lut, 60.0, 58.0, 26.0, 18.0, -0.2, 0.05)
roo.CreateMatlabLookupTable(lut, 60.0, 58.0, 26.0, 18.0, -0.2, 0.05, H_on=[1, 0.01, 0.5, -0.01, 1, 0.2, 1e-5, 0, 1])
raw = roo.Image.from_numpy(np.random.default_rng(1).integers(0, 256, (37, 53), dtype=np.uint8))
rect = roo.Image(53, 37, np.uint8)
roo.Warp(rect, raw, lut)
vol = roo.Volume(53, 37, 19, np.float32)
roo.CostVolumeFromStereoTruncatedAbsAndGrad(vol, img, img, -1.0, 0.9, 0.03, 0.008)
# engine with front end (rectify + one pyramid level) and the median stage
w, h, D, B = 80, 36, 32, 2
yy, xx = np.mgrid[0:2 * h, 0:2 * w].astype(np.float32)
tab = roo.Image.from_numpy(np.ascontiguousarray(np.stack([np.clip(xx + 0.3, 1, 2 * w - 2), np.clip(yy - 0.2, 1, 2 * h - 2)], -1)))
eng = roo.StereoEngine(w, h, D, dodiag=True, subpix=True, lrcheck=True, max_batch=B, fuse_vertical=True, median_size=7,
                       median_maxbad=20, median_iters=2)
eng.set_front_end(1, tab, tab)
rawb = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (B, 2 * h, 2 * w), dtype=np.uint8)).cuda()
d = eng.run_device(rawb, rawb)
torch.cuda.synchronize()
eng.close()
torch.cuda.synchronize()
# round 2: in-place median / FilterDispGrad, CostVolMinimumSquarePenaltySubpix, materialised-cost and generic-sweep code
# paths, the row-strip split engine (two and three strips on this device), the engine's FilterDispGrad stage
img2 = roo.Image.from_numpy(np.random.default_rng(4).random((40, 56), dtype=np.float32) * 20)
roo.MedianFilterRejectNegative5x5(img2, img2, 10)
roo.FilterDispGrad(img2, img2, 3.0)
vol = roo.Volume.from_numpy(np.random.default_rng(5).random((16, 40, 56), dtype=np.float32))
roo.CostVolMinimumSquarePenaltySubpix(roo.Image(56, 40, np.float32), vol, img2, 16, -1.0, 1.0, 2.0)
L, R, _ = stereo_pair(150, 60, 64, config=8)
l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()
for knob in (roo.capi.TUNE_INSWEEP_COST, roo.capi.TUNE_HSWEEP):
    roo.set_tuning(knob, 0)
    e = roo.StereoEngine(150, 60, 64, dodiag=True, subpix=True, lrcheck=True, fuse_vertical=True, filtgrad_threshold=0.5)
    e.run_device(l, r)
    torch.cuda.synchronize()
    e.close()
    roo.set_tuning(knob, 1)
for strips in (2, 3):
    se = roo.SplitStereoEngine(150, 60, 64, devices=[0] * strips, dodiag=True, subpix=True, lrcheck=True)
    out = torch.empty((60, 150), dtype=torch.float32).pin_memory()
    se.run_host(torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory(), out)
    se.close()
rng = np.random.default_rng(3)
for (w, h, D, rad) in ((70, 37, 3, 3), (300, 21, 2, 9), (13, 260, 2, 40)):
    gv = roo.Volume.from_numpy(rng.random((D, h, w), dtype=np.float32))
    gi = roo.Image.from_numpy(rng.random((h, w), dtype=np.float32))
    roo.GuidedFilterVolume(gv, gi, rad, 1e-3, D)
    bo = roo.Image(w, h, np.float32)
    roo.BoxFilter(bo, gi, None, rad)
    roo.ElementwiseMultiplyAdd(bo, gi, gi, bo, -1.0)
    roo.ElementwiseDivision(bo, gi, bo, 0.0, 1e-3)
torch.cuda.synchronize()
print("done")
