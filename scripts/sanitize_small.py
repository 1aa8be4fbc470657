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
torch.cuda.synchronize()
print("done")
