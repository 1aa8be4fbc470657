#!/usr/bin/env python
"""One 3840x2160x256 pair through the fused engine (the single-pair geometry of sgm_vgroup_kernel<8,..>: 12 warps x 2 columns,
250 bands) for `compute-sanitizer --tool memcheck|synccheck python scripts/sanitize_solo.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import roo  # noqa: E402
from bench import make_pairs  # noqa: E402

w, h, D = 3840, 2160, 256
L, R = make_pairs(w, h, D, 5, 1)
e = roo.StereoEngine(w, h, D, dodiag=True, subpix=True, lrcheck=True, max_batch=1)
d = torch.empty((1, h, w), dtype=torch.float32, device="cuda")
e.run_device(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda(), d)
torch.cuda.synchronize()
print("done", float(torch.nan_to_num(d).sum()))
e.close()
