"""A/B of the single-pair geometry (ROO_TUNE_SOLO_GEOMETRY) on BASELINE config 5 and a 1080p single pair."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import roo, capi
from bench import make_pairs
res = []
for (w, h, D) in ((3840, 2160, 256), (1920, 1080, 256)):
    L, R = make_pairs(w, h, D, 5, 1)
    l, r = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    outs = {}
    for knob in (0, 1, 0, 1):
        roo.set_tuning(capi.TUNE_SOLO_GEOMETRY, knob)
        e = roo.StereoEngine(w, h, D, dodiag=True, subpix=True, lrcheck=True, max_batch=1)
        d = torch.empty((1, h, w), dtype=torch.float32, device="cuda")
        for _ in range(3): e.run_device(l, r, d)
        torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(); e.run_device(l, r, d); t1.record(); torch.cuda.synchronize(); ms.append(t0.elapsed_time(t1))
        outs[knob] = d.cpu().numpy().copy()
        res.append({"size": f"{w}x{h}x{D}", "solo_geometry": knob, "ms": round(float(np.median(ms)), 3)})
        e.close()
    a, b = outs[0], outs[1]
    same = np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a).view(np.uint32), np.nan_to_num(b).view(np.uint32))
    res.append({"size": f"{w}x{h}x{D}", "identical": bool(same)})
roo.set_tuning(capi.TUNE_SOLO_GEOMETRY, 1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r2_solo_geometry.json", "w"), indent=1)
print(json.dumps(res))
