#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontback.py tests/test_gpu_cpp_shim.py -q -x 2>&1 | tail -25 > gpurun_out/r2_dense_tests.log
cat gpurun_out/r2_dense_tests.log
timeout 300 python - <<'PY'
import json, time, numpy as np, torch
from kangaroo_b200 import roo
from kangaroo_b200.synth import stereo_pair
res = []
for (w, h, D, rad) in ((640, 480, 64, 2), (1024, 720, 128, 2), (1024, 720, 128, 7)):
    L, R, _ = stereo_pair(w, h, D, config=2)
    l, r, d = roo.Image.from_numpy(L), roo.Image.from_numpy(R), roo.Image(w, h, np.uint8)
    ms = []
    for i in range(8):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); roo.DenseStereo(d, l, r, D, 0.05, rad); t1.record(); torch.cuda.synchronize()
        if i >= 3: ms.append(t0.elapsed_time(t1))
    rec = {"workload": f"{w}x{h}, maxDisp {D}, score_rad {rad}", "ms": round(float(np.median(ms)), 3)}
    if w <= 1024:
        try:
            from oracle import ref_gpu as ref
            ref.dense_stereo(L, R, D, 0.05, rad)
            t = time.perf_counter(); ref.dense_stereo(L, R, D, 0.05, rad); rec["reference_kernel_ms_host_clock"] = round((time.perf_counter() - t) * 1e3, 2)
        except Exception as e:
            rec["reference"] = str(e)
    res.append(rec)
json.dump(res, open("gpurun_out/r2_dense_bench.json", "w"), indent=1)
print(json.dumps(res))
PY
