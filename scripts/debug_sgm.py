import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as ko
from oracle import ref_gpu
from kangaroo_b200 import roo
g = np.load("tests/golden/sgm.npz")
def gpu_sgm(volc, left, md, p1, p2, hz=True, vt=True, rv=True, dg=False):
    d, h, w = volc.shape
    vh = roo.Volume(w, h, d, np.float32); vh.fill_bytes(0x7F)
    roo.SemiGlobalMatching(vh, roo.Volume.from_numpy(volc), roo.Image.from_numpy(left), md, p1, p2, hz, vt, rv, dg)
    return vh.numpy()
for ieee in (0, 1):
    roo.set_ieee_division(bool(ieee))
    H = gpu_sgm(g["volc"], g["left_f32"], 12, 0.01, 0.02, 0, 1, 0)
    G = g["H_h0v1r0"]; O = ko.sgm(g["volc"], g["left_f32"], 12, 0.01, 0.02, 0, 1, 0)
    R = ref_gpu.sgm(g["volc"], g["left_f32"], 12, 0.01, 0.02, 0, 1, 0)
    print("ieee", ieee, "vs golden maxabs", np.abs(H-G).max(), "n_diff", (H!=G).sum(), "/", H.size,
          "| vs oracle n_diff", (H!=O).sum(), "maxabs", np.abs(H-O).max(), "| live ref vs golden", (R!=G).sum(), "| oracle vs golden", (O!=G).sum())
    idx = np.argwhere(H!=G)
    for i in idx[:8]:
        d,y,x = i
        print("  d,y,x", d,y,x, "H", H[d,y,x], "G", G[d,y,x], "O", O[d,y,x])
