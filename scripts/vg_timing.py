import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kangaroo_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libroo_b200_timing.so")
import torch, numpy as np
from kangaroo_b200 import roo
from bench import make_pairs
w,h,D,B=1280,720,128,16
L,R=make_pairs(w,h,D,2,B); l=torch.from_numpy(L).cuda(); r=torch.from_numpy(R).cuda(); d=torch.empty((B,h,w),dtype=torch.float32,device="cuda")
e=roo.StereoEngine(w,h,D,dodiag=True,max_batch=B)
for _ in range(3): e.run_device(l,r,d)
torch.cuda.synchronize()
out=(C.c_ulonglong*15)()
capi.lib().roo_engine_debug_counters(e._h, out, 15, 1)
for _ in range(5): e.run_device(l,r,d)
torch.cuda.synchronize()
capi.lib().roo_engine_debug_counters(e._h, out, 15, 0)
for role,name in enumerate(("lowest warp","interior warps","highest warp")):
    tcp,tup,tdn,tb,rows=[out[role*5+k] for k in range(5)]
    tot=tcp+tup+tdn+tb
    print("%-15s rows %9d  cycles/row: cp.async wait %6.0f | wait up %6.0f | wait down/copied %6.0f | body %6.0f | total %6.0f"%(name,rows,tcp/rows,tup/rows,tdn/rows,tb/rows,tot/rows))
