import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kangaroo_b200 import roo
from kangaroo_b200.synth import stereo_pair
def run(L,R,D,**kw):
    h,w=L.shape
    e=roo.StereoEngine(w,h,D,max_batch=1,keep_volume=True,**kw)
    d=e.run_device(torch.from_numpy(L[None]).cuda(),torch.from_numpy(R[None]).cuda()).cpu().numpy()
    H=e.export_volume(0).numpy(); e.close(); return d,H
for (w,h,D) in [(70,33,40),(130,20,64),(96,40,64),(40,24,12)]:
    L,R,_=stereo_pair(w,h,D,config=11)
    roo.set_ieee_division(True)
    for rev in (False, True):
        df,Hf=run(L,R,D,dodiag=True,doreverse=rev,fuse_vertical=True)
        ds,Hs=run(L,R,D,dodiag=True,doreverse=rev,fuse_vertical=False)
        bad=np.argwhere(Hf!=Hs)
        print((w,h,D),"rev",rev,"ndiff",len(bad), "first", bad[:5].tolist() if len(bad) else None)
        if len(bad):
            ys=np.unique(bad[:,1]); xs=np.unique(bad[:,2]); ds_=np.unique(bad[:,0])
            print("  y range",ys.min(),ys.max(),"x range",xs.min(),xs.max(),"d range",ds_.min(),ds_.max())
