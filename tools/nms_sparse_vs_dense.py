import torch, numpy as np, sys
sys.path.insert(0,".")
import workloads as W
from rs_detection_b200 import core
from rs_detection_b200._lib import NMS_ROTATED
for n in (20000, 50000, 100000):
    for canvas in (1024, 4096):
        d=torch.from_numpy(W.rotated_boxes(n, n, canvas=canvas, smin=8, smax=128)).cuda()
        s=torch.from_numpy(W.distinct_scores(n,n)).cuda()
        z=torch.zeros(n,dtype=torch.int32,device="cuda")
        for thr in (0.1,0.5):
            out=[]
            for lab in (None, z):
                for _ in range(2): core.nms(NMS_ROTATED, d, s, thr, labels=lab)
                torch.cuda.synchronize(); best=1e9
                for _ in range(3):
                    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
                    a.record(); r=core.nms(NMS_ROTATED, d, s, thr, labels=lab); b.record(); torch.cuda.synchronize()
                    best=min(best,a.elapsed_time(b))
                out.append((best,r.count))
            print(f"n={n} canvas={canvas} thr={thr}: default {out[0][0]:.2f} ms (kept {out[0][1]}), dense {out[1][0]:.2f} ms (kept {out[1][1]})", flush=True)
