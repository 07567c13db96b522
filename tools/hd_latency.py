import numpy as np, time, sys, os
sys.path.insert(0,'.')
import orbslam_jpminipc_b200 as pkg
from oracle import pyoracle as po
rng = np.random.default_rng(3)
for (h,w,nf,kind) in [(1080,1920,1000,'noise'),(1080,1920,1000,'synth'),(720,1280,200,'noise')]:
    if kind=='noise': img = rng.integers(0,256,(h,w),dtype=np.uint8)
    else:
        from orbslam_jpminipc_b200.synth import synth_frame
        img = synth_frame(h,w,5)
    ex = pkg.ORBextractor(nf,1.2,8,1,20,device=0,max_width=w,max_height=h,max_batch=1)
    k,d = ex(img); k,d = ex(img)
    t=time.perf_counter()
    for _ in range(20): ex(img)
    ms=(time.perf_counter()-t)/20*1e3
    rk,rd = po.OracleExtractor(nf,1.2,8,1,20)(img)
    print(kind,h,w,nf,"ms/frame",round(ms,3),"exact",len(k)==len(rk) and np.array_equal(d,rd) and np.array_equal(k['x'],rk['x']))
