cd $GRAFT_REPO_ROOT
cp orbslam_jpminipc_b200/liborb_b200.so /tmp/orig.so
trap 'cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so' EXIT     # round-1 script: swaps the product library; always restore it (newer A/B runs use tools/ab_lib.sh + ORB_B200_LIB instead)
for t in 64x32 64x64 128x64 64x128; do
  cp orbslam_jpminipc_b200/liborb_b200_t$t.so orbslam_jpminipc_b200/liborb_b200.so
  python - <<PY
import time, numpy as np, orbslam_jpminipc_b200 as pkg
from orbslam_jpminipc_b200.synth import synth_frame
for (h,w,nf) in [(480,640,1000),(480,752,1000),(376,1241,2000)]:
    ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=1)
    img = synth_frame(h, w, 1000)
    for _ in range(5): ex(img)
    t0=time.perf_counter()
    for _ in range(50): ex(img)
    print("$t", (h,w,nf), "single-frame host-API latency ms %.3f" % ((time.perf_counter()-t0)/50*1e3))
    ex.close()
PY
done
cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so
