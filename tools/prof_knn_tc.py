"""one launch of the tensor-core kNN engine on 2000 queries x 2 M rows (for ncu): python tools/prof_knn_tc.py [popc]"""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
import orbslam_jpminipc_b200 as pkg
from orbslam_jpminipc_b200._lib import check, lib, ptr
L = lib()
ex = pkg.ORBextractor(500, max_width=64, max_height=64, max_batch=1)
check(L.orb_set_knn_engine(ex._h, 0 if len(sys.argv) > 1 else 1), "engine")
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
db = torch.randint(0, 256, (2_000_000, 32), dtype=torch.uint8, device=dev, generator=g)
q = torch.randint(0, 256, (2000, 32), dtype=torch.uint8, device=dev, generator=g)
o = [torch.zeros(2000, dtype=torch.int32, device=dev) for _ in range(3)]
for _ in range(3):
    check(L.orb_hamming_knn2_device(ex._h, ptr(q), 2000, ptr(db), db.shape[0], 1, 0, ptr(o[0]), ptr(o[1]), ptr(o[2]), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "knn")
torch.cuda.synchronize()
print("ok", int(o[1].min()))
