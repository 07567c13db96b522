import torch, time
x = torch.empty(92405760, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
y = torch.empty(15361024, dtype=torch.uint8).pin_memory()
dy = torch.empty_like(y, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True); y.copy_(dy, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("H2D 92.4MB: %.3f ms  %.1f GB/s" % (ms, 92.405760 / ms))
e0.record()
for _ in range(10): y.copy_(dy, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("D2H 15.4MB: %.3f ms  %.1f GB/s" % (ms, 15.361024 / ms))
