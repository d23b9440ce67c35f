"""Measures GPU<->GPU copy bandwidth over NVLink (device 1 -> device 0), the number the multi-GPU design argument in
DESIGN.md §5 rests on (individual sharding would pull 7/8 of every parental row through this path each generation)."""
import torch
a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:1")
b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
print("peer access 0<-1:", torch.cuda.can_device_access_peer(0, 1))
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.device(0):
    s.record()
    for _ in range(10):
        b.copy_(a)
    e.record()
    e.synchronize()
ms = s.elapsed_time(e) / 10
print(f"1 GiB device1 -> device0: {ms:.3f} ms = {(1 << 30) / ms / 1e6:.1f} GB/s")
