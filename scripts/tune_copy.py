import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luisacomputegaussiansplatting_b200 import lcgs
dev = lcgs.Device(0)
n = 17_790_973
d_k = torch.randint(0, 2**45, (n,), dtype=torch.int64, device="cuda"); d_v = torch.arange(n, dtype=torch.int32, device="cuda")
d_ko, d_vo = torch.zeros_like(d_k), torch.zeros_like(d_v)
s = lcgs.DeviceRadixSort(); s.create(dev)
for bits in (0, 7, 9):
    for _ in range(3): s.SortPairs(None, d_k, d_ko, d_v, d_vo, n, 0, bits)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): s.SortPairs(None, d_k, d_ko, d_v, d_vo, n, 0, bits)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("bits %d: %.4f ms -> %.0f GB/s (24 B/pair%s)" % (bits, ms, n * 24 / ms / 1e6, "" if bits == 0 else " + 8 B histogram read"))
# torch copy for reference
a = torch.empty(n * 3, dtype=torch.int32, device="cuda"); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): b.copy_(a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("torch copy of the same bytes: %.4f ms -> %.0f GB/s" % (ms, n * 24 / ms / 1e6))
