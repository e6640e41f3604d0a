#!/usr/bin/env python
"""Times lcgs_b200_sort_pairs_u64_u32 on the C3 frame's real (tile<<32|depth) keys for one
LCGS_SORT_VARIANT (set in the environment before the process starts) and checks the result."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luisacomputegaussiansplatting_b200 import lcgs, scenes  # noqa: E402

cache = "/tmp/c3_keys.npz"
dev = lcgs.Device(0)
if os.path.exists(cache):
    z = np.load(cache)
    keys, vals = z["keys"], z["vals"]
else:
    sc, cfg = scenes.make_config_scene("C3")
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, cfg.W, cfg.H)
    n = r.render(lcgs.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), cfg.W, cfg.H))
    keys = r.keys_unsorted[:n].cpu().numpy()
    vals = r.vals_unsorted[:n].cpu().numpy()
    np.savez(cache, keys=keys, vals=vals)
    del r
n = keys.shape[0]
d_k, d_v = torch.from_numpy(keys).cuda(), torch.from_numpy(vals).cuda()
d_ko, d_vo = torch.zeros_like(d_k), torch.zeros_like(d_v)
s = lcgs.DeviceRadixSort()
s.create(dev)
end_bit = 45
for _ in range(3):
    s.SortPairs(None, d_k, d_ko, d_v, d_vo, n, 0, end_bit)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    s.SortPairs(None, d_k, d_ko, d_v, d_vo, n, 0, end_bit)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
ok = bool((d_ko[1:] >= d_ko[:-1]).all().item())
ku = keys.view(np.uint64)
order = np.argsort(ku, kind="stable")
exact = np.array_equal(d_ko.cpu().numpy().view(np.uint64), ku[order]) and np.array_equal(d_vo.cpu().numpy(), vals[order])
print("variant %s: n=%d sort %.4f ms (%.1f GB/s algorithmic, 6 passes + histogram) sorted=%s exact=%s" % (
    os.environ.get("LCGS_SORT_VARIANT", "default"), n, ms, n * (8 + 24 * 6) / ms / 1e6, ok, exact), flush=True)
