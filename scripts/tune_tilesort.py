#!/usr/bin/env python
"""Times the fused frame's tile sort (SortPairs over key bits [32, 45) of the C3 frame's depth-ordered
instance list) for the LCGS_SORT_VARIANT / LCGS_SORT_ABLATE set in the environment.  Ablations give
wrong results on purpose; they measure what one phase of the onesweep pass costs."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luisacomputegaussiansplatting_b200 import lcgs, scenes  # noqa: E402

cache = "/tmp/c3_tilekeys.npz"
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
b0, b1 = 32, 45
for _ in range(3):
    s.SortPairs(None, d_k, d_ko, d_v, d_vo, n, b0, b1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    s.SortPairs(None, d_k, d_ko, d_v, d_vo, n, b0, b1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
exact = None
if not os.environ.get("LCGS_SORT_ABLATE"):
    ku = keys.view(np.uint64)
    order = np.argsort(ku >> np.uint64(32), kind="stable")
    exact = np.array_equal(d_ko.cpu().numpy().view(np.uint64), ku[order]) and np.array_equal(d_vo.cpu().numpy(), vals[order])
print("variant %s ablate %s: n=%d tile sort %.4f ms (histogram + 2 passes) exact=%s" % (
    os.environ.get("LCGS_SORT_VARIANT", "default"), os.environ.get("LCGS_SORT_ABLATE", "0"), n, ms, exact), flush=True)
