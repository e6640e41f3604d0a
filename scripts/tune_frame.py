#!/usr/bin/env python
"""Per-stage times of the fused C3 frame for the LCGS_* tuning variables set in the environment."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luisacomputegaussiansplatting_b200 import lcgs, scenes  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "C3"
import numpy as np
cache = "/tmp/scene_%s.npz" % key
sc, cfg = None, scenes.CONFIGS[key]
if os.path.exists(cache):
    z = np.load(cache)
    arrs = [z[k] for k in ("pos", "scale", "rotq", "sh", "opacity")]
else:
    sc, cfg = scenes.make_config_scene(key)
    arrs = [sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity]
    np.savez(cache, pos=sc.pos, scale=sc.scale, rotq=sc.rotq, sh=sc.sh, opacity=sc.opacity)
dev = lcgs.Device(0)
r = lcgs.Renderer(dev, *arrs, cfg.W, cfg.H)
cam = lcgs.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), cfg.W, cfg.H)
dev.set_profiling(True)
acc = None
for it in range(13):
    r.render(cam)
    if it >= 3:
        st = dev.stage_times()
        acc = st if acc is None else {k: acc[k] + st[k] for k in st}
avg = {k: round(v / 10, 4) for k, v in acc.items()}
import torch
n = dev.num_rendered()
pos = torch.arange(n, device="cuda", dtype=torch.int64) % 1000003
sig = (int((r.keys[:n] * (pos + 1)).sum().item()) & 0xFFFFFFFFFFFF, int((r.vals[:n].to(torch.int64) * (pos + 7)).sum().item()) & 0xFFFFFFFFFFFF,
       int(r.img.view(torch.int32).to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF, int(r.ranges.to(torch.int64).sum().item()))
env = {k: v for k, v in os.environ.items() if k.startswith("LCGS_")}
print("%s total %.4f %s sig %s" % (env, sum(avg.values()), avg, "%x.%x.%x.%x" % sig), flush=True)
