#!/usr/bin/env python
"""Repeat the full C3 frame and a sweep of orbit views many times: sorted keys, sorted values, ranges and the image
must be bit-identical on every repetition (a scheduling-dependent bug in the look-backs or in the warp ranking of the
sort -- which relies on a warp's same-address shared-memory atomics executing in program order -- would show here)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luisacomputegaussiansplatting_b200 import lcgs, scenes  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sc, cfg = scenes.make_config_scene("C3")
dev = lcgs.Device(0)
r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, cfg.W, cfg.H, keep_intermediates=False)
bad = 0
for view in (None, 3, 77, 140, 201):
    pose = (scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world)) if view is None else scenes.orbit_pose(view)
    cam = lcgs.make_camera(*pose, cfg.W, cfg.H)
    n = r.render(cam)
    ref = [t.clone() for t in (r.keys[:n], r.vals[:n], r.ranges, r.img)]
    for i in range(reps):
        m = r.render(cam)
        same = m == n and all(torch.equal(a, b) for a, b in zip(ref, (r.keys[:n], r.vals[:n], r.ranges, r.img)))
        bad += 0 if same else 1
    print("view", view, "N", n, "repetitions", reps, "mismatches so far", bad, flush=True)
print("DETERMINISTIC" if bad == 0 else "NON-DETERMINISTIC: %d" % bad)
sys.exit(0 if bad == 0 else 1)
