#!/usr/bin/env python
"""Ablation timings on the real C3 frame: for each mask (csrc/common.cuh kAblate*) render a few frames
with that part of the work dropped (results are wrong on purpose) and print the per-stage times, so the
cost of one phase of a kernel can be read off as a difference.  Tuning only.

usage: python scripts/ablate.py [config] [mask ...]      (masks may be ORed, e.g. 5)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luisacomputegaussiansplatting_b200 import _capi, lcgs, scenes  # noqa: E402

NAMES = {0: "baseline", 1: "sort: no global stores", 2: "sort: no look-back", 4: "dup: no histograms", 8: "dup: no stores",
         16: "dup: coalesced instead of gathered rects", 32: "compact: no histograms", 64: "compact: no depth loads",
         128: "gather-scan: coalesced instead of gathered rects"}
key = sys.argv[1] if len(sys.argv) > 1 else "C3"
masks = [int(a) for a in sys.argv[2:]] or [0, 1, 2, 3, 4, 8, 16, 32, 64, 128, 0]
sc, cfg = scenes.make_config_scene(key)
dev = lcgs.Device(0)
r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, cfg.W, cfg.H)
cam = lcgs.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), cfg.W, cfg.H)
dev.set_profiling(True)
lib = _capi.load()
for m in masks:
    lib.lcgs_b200_debug_ablate(m)
    acc = None
    for it in range(8):
        r.render(cam)
        if it >= 3:
            st = dev.stage_times()
            acc = st if acc is None else {k: acc[k] + st[k] for k in st}
    avg = {k: round(v / 5, 4) for k, v in acc.items()}
    print("mask %3d %-50s total %.4f  %s" % (m, " + ".join(NAMES[b] for b in NAMES if b and (m & b)) or NAMES[0],
                                           sum(avg.values()), avg), flush=True)
lib.lcgs_b200_debug_ablate(0)
