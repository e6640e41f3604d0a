#!/usr/bin/env python
"""Experiment: frames of a sweep rendered alternately through TWO contexts on two streams (own workspaces and frame
buffers, shared scene), so that the HBM-bound front of frame i+1 (preprocess, binning, sorts) can overlap the
issue-bound blend of frame i.  Prints ms/frame for one stream and for two (with and without stream priorities)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luisacomputegaussiansplatting_b200 import lcgs, scenes  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "C3"
cache = "/tmp/scene_%s.npz" % key
cfg = scenes.CONFIGS[key]
if os.path.exists(cache):
    z = np.load(cache)
    arrs = [z[k] for k in ("pos", "scale", "rotq", "sh", "opacity")]
else:
    sc, cfg = scenes.make_config_scene(key)
    arrs = [sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity]
    np.savez(cache, pos=sc.pos, scale=sc.scale, rotq=sc.rotq, sh=sc.sh, opacity=sc.opacity)
W, H = cfg.W, cfg.H
devs = [lcgs.Device(0), lcgs.Device(0)]
d_arrs = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]
rs = [lcgs.Renderer(d, *d_arrs, W, H, rgb8=True) for d in devs]
vp = lcgs.view_params(lcgs.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), W, H))


def run(streams, steps=40):
    for i in range(6):
        rs[i % len(streams)].render_async(vp, stream=streams[i % len(streams)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]:
        s.wait_event(e0)
    for i in range(steps):
        rs[i % len(streams)].render_async(vp, stream=streams[i % len(streams)])
    for s in streams[1:]:
        ev = torch.cuda.Event()
        ev.record(s)
        streams[0].wait_event(ev)
    e1.record(streams[0])
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) / steps * 1e3


s0 = torch.cuda.Stream()
print("one stream          : %.4f ms/frame (wall %.4f)" % run([s0]), flush=True)
print("two streams         : %.4f ms/frame (wall %.4f)" % run([torch.cuda.Stream(), torch.cuda.Stream()]), flush=True)
print("two streams, prio   : %.4f ms/frame (wall %.4f)" % run([torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)]), flush=True)
devs.append(lcgs.Device(0))
rs.append(lcgs.Renderer(devs[2], *d_arrs, W, H, rgb8=True))
print("three streams       : %.4f ms/frame (wall %.4f)" % run([torch.cuda.Stream() for _ in range(3)]), flush=True)
n0, n1 = devs[0].num_rendered(), devs[1].num_rendered()
same = bool(torch.equal(rs[0].img, rs[1].img)) and bool(torch.equal(rs[0].vals[:n0], rs[1].vals[:n1]))
print("num_rendered", n0, n1, "frames identical:", same)
