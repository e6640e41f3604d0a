#!/usr/bin/env python
"""Generates tests/golden/tiny_frame.npz: a 300-Gaussian frame and every intermediate the oracle
produces for it.  The reference cannot run here, so the fixture pins the ORACLE
(regression) and gives the GPU tests a committed vector set independent of the generator code.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from luisacomputegaussiansplatting_b200 import scenes  # noqa: E402
from oracle import oracle as orc  # noqa: E402

W, H, P = 96, 64, 300
sc, cfg = scenes.make_config_scene("C3", P=P)
# enlarge the splats so the tiny frame has overlap, saturation and multi-tile Gaussians
scale = (sc.scale * np.float32(40.0)).astype(np.float32)
pose = (scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world))
cam = orc.make_camera(*pose, W, H)
vp = orc.view_params(cam)
fr = orc.forward(sc.pos, scale, sc.rotq, sc.sh, sc.opacity, vp, bg=(0.1, 0.2, 0.3))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_frame.npz")
np.savez_compressed(
    out, W=W, H=H, bg=np.array([0.1, 0.2, 0.3], np.float32), pos=sc.pos, scale=scale, rotq=sc.rotq, sh=sc.sh,
    opacity=sc.opacity, view_params=np.frombuffer(bytes(vp), np.uint8), num_rendered=fr.num_rendered, color=fr.color,
    means_2d=fr.means_2d, depth=fr.depth, conic=fr.conic, tiles_touched=fr.tiles_touched, radii=fr.radii,
    offsets=fr.offsets, keys_sorted=fr.keys_sorted, vals_sorted=fr.vals_sorted, ranges=fr.ranges, img=fr.img,
    n_examined=fr.n_examined)
print(out, os.path.getsize(out), "bytes; N =", fr.num_rendered, "visible", int((fr.depth >= 0.2).sum()),
      "saturated pixels", int((fr.n_examined < (fr.ranges[:, 1] - fr.ranges[:, 0]).max()).sum()))
