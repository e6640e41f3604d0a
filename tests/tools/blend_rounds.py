"""TEST-ONLY analysis tool: work of the two-pixel blend schedule as a function of the round size (tests/tools/blend_rounds.cpp).

    python tests/tools/blend_rounds.py [C3|C2|C1] [P]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from luisacomputegaussiansplatting_b200 import scenes  # noqa: E402
from oracle import oracle as orc  # noqa: E402

SRC, SO = os.path.join(HERE, "blend_rounds.cpp"), os.path.join(HERE, "libblend_rounds.so")
key = sys.argv[1] if len(sys.argv) > 1 else "C3"
P = int(sys.argv[2]) if len(sys.argv) > 2 else None
subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-x", "c++",
                SRC, "-o", SO], check=True)
lib = C.CDLL(SO)
sc, cfg = scenes.make_config_scene(key, P=P)
vp = orc.view_params(orc.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), cfg.W, cfg.H))
fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, vp)
uo = np.unique(sc.opacity)
thr = np.array([orc.alpha_threshold(float(o)) for o in uo], np.float32)[np.searchsorted(uo, sc.opacity)].astype(np.float32)
Rs = np.array([32, 64, 128, 256], np.int32)
out = np.zeros((len(Rs), 5), np.uint64)
p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
lib.br_run(C.c_int(cfg.W), C.c_int(cfg.H), p(np.ascontiguousarray(fr.ranges, np.uint32)), p(np.ascontiguousarray(fr.vals_sorted, np.uint32)),
           p(np.ascontiguousarray(fr.means_2d, np.float32)), p(np.ascontiguousarray(fr.conic, np.float32)),
           p(np.ascontiguousarray(sc.opacity, np.float32)), p(thr), C.c_int(len(Rs)), p(Rs), p(out))
print("round size   rounds  warp-rounds  segment walks   hit evaluations   candidates staged")
for R, row in zip(Rs, out):
    print("%9d %9d %11d %13d %17d %18d" % (R, *[int(x) for x in row]))
