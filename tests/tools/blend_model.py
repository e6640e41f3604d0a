"""TEST-ONLY analysis tool: counts the work of the blend kernel's scheduling variants on an oracle frame.

    python tests/tools/blend_model.py [C3|C2|C1] [P]

Replays csrc/blend.cu's round / segment / patch structure on the CPU (tests/tools/blend_model.cpp includes the
device math header) and prints how many segment walks and hit evaluations each variant would execute.
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

SRC = os.path.join(HERE, "blend_model.cpp")
SO = os.path.join(HERE, "libblend_model.so")
NAMES = ["rounds", "warp_rounds", "seg_walks", "hits", "lane_ok", "lane_blend", "tile_survivors", "instances_seen",
         "seg_walks_dense", "hits_half_lr", "hits_half_tb", "hits_quarter", "hits_rows", "hits_8x8", "seg_walks_8x8",
         "hits_exact", "warp_rounds_8x8", "hits_16x4", "seg_walks_16x4", "warp_rounds_16x4",
         "hits_8x8_exact", "hits_16x4_exact"]


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "C3"
    P = int(sys.argv[2]) if len(sys.argv) > 2 else None
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
                    "-x", "c++", SRC, "-o", SO], check=True)
    lib = C.CDLL(SO)
    sc, cfg = scenes.make_config_scene(key, P=P)
    pose = (scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world))
    vp = orc.view_params(orc.make_camera(*pose, cfg.W, cfg.H))
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, vp)
    thr = np.array([orc.alpha_threshold(float(o)) for o in np.unique(sc.opacity)], np.float32)
    thr = thr[np.searchsorted(np.unique(sc.opacity), sc.opacity)].astype(np.float32)
    out = (C.c_ulonglong * len(NAMES))()
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    ranges = np.ascontiguousarray(fr.ranges, np.uint32)
    pl = np.ascontiguousarray(fr.vals_sorted, np.uint32)
    means = np.ascontiguousarray(fr.means_2d, np.float32)
    conic = np.ascontiguousarray(fr.conic, np.float32)
    op = np.ascontiguousarray(sc.opacity, np.float32)
    lib.bm_run(C.c_int(cfg.W), C.c_int(cfg.H), p(ranges), p(pl), p(means), p(conic), p(op), p(thr), out)
    d = dict(zip(NAMES, [int(x) for x in out]))
    print("N", fr.num_rendered, orc.blend_stats())
    for k, v in d.items():
        print("%-18s %14d" % (k, v))
    walk, hit = 40, 32
    print("model instr: walk %.1f M + hits %.1f M" % (d["seg_walks"] * walk / 1e6, d["hits"] * hit / 1e6))


if __name__ == "__main__":
    main()
