"""Runs the CUDA path's math header (csrc/lcgs_math.cuh) on the CPU and checks it against the oracle.

This validates the transcription of the per-Gaussian arithmetic (projection, EWA, conic/radius,
tile rect, SH colour, alpha threshold) without a GPU.  The mirror is test-only code: the product
library never executes these functions on the host.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from luisacomputegaussiansplatting_b200 import scenes
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_mirror", "host_mirror.cpp")
SO = os.path.join(HERE, "host_mirror", "libhost_mirror.so")
HDR = os.path.join(HERE, "..", "luisacomputegaussiansplatting_b200", "csrc", "lcgs_math.cuh")


@pytest.fixture(scope="module")
def hm():
    newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-mavx2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared",
                        "-fvisibility=hidden", "-x", "c++", SRC, "-o", SO], check=True)
    lib = C.CDLL(SO)
    lib.hm_exp.restype = C.c_float
    lib.hm_exp.argtypes = [C.c_float]
    lib.hm_alpha_threshold.restype = C.c_float
    lib.hm_alpha_threshold.argtypes = [C.c_float]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("key,P,W,H,band", [("C3", 6000, 480, 272, (0, -1)), ("C1", 3000, 200, 200, (0, -1)),
                                            ("C3", 3000, 333, 111, (2, 5))])
def test_device_math_matches_oracle(hm, key, P, W, H, band):
    sc, cfg = scenes.make_config_scene(key, P=P)
    cam = orc.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), W, H)
    vp = orc.view_params(cam)
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, vp, row0=band[0], row1=band[1])
    means = np.zeros((P, 2), np.float32)
    depth = np.zeros(P, np.float32)
    conic = np.zeros((P, 3), np.float32)
    color = np.zeros((P, 3), np.float32)
    radii = np.zeros(P, np.int32)
    tiles = np.zeros(P, np.uint32)
    thr = np.zeros(P, np.float32)
    hm.hm_preprocess(C.c_int(P), C.c_int(3), _p(sc.pos), _p(sc.scale), _p(sc.rotq), _p(sc.sh), _p(sc.opacity),
                     C.c_float(1.0), C.byref(vp), C.c_int(band[0]), C.c_int(band[1]), _p(means), _p(depth), _p(conic),
                     _p(color), _p(radii), _p(tiles), _p(thr))
    vis = fr.depth >= 0.2
    assert np.array_equal(depth.view(np.uint32), fr.depth.view(np.uint32))
    assert np.array_equal(radii, fr.radii)
    assert np.array_equal(tiles, fr.tiles_touched)
    assert np.array_equal(means.view(np.uint32), fr.means_2d.view(np.uint32))
    assert np.array_equal(conic.view(np.uint32), fr.conic.view(np.uint32))
    assert np.array_equal(color[vis].view(np.uint32), fr.color[vis].view(np.uint32))
    for i in np.nonzero(vis)[0][:400]:
        assert thr[i] == orc.alpha_threshold(float(sc.opacity[i]))


def test_exp_and_threshold_edge_cases(hm):
    for x in [0.0, -0.0, -1e-38, -1.0, -5.541, -87.0, -103.5, -104.5, -1e9, 3.0, 88.9, float("inf"), float("-inf")]:
        assert hm.hm_exp(x) == orc.exp(x)
    for op in [0.0, 1e-9, 0.0039, 0.003921569, 0.00392157, 0.01, 0.5, 0.99, 1.0, 2.0]:
        assert hm.hm_alpha_threshold(op) == orc.alpha_threshold(op)


def test_cull_rect_is_conservative(hm):
    """cull_rect may only drop a (Gaussian, rectangle) pair if no pixel in the rectangle passes."""
    W, H, P = 1920, 1080, 120_000
    sc, cfg = scenes.make_config_scene("C3", P=P)
    cam = orc.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), W, H)
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, orc.view_params(cam))
    n = fr.num_rendered
    gx = (W + 15) // 16
    tiles = (fr.keys_sorted >> np.uint64(32)).astype(np.int64)
    ids = fr.vals_sorted.astype(np.int64)
    rng = np.random.default_rng(0)
    thr_all = np.array([orc.alpha_threshold(float(o)) for o in sc.opacity[:P]], np.float32)
    stats = {}
    for name, (w, h) in {"tile16x16": (16, 16), "patch8x4": (8, 4)}.items():
        tx, ty = (tiles % gx) * 16, (tiles // gx) * 16
        ox = rng.integers(0, 16 // w, n) * w
        oy = rng.integers(0, 16 // h, n) * h
        rect = np.stack([tx + ox, ty + oy, tx + ox + w - 1, ty + oy + h - 1], axis=1).astype(np.int32)
        mean = np.ascontiguousarray(fr.means_2d[ids])
        conic = np.ascontiguousarray(fr.conic[ids])
        thr = np.ascontiguousarray(thr_all[ids])
        for fast in (0, 1):  # the general test and the hoisted-division variant the blend kernel runs
            out = np.zeros(n, np.uint8)
            hm.hm_cull_check(C.c_long(n), _p(mean), _p(conic), _p(thr), _p(rect), C.c_int(fast), _p(out))
            assert not np.any(out == 3), "cull_rect dropped a contributing pair (%s, fast=%d)" % (name, fast)
            dropped, empty = np.mean(out & 1), np.mean((out & 2) == 0)
            stats[name, fast] = (dropped, empty)
            assert dropped > 0.8 * empty - 0.01  # and it is tight: it finds most of the empty rectangles
    print(stats)
    # adversarial: extremely elongated / huge / tiny Gaussians around random rectangles
    m = 200_000
    ang = rng.uniform(0, np.pi, m)
    l1 = 10 ** rng.uniform(-0.5, 7, m)
    l2 = 10 ** rng.uniform(-0.5, 1, m)
    ca, sa = np.cos(ang), np.sin(ang)
    cxx, cxy, cyy = ca * ca * l1 + sa * sa * l2, ca * sa * (l1 - l2), sa * sa * l1 + ca * ca * l2
    det = cxx * cyy - cxy * cxy
    conic = np.stack([cyy / det, -cxy / det, cxx / det], axis=1).astype(np.float32)
    mean = rng.uniform(-3000, 5000, (m, 2)).astype(np.float32)
    x0 = rng.integers(0, 1900, m)
    y0 = rng.integers(0, 1060, m)
    rect = np.stack([x0, y0, x0 + rng.integers(0, 16, m), y0 + rng.integers(0, 16, m)], axis=1).astype(np.int32)
    thr = (-10 ** rng.uniform(-3, 0.75, m)).astype(np.float32)
    for fast in (0, 1):
        out = np.zeros(m, np.uint8)
        hm.hm_cull_check(C.c_long(m), _p(mean), _p(conic), _p(thr), _p(rect), C.c_int(fast), _p(out))
        assert not np.any(out == 3)
    # degenerate inputs are never culled by the fast variant: NaN threshold / mean, non-concave conic
    bad_mean = np.array([[np.nan, 5.0], [np.inf, 5.0], [5.0, 5.0], [5.0, 5.0], [5000.0, 5000.0]], np.float32)
    bad_conic = np.array([[1, 0, 1], [1, 0, 1], [-1, 0, 1], [1, 5, 1], [np.nan, 0, 1]], np.float32)
    bad_thr = np.array([-1, -1, -1, -1, -1], np.float32)
    bad_rect = np.tile(np.array([[100, 100, 107, 103]], np.int32), (5, 1))
    out = np.zeros(5, np.uint8)
    hm.hm_cull_check(C.c_long(5), _p(bad_mean), _p(bad_conic), _p(bad_thr), _p(bad_rect), C.c_int(1), _p(out))
    assert not np.any(out & 1)
