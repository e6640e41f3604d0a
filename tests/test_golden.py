"""Committed golden vectors (tests/golden/tiny_frame.npz, made by tests/golden/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as orc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_frame.npz"))
KEYS = ["color", "means_2d", "depth", "conic", "tiles_touched", "radii", "offsets", "keys_sorted", "vals_sorted",
        "ranges", "img", "n_examined"]


def _vp():
    vp = orc.ViewParams()
    C.memmove(C.byref(vp), G["view_params"].tobytes(), C.sizeof(vp))
    return vp


def test_oracle_reproduces_golden_frame():
    fr = orc.forward(G["pos"], G["scale"], G["rotq"], G["sh"], G["opacity"], _vp(), bg=G["bg"])
    assert fr.num_rendered == int(G["num_rendered"])
    for k in KEYS:
        assert np.array_equal(getattr(fr, k).view(np.uint8), G[k].view(np.uint8)), k


@pytest.mark.gpu
def test_gpu_reproduces_golden_frame():
    from luisacomputegaussiansplatting_b200 import lcgs
    W, H = int(G["W"]), int(G["H"])
    dev = lcgs.Device(0)
    vp = lcgs.ViewParams()
    C.memmove(C.byref(vp), G["view_params"].tobytes(), C.sizeof(vp))
    r = lcgs.Renderer(dev, G["pos"], G["scale"], G["rotq"], G["sh"], G["opacity"], W, H, list_capacity=100_000,
                      bg_color=[float(x) for x in G["bg"]])
    r.render_async(vp)
    n = dev.num_rendered()
    g = r.intermediates(n)
    assert n == int(G["num_rendered"])
    for k in ["means_2d", "depth", "conic", "tiles_touched", "radii", "offsets", "keys_sorted", "vals_sorted", "ranges"]:
        assert np.array_equal(np.ascontiguousarray(g[k]).view(np.uint8), G[k].view(np.uint8)), k
    sel = G["tiles_touched"] > 0
    assert np.array_equal(g["color"][sel].view(np.uint32), G["color"][sel].view(np.uint32))
    assert float(np.abs(g["img"] - G["img"]).max()) <= 2e-3
    dev.close()
