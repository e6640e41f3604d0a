"""A pin held by the reference itself: the frame geometry of its PUBLISHED render.

tests/golden/reference_doc_borders.npz holds per-column / per-row maxima of the border of
/root/reference/doc/mip360_bicycle_30000_cuda.png (made by tests/golden/make_reference_doc_fixture.py; the trained
.ply behind the picture is not in the repository, so pixels cannot be compared).  Whatever the scene, that picture
shows three properties of the reference's output path, and the oracle and the CUDA path must show the same ones on a
scene that covers the whole frame:
  Q1   the last tile column and the last tile row are never rendered (gs_tile_splatter/shader.cpp:102-163),
  flip the app writes the image upside down (app/main.cpp:322-337): the unrendered 7-pixel tile row of a 1063-pixel-high
       frame is at the TOP of the PNG,
  bg   the untouched pixels are exactly 0 and the bands next to them are rendered,
  255  saturated areas stop at 252 (one alpha-0.99 Gaussian of colour >= 1) and 254, never 255: colour clamp to 1,
       accumulated weight < 1, truncating uint8(v * 255).
"""
import os

import numpy as np
import pytest

from luisacomputegaussiansplatting_b200 import scenes
from oracle import oracle as orc

FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_doc_borders.npz"))
W, H = int(FIX["W"]), int(FIX["H"])


def band_geometry(rgb8):
    """(zero columns at the right, zero rows at the top, zero rows at the bottom, zero columns at the left) of an HWC image."""
    col = rgb8.max(axis=(0, 2))
    row = rgb8.max(axis=(1, 2))
    count = lambda v: int(np.argmax(v != 0)) if (v != 0).any() else len(v)  # noqa: E731
    return count(col[::-1]), count(row), count(row[::-1]), count(col)


def fixture_geometry():
    count = lambda v: int(np.argmax(v != 0)) if (v != 0).any() else len(v)  # noqa: E731
    return (count(FIX["col_max_right48"][::-1]), count(FIX["row_max_top24"]), count(FIX["row_max_bottom24"][::-1]),
            count(FIX["col_max_left24"]))


def covering_scene():
    sc, cfg = scenes.make_config_scene("C2", P=40_000)
    return sc, (sc.scale * np.float32(25.0)).astype(np.float32), cfg


def test_published_render_shows_q1_and_the_flip():
    assert (W, H) == (1600, 1063)
    # 1600 = 100 tiles exactly: the last tile column is 16 pixels; 1063 = 66 * 16 + 7: the last tile row is 7 pixels high
    assert fixture_geometry() == (16, H - 16 * ((H + 15) // 16 - 1), 0, 0) == (16, 7, 0, 0)


def test_published_renders_never_reach_255():
    for k in ("hist_top16_bicycle", "hist_top16_lego"):
        assert FIX[k][15] == 0 and FIX[k][:15].sum() > 0
    assert FIX["hist_top16_bicycle"][12] > 20_000 and FIX["hist_top16_bicycle"][13:].sum() == 0  # saturates at 252
    assert FIX["hist_top16_lego"][14] > 1_000                                                    # saturates at 254


def test_oracle_saturates_like_the_published_renders():
    # bright, opaque Gaussians stacked in front of the camera: every covered pixel saturates
    n = 64
    rng = np.random.default_rng(5)
    pos = (np.array(scenes.CAM_TARGET, np.float32) + rng.normal(0, 0.05, (n, 3))).astype(np.float32)
    scale = np.full((n, 3), 0.6, np.float32)
    rotq = np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1))
    sh = np.zeros((n, 16, 3), np.float32)
    sh[:, 0, :] = 10.0  # far above 1 after the +0.5: clamped to 1 by the SH stage
    cam = orc.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up("colmap"), 160, 96)
    one = orc.forward(pos[:1], scale[:1], rotq[:1], sh[:1], np.full(1, 0.999, np.float32), orc.view_params(cam))
    many = orc.forward(pos, scale, rotq, sh, np.full(n, 0.6, np.float32), orc.view_params(cam))
    assert orc.image_to_rgb8(one.img).max() == 252   # a single alpha-0.99 Gaussian: floor(255 * 0.99)
    assert orc.image_to_rgb8(many.img).max() == 254  # accumulated weight 1 - T with T >= 1e-4 -> floor(254.97)


def test_oracle_frame_has_the_published_geometry():
    sc, scale, cfg = covering_scene()
    cam = orc.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), W, H)
    fr = orc.forward(sc.pos, scale, sc.rotq, sc.sh, sc.opacity, orc.view_params(cam))
    assert band_geometry(orc.image_to_rgb8(fr.img)) == fixture_geometry()


@pytest.mark.gpu
def test_gpu_frame_has_the_published_geometry():
    from luisacomputegaussiansplatting_b200 import lcgs
    sc, scale, cfg = covering_scene()
    dev = lcgs.Device(0)
    r = lcgs.Renderer(dev, sc.pos, scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=30_000_000, rgb8=True)
    r.render(lcgs.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), W, H))
    assert band_geometry(r.rgb8.cpu().numpy().reshape(H, W, 3)) == fixture_geometry()
    dev.close()
