"""Cross-checks the C oracle against an independent numpy restatement and hand-computed cases.

The reference has no tests for these stages ("parity unpinned"); two independently
written restatements agreeing bit for bit is the strongest pin available offline.
"""
import math

import numpy as np
import pytest

import np_restatement as npr
from luisacomputegaussiansplatting_b200 import scenes
from oracle import oracle as orc


def _scene(P=4000, W=320, H=200, key="C3"):
    sc, cfg = scenes.make_config_scene(key, P=P)
    cam = orc.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), W, H)
    return sc, orc.view_params(cam)


@pytest.fixture(scope="module")
def frame():
    sc, vp = _scene()
    return sc, vp, orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, vp)


def test_sh_matches_numpy(frame):
    sc, vp, fr = frame
    ref = npr.sh_color(sc.pos, sc.sh, list(vp.cam_pos))
    assert np.array_equal(ref.view(np.uint32), fr.color.view(np.uint32))
    assert fr.color.min() >= 0.0 and fr.color.max() <= 1.0  # Q3: clamp to [0,1]


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_degrees(deg):
    sc, vp = _scene(P=500)
    feat = (deg + 1) ** 2
    sh = np.ascontiguousarray(sc.sh[:, :feat, :])
    got = orc.sh_process(sc.pos, sh, list(vp.cam_pos), deg)
    ref = npr.sh_color(sc.pos, sh, list(vp.cam_pos), deg)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))


def test_sh_closed_form_dc_only():
    # a Gaussian whose higher bands are zero: colour = clamp(0.2820948*dc + 0.5)
    pos = np.array([[0.0, 0.0, 1.0]], np.float32)
    sh = np.zeros((1, 16, 3), np.float32)
    sh[0, 0] = [1.0, -3.0, 3.0]
    c = orc.sh_process(pos, sh, [0.0, 0.0, 0.0])
    assert np.allclose(c[0], [0.5 + 0.28209479, 0.0, 1.0], atol=1e-6)


def test_projection_and_tiles_match_numpy(frame):
    sc, vp, fr = frame
    ndc, depth, cov, vis = npr.project(sc.pos, sc.scale, sc.rotq, list(vp.view), list(vp.proj), vp.tanfovx,
                                       vp.tanfovy, vp.focalx, vp.focaly)
    o_ndc, o_depth, o_cov = orc.project(sc.pos, sc.scale, sc.rotq, vp)
    assert np.array_equal(depth.view(np.uint32), o_depth.view(np.uint32))
    assert np.array_equal(ndc.view(np.uint32), o_ndc.view(np.uint32))
    assert np.array_equal(cov.view(np.uint32), o_cov.view(np.uint32))
    assert np.array_equal(vis, o_depth >= 0.2)
    pix, conic, tiles, radii = npr.allocate_tiles(vp.width, vp.height, depth, ndc, cov)
    assert np.array_equal(radii, fr.radii)
    assert np.array_equal(tiles, fr.tiles_touched)
    assert np.array_equal(pix.view(np.uint32), fr.means_2d.view(np.uint32))
    assert np.array_equal(conic.view(np.uint32), fr.conic.view(np.uint32))
    assert 0 < vis.sum() < sc.num_gaussians  # the scene exercises both sides of the near cull


def test_keys_sort_ranges_match_numpy(frame):
    sc, vp, fr = frame
    assert fr.num_rendered == int(fr.tiles_touched.sum()) == int(fr.offsets[-1])
    assert np.array_equal(fr.offsets, np.cumsum(fr.tiles_touched, dtype=np.uint64).astype(np.uint32))
    keys, vals = npr.keys_and_values(vp.width, vp.height, fr.means_2d, fr.radii, fr.depth, fr.tiles_touched)
    assert np.array_equal(keys, fr.keys_unsorted) and np.array_equal(vals, fr.vals_unsorted)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(keys[order], fr.keys_sorted) and np.array_equal(vals[order], fr.vals_sorted)
    gx, gy = orc.grids(vp.width, vp.height)
    assert np.array_equal(npr.tile_ranges(fr.keys_sorted, gx * gy), fr.ranges)
    # Q1: the last tile column and row never receive a Gaussian
    tiles = (fr.keys_sorted >> np.uint64(32)).astype(np.int64)
    assert (tiles % gx).max() < gx - 1 and (tiles // gx).max() < gy - 1


def test_blend_matches_python_pixels(frame):
    sc, vp, fr = frame
    W, H = vp.width, vp.height
    gx, _ = orc.grids(W, H)
    rng = np.random.default_rng(5)
    # pixels from the busiest tiles plus random ones
    lens = fr.ranges[:, 1].astype(np.int64) - fr.ranges[:, 0]
    busy = np.argsort(lens)[-6:]
    pts = [(int(t % gx) * 16 + int(rng.integers(16)), int(t // gx) * 16 + int(rng.integers(16))) for t in busy]
    pts += [(int(rng.integers(W)), int(rng.integers(H))) for _ in range(40)]
    checked = 0
    for px, py in pts:
        if px >= W or py >= H:
            continue
        t = (px // 16) + (py // 16) * gx
        rgb, ex = npr.blend_pixel(px, py, fr.ranges[t, 0], fr.ranges[t, 1], fr.vals_sorted, fr.means_2d, fr.conic,
                                  sc.opacity, fr.color, (0.0, 0.0, 0.0))
        assert ex == fr.n_examined[py, px]
        assert np.allclose(rgb, fr.img[:, py, px], atol=2e-6)
        checked += 1
    assert checked >= 30
    # Q1 again, at the image level: last tile column/row are pure background
    assert not fr.img[:, :, (gx - 1) * 16:].any()


def test_quirk_pixel_centres_have_no_half_offset():
    # Q2: ndc -1 maps to pixel -0.5; a Gaussian projecting to ndc (0,0) lands at ((W-1)/2,(H-1)/2)
    depth = np.array([1.0], np.float32)
    ndc = np.array([[0.0, 0.0]], np.float32)
    cov = np.array([[4.0, 0.0, 4.0]], np.float32)
    pix, conic, tiles, radii = orc.allocate_tiles(64, 48, depth, ndc, cov)
    assert pix[0, 0] == np.float32(31.5) and pix[0, 1] == np.float32(23.5)
    # Q5: low-pass +0.3, det+1e-6: conic.x = c/(a*c-b*b+1e-6) with a=c=4.3
    a = np.float32(4.3)
    det = a * a
    assert conic[0, 0] == np.float32(1.0) / (det + np.float32(1e-6)) * a
    # radius = ceil(3*sqrt(lambda_max)), lambda = mid + sqrt(max(0.1, mid^2-det)) = 4.3 + sqrt(0.1)
    lam = np.float32(4.3) + np.sqrt(np.float32(0.1))
    assert radii[0] == int(math.ceil(3.0 * math.sqrt(float(lam))))
    # rect: min = uint((31.5-7)/16)=1, max = uint(31.5+7+16-1)/16 = 53/16 = 3, clamped to grids-1 = 3
    #       y: min = uint((23.5-7)/16)=1, max = uint(45.5)/16 = 2 (grids.y-1 = 2)  -> (3-1)*(2-1) = 2
    assert tiles[0] == 2


def test_quirk_offscreen_and_saturating_casts():
    # Q4: no frustum cull besides z<0.2; a Gaussian far left of the screen gets radius>0, tiles==0,
    # and the negative float->uint conversion saturates to 0 instead of wrapping.
    depth = np.array([1.0, 0.1, 5.0], np.float32)
    ndc = np.array([[-50.0, 0.0], [0.0, 0.0], [1e30, -1e30]], np.float32)
    cov = np.array([[1.0, 0.0, 1.0]] * 3, np.float32)
    pix, conic, tiles, radii = orc.allocate_tiles(640, 480, depth, ndc, cov)
    assert radii[0] > 0 and tiles[0] == 0
    assert radii[1] == 0 and tiles[1] == 0          # near-culled
    assert tiles[2] == 0 and radii[2] > 0           # +huge x saturates to 2^32-1 -> clamped to grids-1


def test_near_plane_gaussian_covers_every_binnable_tile():
    # Q4: huge covariance -> rect spans [0, grids-1) in both axes
    depth = np.array([0.25], np.float32)
    ndc = np.array([[0.0, 0.0]], np.float32)
    cov = np.array([[1e8, 0.0, 1e8]], np.float32)
    _, _, tiles, radii = orc.allocate_tiles(640, 480, depth, ndc, cov)
    assert tiles[0] == (40 - 1) * (30 - 1) and radii[0] == 30000


def test_empty_frame_leaves_image_untouched():
    # Q10: num_rendered == 0 -> forward returns 0 and never clears the image (impl.cpp:109)
    sc, vp = _scene(P=64)
    pos = sc.pos.copy()
    pos[:] = np.array(scenes.CAM_POS, np.float32) - 10.0 * np.array(list(vp.view)[2:12:4], np.float32)  # behind
    img = np.full((3, vp.height, vp.width), 0.25, np.float32)
    fr = orc.forward(pos, sc.scale, sc.rotq, sc.sh, sc.opacity, vp, img=img, capacity=16)
    assert fr.num_rendered == 0 and np.all(fr.img == 0.25)


def test_exp_is_correctly_rounded_and_monotone():
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-12.0, 0.0, 20000), rng.uniform(-104.0, 89.0, 2000), [0.0, -0.0, -1e-30]])
    bad = 0
    for x in xs.astype(np.float32):
        got = orc.exp(float(x))
        want = np.float32(math.exp(float(x)))
        bad += got != float(want)
    assert bad == 0
    assert orc.exp(-200.0) == 0.0 and orc.exp(100.0) == math.inf and math.isnan(orc.exp(math.nan))
    # exhaustive over every float in [-6, -2^-10]: the range the alpha threshold lives in
    assert orc.lib().orc_exp_monotonicity_violations(-6.0, -0.0009765625) == 0


def test_alpha_threshold_is_the_decision_boundary():
    rng = np.random.default_rng(2)
    inv255 = np.float32(1.0) / np.float32(255.0)
    for op in np.concatenate([rng.uniform(0.0, 1.0, 300), [1.0, 0.99, 1 / 255, 0.003921569, 0.0039, 0.5, 1e-3]]):
        op = np.float32(op)
        thr = orc.alpha_threshold(float(op))
        if op < inv255:
            assert thr == math.inf
            continue
        assert thr <= 0.0
        a_at = min(np.float32(0.99), np.float32(op * np.float32(orc.exp(thr))))
        assert not a_at < inv255
        if thr > -100.0:
            below = np.nextafter(np.float32(thr), np.float32(-np.inf))
            a_below = min(np.float32(0.99), np.float32(op * np.float32(orc.exp(float(below)))))
            assert a_below < inv255
            assert abs(thr - math.log(1.0 / (255.0 * float(op)))) < 1e-5 + 1e-6 * abs(thr)


def test_sort_is_stable_and_handles_edge_sizes():
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 255, 256, 257, 10007):
        keys = (rng.integers(0, 5, n).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 3, n).astype(np.uint64)
        vals = np.arange(n, dtype=np.uint32)
        ko, vo = orc.sort_pairs(keys, vals)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])
    # full 64-bit keys
    keys = rng.integers(0, 2**63, 5000, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, 5000).astype(np.uint64)
    ko, vo = orc.sort_pairs(keys, np.arange(5000, dtype=np.uint32))
    assert np.array_equal(ko, np.sort(keys))


def test_ranges_of_single_tile_and_gaps():
    keys = (np.array([2, 2, 2, 5, 7, 7], np.uint64) << np.uint64(32)) | np.uint64(1)
    r = orc.get_ranges(keys, 9)
    assert r[2].tolist() == [0, 3] and r[5].tolist() == [3, 4] and r[7].tolist() == [4, 6]
    assert not r[[0, 1, 3, 4, 6, 8]].any()
    one = orc.get_ranges(keys[:1], 4)
    assert one[2].tolist() == [0, 1]


def test_tile_row_bands_partition_the_frame(frame):
    """Extension for tile-row sharding: band renders tile exactly the single-device frame."""
    sc, vp, fr = frame
    gx, gy = orc.grids(vp.width, vp.height)
    cuts = [0, 3, 4, gy]
    total = 0
    img = np.zeros_like(fr.img)
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        band = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, vp, row0=r0, row1=r1)
        total += band.num_rendered
        y0, y1 = r0 * 16, min(vp.height, r1 * 16)
        img[:, y0:y1] = band.img[:, y0:y1]
        # band-local ranges equal the global ones shifted by the band's first instance
        g = fr.ranges[r0 * gx:r1 * gx].astype(np.int64)
        b = band.ranges.astype(np.int64)
        nz = g[:, 1] > g[:, 0]
        assert np.array_equal(nz, b[:, 1] > b[:, 0])
        assert np.array_equal((g[nz, 1] - g[nz, 0]), (b[nz, 1] - b[nz, 0]))
    assert total == fr.num_rendered
    assert np.array_equal(img.view(np.uint32), fr.img.view(np.uint32))


def test_image_to_rgb8_flips_and_truncates():
    img = np.zeros((3, 2, 2), np.float32)
    img[0, 0, 0] = 0.999   # bottom row after the flip (main.cpp:331)
    img[1, 1, 1] = 1.0
    out = orc.image_to_rgb8(img)
    assert out[1, 0, 0] == 254 and out[0, 1, 1] == 255 and out.sum() == 254 + 255
