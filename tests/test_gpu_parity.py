"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): visibility masks, tile counts, sorted keys and tile ranges
bit-exact; rendered RGB within max-abs 2e-3 per channel and PSNR >= 50 dB.  We additionally hold
depth / pixel means / conic / colour / radii / offsets / unsorted lists / sorted values bit-exact.
"""
import math

import numpy as np
import pytest

from luisacomputegaussiansplatting_b200 import scenes
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

MAX_ABS = 2e-3
MIN_PSNR = 50.0


@pytest.fixture(scope="module")
def lcgs():
    from luisacomputegaussiansplatting_b200 import lcgs as m
    return m


@pytest.fixture(scope="module")
def dev(lcgs):
    d = lcgs.Device(0)
    yield d
    d.close()


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return math.inf if mse == 0 else 10.0 * math.log10(1.0 / mse)


def assert_image_close(got, want):
    err = float(np.abs(got - want).max())
    p = psnr(got, want)
    assert err <= MAX_ABS, "max-abs %.3g" % err
    assert p >= MIN_PSNR, "psnr %.2f" % p
    return err, p


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_frame_matches(gpu: dict, fr: "orc.Frame", fused: bool):
    n = fr.num_rendered
    assert gpu["num_rendered"] == n
    assert np.array_equal(bits(gpu["depth"]) , bits(fr.depth)), "depth / visibility mask"
    assert np.array_equal(gpu["depth"] >= 0.2, fr.depth >= 0.2)
    assert np.array_equal(gpu["radii"], fr.radii), "radii"
    assert np.array_equal(gpu["tiles_touched"], fr.tiles_touched), "tile counts"
    assert np.array_equal(bits(gpu["means_2d"]), bits(fr.means_2d)), "pixel means"
    assert np.array_equal(bits(gpu["conic"]), bits(fr.conic)), "conic"
    sel = fr.tiles_touched > 0 if fused else np.ones_like(fr.tiles_touched, bool)
    assert np.array_equal(bits(gpu["color"][sel]), bits(fr.color[sel])), "colour"
    assert np.array_equal(gpu["offsets"], fr.offsets), "inclusive sum"
    if fused:
        # the fused frame emits the instances in depth order (same pairs, different order)
        go = np.lexsort((gpu["vals_unsorted"], gpu["keys_unsorted"]))
        fo = np.lexsort((fr.vals_unsorted, fr.keys_unsorted))
        assert np.array_equal(gpu["keys_unsorted"][go], fr.keys_unsorted[fo]), "unsorted keys (as a multiset)"
        assert np.array_equal(gpu["vals_unsorted"][go], fr.vals_unsorted[fo]), "unsorted values (as a multiset)"
    else:
        assert np.array_equal(gpu["keys_unsorted"], fr.keys_unsorted), "unsorted keys"
        assert np.array_equal(gpu["vals_unsorted"], fr.vals_unsorted), "unsorted values"
    assert np.array_equal(gpu["keys_sorted"], fr.keys_sorted), "sorted keys"
    assert np.array_equal(gpu["vals_sorted"], fr.vals_sorted), "sorted values (stability)"
    assert np.array_equal(gpu["ranges"], fr.ranges), "tile ranges"
    return assert_image_close(gpu["img"], fr.img)


def make_case(key, P, W, H, pose=None):
    sc, cfg = scenes.make_config_scene(key, P=P)
    if pose is None:
        pose = (scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world))
    return sc, pose


CASES = [
    ("C3", 20_000, 640, 360),     # both dims multiples of 8 but H not of 16
    ("C1", 10_000, 200, 200),     # blender world, ragged tiles
    ("C2", 50_000, 1237, 822),    # the bicycle resolution (odd sizes)
    ("C3", 3_000, 97, 61),        # tiny ragged frame
    ("C3", 1, 64, 64),            # single Gaussian
    ("C3", 4097, 320, 240),       # one more than a scan tile
]


@pytest.mark.parametrize("key,P,W,H", CASES)
def test_fused_render_matches_oracle(lcgs, dev, key, P, W, H):
    sc, pose = make_case(key, P, W, H)
    cam = lcgs.make_camera(*pose, W, H)
    vp = lcgs.view_params(cam)
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, orc.view_params(orc.make_camera(*pose, W, H)))
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=max(fr.num_rendered, 1) + 17)
    n = r.render(cam)
    assert bytes(vp) == bytes(orc.view_params(orc.make_camera(*pose, W, H)))
    assert_frame_matches(r.intermediates(n), fr, fused=True)
    # a second frame through the same context/buffers gives the same bits (no stale state)
    img1 = r.image().cpu().numpy().copy()
    assert r.render(cam) == n
    assert np.array_equal(bits(r.image().cpu().numpy()), bits(img1))


def test_reference_style_api_matches_oracle(lcgs, dev):
    """SHProcessor.process + GSProjector.forward + GSTileSplatter.forward, as app/main.cpp:266-299."""
    import torch
    W, H, P = 512, 288, 30_000
    sc, pose = make_case("C3", P, W, H)
    cam = lcgs.make_camera(*pose, W, H)
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, orc.view_params(orc.make_camera(*pose, W, H)))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    d_pos, d_scale, d_rotq, d_sh, d_opacity = t(sc.pos), t(sc.scale), t(sc.rotq), t(sc.sh), t(sc.opacity)
    d_color = dev.create_buffer(torch.float32, 3 * P)
    d_means, d_depth, d_cov = (dev.create_buffer(torch.float32, k * P) for k in (2, 1, 3))
    L = fr.num_rendered + 100
    gx, gy = (W + 15) // 16, (H + 15) // 16
    accel = lcgs.GSTileSplatterAccelProxy(
        dev.create_buffer(torch.int32, P), dev.create_buffer(torch.int32, P), dev.create_buffer(torch.int64, L),
        dev.create_buffer(torch.int32, L), dev.create_buffer(torch.int64, L), dev.create_buffer(torch.int32, L),
        dev.create_buffer(torch.int32, 2 * gx * gy))
    d_img = dev.create_buffer(torch.float32, 3 * W * H)
    d_radii = dev.create_buffer(torch.int32, P)

    shp, proj, spl = lcgs.SHProcessor(), lcgs.GSProjector(), lcgs.GSTileSplatter()
    shp.create(dev), proj.create(dev), spl.create(dev)
    bf, scan, sort = lcgs.BufferFiller(), lcgs.DeviceScan(), lcgs.DeviceRadixSort()
    scan.create(dev), sort.create(dev)
    spl.set_buffer_filler(bf), spl.set_device_scan(scan), spl.set_device_radix_sort(sort)

    shp.process(None, lcgs.GPUPointsProxy(P, 3, d_pos), cam, d_sh, d_color, 3, 3)
    proj.forward(None, lcgs.GSProjectorInputProxy(P, d_pos, d_scale, d_rotq, 1.0),
                 lcgs.GSProjectorOutputProxy(d_means, d_cov, d_depth), cam)
    # after K1/K2: colours for every Gaussian, NDC means and pixel^2 covariances
    o_ndc, o_depth, o_cov = orc.project(sc.pos, sc.scale, sc.rotq, orc.view_params(orc.make_camera(*pose, W, H)))
    assert np.array_equal(bits(d_color.cpu().numpy().reshape(-1, 3)), bits(fr.color))
    assert np.array_equal(bits(d_means.cpu().numpy().reshape(-1, 2)), bits(o_ndc))
    assert np.array_equal(bits(d_cov.cpu().numpy().reshape(-1, 3)), bits(o_cov))
    assert np.array_equal(bits(d_depth.cpu().numpy()), bits(o_depth))

    inp = lcgs.GSTileSplatterInputProxy(P, (0.0, 0.0, 0.0), d_means, d_depth, d_cov, d_color, d_opacity)
    out = lcgs.GSSplatForwardOutputProxy(H, W, d_img, d_radii)
    n = spl.forward(dev, None, accel, inp, out)
    u32 = lambda x: x.cpu().numpy().view(np.uint32)  # noqa: E731
    gpu = dict(num_rendered=n, depth=d_depth.cpu().numpy(), means_2d=d_means.cpu().numpy().reshape(-1, 2),
               conic=d_cov.cpu().numpy().reshape(-1, 3), color=d_color.cpu().numpy().reshape(-1, 3),
               tiles_touched=u32(accel.tiles_touched), radii=d_radii.cpu().numpy(), offsets=u32(accel.point_offsets),
               keys_unsorted=accel.point_list_keys_unsorted[:n].cpu().numpy().view(np.uint64),
               vals_unsorted=u32(accel.point_list_unsorted[:n]),
               keys_sorted=accel.point_list_keys[:n].cpu().numpy().view(np.uint64),
               vals_sorted=u32(accel.point_list[:n]), ranges=u32(accel.ranges).reshape(-1, 2),
               img=d_img.cpu().numpy().reshape(3, H, W))
    assert_frame_matches(gpu, fr, fused=False)


@pytest.mark.parametrize("deg", [0, 1, 2])
def test_lower_sh_degrees(lcgs, dev, deg):
    import torch
    sc, pose = make_case("C3", 5000, 64, 64)
    cam = lcgs.make_camera(*pose, 64, 64)
    sh = np.ascontiguousarray(sc.sh[:, :(deg + 1) ** 2, :])
    want = orc.sh_process(sc.pos, sh, list(cam.position), deg)
    d_color = dev.create_buffer(torch.float32, 3 * 5000)
    p = lcgs.SHProcessor()
    p.create(dev)
    p.process(None, lcgs.GPUPointsProxy(5000, 3, torch.from_numpy(sc.pos).cuda()), cam, torch.from_numpy(sh).cuda(),
              d_color, deg, 3)
    assert np.array_equal(bits(d_color.cpu().numpy().reshape(-1, 3)), bits(want))


def test_tile_row_bands_reassemble_the_frame(lcgs, dev):
    W, H, P = 640, 360, 20_000
    sc, pose = make_case("C3", P, W, H)
    cam = lcgs.make_camera(*pose, W, H)
    ovp = orc.view_params(orc.make_camera(*pose, W, H))
    full = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, ovp)
    gy = (H + 15) // 16
    img = np.zeros_like(full.img)
    total = 0
    for r0, r1 in [(0, 7), (7, 8), (8, gy)]:
        band = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, ovp, row0=r0, row1=r1)
        r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H,
                          list_capacity=max(band.num_rendered, 1), tile_rows=(r0, r1))
        n = r.render(cam)
        assert_frame_matches(r.intermediates(n), band, fused=True)
        total += n
        img[:, r0 * 16:min(H, r1 * 16)] = r.image().cpu().numpy()[:, r0 * 16:min(H, r1 * 16)]
    assert total == full.num_rendered
    assert_image_close(img, full.img)


def test_empty_frame_and_capacity_overflow(lcgs, dev):
    W, H, P = 128, 96, 2000
    sc, pose = make_case("C3", P, W, H)
    cam = lcgs.make_camera(*pose, W, H)
    # everything behind the camera: num_rendered == 0 and the image is left untouched (Q10)
    behind = sc.pos.copy()
    behind[:] = np.array(scenes.CAM_POS, np.float32) - 5.0 * np.array(list(cam.front), np.float32)
    r = lcgs.Renderer(dev, behind, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=1000)
    r.img.fill_(0.25)
    assert r.render(cam) == 0
    assert bool((r.img == 0.25).all())
    # overflow is reported, not undefined behaviour (reference: unchecked, main.cpp:245)
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, orc.view_params(orc.make_camera(*pose, W, H)))
    assert fr.num_rendered > 64
    r2 = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=64)
    with pytest.raises(lcgs.CapacityError):
        r2.render(cam)
    # and the context stays usable
    r3 = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=fr.num_rendered)
    assert r3.render(cam) == fr.num_rendered
    assert_image_close(r3.image().cpu().numpy(), fr.img)


@pytest.mark.parametrize("n", [0, 1, 3, 4095, 4096, 4097, 100_003, 3_000_000])
def test_scan_primitive(lcgs, dev, n):
    import torch
    rng = np.random.default_rng(n)
    x = rng.integers(0, 50, size=n, dtype=np.uint32)
    d_in = torch.from_numpy(x.view(np.int32)).cuda()
    d_out = torch.zeros(n, dtype=torch.int32, device="cuda")
    s = lcgs.DeviceScan()
    s.create(dev)
    s.InclusiveSum(None, d_in, d_out, n)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), np.cumsum(x, dtype=np.uint64).astype(np.uint32))


def test_scan_wraps_like_uint32(lcgs, dev):
    import torch
    x = np.full(10_000, 0x7FFFFFF, np.uint32)
    d_in = torch.from_numpy(x.view(np.int32)).cuda()
    d_out = torch.zeros_like(d_in)
    s = lcgs.DeviceScan()
    s.create(dev)
    s.InclusiveSum(None, d_in[1:], d_out[1:], 9_999)  # also an unaligned (4-byte offset) view
    want = np.cumsum(x[1:], dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(d_out[1:].cpu().numpy().view(np.uint32), want)


@pytest.mark.parametrize("n,bits_range,key_bits", [(0, (0, 64), 64), (1, (0, 64), 64), (255, (0, 64), 64),
                                                   (4096, (0, 64), 64), (4097, (0, 64), 64), (100_003, (0, 45), 45),
                                                   (1_000_000, (0, 64), 64), (2_000_000, (0, 44), 44),
                                                   (300_000, (8, 40), 64), (50_000, (0, 3), 64)])
def test_sort_primitive(lcgs, dev, n, bits_range, key_bits):
    import torch
    rng = np.random.default_rng(n + 1)
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64)
    if key_bits < 64:
        keys &= np.uint64((1 << key_bits) - 1)
    if n > 10:  # plenty of duplicate keys to exercise stability
        keys[rng.integers(0, n, n // 3)] = keys[rng.integers(0, n, n // 3)]
    vals = np.arange(n, dtype=np.uint32)
    d_k = torch.from_numpy(keys.view(np.int64)).cuda()
    d_v = torch.from_numpy(vals.view(np.int32)).cuda()
    d_ko, d_vo = torch.zeros_like(d_k), torch.zeros_like(d_v)
    s = lcgs.DeviceRadixSort()
    s.create(dev)
    s.SortPairs(None, d_k, d_ko, d_v, d_vo, n, *bits_range)
    b, e = bits_range
    mask = np.uint64(((1 << (e - b)) - 1) << b) if e - b < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    assert np.array_equal(d_ko.cpu().numpy().view(np.uint64), keys[order])
    assert np.array_equal(d_vo.cpu().numpy().view(np.uint32), vals[order])
    # inputs are preserved
    assert np.array_equal(d_k.cpu().numpy().view(np.uint64), keys)


def test_sorted_input_and_all_equal_keys(lcgs, dev):
    import torch
    s = lcgs.DeviceRadixSort()
    s.create(dev)
    for keys in (np.arange(70_000, dtype=np.uint64) << np.uint64(20), np.full(70_000, 12345, np.uint64),
                 (np.arange(70_000, dtype=np.uint64)[::-1] << np.uint64(33)).copy()):
        vals = np.arange(keys.size, dtype=np.uint32)
        d_k, d_v = torch.from_numpy(keys.view(np.int64)).cuda(), torch.from_numpy(vals.view(np.int32)).cuda()
        d_ko, d_vo = torch.zeros_like(d_k), torch.zeros_like(d_v)
        s.SortPairs(None, d_k, d_ko, d_v, d_vo, keys.size)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(d_ko.cpu().numpy().view(np.uint64), keys[order])
        assert np.array_equal(d_vo.cpu().numpy().view(np.uint32), vals[order])


def test_buffer_filler(lcgs, dev):
    import torch
    bf = lcgs.BufferFiller()
    a = torch.zeros(100_001, dtype=torch.int32, device="cuda")
    b = torch.zeros(33, dtype=torch.int64, device="cuda")
    c = torch.zeros(7, dtype=torch.float32, device="cuda")
    bf.fill(dev, a, 7), bf.fill(dev, b, 1 << 40), bf.fill(dev, c, 0.5)
    assert bool((a == 7).all()) and bool((b == (1 << 40)).all()) and bool((c == 0.5).all())


def test_frame_is_capturable_in_a_cuda_graph(lcgs, dev):
    """No allocation and no synchronisation inside lcgs_b200_render: the frame replays from a CUDA graph
     and gives the same bits as the eager call."""
    import torch
    W, H, P = 400, 240, 12_000
    sc, pose = make_case("C3", P, W, H)
    vp = lcgs.view_params(lcgs.make_camera(*pose, W, H))
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=200_000)
    r.render_async(vp)                      # warm-up: workspace reservation and function attributes
    n = dev.num_rendered()
    eager = r.intermediates(n)
    r.img.zero_(), r.keys.zero_(), r.vals.zero_(), r.ranges.zero_()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            r.render_async(vp, stream=s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    replay = r.intermediates(n)
    for k in ("keys_sorted", "vals_sorted", "ranges", "tiles_touched", "radii"):
        assert np.array_equal(eager[k], replay[k]), k
    assert np.array_equal(bits(eager["img"]), bits(replay["img"]))


def _render_and_compare(lcgs, dev, sc, pose, W, H, sh=None, sh_deg=3, scale_modifier=1.0, bg=(0.0, 0.0, 0.0), opacity=None):
    sh = sc.sh if sh is None else sh
    opacity = sc.opacity if opacity is None else opacity
    ovp = orc.view_params(orc.make_camera(*pose, W, H))
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sh, opacity, ovp, bg=bg, sh_deg=sh_deg, scale_modifier=scale_modifier)
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sh, opacity, W, H, list_capacity=max(fr.num_rendered, 1) + 5,
                      sh_deg=sh_deg, scale_modifier=scale_modifier, bg_color=bg)
    n = r.render(lcgs.make_camera(*pose, W, H))
    assert_frame_matches(r.intermediates(n), fr, fused=True)
    return fr


def test_heavy_tailed_tiles_and_huge_splats(lcgs, dev):
    """scale_modifier 25: Gaussians cover hundreds of tiles each (the warp-cooperative key emission and
    long, saturating tile lists), non-black background."""
    sc, pose = make_case("C3", 2500, 640, 360)
    fr = _render_and_compare(lcgs, dev, sc, pose, 640, 360, scale_modifier=25.0, bg=(0.2, 0.4, 0.6))
    assert fr.tiles_touched.max() > 300 and fr.num_rendered > 100_000
    lens = fr.ranges[:, 1].astype(np.int64) - fr.ranges[:, 0]
    assert lens.max() > 500


def test_big_gaussians_take_the_piece_balanced_emission_path(lcgs, dev):
    """Frames / bands of >= 12288 tiles expand Gaussians that cover more than 128 tiles in emit_big_kernel (pieces of 2048
    instances, one CTA each) instead of in the warp that owns them; offsets, unsorted multiset, sorted lists, ranges and
    image must not change.  2048 x 1536 = 128 x 96 tiles, scale_modifier 30: most instances come from such Gaussians."""
    sc, pose = make_case("C3", 3000, 2048, 1536)
    fr = _render_and_compare(lcgs, dev, sc, pose, 2048, 1536, scale_modifier=30.0, bg=(0.1, 0.2, 0.3))
    big = fr.tiles_touched > 128
    assert big.sum() > 200 and fr.tiles_touched[big].sum() > 0.5 * fr.num_rendered
    assert (fr.tiles_touched > 2048).sum() > 10, "some Gaussians must span several pieces"


@pytest.mark.parametrize("deg", [0, 1, 2])
def test_fused_render_with_lower_sh_degree(lcgs, dev, deg):
    sc, pose = make_case("C3", 6000, 256, 160)
    sh = np.ascontiguousarray(sc.sh[:, :(deg + 1) ** 2, :])
    _render_and_compare(lcgs, dev, sc, pose, 256, 160, sh=sh, sh_deg=deg)


def test_opacity_extremes(lcgs, dev):
    """opacity 0, below 1/255 (never contributes: threshold +inf), exactly 1/255, and 1."""
    sc, pose = make_case("C3", 8000, 320, 200)
    op = sc.opacity.copy()
    op[0::5] = 0.0
    op[1::5] = 0.0039
    op[2::5] = np.float32(1.0) / np.float32(255.0)
    op[3::5] = 1.0
    _render_and_compare(lcgs, dev, sc, pose, 320, 200, opacity=op, scale_modifier=3.0)


def test_far_gaussians_run_the_fourth_depth_pass(lcgs, dev):
    """The fused frame sorts depth keys relative to bits(0.2f); below depth 13107 they fit three 9-bit
    digits and the fourth pass skips itself.  Push a third of the Gaussians beyond that depth (scaled
    up so that they still cover tiles) so the fourth pass has to run, mixed with near ones."""
    sc, pose = make_case("C3", 6000, 320, 200)
    pos, scale = sc.pos.copy(), sc.scale.copy()
    cam = np.asarray(pose[0], np.float32)
    far = np.arange(pos.shape[0]) % 3 == 0
    k = np.float32(9000.0)
    pos[far] = cam + (pos[far] - cam) * k
    scale[far] = scale[far] * k
    sc2 = type(sc)(**{**sc.__dict__, "pos": np.ascontiguousarray(pos), "scale": np.ascontiguousarray(scale)})
    fr = _render_and_compare(lcgs, dev, sc2, pose, 320, 200)
    touching = fr.tiles_touched > 0
    assert (fr.depth[touching] > 13200.0).sum() > 100 and (fr.depth[touching] < 100.0).sum() > 100


def test_extreme_aspect_frame_takes_the_general_emission_path(lcgs, dev):
    """The instance emission normally maps a slot of a Gaussian's rect to (x, y) by multiply-high and finds the
    owning Gaussian with a ballot; grids with gx * gx * gy >= 2^32 take integer division and a shuffle search
    instead.  A 640000 x 48 frame (40000 x 3 tiles: 4.8e9) reaches that path in the production build; it also
    exercises tile columns far beyond 8 bits in the packed 16-bit rect fields and 17-bit tile ids in the sort."""
    sc, pose = make_case("C3", 4000, 640_000, 48)
    fr = _render_and_compare(lcgs, dev, sc, pose, 640_000, 48, scale_modifier=40.0)
    assert fr.num_rendered > 10_000 and int(fr.keys_sorted[-1] >> np.uint64(32)) > 65_536


def test_frames_wider_than_65535_tiles_are_rejected(lcgs, dev):
    """The fused path packs tile coordinates into 16 bits; frame_geom refuses anything beyond instead of corrupting keys."""
    sc, pose = make_case("C3", 100, 64, 64)
    W = 16 * 65_536
    with pytest.raises(lcgs.LcgsError):
        r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, 64, 64, list_capacity=1000)
        r.c_frame.width = W
        vp = lcgs.view_params(lcgs.make_camera(*pose, W, 64))
        r.render_async(vp)


def test_rgb8_epilogue_is_the_apps_post_process(lcgs, dev):
    """target_rgb8 (fused HWC / v-flip / truncating *255 epilogue of the blend kernel) must equal the app's host
    post-process (app/main.cpp:322-337, restated by orc_image_to_rgb8) applied to the same float image, bit for
    bit, and stay within one 8-bit level of the oracle's own frame."""
    W, H, P = 637, 355, 30_000   # ragged in both directions
    sc, pose = make_case("C3", P, W, H)
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, orc.view_params(orc.make_camera(*pose, W, H)),
                     bg=(0.1, 0.3, 0.9))
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=fr.num_rendered + 1,
                      bg_color=(0.1, 0.3, 0.9), rgb8=True)
    assert r.render(lcgs.make_camera(*pose, W, H)) == fr.num_rendered
    got = r.rgb8.cpu().numpy().reshape(H, W, 3)
    assert np.array_equal(got, orc.image_to_rgb8(r.image().cpu().numpy()))
    diff = np.abs(got.astype(np.int16) - orc.image_to_rgb8(fr.img).astype(np.int16))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
    # and through the read-back entry point
    import torch
    host = torch.empty(3 * W * H, dtype=torch.uint8).pin_memory()
    r.read_image_rgb8(host)
    torch.cuda.synchronize()
    assert np.array_equal(host.numpy().reshape(H, W, 3), got)


def test_transpose_rgba8_matches_the_viewer_shader(lcgs, dev):
    """Display::_transpose_shader (app/display.cpp:30-39): CHW float -> RGBA8 unorm, alpha 255, no flip."""
    import torch
    W, H = 333, 77
    rng = np.random.default_rng(5)
    img = rng.uniform(-0.2, 1.2, size=(3, H, W)).astype(np.float32)
    img[:, 0, :4] = np.array([0.0, 1.0, 0.5, 0.49803922], np.float32)  # exact ends and a tie-ish value
    out = lcgs.transpose_rgba8(dev, torch.from_numpy(img).cuda().reshape(-1), W, H).cpu().numpy()
    want = np.empty((H, W, 4), np.uint8)
    want[..., :3] = np.rint(np.clip(img, 0.0, 1.0) * np.float32(255.0)).astype(np.uint8).transpose(1, 2, 0)
    want[..., 3] = 255
    assert np.array_equal(out, want)


def test_scene_constants_do_not_change_a_bit(lcgs, dev):
    """lcgs_b200_scene_prepare moves the alpha-test constants out of the frame; frames with and without it agree bit for bit."""
    W, H, P = 512, 288, 25_000
    sc, pose = make_case("C3", P, W, H)
    cam = lcgs.make_camera(*pose, W, H)
    a = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=1_000_000, prepare_scene=True)
    b = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=1_000_000, prepare_scene=False)
    na, nb = a.render(cam), b.render(cam)
    assert na == nb > 0
    assert np.array_equal(bits(a.image().cpu().numpy()), bits(b.image().cpu().numpy()))
    thr = a.alpha_consts.cpu().numpy().reshape(-1, 2)[:, 0]
    want = np.array([orc.alpha_threshold(float(o)) for o in sc.opacity[:2000]], np.float32)
    assert np.array_equal(bits(thr[:2000]), bits(want))


def test_moving_camera_resets_every_frame(lcgs, dev):
    """The viewer moves the camera between frames (app/display.cpp:49-153).  The reference then reads stale depth /
    means / conic of Gaussians that were culled in the new frame; this build defines culled = zero
    every frame, so a frame never depends on the frames before it: A, B, A through the same buffers equal the
    oracle's fresh-buffer results for A, B, A."""
    W, H, P = 480, 272, 20_000
    sc, _ = make_case("C3", P, W, H)
    poses = [scenes.orbit_pose(0), scenes.orbit_pose(97), (scenes.CAM_POS, scenes.CAM_TARGET, scenes.WORLD_UP_COLMAP),
             scenes.orbit_pose(0)]
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=2_000_000)
    culled = []
    for pose in poses:
        fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, orc.view_params(orc.make_camera(*pose, W, H)))
        n = r.render(lcgs.make_camera(*pose, W, H))
        assert_frame_matches(r.intermediates(n), fr, fused=True)
        culled.append(fr.depth < 0.2)
    assert (culled[0] != culled[1]).sum() > 100, "the poses must cull different Gaussians for this test to mean anything"


def test_empty_band_still_writes_background(lcgs, dev):
    """A band of tile rows whose own instance count is 0 (e.g. only the last tile row, which is never binned: Q1)
    must still write bg * T to its tiles: quirk Q10 (image untouched) is about the FRAME's count, and a single-GPU
    frame with num_rendered > 0 does write those pixels."""
    W, H, P = 320, 200, 6000
    sc, pose = make_case("C3", P, W, H)
    cam = lcgs.make_camera(*pose, W, H)
    bg = (0.25, 0.5, 0.75)
    full = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=500_000, bg_color=bg)
    assert full.render(cam) > 0
    gy = (H + 15) // 16
    band = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=500_000, bg_color=bg,
                         tile_rows=(gy - 1, gy))
    band.img.fill_(-7.0)   # stale pixels from an earlier frame
    assert band.render(cam) == 0
    y0 = (gy - 1) * 16
    got, want = band.image().cpu().numpy(), full.image().cpu().numpy()
    assert np.array_equal(bits(got[:, y0:]), bits(want[:, y0:]))
    assert bool((got[:, :y0] == -7.0).all()), "a band only writes its own rows"


def test_pipelined_frames_report_their_own_capacity(lcgs, dev):
    """Two frames with different list capacities enqueued back to back on one stream: each frame's overflow is decided
    on the device against its own capacity (the old host-side compare used whichever capacity was enqueued last)."""
    W, H, P = 256, 160, 5000
    sc, pose = make_case("C3", P, W, H)
    vp = lcgs.view_params(lcgs.make_camera(*pose, W, H))
    big = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=400_000)
    n = big.render(lcgs.make_camera(*pose, W, H))
    assert n > 100
    small = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=50)
    small.render_async(vp)   # overflows
    big.render_async(vp)     # does not: the last enqueued frame decides what num_rendered reports
    assert dev.num_rendered() == n
    big.render_async(vp)
    small.render_async(vp)
    with pytest.raises(lcgs.CapacityError):
        dev.num_rendered()


def test_sub_allocated_frame_buffers(lcgs, dev):
    """The header promises 8-byte alignment requirements only: tiles_touched / point_offsets / depth carved out of a
    larger allocation at a 4-byte offset take the scalar path of the scan and must give the same frame."""
    import torch
    W, H, P = 300, 180, 9001
    sc, pose = make_case("C3", P, W, H)
    cam = lcgs.make_camera(*pose, W, H)
    fr = orc.forward(sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, orc.view_params(orc.make_camera(*pose, W, H)))
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=fr.num_rendered + 3)
    pool = torch.zeros(3 * (P + 4), dtype=torch.int32, device="cuda")
    r.tiles_touched = pool[1:1 + P]
    r.point_offsets = pool[P + 6:2 * P + 6]
    r.depth = pool[2 * P + 11:3 * P + 11].view(torch.float32)
    assert all(t.data_ptr() % 16 != 0 and t.data_ptr() % 4 == 0 for t in (r.tiles_touched, r.point_offsets, r.depth))
    f = r.c_frame
    f.tiles_touched, f.point_offsets, f.depth = r.tiles_touched.data_ptr(), r.point_offsets.data_ptr(), r.depth.data_ptr()
    n = r.render(cam)
    assert_frame_matches(r.intermediates(n), fr, fused=True)
