"""The gather fused into the render (peer stores), on ONE GPU: two processes share cuda:0, the process group is
gloo (NCCL refuses two ranks on one device), and rank 1 blends its frames / its band straight into rank 0's
buffer through CUDA IPC -- the same code path that crosses NVLink on a multi-GPU box
(scripts/multi_gpu_check.py checks that one).  Results must equal single-process frames bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

W, H, P, VIEWS = 320, 200, 6000, 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from luisacomputegaussiansplatting_b200 import distributed as D
    from luisacomputegaussiansplatting_b200 import lcgs, scenes

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    try:
        sc, cfg = scenes.make_config_scene("C3", P=P)
        dev = lcgs.Device(0)
        r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=400_000, keep_intermediates=False)
        ring = D.PeerFrameRing(dev, W, H, slots=VIEWS)

        def view_into(k, ptr):
            r.set_target_ptr(ptr)
            r.render_async(lcgs.view_params(lcgs.make_camera(*scenes.orbit_pose(k * 13), W, H)))

        frames = D.render_sweep_view_sharded_peer(view_into, VIEWS, ring)
        r.set_target(r.img)
        pose = (scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world))
        vp = lcgs.view_params(lcgs.make_camera(*pose, W, H))
        bands = D.split_tile_rows((H + 15) // 16, world)
        r0, r1 = bands[rank]
        rb = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=400_000, tile_rows=(r0, r1),
                           keep_intermediates=False)

        def band_into(a, b, ptr):
            rb.set_target_ptr(ptr)
            rb.render_async(vp)

        img, _ = D.render_frame_tile_row_sharded_peer(band_into, H, ring, slot=0)
        if rank == 0:
            want_views = []
            for k in range(VIEWS):
                r.render(lcgs.make_camera(*scenes.orbit_pose(k * 13), W, H))
                want_views.append(r.image().cpu().numpy().copy())
            r.render(lcgs.make_camera(*pose, W, H))
            q.put((frames, want_views, img, r.image().cpu().numpy().copy()))
        ring.close()
        dev.close()
    finally:
        dist.destroy_process_group()


def test_peer_stores_assemble_views_and_tile_row_bands():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    frames, want_views, img, want_img = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    def diff(a, b):
        d = (a.view(np.uint32) != b.view(np.uint32)).any(axis=0)
        ys, xs = np.nonzero(d)
        return "%d pixels differ, rows %s..%s cols %s..%s, max abs %.3g" % (
            int(d.sum()), ys.min() if ys.size else None, ys.max() if ys.size else None, xs.min() if xs.size else None,
            xs.max() if xs.size else None, float(np.abs(a - b).max()))

    bad = {k: diff(frames[k], want_views[k]) for k in range(VIEWS) if not np.array_equal(frames[k].view(np.uint32), want_views[k].view(np.uint32))}
    assert not bad, bad
    assert np.array_equal(img.view(np.uint32), want_img.view(np.uint32)), diff(img, want_img)
    assert float(np.abs(want_img).max()) > 0.0
