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


def test_stream_ordered_flags_and_consumer_checksum():
    """The device-side hand-over (peer_signal / peer_wait / checksum_u32) in one process: a render stream produces 9 frames
    into a 2-slot ring, a consumer stream checksums each one and releases the slot, nothing synchronises in between;
    the checksums equal those of frames rendered one by one.  Then a wait on a flag nobody signals gives up after its
    timeout and is reported."""
    import torch

    from luisacomputegaussiansplatting_b200 import lcgs, scenes

    sc, cfg = scenes.make_config_scene("C3", P=P)
    dev = lcgs.Device(0)
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=400_000)
    n_words, frames, slots = 3 * W * H, 9, 2
    frame_bytes = ((4 * n_words + 255) // 256) * 256
    mem = torch.zeros(slots * frame_bytes + 4 * slots * 2, dtype=torch.uint8, device="cuda")
    base = mem.data_ptr()
    ready = [base + slots * frame_bytes + 4 * s for s in range(slots)]
    consumed = [base + slots * frame_bytes + 4 * (slots + s) for s in range(slots)]
    sums = torch.zeros(frames, dtype=torch.int64, device="cuda")
    render, consume = torch.cuda.Stream(), torch.cuda.Stream()
    vps = [lcgs.view_params(lcgs.make_camera(*scenes.orbit_pose(k * 11), W, H)) for k in range(frames)]
    torch.cuda.synchronize()
    for k in range(frames):
        slot, seq = k % slots, k // slots + 1
        if k >= slots:
            dev.peer_wait(consumed[slot], seq - 1, stream=render)
        r.set_target_ptr(base + slot * frame_bytes)
        r.render_async(vps[k], stream=render)
        dev.peer_signal(ready[slot], seq, stream=render)
        dev.peer_wait(ready[slot], seq, stream=consume)
        dev.checksum_u32(base + slot * frame_bytes, n_words, sums[k:k + 1], stream=consume)
        dev.peer_signal(consumed[slot], seq, stream=consume)
    torch.cuda.synchronize()
    assert dev.peer_timeouts() == 0
    r.set_target(r.img)
    one = torch.zeros(1, dtype=torch.int64, device="cuda")
    for k in range(frames):
        r.render_async(vps[k])
        dev.checksum_u32(r.img.data_ptr(), n_words, one)
        torch.cuda.synchronize()
        want = int(r.img.view(torch.int32).to(torch.int64).bitwise_and(0xFFFFFFFF).sum().item())
        assert int(one.item()) == want, "checksum kernel"
        assert int(sums[k].item()) == want, "frame %d consumed before it was complete, or overwritten before it was read" % k
    # a dead peer cannot hang the GPU
    dev.peer_wait(ready[0], 1000, timeout_ms=50)
    torch.cuda.synchronize()
    assert dev.peer_timeouts() == 1
    dev.close()
