#!/usr/bin/env python
"""bench.py -- forward splat-render benchmark (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C3]

One "step" = one frame of the hot path (fused preprocess -> scan -> key duplication -> onesweep
sort -> tile ranges -> blend) of the mip360_garden-shaped synthetic scene (C3: 5.8 M SH3 Gaussians,
1920x1080, the reference's hard-coded camera), the configuration BASELINE.json's metric is quoted
on.  Prints ONE JSON line on rank 0.

  value      Gaussians/s, whole job, device-timed (CUDA events, max over ranks), scene resident in HBM
  e2e        same metric through the public API with the per-frame host traffic inside the timed
             region: camera parameters from host memory in, finished planar image D2H into pinned
             host memory + num_rendered out, host wall clock.  (The Gaussian set is uploaded once
             before the clock starts, as app/main.cpp:216-226 does; e2e_with_scene_upload
             additionally re-uploads the 1.4 GB scene from pinned memory every frame.)
  roofline   onesweep pass kernel (largest HBM stage): 24 B per instance per launch / average launch
             duration measured with CUDA events on the launching stream
  cpu_baseline  the CPU oracle (a port of the reference's algorithm; the reference itself cannot be
             built offline) on the host cores of this box, same scene and camera

N > 1 (torchrun): the Gaussian set is replicated, every rank renders its own frame per step and the
finished frames reach rank 0 inside the timed region (view sharding, weak scaling): by default every
rank's blend kernel stores straight into rank 0's peer-mapped ring over NVLink (--gather peer, the
gather is fused into the render); --gather nccl uses an overlapped dist.gather instead.  --impl reference times the CPU oracle instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gaussians_per_second_forward_render"
UNIT = "Gaussians/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--gaussians", type=int, default=None, help="override P (debugging only; invalidates the number)")
    ap.add_argument("--orbit", action="store_true", help="rank r / step s renders orbit view s*N+r instead of the fixed pose")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how finished frames reach rank 0 -- peer: every rank's blend kernel stores straight into "
                         "rank 0's peer-mapped ring over NVLink (the gather is fused into the render); nccl: dist.gather")
    ap.add_argument("--capacity", type=int, default=None,
                    help="instance list capacity L (default 20 000 000 = app/main.cpp:245; 260 000 000 for C5)")
    args = ap.parse_args()
    if args.capacity is None:
        args.capacity = 260_000_000 if args.config == "C5" else 20_000_000
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(args, cfg, P, extra=None):
    c = {"workload": "%s: %s, reference camera (app/main.cpp:191-202), list capacity %d" % (cfg.key, cfg.description,
                                                                                        args.capacity),
         "gaussians": P, "width": cfg.W, "height": cfg.H, "sh_degree": 3,
         "l2": "inputs (%.2f GB Gaussian set + %.2f GB instance lists) exceed the 126 MB L2; no explicit flush" % (
             P * 236 / 1e9, args.capacity * 24 / 1e9),
         "parallelism": ("view-sharded x%d (scene replicated, %s)" % (
             args.gpus, "frames blended straight into rank 0's peer-mapped ring over NVLink" if args.gather == "peer"
             else "NCCL frame gather to rank 0")) if args.gpus > 1
         else "single GPU"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle): cpu_baseline of the b200 arm and the whole --impl reference arm
# ------------------------------------------------------------------------------------------------

def oracle_frame_seconds(sc, cfg, P_sample, reps, warm):
    """Times the oracle's whole frame on the first P_sample Gaussians of the scene."""
    from luisacomputegaussiansplatting_b200 import scenes
    from oracle import oracle as orc

    cam = orc.make_camera(scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world), cfg.W, cfg.H)
    vp = orc.view_params(cam)
    sl = slice(0, P_sample)
    args = (sc.pos[sl], sc.scale[sl], sc.rotq[sl], sc.sh[sl], sc.opacity[sl], vp)
    times, stages, n = [], None, 0
    for i in range(warm + reps):
        t0 = time.perf_counter()
        fr = orc.forward(*args, capacity=40_000_000)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
            stages = fr.stage_ms
            n = fr.num_rendered
    return times, stages, n, orc.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from luisacomputegaussiansplatting_b200 import scenes

    sc, cfg = scenes.make_config_scene(args.config, P=args.gaussians)
    P = sc.num_gaussians
    # bounded sample: the full frame unless K+W frames would take more than ~4 minutes
    probe, _, _, cores = oracle_frame_seconds(sc, cfg, P, 1, 0)
    total = args.steps + args.warmup
    P_sample = P
    if probe[0] * total > 240.0:
        P_sample = max(10_000, int(P * 240.0 / (probe[0] * total)))
    times, stages, n, cores = oracle_frame_seconds(sc, cfg, P_sample, args.steps, args.warmup)
    sec = float(np.sum(times))
    value = P_sample * args.steps / sec
    sample = "whole frame of the first %d of %d Gaussians, %d timed frames, oracle stages %s" % (
        P_sample, P, args.steps, json.dumps({k: round(v, 1) for k, v in stages.items()}))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, P, {"note": "CPU oracle port of the reference path on host cores; the "
                                                          "reference itself cannot be built offline (LuisaCompute + lcpp "
                                                          "are network dependencies)"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "num_rendered": n,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm), reasons=sorted(reasons))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------

# preprocess, scan+compact (+depth histograms), [depth sort: 4 passes, the last one returns at once], duplicate_keys
# (+chained offsets, tile histograms), [tile sort: 2 passes], ranges, blend
KERNELS_PER_FRAME = 11


def run_b200(args):
    import torch
    import torch.distributed as dist

    from luisacomputegaussiansplatting_b200 import lcgs, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.gpus != world and rank == 0:
        sys.stderr.write("warning: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE\n" % (args.gpus, world))
    args.gpus = world

    sc, cfg = scenes.make_config_scene(args.config, P=args.gaussians)
    P, W, H = sc.num_gaussians, cfg.W, cfg.H
    dev = lcgs.Device(local)
    t0 = time.perf_counter()
    r = lcgs.Renderer(dev, sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity, W, H, list_capacity=args.capacity,
                      keep_intermediates=False)
    torch.cuda.synchronize()
    upload_s = time.perf_counter() - t0

    def pose(step):
        if args.orbit:
            return scenes.orbit_pose(step * world + rank)
        return scenes.CAM_POS, scenes.CAM_TARGET, scenes.world_up(cfg.world)

    # N > 1, --gather peer (default): rank 0 owns a ring of 2 x N frames every rank can write over NVLink;
    # rank r blends frame i straight into slot (i & 1) * N + r, so the gather IS the blend's stores.
    # --gather nccl: frames are rendered into two alternating device images so that the NCCL gather of
    # frame i (to rank 0) overlaps the render of frame i+1.
    use_peer = world > 1 and args.gather == "peer"
    ring = None
    if use_peer:
        from luisacomputegaussiansplatting_b200 import distributed as D
        ok = torch.ones(1, device="cuda")
        try:
            ring = D.PeerFrameRing(dev, W, H, slots=2 * world)
        except Exception as e:  # e.g. no peer access between two GPUs of the box: every rank falls back together
            sys.stderr.write("rank %d: peer ring unavailable (%s); falling back to the NCCL gather\n" % (rank, e))
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            if ring is not None:
                ring.buf.close()
            ring, use_peer, args.gather = None, False, "nccl"
    frame_imgs = [r.img, torch.empty_like(r.img)] if world > 1 else [r.img]
    gather_lists = [None, None]
    if world > 1 and rank == 0 and not use_peer:
        gather_lists = [[torch.empty_like(r.img) for _ in range(world)] for _ in range(2)]
    pending = [None, None]

    def step_device(i):
        cam = lcgs.make_camera(*pose(i), W, H)
        if use_peer:
            r.set_target_ptr(ring.ptr((i & 1) * world + rank))
            r.render_async(lcgs.view_params(cam))
        elif world > 1:
            b = i & 1
            if pending[b] is not None:
                pending[b].wait()                      # buffer b has been sent: the stream may overwrite it
            r.set_target(frame_imgs[b])
            r.render_async(lcgs.view_params(cam))
            pending[b] = dist.gather(frame_imgs[b], gather_lists[b], dst=0, async_op=True)
        else:
            r.render_async(lcgs.view_params(cam))

    def drain():
        for b in range(2):
            if pending[b] is not None:
                pending[b].wait()
                pending[b] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the device-timed region -------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step_device(i)
    drain()
    r.set_target(frame_imgs[0])
    n_rendered = dev.num_rendered()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_device(i)
    drain()
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    r.set_target(frame_imgs[0])
    # ---- per-stage breakdown + roofline of the onesweep pass kernel (rank 0, N=1 semantics) -------
    dev.set_profiling(True)
    stage_acc, sort_acc = {}, {"histogram_ms": 0.0, "passes_ms": 0.0}
    reps = max(3, min(args.steps, 10))
    for i in range(reps):
        cam = lcgs.make_camera(*pose(i), W, H)
        r.render_async(lcgs.view_params(cam))
        st = dev.stage_times()
        sb = dev.sort_breakdown()
        for k, v in st.items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / reps
        sort_acc["histogram_ms"] += sb["histogram_ms"] / reps
        sort_acc["passes_ms"] += sb["passes_ms"] / reps
        sort_acc["num_passes"] = sb["num_passes"]
    dev.set_profiling(False)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- end to end through the public API: camera in (host), image + count out (pinned host) ----
    # Streaming use of the API: frame i+1 is enqueued while frame i's image travels to the host on a
    # copy stream (two device image buffers, two pinned host buffers); the host consumes frame i-1
    # before it enqueues frame i+1, so at most two frames are in flight.  Every frame's image and
    # count are in host memory when the clock stops.
    main_stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    dev_imgs = [r.img, torch.empty_like(r.img)]
    host_imgs = [torch.empty(3 * W * H, dtype=torch.float32).pin_memory() for _ in range(2)]
    host_counts = torch.zeros(args.steps, dtype=torch.int32).pin_memory()
    ev_render = [torch.cuda.Event() for _ in range(2)]
    ev_copy = [torch.cuda.Event() for _ in range(2)]
    checksum = 0.0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        b = i & 1
        if i >= 2:
            main_stream.wait_event(ev_copy[b])          # device buffer b has been read out
        cam = lcgs.make_camera(*pose(i), W, H)           # host-side camera -> kernel parameters
        r.set_target(dev_imgs[b])
        r.render_async(lcgs.view_params(cam))
        r.read_num_rendered_async(host_counts[i:i + 1])
        ev_render[b].record(main_stream)
        copy_stream.wait_event(ev_render[b])
        r.read_image(host_imgs[b], stream=copy_stream)
        ev_copy[b].record(copy_stream)
        if i >= 1:
            ev_copy[b ^ 1].synchronize()                 # frame i-1 is on the host: consume it
            checksum += float(host_imgs[b ^ 1][12345])
    ev_copy[(args.steps - 1) & 1].synchronize()
    torch.cuda.synchronize()
    checksum += float(host_imgs[(args.steps - 1) & 1][12345])
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    assert int(host_counts.min()) == n_rendered or args.orbit, (host_counts.tolist(), n_rendered)
    r.set_target(dev_imgs[0])
    host_img = host_imgs[0]

    # ---- variant: re-upload the whole Gaussian set from pinned host memory every frame -----------
    e2e_cold = None
    if rank == 0 and world == 1:
        pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in
                  (sc.pos, sc.scale, sc.rotq, sc.sh, sc.opacity)]
        dsts = [r.pos, r.scale, r.rotq, r.sh, r.opacity]
        k = max(2, min(args.steps, 5))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(k):
            for d, s in zip(dsts, pinned):
                d.view(-1).copy_(s.view(-1), non_blocking=True)
            r.render_async(lcgs.view_params(lcgs.make_camera(*pose(i), W, H)))
            r.read_image(host_img)
            dev.num_rendered()
        e2e_cold = {"value": P * k / (time.perf_counter() - t0), "unit": UNIT,
                    "h2d_bytes_per_step": int(sum(p.numel() * 4 for p in pinned)), "d2h_bytes_per_step": 3 * W * H * 4 + 8,
                    "steps": k}
        del pinned

    if ring is not None:
        ring.close()  # the owner frees, the others unmap (after a device sync)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- assemble the line ---------------------------------------------------------------------
    ms_per_step = ms_total / args.steps
    value = world * P * args.steps / (ms_total * 1e-3)
    peak, peak_src = peaks()
    N = n_rendered
    V = int((r.depth >= 0.2).sum().item())
    touching = int((r.tiles_touched > 0).sum().item())
    T = r.num_tiles
    passes = sort_acc["num_passes"]
    pass_ms = sort_acc["passes_ms"] / passes
    achieved = 24.0 * N / (pass_ms * 1e-3) / 1e9
    # algorithmic bytes per stage of the data flow that actually runs (M = Gaussians touching a tile): the Gaussians
    # are depth-sorted first (8-byte pairs), so the 12-byte instance pairs need `passes` tile-bit passes only.
    # The reference's data flow would move N*(8+24*ceil((32+log2 T)/8)) bytes in the sort alone.
    M = touching
    alg = {"preprocess": 48.0 * P + 228.0 * V, "scan": 8.0 * P + 12.0 * M, "depth_sort": M * (4.0 + 16.0 * 4) + 16.0 * M,
           "duplicate_keys": 20.0 * M + 12.0 * N, "sort": N * (8.0 + 24.0 * passes), "ranges": 8.0 * N + 16.0 * T,
           "blend": 40.0 * N + 12.0 * W * H}
    stages = {k: {"ms": round(v, 4), "alg_GB": round(alg[k] / 1e9, 4), "GBps": round(alg[k] / (v * 1e-3) / 1e9, 1),
                  "frac_of_hbm_peak": round(alg[k] / (v * 1e-3) / 1e9 / peak, 3)} for k, v in stage_acc.items()}
    stages["blend"]["bound"] = "fp32+shared-memory (not HBM)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, cfg, P),
        "frames_per_second": world * args.steps / (ms_total * 1e-3),
        "num_rendered": N, "visible": V, "touching": touching,
        "e2e": {"value": world * P * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 164,
                "d2h_bytes_per_step": 3 * W * H * 4 + 8, "ms_per_step": e2e_s / args.steps * 1e3,
                "note": "camera parameters in from host memory, planar image + num_rendered out to pinned host memory "
                        "every frame (D2H of frame i overlaps the render of frame i+1, <= 2 frames in flight); Gaussian "
                        "set uploaded once (%.0f ms) as in app/main.cpp:216-226" % (upload_s * 1e3)},
        "e2e_with_scene_upload": e2e_cold,
        "gpu_launches": KERNELS_PER_FRAME * args.steps * world,
        "roofline": {"kernel": "onesweep_pass_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "alg_bytes_per_launch": 24.0 * N, "launch_ms": pass_ms, "launches_per_step": passes},
        "stages": stages, "sort_breakdown": sort_acc,
        "clocks": clock_info,
    }
    # traffic from the committed ncu --set full capture, if it has been summarised
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            line["roofline"]["traffic"] = json.load(open(prof)).get("onesweep_pass_kernel_dram_bytes_per_launch")
        except Exception:
            pass

    if not args.no_cpu_baseline and world == 1:
        # the full C3 frame costs the oracle a few seconds: 1 warm-up + 3 timed frames ~ 15-30 s of CPU work
        times, ostages, on, cores = oracle_frame_seconds(sc, cfg, P, 3, 1)
        line["cpu_baseline"] = {"value": P * len(times) / float(np.sum(times)), "unit": UNIT, "cores": cores,
                                "kind": "port",
                                "sample": "3 whole frames of the same scene/camera (all %d Gaussians), oracle stage ms %s" % (
                                    P, json.dumps({k: round(v, 1) for k, v in ostages.items()})),
                                "ms_per_frame": float(np.mean(times)) * 1e3, "num_rendered": on}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
