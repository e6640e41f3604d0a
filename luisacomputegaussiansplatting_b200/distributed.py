"""Multi-GPU partitioning of the forward render path (one process per GPU, torch.distributed).

The path shards in the two ways BASELINE.json names, with no mid-pipeline exchange:

* view sharding   -- a camera sweep is split by view (view k -> rank k mod G), the Gaussian set is
                     replicated, every rank runs the whole pipeline on its own GPU; the only
                     collective gathers the finished frames on rank 0.
* tile-row sharding -- one very large frame is split into contiguous bands of 16-pixel tile rows;
                     every rank preprocesses all Gaussians but bins/sorts/blends only its band (the
                     kernels clip tile rects to the band, so instance sets partition exactly); the
                     only collective gathers the finished strips on rank 0.

Everything here is host logic over torch.distributed and works with the `nccl` backend (CUDA
tensors, NVLink) and the `gloo` backend (CPU tensors, used by the CPU tests).  The renderer itself is
passed in as a callable, so nothing in this file touches a kernel.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# --------------------------------------------------------------------------------------------------
# partitioning
# --------------------------------------------------------------------------------------------------

def shard_views(num_views: int, world: int, rank: int) -> List[int]:
    """Views of a sweep owned by `rank`: k with k mod world == rank (round-robin keeps neighbouring,
    similarly expensive views on different GPUs)."""
    return list(range(rank, num_views, world))


def split_tile_rows(num_rows: int, world: int, weights: Optional[Sequence[float]] = None) -> List[Tuple[int, int]]:
    """Contiguous bands [r0, r1) of tile rows, one per rank, covering [0, num_rows).

    With `weights` (e.g. instances per tile row from a previous frame) the bands are balanced by
    cumulative weight instead of by row count: tile populations are heavy-tailed.  Bands may be
    empty when world > num_rows."""
    if weights is None:
        weights = [1.0] * num_rows
    assert len(weights) == num_rows
    cum = [0.0]
    for w in weights:
        cum.append(cum[-1] + max(float(w), 0.0))
    if cum[-1] <= 0.0:
        cum = [float(i) for i in range(num_rows + 1)]
    total = cum[-1]
    cuts = [0]
    for g in range(1, world):
        target = total * g / world
        r = cuts[-1]
        while r < num_rows and abs(cum[r + 1] - target) <= abs(cum[r] - target):
            r += 1
        cuts.append(r)
    cuts.append(num_rows)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def row_weights_from_ranges(ranges: torch.Tensor, gx: int) -> List[float]:
    """Instances per tile row from a frame's ranges buffer ([tiles][2], start/end)."""
    r = ranges.view(-1, 2).to(torch.int64)
    per_tile = (r[:, 1] - r[:, 0]).clamp_(min=0)
    rows = per_tile.numel() // gx
    return per_tile[: rows * gx].view(rows, gx).sum(dim=1).to(torch.float64).cpu().tolist()


def fit_band_cost(samples: Sequence[Tuple[float, float, float]]) -> Tuple[float, float]:
    """Least-squares fit of measured band times, ms ~ a * instances + b * tile_rows + c, over `samples` =
    (instances, tile_rows, ms) of several bands (ideally from two different splits, so that instances and rows
    vary independently).  Returns (a, b) clipped to >= 0: the cost of one more instance and of one more tile row.
    Balancing by instance count alone leaves wide, sparse bands up to 1.8x slower than narrow dense ones
    (per-tile and per-pixel work in the blend, fewer and bigger Gaussians in the emission)."""
    import numpy as np

    A = np.array([[n, r, 1.0] for n, r, _ in samples], np.float64)
    y = np.array([t for _, _, t in samples], np.float64)
    if len(samples) < 3 or np.linalg.matrix_rank(A) < 3:
        return 1.0, 0.0
    (a, b, _c), *_ = np.linalg.lstsq(A, y, rcond=None)
    a, b = max(float(a), 0.0), max(float(b), 0.0)
    return (a, b) if (a > 0.0 or b > 0.0) else (1.0, 0.0)


def split_tile_rows_by_cost(num_rows: int, world: int, weights: Sequence[float], a: float, b: float) -> List[Tuple[int, int]]:
    """Bands balanced by the fitted cost a * instances_in_row + b per tile row (see fit_band_cost)."""
    return split_tile_rows(num_rows, world, [a * float(w) + b for w in weights])


# --------------------------------------------------------------------------------------------------
# collectives (the only exchange steps of the path)
# --------------------------------------------------------------------------------------------------

def gather_frames(frame: torch.Tensor, dst: int = 0, group=None) -> Optional[List[torch.Tensor]]:
    """Gather one equally-sized frame per rank on `dst` (NCCL has no native gather; torch lowers this
    to grouped send/recv).  Returns the list on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = [torch.empty_like(frame) for _ in range(world)] if rank == dst else None
    dist.gather(frame, out, dst=dst, group=group)
    return out


def gather_strips(img: torch.Tensor, bands: Sequence[Tuple[int, int]], height: int, dst: int = 0,
                  group=None) -> Optional[torch.Tensor]:
    """Assemble a tile-row-sharded frame on `dst`.

    `img` is this rank's planar [3, H, W] image of which only pixel rows [16*r0, min(H, 16*r1)) of its
    band are valid.  Each rank sends its strip (3 contiguous channel chunks packed into one message);
    dst writes them into its own image and returns it."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = [(min(height, 16 * r0), min(height, 16 * r1)) for r0, r1 in bands]
    if rank != dst:
        y0, y1 = rows[rank]
        if y1 > y0:
            dist.send(img[:, y0:y1, :].contiguous(), dst=dst, group=group)
        return None
    for src in range(world):
        if src == dst:
            continue
        y0, y1 = rows[src]
        if y1 > y0:
            buf = torch.empty((3, y1 - y0, img.shape[2]), dtype=img.dtype, device=img.device)
            dist.recv(buf, src=src, group=group)
            img[:, y0:y1, :] = buf
    return img


# --------------------------------------------------------------------------------------------------
# the exchange fused into the render: peer stores instead of a collective
# --------------------------------------------------------------------------------------------------

class PeerFrameRing:
    """`slots` frames in ONE buffer on rank `dst` that every rank can write over NVLink (lcgs.PeerBuffer).
    A slot is a planar [3,H,W] float32 image, optionally followed by the uint8 HWC image the app saves
    (`rgb8=True`).  A rank renders straight into `ptr(slot)` -- or, for tile-row sharding, all ranks render
    their bands into the same slot -- so the blend kernel's stores are the gather.

    Two ways to hand frames over:
      * batch:  `complete()` (stream sync + barrier), dst reads, `release()` (barrier);
      * stream: per-slot 32-bit sequence flags inside the buffer (`ready`: written by the renderer after its frame,
        polled by dst; `consumed`: written by dst after it has read the slot, polled by the renderer before it
        overwrites the slot), all in stream order on the devices -- no host synchronisation between frames
        (`signal_ready` / `wait_ready` / `signal_consumed` / `wait_consumed`).  For tile-row sharding every
        writer has its own ready flag per slot (`writer`)."""

    def __init__(self, device, width: int, height: int, slots: int, dst: int = 0, group=None, rgb8: bool = False,
                 writers: int = 1, float_image: bool = True):
        """float_image=False: slots hold only the uint8 image (the app's final product: 3 bytes per pixel cross NVLink
        instead of 15); the planar float image then stays in the renderer's local buffer."""
        from . import lcgs

        assert float_image or rgb8
        self.device = device
        self.W, self.H, self.slots, self.dst, self.group = width, height, slots, dst, group
        self.img_bytes = 3 * width * height * 4 if float_image else 0
        self.rgb8_bytes = ((3 * width * height + 255) // 256) * 256 if rgb8 else 0
        self.frame_bytes = self.img_bytes + self.rgb8_bytes          # a multiple of 4 (and of 16 for W*H % 4 == 0)
        self.frame_bytes = ((self.frame_bytes + 255) // 256) * 256
        self.writers = writers
        self.flags_offset = self.frame_bytes * slots
        # [slots][writers] ready words, then [slots] consumed words
        self.buf = lcgs.PeerBuffer(device, self.flags_offset + 4 * slots * (writers + 1) + 256, owner=dst, group=group)

    def ptr(self, slot: int) -> int:
        """The slot's planar float image (its start; with float_image=False the slot starts with the uint8 image)."""
        assert 0 <= slot < self.slots
        return self.buf.ptr + slot * self.frame_bytes

    def rgb8_ptr(self, slot: int) -> int:
        assert self.rgb8_bytes
        return self.ptr(slot) + self.img_bytes

    def _ready_ptr(self, slot: int, writer: int = 0) -> int:
        return self.buf.ptr + self.flags_offset + 4 * (slot * self.writers + writer)

    def _consumed_ptr(self, slot: int) -> int:
        return self.buf.ptr + self.flags_offset + 4 * (self.slots * self.writers + slot)

    # ---- stream-ordered hand-over ----
    def signal_ready(self, slot: int, seq: int, writer: int = 0, stream=None):
        """Renderer: frame number `seq` (1, 2, ...) of this slot is complete once the stream gets here."""
        self.device.peer_signal(self._ready_ptr(slot, writer), seq, stream)

    def wait_ready(self, slot: int, seq: int, writer: int = 0, stream=None, timeout_ms: int = 5000):
        self.device.peer_wait(self._ready_ptr(slot, writer), seq, timeout_ms, stream)

    def signal_consumed(self, slot: int, seq: int, stream=None):
        """dst: frame `seq` of this slot has been read; the slot may be overwritten."""
        self.device.peer_signal(self._consumed_ptr(slot), seq, stream)

    def wait_consumed(self, slot: int, seq: int, stream=None, timeout_ms: int = 5000):
        self.device.peer_wait(self._consumed_ptr(slot), seq, timeout_ms, stream)

    # ---- batch hand-over ----
    def complete(self):
        """All frames enqueued so far (on every rank) are visible on dst when this returns."""
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def frame(self, slot: int):
        """dst only: host copy of a slot as [3, H, W] float32."""
        return self.buf.to_host(slot * self.frame_bytes, 3 * self.W * self.H).reshape(3, self.H, self.W)

    def release(self):
        """dst has read what it wanted: the writers may reuse the slots when this returns (collective)."""
        dist.barrier(group=self.group)

    def close(self):
        """Collective: the writers unmap, then the owner frees."""
        if not self.buf.is_owner:
            self.buf.close()
        dist.barrier(group=self.group)
        if self.buf.is_owner:
            self.buf.close()


def render_sweep_view_sharded_peer(render_view_into: Callable[[int, int], None], num_views: int, ring: PeerFrameRing):
    """View sharding with the gather fused into the render: view k is rendered by rank k mod G directly
    into slot k of `ring` (ring.slots >= num_views) on the destination rank.  `render_view_into(k, ptr)`
    enqueues the frame of view k with target pointer `ptr`.  Returns the frames (host arrays) on dst."""
    world = dist.get_world_size(ring.group)
    rank = dist.get_rank(ring.group)
    assert ring.slots >= num_views
    for k in shard_views(num_views, world, rank):
        render_view_into(k, ring.ptr(k))
    ring.complete()
    frames = [ring.frame(k) for k in range(num_views)] if rank == ring.dst else None
    ring.release()
    return frames


def render_frame_tile_row_sharded_peer(render_band_into: Callable[[int, int, int], None], height: int, ring: PeerFrameRing,
                                       slot: int = 0, weights: Optional[Sequence[float]] = None,
                                       bands: Optional[Sequence[Tuple[int, int]]] = None):
    """Tile-row sharding with the gather fused into the render: every rank renders its band of tile rows
    directly into the SAME image (slot `slot` of `ring`) on the destination rank.
    `render_band_into(r0, r1, ptr)` enqueues the band [r0, r1) with target pointer `ptr`.
    Returns (assembled host image on dst / None, bands)."""
    world = dist.get_world_size(ring.group)
    rank = dist.get_rank(ring.group)
    if bands is None:
        bands = split_tile_rows((height + 15) // 16, world, weights)
    assert len(bands) == world
    r0, r1 = bands[rank]
    if r1 > r0:
        render_band_into(r0, r1, ring.ptr(slot))
    ring.complete()
    img = ring.frame(slot) if rank == ring.dst else None
    ring.release()
    return img, bands


# --------------------------------------------------------------------------------------------------
# drivers
# --------------------------------------------------------------------------------------------------

def render_sweep_view_sharded(render_view: Callable[[int], torch.Tensor], num_views: int, dst: int = 0,
                              group=None) -> Optional[List[torch.Tensor]]:
    """Render `num_views` views, view k on rank k mod G, and gather them in view order on `dst`.

    `render_view(k)` returns this rank's finished frame for view k (any fixed shape).  Every rank
    takes part in ceil(num_views / G) gather rounds (ranks without a view in the last round send a
    dummy frame that dst drops)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_views(num_views, world, rank)
    rounds = (num_views + world - 1) // world
    frames: List[Optional[torch.Tensor]] = [None] * num_views if rank == dst else []
    template = None
    for i in range(rounds):
        if i < len(mine):
            frame = render_view(mine[i])
            template = frame
        else:
            assert template is not None or rounds == 1
            frame = torch.zeros_like(template) if template is not None else render_view(0) * 0
        got = gather_frames(frame, dst=dst, group=group)
        if rank == dst:
            for r in range(world):
                k = i * world + r
                if k < num_views:
                    frames[k] = got[r].clone()
    return frames if rank == dst else None


def render_frame_tile_row_sharded(render_band: Callable[[int, int], torch.Tensor], height: int,
                                  weights: Optional[Sequence[float]] = None, dst: int = 0, group=None,
                                  bands: Optional[Sequence[Tuple[int, int]]] = None):
    """Render one frame split by tile rows: `render_band(r0, r1)` returns this rank's [3,H,W] image
    with its band rendered.  Returns (assembled image on dst / None, bands)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gy = (height + 15) // 16
    if bands is None:
        bands = split_tile_rows(gy, world, weights)
    r0, r1 = bands[rank]
    img = render_band(r0, r1)
    return gather_strips(img, bands, height, dst=dst, group=group), bands
