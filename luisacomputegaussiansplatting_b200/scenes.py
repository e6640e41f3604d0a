"""Synthetic Gaussian clouds with the shapes of the reference's release scenes.

The release checkpoints (README.md:26-29 of the reference) are not available offline, so the
benchmark scenes are seeded synthetic clouds.  Arrays are the *post-activation* host
arrays that ``read_gs_ply`` produces (app/gaussians.cpp:75-171): pos [P,3], scale [P,3] (= exp of
the stored log-scale), rotq [P,4] normalised (r,x,y,z), opacity [P] (= sigmoid of the stored
logit), sh [P,16,3] coefficient-major / RGB-interleaved.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# the reference's hard-coded camera (app/main.cpp:191-202)
CAM_POS = (-3.0, -0.5, 3.3)
CAM_TARGET = (0.0, 3.0, 0.5)
WORLD_UP_COLMAP = (0.0, -1.0, -1.0)
WORLD_UP_BLENDER = (0.0, 0.0, 1.0)


@dataclass
class Scene:
    name: str
    pos: np.ndarray
    scale: np.ndarray
    rotq: np.ndarray
    opacity: np.ndarray
    sh: np.ndarray
    # pre-activation values, as stored in an INRIA .ply
    log_scale: np.ndarray
    raw_rot: np.ndarray
    logit_opacity: np.ndarray
    world: str = "colmap"

    @property
    def num_gaussians(self) -> int:
        return int(self.pos.shape[0])

    def nbytes(self) -> int:
        return sum(a.nbytes for a in (self.pos, self.scale, self.rotq, self.opacity, self.sh))


@dataclass(frozen=True)
class Config:
    """One row of BASELINE.json:configs."""

    key: str
    shape: str  # "object" | "scene360"
    P: int
    W: int
    H: int
    seed: int
    mu: float
    sigma: float
    world: str
    description: str


CONFIGS = {
    "C1": Config("C1", "object", 300_000, 800, 800, 1, -4.6, 0.7, "blender",
                 "nerf_blender_lego-shaped, 300k SH3, 800x800, --world=blender"),
    "C2": Config("C2", "scene360", 6_100_000, 1237, 822, 2, -5.0, 0.8, "colmap",
                 "mip360_bicycle-shaped, 6.1M SH3, 1237x822"),
    "C3": Config("C3", "scene360", 5_800_000, 1920, 1080, 3, -5.0, 0.8, "colmap",
                 "mip360_garden-shaped, 5.8M SH3, 1920x1080"),
    "C5": Config("C5", "scene360", 10_000_000, 7680, 4320, 4, -5.0, 0.8, "colmap",
                 "10M SH3 synthetic scene, 7680x4320, tile-row sharded"),
}


def _finish(name, rng, pos, scale_mult, mu, sigma, world) -> Scene:
    P = pos.shape[0]
    log_scale = rng.normal(mu, sigma, size=(P, 3)).astype(np.float32)
    if scale_mult is not None:
        log_scale = (log_scale + np.log(scale_mult, dtype=np.float32)[:, None]).astype(np.float32)
    raw_rot = rng.standard_normal(size=(P, 4), dtype=np.float32)
    logit = rng.normal(0.0, 2.0, size=P).astype(np.float32)
    sh = np.empty((P, 16, 3), np.float32)
    sh[:, 0, :] = rng.standard_normal(size=(P, 3), dtype=np.float32)
    rest = rng.standard_normal(size=(P, 15, 3), dtype=np.float32)
    rest *= np.float32(0.15)
    sh[:, 1:, :] = rest
    del rest
    # activations (app/gaussians.cpp:15-35)
    scale = np.exp(log_scale, dtype=np.float32)
    norm = np.sqrt((raw_rot * raw_rot).sum(axis=1, dtype=np.float32), dtype=np.float32)
    rotq = (raw_rot / norm[:, None]).astype(np.float32)
    opacity = (np.float32(1.0) / (np.float32(1.0) + np.exp(-logit, dtype=np.float32))).astype(np.float32)
    return Scene(name, np.ascontiguousarray(pos, np.float32), scale, rotq, opacity, sh, log_scale, raw_rot, logit,
                 world)


def object_scene(P: int, seed: int, mu: float = -4.6, sigma: float = 0.7, name: str = "object") -> Scene:
    """Blender-object-shaped cloud: uniform box of half-extents (0.9,0.9,0.6) around the target."""
    rng = np.random.default_rng(seed)
    half = np.array([0.9, 0.9, 0.6], np.float32)
    pos = (rng.uniform(-1.0, 1.0, size=(P, 3)).astype(np.float32) * half + np.array(CAM_TARGET, np.float32))
    return _finish(name, rng, pos.astype(np.float32), None, mu, sigma, "blender")


def scene360(P: int, seed: int, mu: float = -5.0, sigma: float = 0.8, name: str = "scene360") -> Scene:
    """Mip-NeRF-360-shaped cloud: 25 % object blob, 45 % ground disc, 30 % far background shell."""
    rng = np.random.default_rng(seed)
    target = np.array(CAM_TARGET, np.float32)
    up = np.array([0.0, -1.0, 0.0], np.float32)
    cat = rng.random(P)
    n_obj = int((cat < 0.25).sum())
    n_gnd = int(((cat >= 0.25) & (cat < 0.70)).sum())
    n_bg = P - n_obj - n_gnd
    pos = np.empty((P, 3), np.float32)
    # object: N(target, 0.6^2 I)
    pos[:n_obj] = target + rng.normal(0.0, 0.6, size=(n_obj, 3)).astype(np.float32)
    # ground disc of radius 6 through target - 0.8*up, thickness sigma 0.03
    e1 = np.cross(up, np.array([1.0, 0.0, 0.3], np.float32))
    e1 = (e1 / np.linalg.norm(e1)).astype(np.float32)
    e2 = np.cross(up, e1).astype(np.float32)
    r = 6.0 * np.sqrt(rng.random(n_gnd))
    th = rng.uniform(0.0, 2.0 * np.pi, size=n_gnd)
    hgt = rng.normal(0.0, 0.03, size=n_gnd)
    centre = target - np.float32(0.8) * up
    pos[n_obj:n_obj + n_gnd] = (centre + (r * np.cos(th))[:, None] * e1 + (r * np.sin(th))[:, None] * e2
                                + hgt[:, None] * up).astype(np.float32)
    # background shell: direction uniform on the sphere, radius U(6,30)
    d = rng.standard_normal(size=(n_bg, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rad = rng.uniform(6.0, 30.0, size=n_bg)
    pos[n_obj + n_gnd:] = (target + d * rad[:, None]).astype(np.float32)
    # shuffle so that categories are interleaved in memory like a trained checkpoint
    perm = rng.permutation(P)
    pos = pos[perm]
    dist = np.linalg.norm(pos - target, axis=1).astype(np.float32)
    mult = np.maximum(np.float32(1.0), dist / np.float32(3.0)).astype(np.float32)
    return _finish(name, rng, pos, mult, mu, sigma, "colmap")


def make_config_scene(key: str, P: int | None = None) -> tuple[Scene, Config]:
    """Scene of BASELINE.json config `key` (C1/C2/C3/C5); P overrides the Gaussian count (tests)."""
    cfg = CONFIGS[key]
    n = cfg.P if P is None else P
    if cfg.shape == "object":
        sc = object_scene(n, cfg.seed, cfg.mu, cfg.sigma, name=cfg.key)
    else:
        sc = scene360(n, cfg.seed, cfg.mu, cfg.sigma, name=cfg.key)
    return sc, cfg


def world_up(world: str):
    return WORLD_UP_BLENDER if world == "blender" else WORLD_UP_COLMAP


def orbit_pose(k: int, n_views: int = 256):
    """View k of the C4 orbit: returns (pos, target, world_up) as float32 triples."""
    up = np.array([0.0, -1.0, 0.0], np.float64)
    e1 = np.cross(up, np.array([1.0, 0.0, 0.3]))
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(up, e1)
    target = np.array(CAM_TARGET, np.float64)
    th = 2.0 * np.pi * (k % n_views) / n_views
    pos = target + 4.6 * (np.cos(th) * e1 + np.sin(th) * e2) + 1.5 * up
    return pos.astype(np.float32), target.astype(np.float32), up.astype(np.float32)
