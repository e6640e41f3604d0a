"""ctypes front-end of the CPU oracle (oracle/lcgs_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under luisacomputegaussiansplatting_b200/ imports it.

Parity status: "parity unpinned" except for the camera helpers (see the header of lcgs_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblcgs_oracle.so")
_SRC = os.path.join(_HERE, "lcgs_oracle.c")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile if the .so is missing or stale."""
    stale = (not os.path.exists(_SO)) or (
        os.path.exists(_SRC) and os.path.getmtime(_SRC) > os.path.getmtime(_SO)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "liblcgs_oracle.so"], check=True)
    return _SO


class Camera(C.Structure):
    """orc_camera == lcgs::Camera (lcgs/include/lcgs/util/camera.h:15-25)."""

    _fields_ = [
        ("position", C.c_float * 3),
        ("front", C.c_float * 3),
        ("up", C.c_float * 3),
        ("right", C.c_float * 3),
        ("fov", C.c_float),
        ("aspect_ratio", C.c_float),
        ("width", C.c_int),
        ("height", C.c_int),
    ]


class ViewParams(C.Structure):
    """orc_view_params: host-derived kernel parameters (gs_projector/impl.cpp:34-42)."""

    _fields_ = [
        ("view", C.c_float * 16),
        ("proj", C.c_float * 16),
        ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("focalx", C.c_float),
        ("focaly", C.c_float),
        ("cam_pos", C.c_float * 3),
        ("width", C.c_int),
        ("height", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_exp.restype = C.c_float
        _lib.orc_exp.argtypes = [C.c_float]
        _lib.orc_alpha_threshold.restype = C.c_float
        _lib.orc_alpha_threshold.argtypes = [C.c_float]
        _lib.orc_exp_monotonicity_violations.restype = C.c_long
        _lib.orc_exp_monotonicity_violations.argtypes = [C.c_float, C.c_float]
        _lib.orc_forward.restype = C.c_long
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_set_num_threads.argtypes = [C.c_int]
        _lib.orc_set_num_threads.restype = None
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# ------------------------------------------------------------------------------------------------
# camera
# ------------------------------------------------------------------------------------------------

def get_lookat_cam(pos, target, world_up) -> Camera:
    cam = Camera()
    lib().orc_get_lookat_cam(_p(_f32(pos)), _p(_f32(target)), _p(_f32(world_up)), C.byref(cam))
    return cam


def make_camera(pos, target, world_up, width: int, height: int, fov: float = 60.0) -> Camera:
    """get_lookat_cam + the three assignments of app/main.cpp:204-207."""
    cam = get_lookat_cam(pos, target, world_up)
    cam.fov = fov
    cam.aspect_ratio = np.float32(width) / np.float32(height)
    cam.width = width
    cam.height = height
    return cam


def world_to_local_matrix(cam: Camera) -> np.ndarray:
    m = np.zeros(16, np.float32)
    lib().orc_world_to_local_matrix(C.byref(cam), _p(m))
    return m


def local_to_world_matrix(cam: Camera) -> np.ndarray:
    m = np.zeros(16, np.float32)
    lib().orc_local_to_world_matrix(C.byref(cam), _p(m))
    return m


def projection_matrix(tanfovx, tanfovy, znear=0.1, zfar=100.0) -> np.ndarray:
    m = np.zeros(16, np.float32)
    lib().orc_projection_matrix(C.c_float(tanfovx), C.c_float(tanfovy), C.c_float(znear), C.c_float(zfar), _p(m))
    return m


def mat4_mul_vec4(m, v) -> np.ndarray:
    o = np.zeros(4, np.float32)
    lib().orc_mat4_mul_vec4(_p(_f32(m)), _p(_f32(v)), _p(o))
    return o


def view_params(cam: Camera) -> ViewParams:
    vp = ViewParams()
    lib().orc_view_params_from_camera(C.byref(cam), C.byref(vp))
    return vp


# ------------------------------------------------------------------------------------------------
# stages
# ------------------------------------------------------------------------------------------------

def grids(W: int, H: int):
    return (W + 15) // 16, (H + 15) // 16


def sh_process(pos, sh, cam_pos, deg: int = 3) -> np.ndarray:
    pos, sh = _f32(pos), _f32(sh)
    P = pos.shape[0]
    color = np.zeros((P, 3), np.float32)
    lib().orc_sh_process(C.c_int(P), C.c_int(deg), _p(_f32(cam_pos)), _p(pos), _p(sh), _p(color))
    return color


def project(pos, scale, rotq, vp: ViewParams, scale_modifier: float = 1.0):
    """K2 on zero-initialised outputs (= the build's defined behaviour for culled Gaussians)."""
    pos, scale, rotq = _f32(pos), _f32(scale), _f32(rotq)
    P = pos.shape[0]
    means = np.zeros((P, 2), np.float32)
    depth = np.zeros(P, np.float32)
    cov = np.zeros((P, 3), np.float32)
    lib().orc_project(C.c_int(P), _p(pos), _p(scale), _p(rotq), C.c_float(scale_modifier), C.byref(vp), _p(means),
                      _p(depth), _p(cov))
    return means, depth, cov


def allocate_tiles(W, H, depth, means_ndc, cov, row0: int = 0, row1: int = -1):
    """K3: returns (means_pix, conic, tiles_touched, radii); inputs are not modified."""
    depth = _f32(depth)
    P = depth.shape[0]
    means = _f32(means_ndc).copy()
    conic = _f32(cov).copy()
    tiles = np.zeros(P, np.uint32)
    radii = np.zeros(P, np.int32)
    lib().orc_allocate_tiles(C.c_int(P), C.c_int(W), C.c_int(H), _p(depth), _p(means), _p(conic), _p(tiles), _p(radii),
                             C.c_int(row0), C.c_int(row1))
    return means, conic, tiles, radii


def inclusive_sum(x) -> np.ndarray:
    x = np.ascontiguousarray(x, np.uint32)
    out = np.zeros_like(x)
    lib().orc_inclusive_sum_u32(_p(x), _p(out), C.c_int(x.shape[0]))
    return out


def copy_with_keys(W, H, means_pix, offsets, radii, depth, row0: int = 0, row1: int = -1):
    offsets = np.ascontiguousarray(offsets, np.uint32)
    P = offsets.shape[0]
    n = int(offsets[-1]) if P else 0
    keys = np.zeros(n, np.uint64)
    vals = np.zeros(n, np.uint32)
    lib().orc_copy_with_keys(C.c_int(P), C.c_int(W), C.c_int(H), _p(_f32(means_pix)), _p(offsets),
                             _p(np.ascontiguousarray(radii, np.int32)), _p(_f32(depth)), _p(keys), _p(vals),
                             C.c_int(row0), C.c_int(row1))
    return keys, vals


def sort_pairs(keys, vals):
    keys = np.ascontiguousarray(keys, np.uint64)
    vals = np.ascontiguousarray(vals, np.uint32)
    ko, vo = np.zeros_like(keys), np.zeros_like(vals)
    lib().orc_sort_pairs_u64_u32(_p(keys), _p(vals), _p(ko), _p(vo), C.c_size_t(keys.shape[0]))
    return ko, vo


def get_ranges(keys_sorted, num_tiles: int) -> np.ndarray:
    keys_sorted = np.ascontiguousarray(keys_sorted, np.uint64)
    ranges = np.zeros((num_tiles, 2), np.uint32)
    lib().orc_get_ranges(C.c_size_t(keys_sorted.shape[0]), _p(keys_sorted), _p(ranges), C.c_int(num_tiles))
    return ranges


def blend(W, H, bg, ranges, point_list, means_pix, conic, opacity, color, row0: int = 0, row1: int = -1,
          img: np.ndarray | None = None):
    """K9: returns (img CHW float32 [3,H,W], n_examined [H,W])."""
    if img is None:
        img = np.zeros((3, H, W), np.float32)
    nex = np.zeros((H, W), np.uint32)
    lib().orc_blend(C.c_int(W), C.c_int(H), _p(_f32(bg)), _p(np.ascontiguousarray(ranges, np.uint32)),
                    _p(np.ascontiguousarray(point_list, np.uint32)), _p(_f32(means_pix)), _p(_f32(conic)),
                    _p(_f32(opacity)), _p(_f32(color)), _p(img), _p(nex), C.c_int(row0), C.c_int(row1))
    return img, nex


def exp(x: float) -> float:
    return float(lib().orc_exp(C.c_float(x)))


def alpha_threshold(op: float) -> float:
    return float(lib().orc_alpha_threshold(C.c_float(op)))


def image_to_rgb8(img_chw: np.ndarray) -> np.ndarray:
    _, H, W = img_chw.shape
    out = np.zeros((H, W, 3), np.uint8)
    lib().orc_image_to_rgb8(C.c_int(W), C.c_int(H), _p(_f32(img_chw)), _p(out))
    return out


# ------------------------------------------------------------------------------------------------
# whole frame
# ------------------------------------------------------------------------------------------------

STAGES = ("sh", "project", "allocate_tiles", "scan", "copy_with_keys", "sort", "ranges", "blend")


@dataclass
class Frame:
    """Every intermediate of one frame, as the reference's buffers would hold them."""

    num_rendered: int
    color: np.ndarray
    means_2d: np.ndarray      # pixel coordinates after K3 (zeros where culled)
    depth: np.ndarray
    conic: np.ndarray
    tiles_touched: np.ndarray
    radii: np.ndarray
    offsets: np.ndarray
    keys_unsorted: np.ndarray
    vals_unsorted: np.ndarray
    keys_sorted: np.ndarray
    vals_sorted: np.ndarray
    ranges: np.ndarray
    img: np.ndarray           # [3,H,W]
    n_examined: np.ndarray    # [H,W]
    stage_ms: dict = field(default_factory=dict)


def forward(pos, scale, rotq, sh, opacity, vp: ViewParams, bg=(0.0, 0.0, 0.0), sh_deg: int = 3,
            scale_modifier: float = 1.0, capacity: int | None = None, row0: int = 0, row1: int = -1,
            img: np.ndarray | None = None) -> Frame:
    pos, scale, rotq, sh, opacity = _f32(pos), _f32(scale), _f32(rotq), _f32(sh), _f32(opacity)
    P = pos.shape[0]
    W, H = vp.width, vp.height
    gx, gy = grids(W, H)
    r1 = gy if row1 < 0 else row1
    ntile = gx * (r1 - row0)
    if capacity is None:
        # two-pass: the oracle needs the lists sized; run the cheap stages first to learn N
        m, d, c = project(pos, scale, rotq, vp, scale_modifier)
        _, _, tiles, _ = allocate_tiles(W, H, d, m, c, row0, row1)
        capacity = max(int(tiles.astype(np.uint64).sum()), 1)
    color = np.zeros((P, 3), np.float32)
    means = np.zeros((P, 2), np.float32)
    depth = np.zeros(P, np.float32)
    conic = np.zeros((P, 3), np.float32)
    tiles = np.zeros(P, np.uint32)
    radii = np.zeros(P, np.int32)
    offsets = np.zeros(P, np.uint32)
    ku = np.zeros(capacity, np.uint64)
    vu = np.zeros(capacity, np.uint32)
    ks = np.zeros(capacity, np.uint64)
    vs = np.zeros(capacity, np.uint32)
    ranges = np.zeros((ntile, 2), np.uint32)
    if img is None:
        img = np.zeros((3, H, W), np.float32)
    nex = np.zeros((H, W), np.uint32)
    ms = (C.c_double * 8)()
    n = lib().orc_forward(
        C.c_int(P), C.c_int(sh_deg), _p(pos), _p(scale), _p(rotq), _p(sh), _p(opacity), C.c_float(scale_modifier),
        C.byref(vp), _p(_f32(bg)), _p(color), _p(means), _p(depth), _p(conic), _p(tiles), _p(radii), _p(offsets),
        _p(ku), _p(vu), _p(ks), _p(vs), C.c_size_t(capacity), _p(ranges), _p(img), _p(nex), C.c_int(row0),
        C.c_int(row1), ms)
    if n < 0:
        raise RuntimeError("oracle: num_rendered exceeds capacity %d" % capacity)
    n = int(n)
    return Frame(n, color, means, depth, conic, tiles, radii, offsets, ku[:n], vu[:n], ks[:n], vs[:n], ranges, img,
                 nex, dict(zip(STAGES, [float(x) for x in ms])))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    """Use `n` OpenMP threads from now on (overrides OMP_NUM_THREADS, which torchrun sets to 1)."""
    lib().orc_set_num_threads(int(n))


def blend_stats() -> dict:
    """Totals of the last blend: E examined (pixel, entry) pairs, E_alpha passed the alpha test, E_contrib blended."""
    out = (C.c_ulonglong * 3)()
    lib().orc_blend_stats(out)
    return {"E": int(out[0]), "E_alpha": int(out[1]), "E_contrib": int(out[2])}
