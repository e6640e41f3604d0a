"""Pins the oracle's camera helpers to the reference's own known-answer tests.

Vectors restated from /root/reference/test/test_camera.cpp:48-144 (tolerance 1e-5 as upstream).
These are the only results on the path that the reference pins.
"""
import math

import numpy as np

from oracle import oracle as orc

EPS = 1e-5
SQRT2 = np.float32(1.41421356)


def test_lookat_front_and_orthogonality():  # test_camera.cpp:51-67
    cam = orc.get_lookat_cam((-4.0, -4.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0))
    assert np.allclose(list(cam.position), [-4.0, -4.0, 0.0], atol=EPS)
    assert np.allclose(list(cam.front), [1.0 / SQRT2, 1.0 / SQRT2, 0.0], atol=EPS)
    up, right, front = (np.array(list(v), np.float32) for v in (cam.up, cam.right, cam.front))
    assert abs(float(up @ right)) < EPS
    assert abs(float(up @ front)) < EPS
    assert abs(float(right @ front)) < EPS


def test_local_world_round_trip():  # test_camera.cpp:70-86
    cam = orc.get_lookat_cam((-4.0, -4.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0))
    l2w = orc.local_to_world_matrix(cam)
    w2l = orc.world_to_local_matrix(cam)
    local = np.array([2.0 * SQRT2, 3.0, 2.0 * SQRT2, 1.0], np.float32)
    world = orc.mat4_mul_vec4(l2w, local)
    assert np.allclose(world[:3], [0.0, -4.0, 3.0], atol=EPS)
    back = orc.mat4_mul_vec4(w2l, world)
    assert np.all(np.abs(back - local) < EPS)


def test_projection_matrix():  # test_camera.cpp:89-114
    pi = np.float32(3.14159265359)
    fovx = np.float32(60.0) * pi / np.float32(180.0)
    fovy = np.float32(45.0) * pi / np.float32(180.0)
    tanfovx = np.float32(math.tan(fovx / 2.0))
    tanfovy = np.float32(math.tan(fovy / 2.0))
    near, far = 0.1, 100.0
    proj = orc.projection_matrix(tanfovx, tanfovy, near, far)
    n = orc.mat4_mul_vec4(proj, [0.0, 0.0, near, 1.0])
    assert abs(n[2] / n[3]) < EPS
    f = orc.mat4_mul_vec4(proj, [0.0, 0.0, far, 1.0])
    assert abs(f[2] / f[3] - 1.0) < EPS
    p = orc.mat4_mul_vec4(proj, [0.2, 0.3, 2.0, 1.0])
    assert abs(p[0] / p[3] - 0.2 / math.tan(fovx / 2) / 2.0) < 1e-5
    assert abs(p[1] / p[3] - 0.3 / math.tan(fovy / 2) / 2.0) < 1e-5


def test_special_camera_cases():  # test_camera.cpp:117-144
    cam = orc.get_lookat_cam((0.0, 0.0, 5.0), (0.0, 0.0, 10.0), (0.0, 1.0, 0.0))
    # Upstream expects right == (+1,0,0) here (test_camera.cpp:129), which contradicts its own first
    # vector: test_camera.cpp:70-80 (l2w*(2*sqrt2,3,2*sqrt2,1) == (0,-4,3)) only holds for the
    # standard right-handed cross product, and with it cross((0,0,1),(0,1,0)) = (-1,0,0).  doctest's
    # CHECK is non-fatal, and the roadmap lists unit tests as not done (doc/roadmap.md:3), so the
    # inconsistency goes unnoticed upstream.  camera.h:74-82 is what ships; we follow it.
    assert np.allclose(np.abs(list(cam.right)), [1.0, 0.0, 0.0], atol=EPS)
    assert np.allclose(list(cam.right), [-1.0, 0.0, 0.0], atol=EPS)
    l2w = orc.local_to_world_matrix(cam)
    world = orc.mat4_mul_vec4(l2w, [0.0, 0.0, 1.0, 1.0])
    expected = np.array(list(cam.position)) + np.array(list(cam.front))
    assert np.allclose(world[:3], expected, atol=EPS)


def test_view_params_match_reference_host_code():  # gs_projector/impl.cpp:34-42
    cam = orc.make_camera((-3.0, -0.5, 3.3), (0.0, 3.0, 0.5), (0.0, -1.0, -1.0), 1920, 1080)
    vp = orc.view_params(cam)
    fovy = np.float32(60.0) / np.float32(180.0) * np.float32(3.1415926536)
    assert abs(vp.tanfovy - math.tan(float(fovy) * 0.5)) < 1e-6
    assert abs(vp.tanfovx - vp.tanfovy * 1920.0 / 1080.0) < 1e-6
    assert abs(vp.focalx - 1920.0 / (2.0 * vp.tanfovx)) < 1e-3
    assert abs(vp.focaly - 1080.0 / (2.0 * vp.tanfovy)) < 1e-3
    # proj*v = (x/tanfovx, y/tanfovy, a*z+b*w, z)  (camera.h:54-72)
    v = orc.mat4_mul_vec4(np.array(vp.proj, np.float32), [1.0, 2.0, 3.0, 1.0])
    assert abs(v[0] - 1.0 / vp.tanfovx) < 1e-6 and abs(v[1] - 2.0 / vp.tanfovy) < 1e-6 and abs(v[3] - 3.0) < 1e-6
    # view*(p,1).z is the distance along front
    p = np.array([0.0, 3.0, 0.5, 1.0], np.float32)
    pv = orc.mat4_mul_vec4(np.array(vp.view, np.float32), p)
    dist = np.linalg.norm(np.array([0.0, 3.0, 0.5]) - np.array([-3.0, -0.5, 3.3]))
    assert abs(pv[2] - dist) < 1e-4 and abs(pv[0]) < 1e-4 and abs(pv[1]) < 1e-4
