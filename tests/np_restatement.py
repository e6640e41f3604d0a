"""Second, independent restatement of the reference path in vectorised numpy float32.

Written from SURVEY.md section 9 / the reference sources without looking at oracle/lcgs_oracle.c's
structure, to cross-check the C oracle (the reference itself has no tests for these stages).
numpy float32 array arithmetic is IEEE per operation with no FMA contraction, so integer outputs
must agree bit for bit with the oracle and float outputs must be identical too.
Citations are relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32

SH_C0 = F(0.28209479177387814)
SH_C1 = F(0.4886025119029199)
SH_C2 = [F(1.0925484305920792), F(-1.0925484305920792), F(0.31539156525252005), F(-1.0925484305920792),
         F(0.5462742152960396)]
SH_C3 = [F(-0.5900435899266435), F(2.890611442640554), F(-0.4570457994644658), F(0.3731763325901154),
         F(-0.4570457994644658), F(1.445305721320277), F(-0.5900435899266435)]


def sh_color(pos, sh, cam_pos, deg=3):
    """sh_preprocessor.cpp:27-166 + util/sh.hpp; sh is [P,16,3]."""
    pos = pos.astype(F)
    sh = sh.astype(F).reshape(pos.shape[0], -1, 3)
    d = pos - np.asarray(cam_pos, F)[None, :]
    len2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    inv = F(1.0) / np.sqrt(len2)
    x, y, z = d[:, 0] * inv, d[:, 1] * inv, d[:, 2] * inv
    X, Y, Z = x[:, None], y[:, None], z[:, None]
    res = sh[:, 0] * SH_C0
    if deg > 0:
        res = res + (-SH_C1) * ((sh[:, 1] * Y - sh[:, 2] * Z) + sh[:, 3] * X)
    if deg > 1:
        xx, yy, yz, zz, zx, xy = x * x, y * y, y * z, z * z, z * x, x * y
        k = [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * ((F(2.0) * zz - xx) - yy), SH_C2[3] * zx, SH_C2[4] * (xx - yy)]
        l2 = k[0][:, None] * sh[:, 4]
        for j in range(1, 5):
            l2 = l2 + k[j][:, None] * sh[:, 4 + j]
        res = res + l2
    if deg > 2:
        m = [SH_C3[0] * y * (F(3.0) * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * ((F(4.0) * zz - xx) - yy),
             SH_C3[3] * z * ((F(2.0) * zz - F(3.0) * xx) - F(3.0) * yy), SH_C3[4] * x * ((F(4.0) * zz - xx) - yy),
             SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - F(3.0) * yy)]
        l3 = m[0][:, None] * sh[:, 9]
        for j in range(1, 7):
            l3 = l3 + m[j][:, None] * sh[:, 9 + j]
        res = res + l3
    res = res + F(0.5)
    return np.minimum(np.maximum(res, F(0.0)), F(1.0)).astype(F)


def project(pos, scale, rotq, view, proj, tanfovx, tanfovy, focalx, focaly, scale_modifier=1.0):
    """gs_projector/shader.cpp:82-139, util/gaussian.hpp, util/transform.hpp:188-212.

    view/proj: flat column-major 16-vectors.  Returns (ndc[P,2], depth[P], cov[P,3], visible[P]);
    culled rows are zero."""
    V = np.asarray(view, F).reshape(4, 4)  # V[c][r]
    PJ = np.asarray(proj, F).reshape(4, 4)
    p = pos.astype(F)
    pv = [((p[:, 0] * V[0][r] + p[:, 1] * V[1][r]) + p[:, 2] * V[2][r]) + V[3][r] for r in range(3)]
    phx, phy, phw = pv[0] * PJ[0][0], pv[1] * PJ[1][1], pv[2]
    p_w = F(1.0) / (phw + F(1e-6))
    ndc = np.stack([phx * p_w, phy * p_w], axis=1)
    vis = ~(pv[2] < F(0.2))
    sc = F(scale_modifier) * scale.astype(F)
    w, x, y, z = (rotq[:, i].astype(F) for i in range(4))
    two = F(2.0)
    one = F(1.0)
    R = [[(one - (two * y) * y) - (two * z) * z, (two * x) * y + (two * z) * w, (two * x) * z - (two * y) * w],
         [(two * x) * y - (two * z) * w, (one - (two * x) * x) - (two * z) * z, (two * y) * z + (two * x) * w],
         [(two * x) * z + (two * y) * w, (two * y) * z - (two * x) * w, (one - (two * x) * x) - (two * y) * y]]
    M = [[sc[:, c] * R[c][r] for r in range(3)] for c in range(3)]
    Sg = [[(M[0][c] * M[0][r] + M[1][c] * M[1][r]) + M[2][c] * M[2][r] for r in range(3)] for c in range(3)]
    limx, limy = F(1.3) * F(tanfovx), F(1.3) * F(tanfovy)
    with np.errstate(all="ignore"):
        tx = np.minimum(np.maximum(pv[0] / pv[2], -limx), limx) * pv[2]
        ty = np.minimum(np.maximum(pv[1] / pv[2], -limy), limy) * pv[2]
        tz = pv[2]
        J00, J11 = F(focalx) / tz, F(focaly) / tz
        J02, J12 = (-F(focalx) * tx) / (tz * tz), (-F(focaly) * ty) / (tz * tz)
    Wc = [[V[r][c] for r in range(3)] for c in range(3)]  # transpose(mat3(view))
    T0 = [J00 * Wc[0][r] + J02 * Wc[2][r] for r in range(3)]
    T1 = [J11 * Wc[1][r] + J12 * Wc[2][r] for r in range(3)]
    A = [[(Sg[c][0] * T0[0] + Sg[c][1] * T0[1]) + Sg[c][2] * T0[2],
          (Sg[c][0] * T1[0] + Sg[c][1] * T1[1]) + Sg[c][2] * T1[2]] for c in range(3)]
    c00 = (T0[0] * A[0][0] + T0[1] * A[1][0]) + T0[2] * A[2][0]
    c01 = (T0[0] * A[0][1] + T0[1] * A[1][1]) + T0[2] * A[2][1]
    c11 = (T1[0] * A[0][1] + T1[1] * A[1][1]) + T1[2] * A[2][1]
    cov = np.stack([c00, c01, c11], axis=1).astype(F)
    depth = np.where(vis, pv[2], F(0.0)).astype(F)
    ndc = np.where(vis[:, None], ndc, F(0.0)).astype(F)
    cov = np.where(vis[:, None], cov, F(0.0)).astype(F)
    return ndc, depth, cov, vis


def _f2u(f):
    """cvt.rzi.u32.f32 semantics."""
    f = np.asarray(f, F)
    out = np.zeros(f.shape, np.uint64)
    pos = f > 0
    big = f >= F(4294967296.0)
    ok = pos & ~big
    out[ok] = np.floor(f[ok].astype(np.float64)).astype(np.uint64)
    out[big] = 0xFFFFFFFF
    return out.astype(np.uint32)


def get_rect(px, py, radius, gx, gy):
    """module.cpp:22-36"""
    fr = radius.astype(F)
    s16 = F(16.0)
    with np.errstate(all="ignore"):
        mnx = np.minimum(_f2u((px - fr) / s16), np.uint32(gx - 1))
        mny = np.minimum(_f2u((py - fr) / s16), np.uint32(gy - 1))
        mxx = np.minimum(_f2u(((px + fr) + s16) - F(1.0)) // np.uint32(16), np.uint32(gx - 1))
        mxy = np.minimum(_f2u(((py + fr) + s16) - F(1.0)) // np.uint32(16), np.uint32(gy - 1))
    return mnx, mny, mxx, mxy


def allocate_tiles(W, H, depth, ndc, cov):
    """gs_tile_splatter/shader.cpp:102-163; returns pix, conic, tiles, radii (zeros where culled)."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    vis = ~(depth < F(0.2))
    with np.errstate(all="ignore"):
        a = cov[:, 0] + F(0.3)
        b = cov[:, 1]
        c = cov[:, 2] + F(0.3)
        det = a * c - b * b
        inv_det = F(1.0) / (det + F(1e-6))
        conic = np.stack([inv_det * c, inv_det * (-b), inv_det * a], axis=1)
        mid = F(0.5) * (a + c)
        sq = np.sqrt(np.fmax(F(0.1), mid * mid - det))
        lam = np.fmax(mid + sq, mid - sq)
        rf = np.ceil(F(3.0) * np.sqrt(lam))
        rf = np.where(np.isnan(rf), F(0.0), rf)
        radius = np.clip(rf.astype(np.float64), -2147483648.0, 2147483647.0).astype(np.int64).astype(np.int32)
        px = ((ndc[:, 0] + F(1.0)) * F(W) - F(1.0)) * F(0.5)
        py = ((ndc[:, 1] + F(1.0)) * F(H) - F(1.0)) * F(0.5)
    mnx, mny, mxx, mxy = get_rect(px, py, radius, gx, gy)
    tiles = ((mxx - mnx).astype(np.uint32) * (mxy - mny).astype(np.uint32)).astype(np.uint32)
    pix = np.stack([px, py], axis=1).astype(F)
    z = F(0.0)
    return (np.where(vis[:, None], pix, z).astype(F), np.where(vis[:, None], conic, z).astype(F),
            np.where(vis, tiles, 0).astype(np.uint32), np.where(vis, radius, 0).astype(np.int32))


def keys_and_values(W, H, pix, radii, depth, tiles):
    """gs_tile_splatter/shader.cpp:26-69, emitted in Gaussian order then row-major tile order."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    mnx, mny, mxx, mxy = get_rect(pix[:, 0], pix[:, 1], radii, gx, gy)
    keys, vals = [], []
    dbits = depth.astype(F).view(np.uint32).astype(np.uint64)
    for i in np.nonzero((radii > 0) & (tiles > 0))[0]:
        ys = np.arange(mny[i], mxy[i], dtype=np.uint64)
        xs = np.arange(mnx[i], mxx[i], dtype=np.uint64)
        t = (ys[:, None] * np.uint64(gx) + xs[None, :]).reshape(-1)
        keys.append((t << np.uint64(32)) | dbits[i])
        vals.append(np.full(t.shape[0], i, np.uint32))
    if not keys:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    return np.concatenate(keys), np.concatenate(vals)


def tile_ranges(keys_sorted, num_tiles):
    """gs_tile_splatter/shader.cpp:71-100 (empty tiles stay (0,0))."""
    ranges = np.zeros((num_tiles, 2), np.uint32)
    t = (keys_sorted >> np.uint64(32)).astype(np.int64)
    n = t.shape[0]
    if n == 0:
        return ranges
    bnd = np.nonzero(t[1:] != t[:-1])[0] + 1
    starts = np.concatenate([[0], bnd])
    ends = np.concatenate([bnd, [n]])
    ranges[t[starts], 0] = starts
    ranges[t[starts], 1] = ends
    return ranges


def _fma32(a, b, c):
    # exact product in binary64, one binary64 add, then round to binary32
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def blend_pixel(px, py, start, end, point_list, pix, conic, opacity, color, bg):
    """gs_tile_splatter/shader.cpp:249-286 for one pixel; pure Python, small cases only."""
    T = F(1.0)
    C = np.zeros(3, F)
    examined = 0
    pxf, pyf = F(px), F(py)
    for k in range(int(start), int(end)):
        examined += 1
        g = int(point_list[k])
        dx = F(pix[g, 0] - pxf)
        dy = F(pix[g, 1] - pyf)
        cx, cy, cz = conic[g]
        inner = _fma32(F(F(-0.5) * cx) * dx, dx, F(F(F(-0.5) * cz) * dy) * dy)
        power = _fma32(-F(cy * dx), dy, inner)
        if power > 0:
            continue
        e = F(math.exp(float(power)))
        alpha = min(F(0.99), F(opacity[g] * e))
        if alpha < F(1.0) / F(255.0):
            continue
        test_T = F(T * F(F(1.0) - alpha))
        if test_T < F(0.0001):
            break
        w = F(T * alpha)
        C = (C + w * color[g]).astype(F)
        T = test_T
    return (np.asarray(bg, F) * T + C).astype(F), examined
