// lcgs_math.cuh -- per-Gaussian arithmetic of the forward splat path (device side).
//
// Arithmetic contract: IEEE binary32, operations in the order written, no FMA contraction (every
// translation unit of the library is compiled with --fmad=false); fused operations appear only
// where __fmaf_rn / fma() is written.  This is what makes the integer outputs downstream (radii,
// tile counts, keys, ranges) bit-reproducible.  Reference behaviour restated from
// /root/reference (file:line cited per function); matrices are column-major, M*v is summed left to
// right.
//
// The functions are __host__ __device__ so that tests/host_mirror can run the very same source on
// the CPU (g++) against the oracle before any GPU time is spent; the product never calls them on
// the host.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LCGS_HD __host__ __device__ __forceinline__
#else
#define LCGS_HD inline
#endif

namespace lcgs_b200 {

// Host-derived per-view kernel parameters (GSProjector::forward, gs_projector/impl.cpp:34-42).
struct ViewParams {
    float view[16];  // world_to_local_matrix (camera.h:38-51), column-major
    float proj[16];  // projection_matrix(tanfovx, tanfovy, 0.1, 100) (camera.h:54-72)
    float tanfovx, tanfovy;
    float focalx, focaly;
    float cam_pos[3];
    int   width, height;
};

// ---------------------------------------------------------------------------------------------
// conversions with the semantics of cvt.rzi (LuisaCompute's UInt(float)/Int(float) on CUDA)
// ---------------------------------------------------------------------------------------------
LCGS_HD uint32_t f2u_rz(float f)
{
#if defined(__CUDA_ARCH__)
    return __float2uint_rz(f);
#else
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
#endif
}

LCGS_HD int32_t f2i_rz(float f)
{
#if defined(__CUDA_ARCH__)
    return __float2int_rz(f);
#else
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int32_t)0x80000000u;
    return (int32_t)f;
#endif
}

LCGS_HD uint32_t float_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

LCGS_HD float bits_float(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

LCGS_HD uint32_t clampu(uint32_t v, uint32_t lo, uint32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
LCGS_HD float    clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ---------------------------------------------------------------------------------------------
// tile rect (GSModule::mp_get_rect, lcgs/src/module.cpp:22-36; blocks = 16x16)
// Quirk Q1: the upper bound is clamped to grids-1 although the loops over it are exclusive.
// ---------------------------------------------------------------------------------------------
struct TileRect {
    uint32_t x0, y0, x1, y1;
};

LCGS_HD TileRect get_rect(float px, float py, int radius, uint32_t gx, uint32_t gy, uint32_t row0, uint32_t row1)
{
    const float fr = (float)radius;
    TileRect    r;
    r.x0 = clampu(f2u_rz((px - fr) / 16.0f), 0u, gx - 1u);
    r.y0 = clampu(f2u_rz((py - fr) / 16.0f), 0u, gy - 1u);
    r.x1 = clampu(f2u_rz(((px + fr) + 16.0f) - 1.0f) / 16u, 0u, gx - 1u);
    r.y1 = clampu(f2u_rz(((py + fr) + 16.0f) - 1.0f) / 16u, 0u, gy - 1u);
    // tile-row band of the multi-GPU split; identity for [0, gy)
    r.y0 = clampu(r.y0, row0, row1);
    r.y1 = clampu(r.y1, row0, row1);
    return r;
}

// ---------------------------------------------------------------------------------------------
// K2 shad_project_gs_focal (gs_projector/shader.cpp:82-139) + calc_cov (util/gaussian.hpp:15-28)
// + R_from_qvec (util/transform.hpp:188-212) + ewasplat_cov_focal (gaussian.hpp:52-70)
// + mp_cam_clamp (shader.cpp:146-158).
// ---------------------------------------------------------------------------------------------
struct ViewPoint {
    float x, y, z;       // p_view
    float ndc_x, ndc_y;  // p_proj.xy
    bool  visible;       // !(p_view.z < 0.2)
};

LCGS_HD ViewPoint view_transform(const ViewParams& vp, float px, float py, float pz)
{
    const float* V = vp.view;
    ViewPoint    o;
    o.x = ((px * V[0] + py * V[4]) + pz * V[8]) + V[12];
    o.y = ((px * V[1] + py * V[5]) + pz * V[9]) + V[13];
    o.z = ((px * V[2] + py * V[6]) + pz * V[10]) + V[14];
    const float phx = o.x * vp.proj[0];
    const float phy = o.y * vp.proj[5];
    const float p_w = 1.0f / (o.z + 1e-6f);
    o.ndc_x   = phx * p_w;
    o.ndc_y   = phy * p_w;
    o.visible = !(o.z < 0.2f);
    return o;
}

// 2D covariance in pixel^2 units: (cov[0][0], cov[0][1], cov[1][1])
LCGS_HD void ewa_cov2d(const ViewParams& vp, const ViewPoint& pv, float scale_modifier, float s0, float s1, float s2,
                       float qr, float qx, float qy, float qz, float* cov)
{
    const float sc[3] = { scale_modifier * s0, scale_modifier * s1, scale_modifier * s2 };
    const float x = qx, y = qy, z = qz, w = qr;  // rotq.yzwx: stored (r,x,y,z)
    float       R[3][3];
    R[0][0] = (1.0f - (2.0f * y) * y) - (2.0f * z) * z;
    R[0][1] = (2.0f * x) * y + (2.0f * z) * w;
    R[0][2] = (2.0f * x) * z - (2.0f * y) * w;
    R[1][0] = (2.0f * x) * y - (2.0f * z) * w;
    R[1][1] = (1.0f - (2.0f * x) * x) - (2.0f * z) * z;
    R[1][2] = (2.0f * y) * z + (2.0f * x) * w;
    R[2][0] = (2.0f * x) * z + (2.0f * y) * w;
    R[2][1] = (2.0f * y) * z - (2.0f * x) * w;
    R[2][2] = (1.0f - (2.0f * x) * x) - (2.0f * y) * y;
    float M[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) M[c][r] = sc[c] * R[c][r];
    float Sg[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) Sg[c][r] = (M[0][c] * M[0][r] + M[1][c] * M[1][r]) + M[2][c] * M[2][r];

    const float limx = 1.3f * vp.tanfovx;
    const float limy = 1.3f * vp.tanfovy;
    const float txtz = pv.x / pv.z;
    const float tytz = pv.y / pv.z;
    const float tx   = clampf(txtz, -limx, limx) * pv.z;
    const float ty   = clampf(tytz, -limy, limy) * pv.z;
    const float tz   = pv.z;

    const float J00 = vp.focalx / tz;
    const float J11 = vp.focaly / tz;
    const float J02 = (-vp.focalx * tx) / (tz * tz);
    const float J12 = (-vp.focaly * ty) / (tz * tz);
    // columns of transpose(mat3(view)) = camera right / up / front
    const float* V = vp.view;
    float        T0[3], T1[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        T0[r] = J00 * V[r * 4 + 0] + J02 * V[r * 4 + 2];
        T1[r] = J11 * V[r * 4 + 1] + J12 * V[r * 4 + 2];
    }
    float A[3][2];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        A[c][0] = (Sg[c][0] * T0[0] + Sg[c][1] * T0[1]) + Sg[c][2] * T0[2];
        A[c][1] = (Sg[c][0] * T1[0] + Sg[c][1] * T1[1]) + Sg[c][2] * T1[2];
    }
    cov[0] = (T0[0] * A[0][0] + T0[1] * A[1][0]) + T0[2] * A[2][0];
    cov[1] = (T0[0] * A[0][1] + T0[1] * A[1][1]) + T0[2] * A[2][1];
    cov[2] = (T1[0] * A[0][1] + T1[1] * A[1][1]) + T1[2] * A[2][1];
}

// ---------------------------------------------------------------------------------------------
// K3 shad_allocate_tiles (gs_tile_splatter/shader.cpp:102-163) for one visible Gaussian
// ---------------------------------------------------------------------------------------------
struct Splat2D {
    float    px, py;      // pixel-space mean (mp_ndc2pix, module.cpp:18-20; no half-pixel offset, Q2)
    float    conic[3];    // inverse of the low-passed covariance
    float    lambda_max;  // larger eigenvalue of the low-passed covariance
    int32_t  radius;
    TileRect rect;
    uint32_t tiles;
};

LCGS_HD Splat2D splat_from_cov(float ndc_x, float ndc_y, const float* cov, int W, int H, uint32_t gx, uint32_t gy,
                               uint32_t row0, uint32_t row1)
{
    Splat2D     s;
    const float a       = cov[0] + 0.3f;
    const float b       = cov[1];
    const float c       = cov[2] + 0.3f;
    const float det     = a * c - b * b;
    const float inv_det = 1.0f / (det + 1e-6f);
    s.conic[0]          = inv_det * c;
    s.conic[1]          = inv_det * (-b);
    s.conic[2]          = inv_det * a;
    const float mid     = 0.5f * (a + c);
    const float sq      = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float l1      = mid + sq;
    const float l2      = mid - sq;
    s.lambda_max        = fmaxf(l1, l2);
    s.radius            = f2i_rz(ceilf(3.0f * sqrtf(s.lambda_max)));
    s.px                = ((ndc_x + 1.0f) * (float)(uint32_t)W - 1.0f) * 0.5f;
    s.py                = ((ndc_y + 1.0f) * (float)(uint32_t)H - 1.0f) * 0.5f;
    s.rect              = get_rect(s.px, s.py, s.radius, gx, gy, row0, row1);
    s.tiles             = (s.rect.x1 - s.rect.x0) * (s.rect.y1 - s.rect.y0);
    return s;
}

// ---------------------------------------------------------------------------------------------
// K1 SH colour (sh_preprocessor.cpp:27-166, util/sh.hpp:12-138).  `sh` points at this Gaussian's
// coefficients, element (k, c) at sh[(k*3 + c) * stride].  Colour is clamped to [0,1] (Q3).
// ---------------------------------------------------------------------------------------------
#define LCGS_SH_C0 0.28209479177387814f
#define LCGS_SH_C1 0.4886025119029199f

template <typename ShAccessor>
LCGS_HD void sh_color(int deg, const float* cam_pos, float px, float py, float pz, const ShAccessor& sh, float* out)
{
    float res[3] = { sh(0, 0), sh(0, 1), sh(0, 2) };
    if (deg > -1) {
#pragma unroll
        for (int c = 0; c < 3; c++) res[c] = res[c] * LCGS_SH_C0;
        if (deg > 0) {
            const float dx = px - cam_pos[0], dy = py - cam_pos[1], dz = pz - cam_pos[2];
            const float len2 = (dx * dx + dy * dy) + dz * dz;
            const float inv  = 1.0f / sqrtf(len2);
            const float x = dx * inv, y = dy * inv, z = dz * inv;
#pragma unroll
            for (int c = 0; c < 3; c++)
                res[c] = res[c] + (-LCGS_SH_C1) * ((sh(1, c) * y - sh(2, c) * z) + sh(3, c) * x);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, yz = y * z, zz = z * z, zx = z * x, xy = x * y;
                const float k0 = 1.0925484305920792f * xy;
                const float k1 = -1.0925484305920792f * yz;
                const float k2 = 0.31539156525252005f * ((2.0f * zz - xx) - yy);
                const float k3 = -1.0925484305920792f * zx;
                const float k4 = 0.5462742152960396f * (xx - yy);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float l2 = (((k0 * sh(4, c) + k1 * sh(5, c)) + k2 * sh(6, c)) + k3 * sh(7, c)) + k4 * sh(8, c);
                    res[c]         = res[c] + l2;
                }
                if (deg > 2) {
                    const float m0 = -0.5900435899266435f * y * (3.0f * xx - yy);
                    const float m1 = 2.890611442640554f * xy * z;
                    const float m2 = -0.4570457994644658f * y * ((4.0f * zz - xx) - yy);
                    const float m3 = 0.3731763325901154f * z * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
                    const float m4 = -0.4570457994644658f * x * ((4.0f * zz - xx) - yy);
                    const float m5 = 1.445305721320277f * z * (xx - yy);
                    const float m6 = -0.5900435899266435f * x * (xx - 3.0f * yy);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float l3 =
                            (((((m0 * sh(9, c) + m1 * sh(10, c)) + m2 * sh(11, c)) + m3 * sh(12, c)) + m4 * sh(13, c)) +
                             m5 * sh(14, c)) +
                            m6 * sh(15, c);
                        res[c] = res[c] + l3;
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) res[c] = res[c] + 0.5f;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) out[c] = clampf(res[c], 0.0f, 1.0f);
}

// ---------------------------------------------------------------------------------------------
// The blend's alpha test as a threshold on `power`.
//
// The reference skips a (pixel, Gaussian) pair when min(0.99, opacity*exp(power)) < 1/255
// (gs_tile_splatter/shader.cpp:258-259).  exp is monotone, so per Gaussian this is
// `power < thr(opacity)`.  exp_rn() is the deterministic binary64 sequence the oracle uses for exp
// (correctly rounded to binary32 for all practical purposes); alpha_threshold() finds the exact
// boundary by bracketing around log(1/(255*opacity)) and bisecting over float bit patterns.  The
// blend kernel then decides with one compare per pair and uses MUFU.EX2 only for alpha's value.
// ---------------------------------------------------------------------------------------------
LCGS_HD float exp_rn(float xf)
{
    const double x = (double)xf;
    if (x != x) return xf;
    if (x < -104.0) return 0.0f;
    if (x > 89.0) return INFINITY;
    const double t = x * 1.4426950408889634;
    const double n = rint(t);
    double       r = fma(n, -6.93147180369123816490e-01, x);
    r              = fma(n, -1.90821492927058770002e-10, r);
    double p       = 1.6059043836821613e-10;
    p              = fma(p, r, 2.08767569878681e-09);
    p              = fma(p, r, 2.505210838544172e-08);
    p              = fma(p, r, 2.755731922398589e-07);
    p              = fma(p, r, 2.7557319223985893e-06);
    p              = fma(p, r, 2.48015873015873e-05);
    p              = fma(p, r, 0.0001984126984126984);
    p              = fma(p, r, 0.001388888888888889);
    p              = fma(p, r, 0.008333333333333333);
    p              = fma(p, r, 0.041666666666666664);
    p              = fma(p, r, 0.16666666666666666);
    p              = fma(p, r, 0.5);
    p              = fma(p, r, 1.0);
    p              = fma(p, r, 1.0);
    const long long          e    = (long long)n + 1023;
    const unsigned long long bits = (unsigned long long)e << 52;
    double                   s;
#if defined(__CUDA_ARCH__)
    s = __longlong_as_double((long long)bits);
#else
    memcpy(&s, &bits, 8);
#endif
    return (float)(p * s);
}

LCGS_HD bool alpha_passes(float op, float power)
{
    const float alpha = fminf(0.99f, op * exp_rn(power));
    return !(alpha < 1.0f / 255.0f);
}

// Smallest power <= 0 for which the alpha test passes; +inf if none (opacity < 1/255).
LCGS_HD float alpha_threshold(float op)
{
    // power 0: exp_rn(0) == 1 exactly, so the test is on min(0.99, opacity) itself
    if (fminf(0.99f, op) < 1.0f / 255.0f) return INFINITY;
    if (!(op <= 3.0e38f)) return -INFINITY;  // inf / NaN opacity: every power passes
    // bits of non-positive floats grow as the value decreases: 0x80000000 (-0) .. 0xFF800000 (-inf)
    const uint32_t U_ZERO = 0x80000000u, U_NINF = 0xFF800000u;
    float          est    = (float)log((double)(1.0f / 255.0f) / (double)op);
    if (!(est <= 0.0f)) est = -0.0f;
    uint32_t u0 = float_bits(est) | 0x80000000u;
    if (u0 > U_NINF) u0 = U_NINF;
    uint32_t lo, hi;  // passes(lo) == true, passes(hi) == false, lo < hi
    if (alpha_passes(op, bits_float(u0))) {
        lo            = u0;
        uint32_t step = 1u;
        for (;;) {
            const uint32_t c = (U_NINF - lo > step) ? lo + step : U_NINF;
            if (!alpha_passes(op, bits_float(c))) { hi = c; break; }
            if (c == U_NINF) return -INFINITY;
            lo = c;
            step <<= 1;
        }
    } else {
        hi            = u0;
        uint32_t step = 1u;
        for (;;) {
            const uint32_t c = (hi - U_ZERO > step) ? hi - step : U_ZERO;
            if (alpha_passes(op, bits_float(c))) { lo = c; break; }
            hi = c;  // c == U_ZERO cannot fail: checked on entry
            step <<= 1;
        }
    }
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (alpha_passes(op, bits_float(mid))) lo = mid; else hi = mid;
    }
    return bits_float(lo);
}

// ---------------------------------------------------------------------------------------------
// Conservative rectangle culling for the blend kernel.
//
// power(d) = a*dx^2 + b*dx*dy + c*dy^2 with d = mean - pixel, (a,b,c) = (-0.5*conic.x, -conic.y,
// -0.5*conic.z).  blend_power() is the canonical per-pixel evaluation (shared with the oracle).
// cull_rect() returns true only if NO pixel centre inside [x0,x1]x[y0,y1] can satisfy
// power >= thr, i.e. every such pair would be skipped by the alpha test anyway
// (gs_tile_splatter/shader.cpp:259), so dropping the Gaussian for that rectangle changes nothing.
// It maximises the (concave) quadratic over the rectangle exactly -- the maximiser lies on one of
// the at most two edges facing the mean -- and adds a margin that covers the rounding error of the
// per-pixel float evaluation (<= ~1e-6 * (|a|dx^2 + |c|dy^2)).
// ---------------------------------------------------------------------------------------------
LCGS_HD float blend_power(float a, float b, float c, float dx, float dy)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(b * dx, dy, __fmaf_rn(a * dx, dx, (c * dy) * dy));
#else
    return fmaf(b * dx, dy, fmaf(a * dx, dx, (c * dy) * dy));
#endif
}

// quotient used only to locate the maximiser along an edge; an error of a few ulp moves the
// evaluated maximum by O(ulp^2), far inside the margin, so the device may use the fast divide
LCGS_HD float cull_div(float x, float y)
{
#if defined(__CUDA_ARCH__)
    return __fdividef(x, y);
#else
    return x / y;
#endif
}

LCGS_HD bool cull_rect(float mx, float my, float a, float b, float c, float thr, float x0, float y0, float x1, float y1)
{
    if (!(thr <= 0.0f)) return thr > 0.0f;  // +inf: nothing ever passes; NaN: keep
    // only a strictly concave form is culled; anything degenerate is kept
    if (!(a < 0.0f && c < 0.0f && 4.0f * a * c - b * b > 0.0f)) return false;
    const float dxl = mx - x1, dxh = mx - x0;  // dx ranges over [dxl, dxh]
    const float dyl = my - y1, dyh = my - y0;
    const bool  in_x = dxl <= 0.0f && dxh >= 0.0f;
    const bool  in_y = dyl <= 0.0f && dyh >= 0.0f;
    if (in_x && in_y) return false;  // the mean is inside: power reaches 0
    float pmax = -INFINITY;
    if (!in_x) {
        const float xe = dxl > 0.0f ? dxl : dxh;  // edge facing the mean
        float       dy = cull_div(-(b * xe), 2.0f * c);
        dy             = fminf(fmaxf(dy, dyl), dyh);
        pmax           = fmaxf(pmax, (a * xe) * xe + (b * xe) * dy + (c * dy) * dy);
    }
    if (!in_y) {
        const float ye = dyl > 0.0f ? dyl : dyh;
        float       dx = cull_div(-(b * ye), 2.0f * a);
        dx             = fminf(fmaxf(dx, dxl), dxh);
        pmax           = fmaxf(pmax, (a * dx) * dx + (b * dx) * ye + (c * ye) * ye);
    }
    const float ex = fmaxf(dxl * dxl, dxh * dxh), ey = fmaxf(dyl * dyl, dyh * dyh);
    const float margin = 4e-6f * (-(a * ex) - (c * ey)) + 1e-6f;
    return pmax + margin < thr;
}

// ---------------------------------------------------------------------------------------------
// cull_rect() with the two divisions hoisted out: the blend kernel tests every Gaussian of a tile
// against the tile and then against each warp's patch, so the per-Gaussian quotients
//   ry = -b / (2c)   (dy of the maximiser per unit dx along a vertical edge)
//   rx = -b / (2a)   (dx of the maximiser per unit dy along a horizontal edge)
// are computed once in the preprocess and travel in the two spare record slots.  ry = NaN marks a
// Gaussian that must never be culled (non-concave form, non-finite mean/conic, NaN threshold);
// cull_rect_fast() is branch-free otherwise:
//   xe, ye   the offset of smallest magnitude inside [dxl,dxh] / [dyl,dyh] (0 when the mean's
//            column / row crosses the rectangle);
//   p1, p2   the form on the vertical / horizontal line through xe / ye at the clamped maximiser.
// Both points lie inside the rectangle and one of them is the exact maximiser (cull_rect()'s
// argument; when the mean is inside both are 0), so max(p1,p2) + margin bounds every per-pixel
// evaluation from above.  The margin is the same as cull_rect()'s.
// ---------------------------------------------------------------------------------------------
struct CullCoef {
    float ry, rx;
};

LCGS_HD float cull_fma(float x, float y, float z)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(x, y, z);
#else
    return fmaf(x, y, z);
#endif
}

LCGS_HD CullCoef cull_coef(float mx, float my, float a, float b, float c, float thr)
{
    CullCoef k;
    const float big = 1.0e30f;
    const bool  concave = a < 0.0f && c < 0.0f && a > -big && c > -big && 4.0f * a * c - b * b > 0.0f;
    const bool  finite  = fabsf(mx) < big && fabsf(my) < big;
    const bool  thr_ok  = thr <= 0.0f || thr > 0.0f;  // not NaN
    if (concave && finite && thr_ok) {
        k.ry = -b / (2.0f * c);
        k.rx = -b / (2.0f * a);
    } else {
        k.ry = NAN;
        k.rx = 0.0f;
    }
    return k;
}

LCGS_HD bool cull_rect_fast(float mx, float my, float a, float b, float c, float thr, float ry, float rx, float x0, float y0,
                            float x1, float y1)
{
    const float dxl = mx - x1, dxh = mx - x0;  // dx ranges over [dxl, dxh]
    const float dyl = my - y1, dyh = my - y0;
    const float xe = fminf(fmaxf(dxl, 0.0f), dxh);
    const float ye = fminf(fmaxf(dyl, 0.0f), dyh);
    const float ty = fminf(fmaxf(ry * xe, dyl), dyh);
    const float tx = fminf(fmaxf(rx * ye, dxl), dxh);
    const float p1 = cull_fma(c * ty, ty, cull_fma(b * xe, ty, (a * xe) * xe));
    const float p2 = cull_fma(a * tx, tx, cull_fma(b * tx, ye, (c * ye) * ye));
    const float mxa = fmaxf(fabsf(dxl), fabsf(dxh)), mya = fmaxf(fabsf(dyl), fabsf(dyh));
    const float e   = cull_fma(c, mya * mya, a * (mxa * mxa));  // -(|a| ex + |c| ey)
    const float margin = cull_fma(e, -4e-6f, 1e-6f);
    return ry == ry && fmaxf(p1, p2) + margin < thr;
}

}  // namespace lcgs_b200
