/*
 * lcgs_oracle.c -- CPU restatement of LuisaComputeGaussianSplatting's forward splat-render path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The product
 * (luisacomputegaussiansplatting_b200/csrc + include/lcgs_b200.h) never links or imports it.
 *
 * PARITY STATUS: the reference cannot be compiled here (LuisaCompute and lc_parallel_primitive
 * are un-vendored, un-pinned dependencies fetched at configure time: CMakeLists.txt:7-32), so this
 * oracle is a restatement of the reference's DSL kernels, not a build of them.  Only the camera
 * helpers are pinned by the reference's own known-answer tests (test/test_camera.cpp:48-144, see
 * tests/test_oracle_camera.py); the frame geometry of the output (last tile row / column never rendered,
 * vertical flip, zero background) is pinned by the reference's published render doc/mip360_bicycle_30000_cuda.png
 * (tests/test_reference_doc_image.py).  Every other stage is "parity unpinned" by the reference: there is
 * no golden vector, fixture or test for it upstream.  A second, independent numpy restatement
 * (tests/np_restatement.py) cross-checks this file stage by stage.
 *
 * Arithmetic contract (shared with the CUDA path): IEEE binary32, evaluation order exactly as
 * written below, NO fused multiply-add except where fmaf() is written explicitly (build with
 * -ffp-contract=off).  Matrices are column-major: m[c*4+r] (4x4) / m[c][r] (3x3), and
 * M*v = v.x*M[0] + v.y*M[1] + v.z*M[2] (+ v.w*M[3]) summed left to right, which is how
 * LuisaCompute's float3x3/float4x4 operators are defined (the reference relies on that:
 * lcgs/include/lcgs/util/gaussian.hpp:15-28,52-70).
 *
 * All citations "file:line" are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* types                                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* lcgs::Camera, lcgs/include/lcgs/util/camera.h:15-25 */
typedef struct {
    float position[3];
    float front[3];
    float up[3];
    float right[3];
    float fov;          /* degrees, vertical; default 60 */
    float aspect_ratio; /* W/H */
    int   width;
    int   height;
} orc_camera;

/* what GSProjector::forward derives on the host, lcgs/src/gs_projector/impl.cpp:34-42 */
typedef struct {
    float view[16]; /* world_to_local_matrix, column-major */
    float proj[16]; /* projection_matrix(tanfovx, tanfovy, 0.1, 100), column-major */
    float tanfovx, tanfovy;
    float focalx, focaly;
    float cam_pos[3];
    int   width, height;
} orc_view_params;

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                                */
/* ------------------------------------------------------------------------------------------ */

static inline float orc_dot3(const float* a, const float* b)
{
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

static inline void orc_cross3(const float* a, const float* b, float* o)
{
    float x = a[1] * b[2] - a[2] * b[1];
    float y = a[2] * b[0] - a[0] * b[2];
    float z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}

/* normalize(v) := v * (1 / sqrt(dot(v,v))) with IEEE sqrt and divide.  (LuisaCompute's CUDA
 * backend uses v * rsqrt(dot): same expression, unspecified rounding; we fix the rounding.) */
static inline void orc_normalize3(const float* v, float* o)
{
    float len2 = orc_dot3(v, v);
    float inv  = 1.0f / sqrtf(len2);
    o[0] = v[0] * inv; o[1] = v[1] * inv; o[2] = v[2] * inv;
}

static inline float orc_clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

/* UInt(float) as CUDA's cvt.rzi.u32.f32: NaN -> 0, negatives -> 0, >= 2^32 -> 2^32-1.
 * (C's float->unsigned conversion is UB outside the range; the reference's get_rect,
 * lcgs/src/module.cpp:31-35, feeds it negative values for every Gaussian left of the screen.) */
static inline uint32_t orc_f2u(float f)
{
    if (!(f > 0.0f)) return 0u; /* NaN, -x, +-0 */
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}

/* Int x = ceil(float) as cvt.rzi.s32.f32 of an already-integral float: saturating, NaN -> 0. */
static inline int32_t orc_f2i(float f)
{
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT_MAX;
    if (f <= -2147483648.0f) return INT_MIN;
    return (int32_t)f;
}

static inline uint32_t orc_clampu(uint32_t v, uint32_t lo, uint32_t hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

static inline uint32_t orc_float_bits(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

/* ------------------------------------------------------------------------------------------ */
/* 9.1 camera (the only part pinned by upstream tests)                                          */
/* ------------------------------------------------------------------------------------------ */

/* get_lookat_cam, camera.h:74-82 */
ORC_API void orc_get_lookat_cam(const float* pos, const float* target, const float* world_up, orc_camera* cam)
{
    float d[3] = { target[0] - pos[0], target[1] - pos[1], target[2] - pos[2] };
    float c[3];
    memcpy(cam->position, pos, sizeof(float) * 3);
    orc_normalize3(d, cam->front);
    orc_cross3(cam->front, world_up, c);
    orc_normalize3(c, cam->right);
    orc_cross3(cam->right, cam->front, c);
    orc_normalize3(c, cam->up);
    cam->fov          = 60.0f;
    cam->aspect_ratio = 1.0f;
    cam->width        = 512;
    cam->height       = 512;
}

/* local_to_world_matrix, camera.h:27-36: columns right, up, front, position */
ORC_API void orc_local_to_world_matrix(const orc_camera* cam, float* m)
{
    for (int r = 0; r < 3; r++) {
        m[0 * 4 + r] = cam->right[r];
        m[1 * 4 + r] = cam->up[r];
        m[2 * 4 + r] = cam->front[r];
        m[3 * 4 + r] = cam->position[r];
    }
    m[0 * 4 + 3] = 0.0f; m[1 * 4 + 3] = 0.0f; m[2 * 4 + 3] = 0.0f; m[3 * 4 + 3] = 1.0f;
}

/* world_to_local_matrix, camera.h:38-51 */
ORC_API void orc_world_to_local_matrix(const orc_camera* cam, float* m)
{
    float tx = -orc_dot3(cam->position, cam->right);
    float ty = -orc_dot3(cam->position, cam->up);
    float tz = -orc_dot3(cam->position, cam->front);
    for (int c = 0; c < 3; c++) {
        m[c * 4 + 0] = cam->right[c];
        m[c * 4 + 1] = cam->up[c];
        m[c * 4 + 2] = cam->front[c];
        m[c * 4 + 3] = 0.0f;
    }
    m[12] = tx; m[13] = ty; m[14] = tz; m[15] = 1.0f;
}

/* projection_matrix, camera.h:54-72 */
ORC_API void orc_projection_matrix(float tanfovx, float tanfovy, float znear, float zfar, float* m)
{
    float zsign   = 1.0f;
    float fx      = 1.0f / tanfovx;
    float fy      = 1.0f / tanfovy;
    float z_range = zfar - znear;
    float a       = zfar / z_range;
    float b       = -zfar * znear / z_range;
    memset(m, 0, sizeof(float) * 16);
    m[0]  = fx;
    m[5]  = fy;
    m[10] = a * zsign;
    m[11] = zsign;
    m[14] = b;
}

/* float4x4 * float4, column-major, left-to-right sum */
ORC_API void orc_mat4_mul_vec4(const float* m, const float* v, float* o)
{
    float t[4];
    for (int r = 0; r < 4; r++)
        t[r] = ((v[0] * m[0 * 4 + r] + v[1] * m[1 * 4 + r]) + v[2] * m[2 * 4 + r]) + v[3] * m[3 * 4 + r];
    memcpy(o, t, sizeof(t));
}

/* host part of GSProjector::forward, gs_projector/impl.cpp:34-42 (+ SHProcessor::process passing
 * camera.position, sh_preprocessor.cpp:182) */
ORC_API void orc_view_params_from_camera(const orc_camera* cam, orc_view_params* vp)
{
    float fovy    = cam->fov / 180.0f * 3.1415926536f;
    float tanfovy = tanf(fovy * 0.5f);
    float tanfovx = tanfovy * cam->aspect_ratio;
    orc_world_to_local_matrix(cam, vp->view);
    orc_projection_matrix(tanfovx, tanfovy, 0.1f, 100.0f, vp->proj);
    vp->tanfovx = tanfovx;
    vp->tanfovy = tanfovy;
    vp->focalx  = (float)cam->width / (2.0f * tanfovx);
    vp->focaly  = (float)cam->height / (2.0f * tanfovy);
    memcpy(vp->cam_pos, cam->position, sizeof(float) * 3);
    vp->width  = cam->width;
    vp->height = cam->height;
}

/* ------------------------------------------------------------------------------------------ */
/* 9.2 SH colour (K1)  sh_preprocessor.cpp:27-166, util/sh.hpp:12-138                           */
/* ------------------------------------------------------------------------------------------ */

static const float ORC_SH_C0    = 0.28209479177387814f;
static const float ORC_SH_C1    = 0.4886025119029199f;
static const float ORC_SH_C2[5] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                    -1.0925484305920792f, 0.5462742152960396f };
static const float ORC_SH_C3[7] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                    0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                    -0.5900435899266435f };

/* colour of one Gaussian; sh points at its (deg+1)^2 x 3 block */
static void orc_sh_one(int deg, const float* cam_pos, const float* pos, const float* sh, float* out)
{
    float res[3] = { sh[0], sh[1], sh[2] };
    if (deg > -1) {
        for (int c = 0; c < 3; c++) res[c] = sh[c] * ORC_SH_C0; /* sh.hpp:30-34 */
        if (deg > 0) {
            float d[3] = { pos[0] - cam_pos[0], pos[1] - cam_pos[1], pos[2] - cam_pos[2] };
            float dir[3];
            orc_normalize3(d, dir);
            float x = dir[0], y = dir[1], z = dir[2];
            const float* s1 = sh + 3;
            const float* s2 = sh + 6;
            const float* s3 = sh + 9;
            /* -SH_C1 * (sh_10*y - sh_11*z + sh_12*x), sh.hpp:42-50 */
            for (int c = 0; c < 3; c++)
                res[c] = res[c] + (-ORC_SH_C1) * ((s1[c] * y - s2[c] * z) + s3[c] * x);
            if (deg > 1) {
                float xx = x * x, yy = y * y, yz = y * z, zz = z * z, zx = z * x, xy = x * y;
                /* sh.hpp:67-84: scalar coefficient first, then times the float3 */
                float k0 = ORC_SH_C2[0] * xy;
                float k1 = ORC_SH_C2[1] * yz;
                float k2 = ORC_SH_C2[2] * ((2.0f * zz - xx) - yy);
                float k3 = ORC_SH_C2[3] * zx;
                float k4 = ORC_SH_C2[4] * (xx - yy);
                for (int c = 0; c < 3; c++) {
                    float l2 = (((k0 * sh[12 + c] + k1 * sh[15 + c]) + k2 * sh[18 + c]) + k3 * sh[21 + c]) +
                               k4 * sh[24 + c];
                    res[c] = res[c] + l2;
                }
                if (deg > 2) {
                    /* sh.hpp:119-138 */
                    float m0 = ORC_SH_C3[0] * y * (3.0f * xx - yy);
                    float m1 = ORC_SH_C3[1] * xy * z;
                    float m2 = ORC_SH_C3[2] * y * ((4.0f * zz - xx) - yy);
                    float m3 = ORC_SH_C3[3] * z * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
                    float m4 = ORC_SH_C3[4] * x * ((4.0f * zz - xx) - yy);
                    float m5 = ORC_SH_C3[5] * z * (xx - yy);
                    float m6 = ORC_SH_C3[6] * x * (xx - 3.0f * yy);
                    for (int c = 0; c < 3; c++) {
                        float l3 = (((((m0 * sh[27 + c] + m1 * sh[30 + c]) + m2 * sh[33 + c]) + m3 * sh[36 + c]) +
                                     m4 * sh[39 + c]) +
                                    m5 * sh[42 + c]) +
                                   m6 * sh[45 + c];
                        res[c] = res[c] + l3;
                    }
                }
            }
        }
        for (int c = 0; c < 3; c++) res[c] = res[c] + 0.5f; /* sh_preprocessor.cpp:150 */
    }
    for (int c = 0; c < 3; c++) out[c] = orc_clampf(res[c], 0.0f, 1.0f); /* :153 */
}

/* K1 shad_sh_process, sh_preprocessor.cpp:159-166.  sh is [P][(deg+1)^2][3]. */
ORC_API void orc_sh_process(int P, int deg, const float* cam_pos, const float* pos, const float* sh, float* color)
{
    int feat = (deg + 1) * (deg + 1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++)
        orc_sh_one(deg, cam_pos, pos + 3 * (size_t)i, sh + (size_t)i * feat * 3, color + 3 * (size_t)i);
}

/* ------------------------------------------------------------------------------------------ */
/* 9.3 projection (K2, focal variant)  gs_projector/shader.cpp:82-139                           */
/* ------------------------------------------------------------------------------------------ */

/* Returns 1 if the Gaussian passes the near cull and its outputs were written, else 0 (nothing
 * written, like the reference's early $return at shader.cpp:121). */
static int orc_project_one(const orc_view_params* vp, float scale_modifier, const float* pos, const float* s,
                           const float* rotq, float* mean_ndc, float* depth, float* cov2d)
{
    const float* V = vp->view;
    /* p_view_hom = view * (mean,1)  (shader.cpp:109-112) */
    float pv[3];
    for (int r = 0; r < 3; r++)
        pv[r] = ((pos[0] * V[0 * 4 + r] + pos[1] * V[1 * 4 + r]) + pos[2] * V[2 * 4 + r]) + V[3 * 4 + r];
    /* p_proj_hom = proj * p_view_hom: x*fx, y*fy, ., w = z  (camera.h:66-71) */
    float phx = pv[0] * vp->proj[0];
    float phy = pv[1] * vp->proj[5];
    float phw = pv[2];
    float p_w = 1.0f / (phw + 1e-6f); /* shader.cpp:116 */
    float ndc_x = phx * p_w;
    float ndc_y = phy * p_w;
    if (pv[2] < 0.2f) return 0; /* shader.cpp:121 */
    *depth      = pv[2];
    mean_ndc[0] = ndc_x;
    mean_ndc[1] = ndc_y;

    /* 3D covariance: calc_cov, gaussian.hpp:15-28; R_from_qvec, transform.hpp:188-212 */
    float sc[3] = { scale_modifier * s[0], scale_modifier * s[1], scale_modifier * s[2] };
    float x = rotq[1], y = rotq[2], z = rotq[3], w = rotq[0]; /* rotq.yzwx: file order is r,x,y,z */
    float R[3][3];                                              /* R[c][r] */
    R[0][0] = (1.0f - (2.0f * y) * y) - (2.0f * z) * z;
    R[0][1] = (2.0f * x) * y + (2.0f * z) * w;
    R[0][2] = (2.0f * x) * z - (2.0f * y) * w;
    R[1][0] = (2.0f * x) * y - (2.0f * z) * w;
    R[1][1] = (1.0f - (2.0f * x) * x) - (2.0f * z) * z;
    R[1][2] = (2.0f * y) * z + (2.0f * x) * w;
    R[2][0] = (2.0f * x) * z + (2.0f * y) * w;
    R[2][1] = (2.0f * y) * z - (2.0f * x) * w;
    R[2][2] = (1.0f - (2.0f * x) * x) - (2.0f * y) * y;
    float M[3][3]; /* M = R*S: column c of M is scale[c] * R[c] */
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) M[c][r] = sc[c] * R[c][r];
    float Sg[3][3]; /* Sigma = M * M^T: Sigma[c][r] = sum_k M[k][c]*M[k][r] */
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) Sg[c][r] = (M[0][c] * M[0][r] + M[1][c] * M[1][r]) + M[2][c] * M[2][r];

    /* mp_cam_clamp, shader.cpp:146-158 */
    float limx = 1.3f * vp->tanfovx;
    float limy = 1.3f * vp->tanfovy;
    float txtz = pv[0] / pv[2];
    float tytz = pv[1] / pv[2];
    float tx   = orc_clampf(txtz, -limx, limx) * pv[2];
    float ty   = orc_clampf(tytz, -limy, limy) * pv[2];
    float tz   = pv[2];

    /* ewasplat_cov_focal, gaussian.hpp:52-70 */
    float J00 = vp->focalx / tz;
    float J11 = vp->focaly / tz;
    float J02 = (-vp->focalx * tx) / (tz * tz);
    float J12 = (-vp->focaly * ty) / (tz * tz);
    /* W = transpose(mat3(view)): its columns are the camera's right, up, front axes */
    float Wc[3][3];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) Wc[c][r] = V[r * 4 + c];
    /* T = W*J: T[0] = J00*W[0] + J02*W[2], T[1] = J11*W[1] + J12*W[2], T[2] = 0 */
    float T0[3], T1[3];
    for (int r = 0; r < 3; r++) {
        T0[r] = J00 * Wc[0][r] + J02 * Wc[2][r];
        T1[r] = J11 * Wc[1][r] + J12 * Wc[2][r];
    }
    /* A = T^T * Sigma: A[c][r] = sum_k Sigma[c][k] * T[r][k]  (only rows r = 0,1 are non-zero) */
    float A[3][2];
    for (int c = 0; c < 3; c++) {
        A[c][0] = (Sg[c][0] * T0[0] + Sg[c][1] * T0[1]) + Sg[c][2] * T0[2];
        A[c][1] = (Sg[c][0] * T1[0] + Sg[c][1] * T1[1]) + Sg[c][2] * T1[2];
    }
    /* cov = A * T: cov[c][r] = sum_k T[c][k] * A[k][r] */
    cov2d[0] = (T0[0] * A[0][0] + T0[1] * A[1][0]) + T0[2] * A[2][0]; /* cov[0][0] */
    cov2d[1] = (T0[0] * A[0][1] + T0[1] * A[1][1]) + T0[2] * A[2][1]; /* cov[0][1] */
    cov2d[2] = (T1[0] * A[0][1] + T1[1] * A[1][1]) + T1[2] * A[2][1]; /* cov[1][1] */
    return 1;
}

/* K2.  Buffers are caller-owned; culled Gaussians are left untouched (reference behaviour, quirk
 * Q6) -- the harness zero-fills the outputs first, which is the defined behaviour of the build. */
ORC_API void orc_project(int P, const float* pos, const float* scale, const float* rotq, float scale_modifier,
                         const orc_view_params* vp, float* means_2d, float* depth, float* covs_2d)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        size_t k = (size_t)i;
        orc_project_one(vp, scale_modifier, pos + 3 * k, scale + 3 * k, rotq + 4 * k, means_2d + 2 * k, depth + k,
                        covs_2d + 3 * k);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* 9.4 radius / rect / tile count (K3)  gs_tile_splatter/shader.cpp:102-163, module.cpp:18-36   */
/* ------------------------------------------------------------------------------------------ */

/* mp_get_rect, module.cpp:22-36, blocks = (16,16) */
static inline void orc_get_rect(float px, float py, int radius, uint32_t gx, uint32_t gy, uint32_t* mn, uint32_t* mx)
{
    float fr = (float)radius;
    mn[0] = orc_clampu(orc_f2u((px - fr) / 16.0f), 0u, gx - 1u);
    mn[1] = orc_clampu(orc_f2u((py - fr) / 16.0f), 0u, gy - 1u);
    mx[0] = orc_clampu(orc_f2u(((px + fr) + 16.0f) - 1.0f) / 16u, 0u, gx - 1u);
    mx[1] = orc_clampu(orc_f2u(((py + fr) + 16.0f) - 1.0f) / 16u, 0u, gy - 1u);
}

/* Extension used only by the tile-row sharding: restrict the y range of a rect to the
 * band of tile rows [row0,row1).  With row0 = 0, row1 = gy it is the identity. */
static inline void orc_clip_rows(uint32_t* mn, uint32_t* mx, uint32_t row0, uint32_t row1)
{
    mn[1] = orc_clampu(mn[1], row0, row1);
    mx[1] = orc_clampu(mx[1], row0, row1);
}

ORC_API void orc_allocate_tiles(int P, int W, int H, const float* depth, float* means_2d, float* covs_2d,
                                uint32_t* tiles_touched, int32_t* radii, int row0, int row1)
{
    uint32_t gx = (uint32_t)((W + 15) / 16), gy = (uint32_t)((H + 15) / 16);
    if (row1 < 0) row1 = (int)gy;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        size_t k = (size_t)i;
        radii[k]         = 0;
        tiles_touched[k] = 0u;
        if (depth[k] < 0.2f) continue; /* shader.cpp:124 */
        float ndc_x = means_2d[2 * k], ndc_y = means_2d[2 * k + 1];
        float a = covs_2d[3 * k] + 0.3f; /* low-pass, :139-140 */
        float b = covs_2d[3 * k + 1];
        float c = covs_2d[3 * k + 2] + 0.3f;
        float det     = a * c - b * b;
        float inv_det = 1.0f / (det + 1e-6f);
        float conic[3] = { inv_det * c, inv_det * (-b), inv_det * a };
        float mid      = 0.5f * (a + c);
        float sq       = sqrtf(fmaxf(0.1f, mid * mid - det));
        float l1       = mid + sq;
        float l2       = mid - sq;
        int   radius   = orc_f2i(ceilf(3.0f * sqrtf(fmaxf(l1, l2))));
        /* mp_ndc2pix, module.cpp:18-20 */
        float px = ((ndc_x + 1.0f) * (float)(uint32_t)W - 1.0f) * 0.5f;
        float py = ((ndc_y + 1.0f) * (float)(uint32_t)H - 1.0f) * 0.5f;
        uint32_t mn[2], mx[2];
        orc_get_rect(px, py, radius, gx, gy, mn, mx);
        orc_clip_rows(mn, mx, (uint32_t)row0, (uint32_t)row1);
        radii[k]         = radius;
        tiles_touched[k] = (mx[0] - mn[0]) * (mx[1] - mn[1]);
        covs_2d[3 * k]     = conic[0];
        covs_2d[3 * k + 1] = conic[1];
        covs_2d[3 * k + 2] = conic[2];
        means_2d[2 * k]     = px;
        means_2d[2 * k + 1] = py;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* 9.5 scan, keys, sort (K4-K7)                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* lcpp DeviceScan<>::InclusiveSum (gs_tile_splatter/impl.cpp:104): u32, wraps silently */
ORC_API void orc_inclusive_sum_u32(const uint32_t* in, uint32_t* out, int n)
{
    uint32_t acc = 0;
    for (int i = 0; i < n; i++) {
        acc += in[i];
        out[i] = acc;
    }
}

/* K6 shad_copy_with_keys, gs_tile_splatter/shader.cpp:26-69.  Tile ids are band-local
 * ((y-row0)*gx + x); with the full band this is the reference's x + y*grids.x. */
ORC_API void orc_copy_with_keys(int P, int W, int H, const float* means_2d_pix, const uint32_t* offsets,
                                const int32_t* radii, const float* depth, uint64_t* keys, uint32_t* vals, int row0,
                                int row1)
{
    uint32_t gx = (uint32_t)((W + 15) / 16), gy = (uint32_t)((H + 15) / 16);
    if (row1 < 0) row1 = (int)gy;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int i = 0; i < P; i++) {
        size_t k = (size_t)i;
        int radius = radii[k];
        if (radius <= 0) continue;
        uint32_t off = (i >= 1) ? offsets[k - 1] : 0u;
        uint32_t mn[2], mx[2];
        orc_get_rect(means_2d_pix[2 * k], means_2d_pix[2 * k + 1], radius, gx, gy, mn, mx);
        orc_clip_rows(mn, mx, (uint32_t)row0, (uint32_t)row1);
        uint64_t dbits = (uint64_t)orc_float_bits(depth[k]);
        for (uint32_t y = mn[1]; y < mx[1]; y++)
            for (uint32_t x = mn[0]; x < mx[0]; x++) {
                uint64_t key = (uint64_t)(x + (y - (uint32_t)row0) * gx);
                key <<= 32;
                key |= dbits;
                keys[off] = key;
                vals[off] = (uint32_t)i;
                off++;
            }
    }
}

/* lcpp DeviceRadixSort<>::SortPairs<ulong,uint> (gs_tile_splatter/impl.cpp:135-143): ascending by
 * the full 64-bit key, ties keep input order (stable) -- what any LSD radix sort yields.
 * Parallel LSD radix, 8-bit digits, passes whose digit is constant over all keys are skipped
 * (result-identical).  tmp_* are n-sized scratch. */
ORC_API void orc_sort_pairs_u64_u32(const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                                    uint32_t* vals_out, size_t n)
{
    if (n == 0) return;
    uint64_t* kb[2];
    uint32_t* vb[2];
    kb[0] = (uint64_t*)malloc(n * sizeof(uint64_t));
    vb[0] = (uint32_t*)malloc(n * sizeof(uint32_t));
    kb[1] = (uint64_t*)malloc(n * sizeof(uint64_t));
    vb[1] = (uint32_t*)malloc(n * sizeof(uint32_t));
    memcpy(kb[0], keys_in, n * sizeof(uint64_t));
    memcpy(vb[0], vals_in, n * sizeof(uint32_t));
    int cur = 0;
#ifdef _OPENMP
    int nt = omp_get_max_threads();
#else
    int nt = 1;
#endif
    size_t* hist = (size_t*)malloc((size_t)nt * 256 * sizeof(size_t));
    for (int pass = 0; pass < 8; pass++) {
        int shift = pass * 8;
        const uint64_t* ks = kb[cur];
        const uint32_t* vs = vb[cur];
        uint64_t*       kd = kb[cur ^ 1];
        uint32_t*       vd = vb[cur ^ 1];
        memset(hist, 0, (size_t)nt * 256 * sizeof(size_t));
        int trivial = 0;
#pragma omp parallel num_threads(nt)
        {
#ifdef _OPENMP
            int t = omp_get_thread_num();
#else
            int t = 0;
#endif
            size_t  lo = n * (size_t)t / (size_t)nt, hi = n * (size_t)(t + 1) / (size_t)nt;
            size_t* h  = hist + (size_t)t * 256;
            for (size_t i = lo; i < hi; i++) h[(ks[i] >> shift) & 0xFF]++;
#pragma omp barrier
#pragma omp single
            {
                size_t acc = 0;
                for (int d = 0; d < 256; d++) {
                    size_t tot = 0;
                    for (int tt = 0; tt < nt; tt++) {
                        size_t c               = hist[(size_t)tt * 256 + d];
                        hist[(size_t)tt * 256 + d] = acc + tot;
                        tot += c;
                    }
                    if (tot == n) trivial = 1;
                    acc += tot;
                }
            }
            if (!trivial)
                for (size_t i = lo; i < hi; i++) {
                    size_t dst = h[(ks[i] >> shift) & 0xFF]++;
                    kd[dst]    = ks[i];
                    vd[dst]    = vs[i];
                }
        }
        if (!trivial) cur ^= 1;
    }
    memcpy(keys_out, kb[cur], n * sizeof(uint64_t));
    memcpy(vals_out, vb[cur], n * sizeof(uint32_t));
    free(hist);
    free(kb[0]); free(vb[0]); free(kb[1]); free(vb[1]);
}

/* ------------------------------------------------------------------------------------------ */
/* 9.6 ranges (K8) and blend (K9)                                                               */
/* ------------------------------------------------------------------------------------------ */

/* fill ranges=0 (impl.cpp:147) + shad_get_ranges, gs_tile_splatter/shader.cpp:71-100 */
ORC_API void orc_get_ranges(size_t n, const uint64_t* keys, uint32_t* ranges, int num_tiles)
{
    memset(ranges, 0, (size_t)num_tiles * 2 * sizeof(uint32_t));
    for (size_t k = 0; k < n; k++) {
        uint32_t cur = (uint32_t)(keys[k] >> 32);
        if (k == 0) {
            ranges[2 * cur] = 0u;
        } else {
            uint32_t prev = (uint32_t)(keys[k - 1] >> 32);
            if (cur != prev) {
                ranges[2 * prev + 1] = (uint32_t)k;
                ranges[2 * cur]      = (uint32_t)k;
            }
        }
        if (k == n - 1) ranges[2 * cur + 1] = (uint32_t)n;
    }
}

/* exp() of the blend.  The reference calls the DSL's exp (gs_tile_splatter/shader.cpp:258), whose
 * rounding is backend-defined (CUDA expf/__expf, HLSL exp, Metal fast::exp).  The oracle fixes it
 * as a deterministic sequence of IEEE binary64 operations rounded once to binary32 (a correctly
 * rounded expf for all practical purposes, hence monotone), so that the alpha >= 1/255 decision
 * (shader.cpp:259) is a reproducible threshold on `power`.  The CUDA path uses this same sequence
 * once per Gaussian to precompute that threshold, and MUFU.EX2 for alpha's value. */
ORC_API float orc_exp(float xf)
{
    double x = (double)xf;
    if (x != x) return xf;
    if (x < -104.0) return 0.0f;
    if (x > 89.0) return INFINITY;
    double t = x * 1.4426950408889634;
    double n = nearbyint(t);
    double r = fma(n, -6.93147180369123816490e-01, x);
    r        = fma(n, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10; /* 1/13! */
    p        = fma(p, r, 2.08767569878681e-09);  /* 1/12! */
    p        = fma(p, r, 2.505210838544172e-08); /* 1/11! */
    p        = fma(p, r, 2.755731922398589e-07); /* 1/10! */
    p        = fma(p, r, 2.7557319223985893e-06); /* 1/9! */
    p        = fma(p, r, 2.48015873015873e-05);   /* 1/8! */
    p        = fma(p, r, 0.0001984126984126984);  /* 1/7! */
    p        = fma(p, r, 0.001388888888888889);   /* 1/6! */
    p        = fma(p, r, 0.008333333333333333);   /* 1/5! */
    p        = fma(p, r, 0.041666666666666664);   /* 1/4! */
    p        = fma(p, r, 0.16666666666666666);    /* 1/3! */
    p        = fma(p, r, 0.5);
    p        = fma(p, r, 1.0);
    p        = fma(p, r, 1.0);
    int64_t  e    = (int64_t)n + 1023;
    uint64_t bits = (uint64_t)e << 52;
    double   s;
    memcpy(&s, &bits, 8);
    return (float)(p * s);
}

/* the alpha test of the blend: contributes iff min(0.99, op*exp(power)) >= 1/255 */
static inline int orc_alpha_passes(float op, float power, float* alpha_out)
{
    float alpha = fminf(0.99f, op * orc_exp(power));
    *alpha_out  = alpha;
    return !(alpha < 1.0f / 255.0f);
}

/* Smallest power <= 0 (as a float) for which the alpha test passes for this opacity, or +inf if
 * none does.  Brute-force reference for the per-Gaussian threshold the CUDA path precomputes. */
ORC_API float orc_alpha_threshold(float op)
{
    float a;
    if (!orc_alpha_passes(op, 0.0f, &a)) return INFINITY;
    /* bisection over the bit patterns of non-positive floats: bits(-0.0)=0x80000000 .. bits(-inf) */
    uint32_t lo = 0x80000000u; /* passes */
    uint32_t hi = 0xFF800000u; /* -inf: exp = 0 -> fails (op*0 = 0 < 1/255) unless op is inf/nan */
    float    fhi;
    memcpy(&fhi, &hi, 4);
    if (orc_alpha_passes(op, fhi, &a)) return fhi;
    while (hi - lo > 1u) {
        uint32_t mid = lo + (hi - lo) / 2u;
        float    fm;
        memcpy(&fm, &mid, 4);
        if (orc_alpha_passes(op, fm, &a)) lo = mid; else hi = mid;
    }
    float r;
    memcpy(&r, &lo, 4);
    return r;
}

/* Counts, over floats in [lo,hi], adjacent pairs where orc_exp decreases as x increases.  Must be
 * 0 for the threshold formulation to be equivalent to the per-pair test. */
ORC_API long orc_exp_monotonicity_violations(float lo, float hi)
{
    uint32_t blo = orc_float_bits(lo), bhi = orc_float_bits(hi); /* both negative: blo >= bhi */
    long bad = 0;
    if (!(lo < 0.0f) || !(hi < 0.0f) || blo < bhi) return -1;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int64_t b = (int64_t)bhi; b < (int64_t)blo; b++) {
        uint32_t u0 = (uint32_t)b, u1 = (uint32_t)b + 1u; /* u1 is the more negative float */
        float    f0, f1;
        memcpy(&f0, &u0, 4);
        memcpy(&f1, &u1, 4);
        if (orc_exp(f1) > orc_exp(f0)) bad++;
    }
    return bad;
}

/* K9 m_forward_render_shader, gs_tile_splatter/shader.cpp:171-288.  One 16x16 block per tile; the
 * staging through Shared<> arrays does not change any result, so each pixel simply walks its
 * tile's list.  power uses two explicit fused ops (see DESIGN.md "canonical arithmetic"):
 *   power = fma(-(con.y*d.x), d.y, fma((-0.5*con.x)*d.x, d.x, ((-0.5*con.z)*d.y)*d.y))
 * which is the reference's -0.5*(con.x*dx*dx + con.z*dy*dy) - con.y*dx*dy.
 * img is planar CHW; rows [row0*16, min(H,row1*16)) are written, tile ids are band-local.
 * n_examined (optional, W*H) receives the reference's `contributor` counter. */
/* Totals of the last orc_blend call, for the blend roofline of SURVEY.md 8d: E = (pixel, list entry) pairs examined
 * (the reference's `contributor` counter, shader.cpp:219,252), E_alpha = pairs that passed the alpha test,
 * E_contrib = pairs that were blended into the pixel. */
static unsigned long long g_blend_stats[3];
ORC_API void orc_blend_stats(unsigned long long out[3])
{
    out[0] = g_blend_stats[0]; out[1] = g_blend_stats[1]; out[2] = g_blend_stats[2];
}

ORC_API void orc_blend(int W, int H, const float* bg, const uint32_t* ranges, const uint32_t* point_list,
                       const float* means_2d, const float* conic, const float* opacity, const float* color,
                       float* img, uint32_t* n_examined, int row0, int row1)
{
    int gx = (W + 15) / 16, gy = (H + 15) / 16;
    if (row1 < 0) row1 = gy;
    size_t plane = (size_t)W * (size_t)H;
    int    ntile = gx * (row1 - row0);
    unsigned long long tot_e = 0, tot_a = 0, tot_c = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot_e, tot_a, tot_c)
    for (int t = 0; t < ntile; t++) {
        int      bx = t % gx, by = row0 + t / gx;
        uint32_t start = ranges[2 * t], end = ranges[2 * t + 1];
        for (int ly = 0; ly < 16; ly++)
            for (int lx = 0; lx < 16; lx++) {
                int px = bx * 16 + lx, py = by * 16 + ly;
                if (px >= W || py >= H) continue; /* inside == false -> done from the start */
                float    pxf = (float)px, pyf = (float)py;
                float    T = 1.0f, C[3] = { 0.0f, 0.0f, 0.0f };
                uint32_t contributor = 0;
                for (uint32_t k = start; k < end; k++) {
                    contributor++;
                    uint32_t id = point_list[k];
                    float dx = means_2d[2 * (size_t)id] - pxf;
                    float dy = means_2d[2 * (size_t)id + 1] - pyf;
                    float cx = conic[3 * (size_t)id], cy = conic[3 * (size_t)id + 1], cz = conic[3 * (size_t)id + 2];
                    float op = opacity[id];
                    float power = fmaf(-(cy * dx), dy, fmaf((-0.5f * cx) * dx, dx, ((-0.5f * cz) * dy) * dy));
                    if (power > 0.0f) continue;
                    float alpha;
                    if (!orc_alpha_passes(op, power, &alpha)) continue;
                    tot_a++;
                    float test_T = T * (1.0f - alpha);
                    if (test_T < 0.0001f) break; /* done = true; this entry is not blended */
                    tot_c++;
                    float wgt = T * alpha;
                    C[0] = C[0] + wgt * color[3 * (size_t)id];
                    C[1] = C[1] + wgt * color[3 * (size_t)id + 1];
                    C[2] = C[2] + wgt * color[3 * (size_t)id + 2];
                    T    = test_T;
                }
                size_t pix = (size_t)px + (size_t)W * (size_t)py;
                for (int c = 0; c < 3; c++) img[pix + c * plane] = bg[c] * T + C[c];
                if (n_examined) n_examined[pix] = contributor;
                tot_e += contributor;
            }
    }
    g_blend_stats[0] = tot_e; g_blend_stats[1] = tot_a; g_blend_stats[2] = tot_c;
}

/* ------------------------------------------------------------------------------------------ */
/* whole frame (what main.cpp:266-308 + GSTileSplatter::forward do), with per-stage wall time   */
/* ------------------------------------------------------------------------------------------ */

static double orc_now(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

/* stage_ms[8]: sh, project, allocate_tiles, scan, copy_with_keys, sort, ranges, blend.
 * All buffers caller-owned; keys/vals lists have `capacity` entries.  Returns num_rendered, or -1
 * if it exceeds capacity (the reference would overrun its lists: main.cpp:245, impl.cpp:112-115).
 * When num_rendered == 0 the image is left untouched (impl.cpp:109). */
ORC_API long orc_forward(int P, int sh_deg, const float* pos, const float* scale, const float* rotq, const float* sh,
                         const float* opacity, float scale_modifier, const orc_view_params* vp, const float* bg,
                         float* color, float* means_2d, float* depth, float* conic, uint32_t* tiles_touched,
                         int32_t* radii, uint32_t* offsets, uint64_t* keys_unsorted, uint32_t* vals_unsorted,
                         uint64_t* keys_sorted, uint32_t* vals_sorted, size_t capacity, uint32_t* ranges, float* img,
                         uint32_t* n_examined, int row0, int row1, double* stage_ms)
{
    int    W = vp->width, H = vp->height;
    int    gx = (W + 15) / 16, gy = (H + 15) / 16;
    double t0, t1;
    if (row1 < 0) row1 = gy;
    t0 = orc_now();
    orc_sh_process(P, sh_deg, vp->cam_pos, pos, sh, color);
    t1 = orc_now(); if (stage_ms) stage_ms[0] = (t1 - t0) * 1e3; t0 = t1;
    memset(depth, 0, (size_t)P * sizeof(float));
    memset(means_2d, 0, (size_t)P * 2 * sizeof(float));
    memset(conic, 0, (size_t)P * 3 * sizeof(float));
    orc_project(P, pos, scale, rotq, scale_modifier, vp, means_2d, depth, conic);
    t1 = orc_now(); if (stage_ms) stage_ms[1] = (t1 - t0) * 1e3; t0 = t1;
    orc_allocate_tiles(P, W, H, depth, means_2d, conic, tiles_touched, radii, row0, row1);
    t1 = orc_now(); if (stage_ms) stage_ms[2] = (t1 - t0) * 1e3; t0 = t1;
    orc_inclusive_sum_u32(tiles_touched, offsets, P);
    t1 = orc_now(); if (stage_ms) stage_ms[3] = (t1 - t0) * 1e3; t0 = t1;
    long n = P > 0 ? (long)(int32_t)offsets[P - 1] : 0;
    /* impl.cpp:109 is about the FRAME's count.  A band of tile rows (row0 / row1: this build's tile-row sharding, no
     * reference counterpart) whose own count is 0 still belongs to a frame that is rendered, so its tiles get bg. */
    int whole_frame = (row0 == 0 && row1 == gy);
    if (n <= 0 && whole_frame) return 0;
    if (n <= 0) {
        memset(ranges, 0, (size_t)gx * (size_t)(row1 - row0) * 2 * sizeof(uint32_t));
        orc_blend(W, H, bg, ranges, vals_sorted, means_2d, conic, opacity, color, img, n_examined, row0, row1);
        return 0;
    }
    if ((size_t)n > capacity) return -1;
    orc_copy_with_keys(P, W, H, means_2d, offsets, radii, depth, keys_unsorted, vals_unsorted, row0, row1);
    t1 = orc_now(); if (stage_ms) stage_ms[4] = (t1 - t0) * 1e3; t0 = t1;
    orc_sort_pairs_u64_u32(keys_unsorted, vals_unsorted, keys_sorted, vals_sorted, (size_t)n);
    t1 = orc_now(); if (stage_ms) stage_ms[5] = (t1 - t0) * 1e3; t0 = t1;
    orc_get_ranges((size_t)n, keys_sorted, ranges, gx * (row1 - row0));
    t1 = orc_now(); if (stage_ms) stage_ms[6] = (t1 - t0) * 1e3; t0 = t1;
    orc_blend(W, H, bg, ranges, vals_sorted, means_2d, conic, opacity, color, img, n_examined, row0, row1);
    t1 = orc_now(); if (stage_ms) stage_ms[7] = (t1 - t0) * 1e3;
    return n;
}

/* app/main.cpp:322-337: CHW float -> HWC u8 with vertical flip and truncating *255 */
ORC_API void orc_image_to_rgb8(int W, int H, const float* img_chw, uint8_t* rgb)
{
    size_t plane = (size_t)W * (size_t)H;
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            size_t p   = ((size_t)i * W + j) * 3;
            size_t idx = (size_t)(H - i - 1) * W + j;
            for (int c = 0; c < 3; c++) rgb[p + c] = (uint8_t)(img_chw[c * plane + idx] * 255);
        }
}

/* bench.py's CPU legs: torchrun exports OMP_NUM_THREADS=1; the baseline is supposed to use the host's cores */
ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
