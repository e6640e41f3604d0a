// host_mirror.cpp -- TEST-ONLY: runs the device math header (csrc/lcgs_math.cuh) on the CPU so
// that its transcription can be checked against the oracle without a GPU.  Built by
// tests/test_host_mirror.py with g++ -ffp-contract=off; never part of liblcgs_b200.so.
#include "../../luisacomputegaussiansplatting_b200/csrc/lcgs_math.cuh"

using namespace lcgs_b200;

struct PtrSh {
    const float* p;
    float operator()(int k, int c) const { return p[k * 3 + c]; }
};

extern "C" {

__attribute__((visibility("default"))) void hm_preprocess(int P, int sh_deg, const float* pos, const float* scale,
                                                          const float* rotq, const float* sh, const float* opacity,
                                                          float scale_modifier, const ViewParams* vp, int row0, int row1,
                                                          float* means_2d, float* depth, float* conic, float* color,
                                                          int32_t* radii, uint32_t* tiles, float* thr)
{
    const uint32_t gx = (uint32_t)((vp->width + 15) / 16), gy = (uint32_t)((vp->height + 15) / 16);
    const uint32_t r1 = row1 < 0 ? gy : (uint32_t)row1;
    const int      feat = (sh_deg + 1) * (sh_deg + 1);
    for (long i = 0; i < P; i++) {
        const ViewPoint pv = view_transform(*vp, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        depth[i] = 0.f; radii[i] = 0; tiles[i] = 0u;
        means_2d[2 * i] = means_2d[2 * i + 1] = 0.f;
        conic[3 * i] = conic[3 * i + 1] = conic[3 * i + 2] = 0.f;
        color[3 * i] = color[3 * i + 1] = color[3 * i + 2] = 0.f;
        thr[i] = 0.f;
        if (!pv.visible) continue;
        float cov[3];
        ewa_cov2d(*vp, pv, scale_modifier, scale[3 * i], scale[3 * i + 1], scale[3 * i + 2], rotq[4 * i], rotq[4 * i + 1],
                  rotq[4 * i + 2], rotq[4 * i + 3], cov);
        const Splat2D s = splat_from_cov(pv.ndc_x, pv.ndc_y, cov, vp->width, vp->height, gx, gy, (uint32_t)row0, r1);
        depth[i] = pv.z; radii[i] = s.radius; tiles[i] = s.tiles;
        means_2d[2 * i] = s.px; means_2d[2 * i + 1] = s.py;
        conic[3 * i] = s.conic[0]; conic[3 * i + 1] = s.conic[1]; conic[3 * i + 2] = s.conic[2];
        float rgb[3];
        sh_color(sh_deg, vp->cam_pos, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], PtrSh{ sh + (size_t)i * feat * 3 }, rgb);
        color[3 * i] = rgb[0]; color[3 * i + 1] = rgb[1]; color[3 * i + 2] = rgb[2];
        thr[i] = alpha_threshold(opacity[i]);
    }
}

// For n (Gaussian, rectangle) pairs: out[k] bit0 = cull_rect() (fast=0) or cull_rect_fast() (fast=1) says "drop", bit1 = some pixel centre
// of the rectangle passes the exact per-pixel test (0 >= power >= thr).  bit0 && bit1 is a bug.
__attribute__((visibility("default"))) void hm_cull_check(long n, const float* mean, const float* conic, const float* thr,
                                                          const int* rect, int fast, unsigned char* out)
{
    for (long k = 0; k < n; k++) {
        const float a = -0.5f * conic[3 * k], b = -conic[3 * k + 1], c = -0.5f * conic[3 * k + 2];
        const int   x0 = rect[4 * k], y0 = rect[4 * k + 1], x1 = rect[4 * k + 2], y1 = rect[4 * k + 3];
        const bool  drop0 = cull_rect(mean[2 * k], mean[2 * k + 1], a, b, c, thr[k], (float)x0, (float)y0, (float)x1, (float)y1);
        const CullCoef q  = cull_coef(mean[2 * k], mean[2 * k + 1], a, b, c, thr[k]);
        const bool  drop1 = cull_rect_fast(mean[2 * k], mean[2 * k + 1], a, b, c, thr[k], q.ry, q.rx, (float)x0, (float)y0,
                                           (float)x1, (float)y1);
        const bool  drop  = fast ? drop1 : drop0;
        bool        any  = false;
        for (int y = y0; y <= y1 && !any; y++)
            for (int x = x0; x <= x1; x++) {
                const float p = blend_power(a, b, c, mean[2 * k] - (float)x, mean[2 * k + 1] - (float)y);
                if (!(p > 0.0f) && !(p < thr[k])) { any = true; break; }
            }
        out[k] = (unsigned char)((drop ? 1 : 0) | (any ? 2 : 0));
    }
}

__attribute__((visibility("default"))) float hm_exp(float x) { return exp_rn(x); }
__attribute__((visibility("default"))) float hm_alpha_threshold(float op) { return alpha_threshold(op); }
}
