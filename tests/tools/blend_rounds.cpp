// blend_rounds.cpp -- TEST-ONLY analysis tool (see blend_model.cpp): the 8x8-patch, two-pixels-per-lane blend schedule
// without a tile-level cull, replayed for several round sizes R (candidates staged per barrier).  Counts per R: rounds,
// active warp-rounds, 32-candidate segment walks and hit evaluations.  The patch's bounding box of unfinished pixels is
// refreshed once per round, so smaller rounds cull with a tighter box.
#include "../../luisacomputegaussiansplatting_b200/csrc/lcgs_math.cuh"

using namespace lcgs_b200;

struct Rec { float mx, my, a, b, c, thr, l2op, ry, rx; };

extern "C" __attribute__((visibility("default"))) void br_run(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                                                              const float* means, const float* conic, const float* opacity,
                                                              const float* thr, int nR, const int* Rs, unsigned long long* out /* [nR][5] */)
{
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    for (int ri = 0; ri < nR; ri++) {
        const int R = Rs[ri];
        unsigned long long rounds = 0, wrounds = 0, walks = 0, hits = 0, staged = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : rounds, wrounds, walks, hits, staged)
        for (int tile = 0; tile < gx * gy; tile++) {
            const int      tx0 = (tile % gx) * 16, ty0 = (tile / gx) * 16;
            const uint32_t s = ranges[2 * tile], e = ranges[2 * tile + 1], len = e > s ? e - s : 0u;
            float T[256];
            bool  done[256];
            for (int p = 0; p < 256; p++) {
                done[p] = !(tx0 + (p & 15) < W && ty0 + (p >> 4) < H);
                T[p]    = 1.0f;
            }
            for (uint32_t r0 = 0; r0 < len; r0 += R) {
                const uint32_t n = len - r0 < (uint32_t)R ? len - r0 : (uint32_t)R;
                rounds++;
                staged += n;
                for (int w = 0; w < 4; w++) {
                    const int x0 = (w & 1) * 8, y0 = (w >> 1) * 8;
                    float bx0 = 1e9f, by0 = 1e9f, bx1 = -1e9f, by1 = -1e9f;
                    bool  any = false;
                    for (int yy = 0; yy < 8; yy++)
                        for (int xx = 0; xx < 8; xx++)
                            if (!done[(y0 + yy) * 16 + x0 + xx]) {
                                any = true;
                                bx0 = fminf(bx0, (float)(tx0 + x0 + xx)); bx1 = fmaxf(bx1, (float)(tx0 + x0 + xx));
                                by0 = fminf(by0, (float)(ty0 + y0 + yy)); by1 = fmaxf(by1, (float)(ty0 + y0 + yy));
                            }
                    if (!any) continue;
                    wrounds++;
                    bool all_done = false;
                    for (uint32_t sg = 0; sg < n && !all_done; sg += 32) {
                        walks++;
                        for (uint32_t k = sg; k < sg + 32 && k < n; k++) {
                            const uint32_t g = point_list[s + r0 + k];
                            Rec q;
                            q.mx = means[2 * g]; q.my = means[2 * g + 1];
                            q.a = -0.5f * conic[3 * g]; q.b = -conic[3 * g + 1]; q.c = -0.5f * conic[3 * g + 2];
                            q.thr = thr[g]; q.l2op = log2f(opacity[g]);
                            const CullCoef cc = cull_coef(q.mx, q.my, q.a, q.b, q.c, q.thr);
                            if (cull_rect_fast(q.mx, q.my, q.a, q.b, q.c, q.thr, cc.ry, cc.rx, bx0, by0, bx1, by1)) continue;
                            hits++;
                            for (int yy = 0; yy < 8; yy++)
                                for (int xx = 0; xx < 8; xx++) {
                                    const int p = (y0 + yy) * 16 + (x0 + xx);
                                    if (done[p]) continue;
                                    const float dx = q.mx - (float)(tx0 + x0 + xx), dy = q.my - (float)(ty0 + y0 + yy);
                                    const float power = blend_power(q.a, q.b, q.c, dx, dy);
                                    if (power > 0.0f || power < q.thr) continue;
                                    const float alpha = fminf(0.99f, exp2f(power * 1.4426950408889634f + q.l2op));
                                    const float tt    = T[p] * (1.0f - alpha);
                                    if (tt < 0.0001f) { done[p] = true; continue; }
                                    T[p] = tt;
                                }
                        }
                        all_done = true;
                        for (int yy = 0; yy < 8; yy++)
                            for (int xx = 0; xx < 8; xx++) all_done = all_done && done[(y0 + yy) * 16 + x0 + xx];
                    }
                }
                bool tile_done = true;
                for (int p = 0; p < 256; p++) tile_done = tile_done && done[p];
                if (tile_done) break;
            }
        }
        out[ri * 5 + 0] = rounds; out[ri * 5 + 1] = wrounds; out[ri * 5 + 2] = walks; out[ri * 5 + 3] = hits; out[ri * 5 + 4] = staged;
    }
}
