// blend_model.cpp -- TEST-ONLY analysis tool: replays the control flow of csrc/blend.cu on the CPU
// (same culling functions, same round / segment / patch structure) over an oracle frame and counts
// the work each scheduling variant would do.  Used to choose kernel experiments before spending GPU
// time; never part of liblcgs_b200.so.  Built and driven by tests/tools/blend_model.py.
#include "../../luisacomputegaussiansplatting_b200/csrc/lcgs_math.cuh"

#include <algorithm>
#include <vector>

using namespace lcgs_b200;

namespace {

struct Rec {
    float mx, my, a, b, c, thr, l2op, ry, rx;
};

struct Counters {
    // current kernel
    unsigned long long rounds, warp_rounds, seg_walks, hits, lane_ok, lane_blend, tile_survivors, instances_seen;
    // variants
    unsigned long long seg_walks_dense;      // tile survivors densely packed (ceil(n/32) segments per round)
    unsigned long long hits_half_lr;         // two 4x4 halves per warp, own hit list each: iterations = max
    unsigned long long hits_half_tb;         // two 8x2 halves
    unsigned long long hits_quarter;         // four 4x2 quarters: iterations = max of four
    unsigned long long hits_rows;            // four 8x1 rows
    unsigned long long hits_8x8;             // 8x8 patches (2 pixels per lane), 4 warps per tile: iterations (each costs ~1.6x)
    unsigned long long seg_walks_8x8;
    unsigned long long hits_exact;           // warp-hits with at least one ok lane (lower bound for any patch-level cull)
    unsigned long long warp_rounds_8x8;
    unsigned long long hits_16x4, seg_walks_16x4, warp_rounds_16x4;  // 16x4 patches (lane owns (x,y) and (x+8,y))
    unsigned long long hits_8x8_exact, hits_16x4_exact;
};

struct Box {
    float x0, y0, x1, y1;
    bool  any;
};

}  // namespace

extern "C" __attribute__((visibility("default"))) void bm_run(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                                                              const float* means, const float* conic, const float* opacity,
                                                              const float* thr, unsigned long long* out)
{
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    const int ntiles = gx * gy;
    Counters  tot{};
#pragma omp parallel
    {
        Counters         c{};
        std::vector<Rec> surv;
        std::vector<int> segcnt;
#pragma omp for schedule(dynamic, 8)
        for (int tile = 0; tile < ntiles; tile++) {
            const int      tbx = tile % gx, tby = tile / gx;
            const int      tx0 = tbx * 16, ty0 = tby * 16;
            const uint32_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
            const uint32_t len = e > s ? e - s : 0u;
            float          T[256], T8[256], T16[256];
            bool           done[256], done8[256], done16[256];
            for (int p = 0; p < 256; p++) {
                const int px = tx0 + (p & 15), py = ty0 + (p >> 4);
                done[p] = done8[p] = done16[p] = !(px < W && py < H);
                T[p] = T8[p] = T16[p] = 1.0f;
            }
            const uint32_t nrounds = (len + 255) / 256;
            for (uint32_t r = 0; r < nrounds; r++) {
                // ---- produce: tile cull, per-warp segments -------------------------------------------
                Rec seg[8][32];
                int cnt[8];
                int total = 0;
                for (int w = 0; w < 8; w++) {
                    cnt[w] = 0;
                    for (int l = 0; l < 32; l++) {
                        const uint32_t idx = r * 256 + w * 32 + l;
                        if (idx >= len) continue;
                        c.instances_seen++;
                        const uint32_t g = point_list[s + idx];
                        Rec            q;
                        q.mx = means[2 * g]; q.my = means[2 * g + 1];
                        q.a = -0.5f * conic[3 * g]; q.b = -conic[3 * g + 1]; q.c = -0.5f * conic[3 * g + 2];
                        q.thr = thr[g]; q.l2op = log2f(opacity[g]);
                        const CullCoef k = cull_coef(q.mx, q.my, q.a, q.b, q.c, q.thr);
                        q.ry = k.ry; q.rx = k.rx;
                        if (cull_rect_fast(q.mx, q.my, q.a, q.b, q.c, q.thr, q.ry, q.rx, (float)tx0, (float)ty0, (float)(tx0 + 15),
                                           (float)(ty0 + 15)))
                            continue;
                        seg[w][cnt[w]++] = q;
                    }
                    total += cnt[w];
                }
                c.rounds++;
                c.tile_survivors += total;
                const int dense_segs = (total + 31) / 32;

                // generic evaluation of one Gaussian on a set of pixels of the tile; returns #ok lanes
                auto eval = [&](const Rec& q, int x0, int y0, int wdt, int hgt, int& nblend) {
                    int nok = 0;
                    for (int yy = 0; yy < hgt; yy++)
                        for (int xx = 0; xx < wdt; xx++) {
                            const int p = (y0 + yy) * 16 + (x0 + xx);
                            if (done[p]) continue;
                            const float dx = q.mx - (float)(tx0 + x0 + xx), dy = q.my - (float)(ty0 + y0 + yy);
                            const float power = blend_power(q.a, q.b, q.c, dx, dy);
                            const bool  ok = !(power > 0.0f) && !(power < q.thr);
                            if (!ok) continue;
                            nok++;
                            const float alpha = fminf(0.99f, exp2f(power * 1.4426950408889634f + q.l2op));
                            const float tt    = T[p] * (1.0f - alpha);
                            if (tt < 0.0001f) { done[p] = true; continue; }
                            T[p] = tt;
                            nblend++;
                        }
                    return nok;
                };
                auto active_box = [&](int x0, int y0, int wdt, int hgt) {
                    Box b{ 1e9f, 1e9f, -1e9f, -1e9f, false };
                    for (int yy = 0; yy < hgt; yy++)
                        for (int xx = 0; xx < wdt; xx++)
                            if (!done[(y0 + yy) * 16 + x0 + xx]) {
                                b.any = true;
                                b.x0 = fminf(b.x0, (float)(tx0 + x0 + xx)); b.x1 = fmaxf(b.x1, (float)(tx0 + x0 + xx));
                                b.y0 = fminf(b.y0, (float)(ty0 + y0 + yy)); b.y1 = fmaxf(b.y1, (float)(ty0 + y0 + yy));
                            }
                    return b;
                };
                auto hit_box = [&](const Rec& q, const Box& b) {
                    return b.any && !cull_rect_fast(q.mx, q.my, q.a, q.b, q.c, q.thr, q.ry, q.rx, b.x0, b.y0, b.x1, b.y1);
                };

                // ---- consume, warp by warp (patch 8x4: 2 columns x 4 rows of patches) ----------------
                // Pixels of different warps are independent, so warps can be replayed one after the other.
                for (int w = 0; w < 8; w++) {
                    const int x0 = (w & 1) * 8, y0 = (w >> 1) * 4;
                    Box       box = active_box(x0, y0, 8, 4);
                    if (!box.any) continue;
                    c.warp_rounds++;
                    // boxes of the sub-patch variants are taken at round start too
                    const Box bl = active_box(x0, y0, 4, 4), br = active_box(x0 + 4, y0, 4, 4);
                    const Box bt = active_box(x0, y0, 8, 2), bb = active_box(x0, y0 + 2, 8, 2);
                    Box       bq[4], brow[4];
                    for (int k = 0; k < 4; k++) {
                        bq[k]   = active_box(x0 + (k & 1) * 4, y0 + (k >> 1) * 2, 4, 2);
                        brow[k] = active_box(x0, y0 + k, 8, 1);
                    }
                    bool all_done = false;
                    int  walked_slots = 0;
                    for (int sgi = 0; sgi < 8 && !all_done; sgi++) {
                        c.seg_walks++;
                        walked_slots += cnt[sgi];
                        int h = 0, hl = 0, hr = 0, ht = 0, hb = 0, hq[4] = { 0, 0, 0, 0 }, hrw[4] = { 0, 0, 0, 0 };
                        for (int k = 0; k < cnt[sgi]; k++) {
                            const Rec& q = seg[sgi][k];
                            if (!hit_box(q, box)) continue;
                            h++;
                            hl += hit_box(q, bl); hr += hit_box(q, br);
                            ht += hit_box(q, bt); hb += hit_box(q, bb);
                            for (int j = 0; j < 4; j++) { hq[j] += hit_box(q, bq[j]); hrw[j] += hit_box(q, brow[j]); }
                            int       nb  = 0;
                            const int nok = eval(q, x0, y0, 8, 4, nb);
                            c.lane_ok += nok;
                            c.lane_blend += nb;
                            c.hits_exact += nok > 0;
                        }
                        c.hits += h;
                        c.hits_half_lr += std::max(hl, hr);
                        c.hits_half_tb += std::max(ht, hb);
                        c.hits_quarter += std::max(std::max(hq[0], hq[1]), std::max(hq[2], hq[3]));
                        c.hits_rows += std::max(std::max(hrw[0], hrw[1]), std::max(hrw[2], hrw[3]));
                        all_done = !active_box(x0, y0, 8, 4).any;
                    }
                    // dense packing: the warp stops after the segment holding the last slot it walked
                    c.seg_walks_dense += all_done ? (walked_slots + 31) / 32 + (walked_slots == 0) : dense_segs;
                }
                // ---- the same round replayed with 64-pixel patches (2 pixels per lane), own pixel state -------
                for (int shape = 0; shape < 2; shape++) {
                    const int pw = shape == 0 ? 8 : 16, ph = shape == 0 ? 8 : 4;
                    float* Ts = shape == 0 ? T8 : T16;
                    bool*  ds = shape == 0 ? done8 : done16;
                    for (int w = 0; w < 4; w++) {
                        const int x0 = shape == 0 ? (w & 1) * 8 : 0, y0 = shape == 0 ? (w >> 1) * 8 : w * 4;
                        Box box{ 1e9f, 1e9f, -1e9f, -1e9f, false };
                        for (int yy = 0; yy < ph; yy++)
                            for (int xx = 0; xx < pw; xx++)
                                if (!ds[(y0 + yy) * 16 + x0 + xx]) {
                                    box.any = true;
                                    box.x0 = fminf(box.x0, (float)(tx0 + x0 + xx)); box.x1 = fmaxf(box.x1, (float)(tx0 + x0 + xx));
                                    box.y0 = fminf(box.y0, (float)(ty0 + y0 + yy)); box.y1 = fmaxf(box.y1, (float)(ty0 + y0 + yy));
                                }
                        if (!box.any) continue;
                        (shape == 0 ? c.warp_rounds_8x8 : c.warp_rounds_16x4)++;
                        bool all_done = false;
                        for (int sgi = 0; sgi < 8 && !all_done; sgi++) {
                            (shape == 0 ? c.seg_walks_8x8 : c.seg_walks_16x4)++;
                            for (int k = 0; k < cnt[sgi]; k++) {
                                const Rec& q = seg[sgi][k];
                                if (!hit_box(q, box)) continue;
                                (shape == 0 ? c.hits_8x8 : c.hits_16x4)++;
                                int nok = 0;
                                for (int yy = 0; yy < ph; yy++)
                                    for (int xx = 0; xx < pw; xx++) {
                                        const int p = (y0 + yy) * 16 + (x0 + xx);
                                        if (ds[p]) continue;
                                        const float dx = q.mx - (float)(tx0 + x0 + xx), dy = q.my - (float)(ty0 + y0 + yy);
                                        const float power = blend_power(q.a, q.b, q.c, dx, dy);
                                        if (power > 0.0f || power < q.thr) continue;
                                        nok++;
                                        const float alpha = fminf(0.99f, exp2f(power * 1.4426950408889634f + q.l2op));
                                        const float tt    = Ts[p] * (1.0f - alpha);
                                        if (tt < 0.0001f) { ds[p] = true; continue; }
                                        Ts[p] = tt;
                                    }
                                (shape == 0 ? c.hits_8x8_exact : c.hits_16x4_exact) += nok > 0;
                            }
                            all_done = true;
                            for (int yy = 0; yy < ph; yy++)
                                for (int xx = 0; xx < pw; xx++) all_done = all_done && ds[(y0 + yy) * 16 + x0 + xx];
                        }
                    }
                }
                bool tile_done = true;
                for (int p = 0; p < 256; p++) tile_done = tile_done && done[p];
                if (tile_done) break;
            }
        }
#pragma omp critical
        {
            unsigned long long*       t = reinterpret_cast<unsigned long long*>(&tot);
            const unsigned long long* v = reinterpret_cast<const unsigned long long*>(&c);
            for (size_t k = 0; k < sizeof(Counters) / sizeof(unsigned long long); k++) t[k] += v[k];
        }
    }
    const unsigned long long* t = reinterpret_cast<const unsigned long long*>(&tot);
    for (size_t k = 0; k < sizeof(Counters) / sizeof(unsigned long long); k++) out[k] = t[k];
}
