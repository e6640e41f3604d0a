// blend.cu -- stage 5b: per-tile front-to-back alpha blending.
// Replaces m_forward_render_shader (lcgs/src/gs_tile_splatter/shader.cpp:171-288).
//
// One 256-thread CTA per 16x16 tile; each warp owns an 8x4 pixel patch.  The tile's depth-sorted
// list is consumed in rounds of 256 candidates:
//   1. every thread gathers one Gaussian's packed record (pixel mean, pre-scaled conic, alpha-test
//      threshold, opacity, colour: 48 B) and tests it against the TILE rectangle with cull_rect();
//      survivors are compacted, order preserved, into shared memory (ballot + prefix);
//   2. each warp tests 32 survivors at a time, one per lane, against ITS 8x4 PATCH and walks only
//      the set bits of the ballot, all 32 pixels evaluating the same Gaussian with broadcast LDS.
// cull_rect() (lcgs_math.cuh) is conservative with respect to the per-pixel float evaluation, so
// culling never changes a pixel: it only removes pairs the alpha test would have skipped.  On the
// C3 scene 42 % of the (Gaussian, tile) instances the reference's loose rect bins never touch the
// tile, and 78 % of the (Gaussian, patch) pairs are empty -- the reference evaluates all of them
// for all 256 pixels.
// Further differences from the reference that do not change results: colour is staged with the
// batch instead of fetched from global memory per contributing pair (shader.cpp:268-269); the
// alpha >= 1/255 test is a compare against a per-Gaussian power threshold, so rejected pairs cost
// no exp; __syncthreads_and(done) stops fetching once every pixel of the tile has saturated (the
// reference keeps loading and barrier-ing until the list ends, shader.cpp:226-277).
//
// Compute-bound (FP32 issue + shared memory), not HBM-bound; HBM side is 4 B id + 32..48 B record
// per instance (mostly L2 hits) + 12 B per pixel.
#include "common.cuh"

namespace lcgs_b200 {

constexpr int kBlendThreads = 256;
constexpr int kBlendWarps   = kBlendThreads / 32;

__global__ void __launch_bounds__(kBlendThreads)
    blend_kernel(int W, int H, uint32_t gx, uint32_t row0, float bg0, float bg1, float bg2,
                 const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ records, const uint32_t* __restrict__ d_num_rendered,
                 float* __restrict__ img)
{
    // num_rendered == 0: the reference returns before launching the render and leaves the image
    // untouched (lcgs/src/gs_tile_splatter/impl.cpp:109, quirk Q10)
    if (d_num_rendered && *d_num_rendered == 0u) return;

    __shared__ float4   s_a[kBlendThreads];  // pix.x, pix.y, -0.5*conic.x, -conic.y
    __shared__ float4   s_b[kBlendThreads];  // -0.5*conic.z, threshold, opacity, -
    __shared__ float4   s_c[kBlendThreads];  // r, g, b, -
    __shared__ uint32_t s_cnt[kBlendWarps];

    const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL    = 0xFFFFFFFFu;
    const unsigned lt_mask = (1u << lane) - 1u;
    // 8x4 pixel patch per warp: warps tile the 16x16 block as 2 columns x 4 rows of patches
    const int tile_x0 = blockIdx.x * 16, tile_y0 = (row0 + blockIdx.y) * 16;
    const int patch_x0 = tile_x0 + (warp & 1) * 8, patch_y0 = tile_y0 + (warp >> 1) * 4;
    const int px = patch_x0 + (lane & 7);
    const int py = patch_y0 + (lane >> 3);
    const bool  inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;  // no half-pixel offset (Q2)
    const float tx0 = (float)tile_x0, ty0 = (float)tile_y0, tx1 = (float)(tile_x0 + 15), ty1 = (float)(tile_y0 + 15);
    const float wx0 = (float)patch_x0, wy0 = (float)patch_y0, wx1 = (float)(patch_x0 + 7), wy1 = (float)(patch_y0 + 3);

    const uint32_t tile  = blockIdx.x + blockIdx.y * gx;
    const uint2    range = __ldg(ranges + tile);

    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    bool  done = !inside;

    for (uint32_t start = range.x; start < range.y; start += kBlendThreads) {
        // barrier protecting the staging buffers + block-wide early exit
        if (__syncthreads_and(done)) break;

        // ---- gather one candidate per thread, cull against the tile, compact ----------------------
        const uint32_t idx  = start + tid;
        bool           keep = false;
        float4         a, b;
        const float4*  rec = nullptr;
        if (idx < range.y) {
            rec  = records + (size_t)__ldg(point_list + idx) * kRecordFloat4s;
            a    = __ldg(rec);
            b    = __ldg(rec + 1);
            keep = !cull_rect(a.x, a.y, a.z, a.w, b.x, b.y, tx0, ty0, tx1, ty1);
        }
        const unsigned kept = __ballot_sync(FULL, keep);
        if (lane == 0) s_cnt[warp] = __popc(kept);
        __syncthreads();
        uint32_t slot = __popc(kept & lt_mask), cnt = 0;
#pragma unroll
        for (int w = 0; w < kBlendWarps; w++) {
            const uint32_t c = s_cnt[w];
            if (w < warp) slot += c;
            cnt += c;
        }
        if (keep) {
            s_a[slot] = a;
            s_b[slot] = b;
            s_c[slot] = __ldg(rec + 2);
        }
        __syncthreads();

        // ---- per warp: cull 32 survivors at a time against the patch, evaluate the hits ------------
        for (uint32_t base = 0; base < cnt; base += 32) {
            if (__all_sync(FULL, done)) break;
            const uint32_t g   = base + lane;
            bool           hit = false;
            if (g < cnt) {
                const float4 ga = s_a[g];
                const float4 gb = s_b[g];
                hit             = !cull_rect(ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, wx0, wy0, wx1, wy1);
            }
            unsigned hits = __ballot_sync(FULL, hit);
            while (hits) {
                const uint32_t j = base + (uint32_t)__ffs(hits) - 1u;
                hits &= hits - 1u;
                if (done) continue;
                const float4 ea = s_a[j];
                const float4 eb = s_b[j];
                // canonical evaluation (shared with the oracle), two explicit fused ops
                const float power = blend_power(ea.z, ea.w, eb.x, ea.x - pxf, ea.y - pyf);
                if (power > 0.0f || power < eb.y) continue;  // shader.cpp:257,259
                const float alpha  = fminf(0.99f, eb.z * __expf(power));
                const float test_T = T * (1.0f - alpha);
                if (test_T < 0.0001f) {  // shader.cpp:261-265: saturated, this entry is not blended
                    done = true;
                    continue;
                }
                const float4 ec = s_c[j];
                const float  w  = T * alpha;
                C0 = __fmaf_rn(w, ec.x, C0);
                C1 = __fmaf_rn(w, ec.y, C1);
                C2 = __fmaf_rn(w, ec.z, C2);
                T  = test_T;
            }
        }
    }
    if (inside) {
        const size_t plane = (size_t)W * (size_t)H;
        const size_t pix   = (size_t)px + (size_t)W * (size_t)py;
        img[pix]             = __fmaf_rn(bg0, T, C0);
        img[pix + plane]     = __fmaf_rn(bg1, T, C1);
        img[pix + 2 * plane] = __fmaf_rn(bg2, T, C2);
    }
}

int launch_blend(lcgs_b200_ctx* ctx, int W, int H, const float* bg, const uint32_t* ranges, const uint32_t* point_list,
                 const float4* records, const uint32_t* d_num_rendered, float* img, int row0, int row1, cudaStream_t s)
{
    const uint32_t gx = (uint32_t)((W + 15) / 16), gy = (uint32_t)((H + 15) / 16);
    if (row1 < 0) row1 = (int)gy;
    if (W <= 0 || H <= 0 || row1 <= row0) return LCGS_B200_OK;
    dim3 grid(gx, (unsigned)(row1 - row0));
    blend_kernel<<<grid, kBlendThreads, 0, s>>>(W, H, gx, (uint32_t)row0, bg[0], bg[1], bg[2],
                                                reinterpret_cast<const uint2*>(ranges), point_list, records,
                                                d_num_rendered, img);
    LCGS_CUDA_CHECK(ctx, cudaGetLastError());
    return LCGS_B200_OK;
}

}  // namespace lcgs_b200
