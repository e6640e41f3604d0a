// blend.cu -- stage 5b: per-tile front-to-back alpha blending.
// Replaces m_forward_render_shader (lcgs/src/gs_tile_splatter/shader.cpp:171-288).
//
// One 256-thread CTA per 16x16 tile; each warp owns an 8x4 pixel patch (compact patches terminate
// together more often than the reference's 16x2 rows).  The tile's depth-sorted list is consumed in
// batches of 256: every thread gathers one Gaussian's packed 48-byte record (pixel mean, pre-scaled
// conic, alpha-test threshold, opacity, colour) into shared memory, then all pixels walk the batch
// with broadcast LDS.128 reads.  Differences from the reference that do not change results:
//   * colour is staged with the batch instead of being fetched from global memory per
//     contributing pair (shader.cpp:268-269);
//   * the alpha >= 1/255 test is a compare against a per-Gaussian power threshold (see
//     lcgs_math.cuh), so rejected pairs cost no exp at all;
//   * __syncthreads_and(done) stops fetching batches once every pixel of the tile has saturated
//     (the reference keeps loading and barrier-ing until the list ends, shader.cpp:226-277).
//
// Compute-bound (FP32 + shared-memory broadcast), not HBM-bound: ~E examined (pixel, Gaussian)
// pairs per frame at ~12 instructions each; HBM side is 4 B id + 48 B record per instance
// (mostly L2 hits) + 12 B per pixel.
#include "common.cuh"

namespace lcgs_b200 {

constexpr int kBlendThreads = 256;

__global__ void __launch_bounds__(kBlendThreads)
    blend_kernel(int W, int H, uint32_t gx, uint32_t row0, float bg0, float bg1, float bg2,
                 const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ records, const uint32_t* __restrict__ d_num_rendered,
                 float* __restrict__ img)
{
    // num_rendered == 0: the reference returns before launching the render and leaves the image
    // untouched (lcgs/src/gs_tile_splatter/impl.cpp:109, quirk Q10)
    if (d_num_rendered && *d_num_rendered == 0u) return;

    __shared__ float4 s_a[kBlendThreads];  // pix.x, pix.y, -0.5*conic.x, -conic.y
    __shared__ float4 s_b[kBlendThreads];  // -0.5*conic.z, threshold, opacity, cull radius^2
    __shared__ float4 s_c[kBlendThreads];  // r, g, b, -

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 8x4 pixel patch per warp: warps tile the 16x16 block as 2 columns x 4 rows of patches
    const int lx = (warp & 1) * 8 + (lane & 7);
    const int ly = (warp >> 1) * 4 + (lane >> 3);
    const int px = blockIdx.x * 16 + lx;
    const int py = (row0 + blockIdx.y) * 16 + ly;
    const bool  inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;  // no half-pixel offset (Q2)

    const uint32_t tile  = blockIdx.x + blockIdx.y * gx;
    const uint2    range = __ldg(ranges + tile);

    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    bool  done = !inside;

    for (uint32_t start = range.x; start < range.y; start += kBlendThreads) {
        // barrier protecting the staging buffers + block-wide early exit
        if (__syncthreads_and(done)) break;
        const uint32_t idx = start + tid;
        if (idx < range.y) {
            const uint32_t id  = __ldg(point_list + idx);
            const float4*  rec = records + (size_t)id * kRecordFloat4s;
            s_a[tid] = __ldg(rec);
            s_b[tid] = __ldg(rec + 1);
            s_c[tid] = __ldg(rec + 2);
        }
        __syncthreads();
        const int cnt = (int)min((uint32_t)kBlendThreads, range.y - start);
        for (int j = 0; j < cnt && !done; j++) {
            const float4 a  = s_a[j];
            const float4 b  = s_b[j];
            const float  dx = a.x - pxf;
            const float  dy = a.y - pyf;
            // canonical evaluation (shared with the oracle), two explicit fused ops
            const float power = __fmaf_rn(a.w * dx, dy, __fmaf_rn(a.z * dx, dx, (b.x * dy) * dy));
            if (power > 0.0f || power < b.y) continue;  // shader.cpp:257,259
            const float alpha  = fminf(0.99f, b.z * __expf(power));
            const float test_T = T * (1.0f - alpha);
            if (test_T < 0.0001f) {  // shader.cpp:261-265: saturated, this entry is not blended
                done = true;
                continue;
            }
            const float4 c = s_c[j];
            const float  w = T * alpha;
            C0 = __fmaf_rn(w, c.x, C0);
            C1 = __fmaf_rn(w, c.y, C1);
            C2 = __fmaf_rn(w, c.z, C2);
            T  = test_T;
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
