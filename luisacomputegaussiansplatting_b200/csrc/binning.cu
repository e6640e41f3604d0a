// binning.cu -- stage 3 (key duplication) and stage 5a (tile ranges), plus the buffer filler.
//
//  * duplicate_keys_kernel replaces shad_copy_with_keys (lcgs/src/gs_tile_splatter/shader.cpp:26-69)
//    and the two zero-fills in front of it (impl.cpp:117-118, pure overhead: every slot below
//    num_rendered is overwritten).  The reference lets ONE thread write all tiles of a Gaussian
//    serially; near-plane Gaussians cover thousands of tiles, so here a warp expands its 32
//    Gaussians cooperatively: lanes walk the warp's contiguous output range, find the owning
//    Gaussian by a shuffle binary search over the warp-local prefix, and store fully coalesced.
//  * tile_ranges_kernel replaces shad_get_ranges (shader.cpp:71-100).
//
// Both are HBM-bound integer work: 12 bytes written per instance, 8 bytes read per instance.
#include "common.cuh"

namespace lcgs_b200 {

__global__ void __launch_bounds__(256)
    duplicate_keys_kernel(int P, uint32_t gx, uint32_t gy, uint32_t row0, uint32_t row1,
                          const float2* __restrict__ means_pix, const uint32_t* __restrict__ offsets,
                          const int32_t* __restrict__ radii, const float* __restrict__ depth,
                          unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, size_t capacity)
{
    const int      lane        = threadIdx.x & 31;
    const long     warp_global = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long     num_warps   = ((long)gridDim.x * blockDim.x) >> 5;
    const unsigned FULL        = 0xFFFFFFFFu;

    for (long base = warp_global * 32; base < P; base += num_warps * 32) {
        const long i   = base + lane;
        uint32_t   cnt = 0, x0 = 0, y0 = 0, w = 1, dbits = 0, incl = 0;
        if (i < P) {
            incl             = __ldg(offsets + i);
            const int radius = __ldg(radii + i);
            if (radius > 0) {
                const float2   m = __ldg(means_pix + i);
                const TileRect r = get_rect(m.x, m.y, radius, gx, gy, row0, row1);
                x0               = r.x0;
                y0               = r.y0;
                w                = r.x1 - r.x0;
                cnt              = w * (r.y1 - r.y0);
                dbits            = float_bits(__ldg(depth + i));
                if (w == 0) w = 1;
            }
        }
        // exclusive global offset of Gaussian i = offsets[i-1] (0 for i == 0), shader.cpp:42-46
        uint32_t excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = (base > 0) ? __ldg(offsets + base - 1) : 0u;
        // warp-local exclusive prefix of the per-Gaussian counts
        uint32_t x = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        const uint32_t loc   = x - cnt;
        const uint32_t total = __shfl_sync(FULL, x, 31);

        for (uint32_t p0 = 0; p0 < total; p0 += 32) {
            const uint32_t p = p0 + lane;
            // owner = largest lane l with loc_l <= p (zero-count lanes share their successor's loc and
            // are skipped because the search prefers the higher lane)
            int l = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t v = __shfl_sync(FULL, loc, l + step);
                if (v <= p) l += step;
            }
            const uint32_t o_loc   = __shfl_sync(FULL, loc, l);
            const uint32_t o_excl  = __shfl_sync(FULL, excl, l);
            const uint32_t o_x0    = __shfl_sync(FULL, x0, l);
            const uint32_t o_y0    = __shfl_sync(FULL, y0, l);
            const uint32_t o_w     = __shfl_sync(FULL, w, l);
            const uint32_t o_dbits = __shfl_sync(FULL, dbits, l);
            if (p < total) {
                const uint32_t k  = p - o_loc;  // index inside the owner's rect, row-major
                const uint32_t ry = k / o_w;
                const uint32_t rx = k - ry * o_w;
                const uint32_t tile = (o_x0 + rx) + (o_y0 + ry - row0) * gx;
                const size_t   dst  = (size_t)o_excl + k;
                if (dst < capacity) {
                    keys[dst] = ((unsigned long long)tile << 32) | (unsigned long long)o_dbits;
                    vals[dst] = (uint32_t)(base + l);
                }
            }
        }
    }
}

// Fused-path emission: Gaussians are visited in depth order (order[k], sorted depth bits skeys[k]);
// rects are the packed tile rects preprocess wrote (x0 | y0<<16, w | h<<16), offsets2 the inclusive
// sum of the counts in that order.  Same warp-cooperative expansion as above.
__global__ void __launch_bounds__(256)
    duplicate_keys_sorted_kernel(const uint32_t* __restrict__ d_m, uint32_t m_capacity, uint32_t gx, uint32_t row0,
                                 const uint32_t* __restrict__ order, const uint32_t* __restrict__ skeys,
                                 const uint2* __restrict__ rects, const uint32_t* __restrict__ offsets2,
                                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, size_t capacity,
                                 const __grid_constant__ SortDigits digits)
{
    // digit histograms of the tile bits of the emitted keys (two passes at most: the tile sort then skips
    // its histogram kernel).  Consecutive lanes emit consecutive tiles of one Gaussian, so the upper digit
    // is warp-aggregated with match_any.
    __shared__ uint32_t s_hist[2 * 512];
    const bool          do_hist = digits.hist != nullptr;
    const int           nbins   = do_hist ? (digits.num_passes << digits.radix_bits) : 0;
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) s_hist[k] = 0u;
    if (do_hist) __syncthreads();
    const int      lane        = threadIdx.x & 31;
    const long     warp_global = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long     num_warps   = ((long)gridDim.x * blockDim.x) >> 5;
    const unsigned FULL        = 0xFFFFFFFFu;
    uint32_t       M           = *d_m;
    if (M > m_capacity) M = m_capacity;

    for (long base = warp_global * 32; base < M; base += num_warps * 32) {
        const long k   = base + lane;
        uint32_t   cnt = 0, x0 = 0, y0 = 0, w = 1, dbits = 0, incl = 0, idx = 0;
        if (k < M) {
            incl          = __ldg(offsets2 + k);
            idx           = __ldg(order + k);
            dbits         = __ldg(skeys + k);
            const uint2 r = __ldg(rects + idx);
            x0            = r.x & 0xFFFFu;
            y0            = r.x >> 16;
            w             = r.y & 0xFFFFu;
            cnt           = w * (r.y >> 16);
            if (w == 0) w = 1;
        }
        uint32_t excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = (base > 0) ? __ldg(offsets2 + base - 1) : 0u;
        uint32_t x = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        const uint32_t loc   = x - cnt;
        const uint32_t total = __shfl_sync(FULL, x, 31);
        for (uint32_t p0 = 0; p0 < total; p0 += 32) {
            const uint32_t p = p0 + lane;
            int            l = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t v = __shfl_sync(FULL, loc, l + step);
                if (v <= p) l += step;
            }
            const uint32_t o_loc   = __shfl_sync(FULL, loc, l);
            const uint32_t o_excl  = __shfl_sync(FULL, excl, l);
            const uint32_t o_x0    = __shfl_sync(FULL, x0, l);
            const uint32_t o_y0    = __shfl_sync(FULL, y0, l);
            const uint32_t o_w     = __shfl_sync(FULL, w, l);
            const uint32_t o_dbits = __shfl_sync(FULL, dbits, l);
            const uint32_t o_idx   = __shfl_sync(FULL, idx, l);
            if (p < total) {
                const uint32_t j    = p - o_loc;
                const uint32_t ry   = j / o_w;
                const uint32_t rx   = j - ry * o_w;
                const uint32_t tile = (o_x0 + rx) + (o_y0 + ry - row0) * gx;
                const size_t   dst  = (size_t)o_excl + j;
                if (dst < capacity) {
                    keys[dst] = ((unsigned long long)tile << 32) | (unsigned long long)o_dbits;
                    vals[dst] = o_idx;
                    if (do_hist) {
                        atomicAdd(&s_hist[(tile >> (digits.shift[0] - 32)) & digits.mask[0]], 1u);
                        if (digits.num_passes > 1) {
                            const uint32_t d1    = (tile >> (digits.shift[1] - 32)) & digits.mask[1];
                            const unsigned peers = __match_any_sync(__activemask(), d1);
                            if ((peers & ((1u << lane) - 1u)) == 0u) atomicAdd(&s_hist[(1 << digits.radix_bits) + d1], (uint32_t)__popc(peers));
                        }
                    }
                }
            }
        }
    }
    if (do_hist) {
        __syncthreads();
        for (int k = threadIdx.x; k < nbins; k += blockDim.x) {
            const uint32_t c = s_hist[k];
            if (c) atomicAdd(digits.hist + k, c);
        }
    }
}

__global__ void __launch_bounds__(256)
    tile_ranges_kernel(const unsigned long long* __restrict__ keys, size_t n_host, const uint32_t* __restrict__ d_n,
                       size_t capacity, uint32_t* __restrict__ ranges, uint32_t num_tiles)
{
    size_t n = n_host;
    if (d_n) {
        n = *d_n;
        if (n > capacity) n = capacity;
    }
    // boundary between sorted positions k-1 and k (shad_get_ranges, shader.cpp:73-99)
    auto boundary = [&](size_t k, uint32_t prev, uint32_t cur) {
        if (k == 0) {
            if (cur < num_tiles) ranges[2 * cur] = 0u;
        } else if (cur != prev) {
            if (prev < num_tiles) ranges[2 * prev + 1] = (uint32_t)k;
            if (cur < num_tiles) ranges[2 * cur] = (uint32_t)k;
        }
        if (k == n - 1 && cur < num_tiles) ranges[2 * cur + 1] = (uint32_t)n;
    };
    // four keys per thread and iteration: two 16-byte loads plus the predecessor's tile id
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t k = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; k < n; k += stride) {
        if (k + 3 < n) {
            const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(keys + k));
            const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2*>(keys + k) + 1);
            const uint32_t   t0 = (uint32_t)(a.x >> 32), t1 = (uint32_t)(a.y >> 32), t2 = (uint32_t)(b.x >> 32), t3 = (uint32_t)(b.y >> 32);
            const uint32_t   tp = k ? (uint32_t)(__ldg(keys + k - 1) >> 32) : 0u;
            boundary(k, tp, t0);
            boundary(k + 1, t0, t1);
            boundary(k + 2, t1, t2);
            boundary(k + 3, t2, t3);
        } else {
            uint32_t prev = k ? (uint32_t)(__ldg(keys + k - 1) >> 32) : 0u;
            for (size_t j = k; j < n; j++) {
                const uint32_t cur = (uint32_t)(__ldg(keys + j) >> 32);
                boundary(j, prev, cur);
                prev = cur;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T* __restrict__ buf, size_t n, T v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) buf[k] = v;
}

int launch_duplicate_keys(lcgs_b200_ctx* ctx, int P, int W, int H, const float* means_2d, const uint32_t* offsets,
                          const int32_t* radii, const float* depth, uint64_t* keys, uint32_t* vals, size_t capacity,
                          int row0, int row1, cudaStream_t s)
{
    if (P <= 0) return LCGS_B200_OK;
    const uint32_t gx = (uint32_t)((W + 15) / 16), gy = (uint32_t)((H + 15) / 16);
    const long     warps_needed = ((long)P + 31) / 32;
    long           blocks       = (warps_needed + 7) / 8;
    const long     max_blocks   = (long)ctx->num_sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    duplicate_keys_kernel<<<(unsigned)blocks, 256, 0, s>>>(P, gx, gy, (uint32_t)row0, row1 < 0 ? gy : (uint32_t)row1,
                                                          reinterpret_cast<const float2*>(means_2d), offsets, radii,
                                                          depth, reinterpret_cast<unsigned long long*>(keys), vals,
                                                          capacity);
    LCGS_CUDA_CHECK(ctx, cudaGetLastError());
    return LCGS_B200_OK;
}

int launch_duplicate_keys_sorted(lcgs_b200_ctx* ctx, const uint32_t* d_m, int P, int W, const uint32_t* order,
                                 const uint32_t* skeys, const uint2* rects, const uint32_t* offsets2, uint64_t* keys,
                                 uint32_t* vals, size_t capacity, int row0, const SortDigits* digits, cudaStream_t s)
{
    if (P <= 0) return LCGS_B200_OK;
    // fused histograms need the digits to live in the tile id (key bits >= 32) and at most two passes
    SortDigits dg{};
    if (digits && digits->hist && digits->num_passes >= 1 && digits->num_passes <= 2 && digits->radix_bits <= 9 &&
        digits->shift[0] >= 32)
        dg = *digits;
    const uint32_t gx           = (uint32_t)((W + 15) / 16);
    const long     warps_needed = ((long)P + 31) / 32;
    long           blocks       = (warps_needed + 7) / 8;
    const long     max_blocks   = (long)ctx->num_sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    duplicate_keys_sorted_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_m, (uint32_t)P, gx, (uint32_t)row0, order, skeys, rects,
                                                                 offsets2, reinterpret_cast<unsigned long long*>(keys), vals,
                                                                 capacity, dg);
    LCGS_CUDA_CHECK(ctx, cudaGetLastError());
    return LCGS_B200_OK;
}

int launch_ranges(lcgs_b200_ctx* ctx, const uint64_t* keys, size_t n_host, const uint32_t* d_n, size_t capacity,
                  uint32_t* ranges, int num_tiles, cudaStream_t s)
{
    if (num_tiles <= 0) return LCGS_B200_OK;
    LCGS_CUDA_CHECK(ctx, cudaMemsetAsync(ranges, 0, (size_t)num_tiles * 2 * sizeof(uint32_t), s));
    const size_t bound = d_n ? capacity : n_host;
    if (bound == 0) return LCGS_B200_OK;
    LCGS_REQUIRE(ctx, (((uintptr_t)keys) & 15) == 0, "tile_ranges: keys must be 16-byte aligned");
    size_t       blocks     = (bound + 4095) / 4096;
    const size_t max_blocks = (size_t)ctx->num_sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    tile_ranges_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(keys), n_host, d_n,
                                                       capacity, ranges, (uint32_t)num_tiles);
    LCGS_CUDA_CHECK(ctx, cudaGetLastError());
    return LCGS_B200_OK;
}

template <typename T>
static int launch_fill(lcgs_b200_ctx* ctx, T* buf, size_t n, T v, cudaStream_t s)
{
    if (n == 0) return LCGS_B200_OK;
    size_t       blocks     = (n + 1023) / 1024;
    const size_t max_blocks = (size_t)ctx->num_sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    fill_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(buf, n, v);
    LCGS_CUDA_CHECK(ctx, cudaGetLastError());
    return LCGS_B200_OK;
}

int launch_fill_u32(lcgs_b200_ctx* ctx, uint32_t* buf, size_t n, uint32_t v, cudaStream_t s) { return launch_fill(ctx, buf, n, v, s); }
int launch_fill_u64(lcgs_b200_ctx* ctx, uint64_t* buf, size_t n, uint64_t v, cudaStream_t s)
{
    return launch_fill(ctx, reinterpret_cast<unsigned long long*>(buf), n, (unsigned long long)v, s);
}
int launch_fill_f32(lcgs_b200_ctx* ctx, float* buf, size_t n, float v, cudaStream_t s) { return launch_fill(ctx, buf, n, v, s); }

}  // namespace lcgs_b200
