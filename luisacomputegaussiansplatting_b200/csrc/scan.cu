// scan.cu -- stage 2: inclusive prefix sum of tiles_touched (uint32), single pass with decoupled
// look-back.  Replaces lcpp DeviceScan<>::InclusiveSum (call site
// lcgs/src/gs_tile_splatter/impl.cpp:103-107) and the 4-byte D2H read-back of the last element:
// the grand total (num_rendered) is left in device memory for the following stages.
//
// HBM-bound: 8 bytes per item (one 4-byte read, one 4-byte write), 16-byte vector accesses.
#include "common.cuh"

namespace lcgs_b200 {

// tile status word: (flag << 32) | value, flag 0 = empty, 1 = tile aggregate, 2 = inclusive prefix
constexpr unsigned long long kFlagAggregate = 1ull << 32;
constexpr unsigned long long kFlagInclusive = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kScanThreads)
    scan_inclusive_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t num_tiles,
                          unsigned long long* status, uint32_t* ticket, uint32_t* d_total, int vec_ok)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_sum[kScanThreads / 32];
    __shared__ uint32_t s_tile_prefix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);  // tiles start in ticket order: look-back cannot deadlock
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= num_tiles) return;
    const uint32_t base = tile * kScanTile + warp * (kScanItems * 32);

    // warp-striped: chunk j of a warp is 128 consecutive items, 4 per lane
    uint32_t v[kScanItems / 4][4];
#pragma unroll
    for (int j = 0; j < kScanItems / 4; j++) {
        const uint32_t e = base + j * 128 + lane * 4;
        if (vec_ok && e + 3 < n) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + e));
            v[j][0] = q.x; v[j][1] = q.y; v[j][2] = q.z; v[j][3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) v[j][k] = (e + k < n) ? __ldg(in + e + k) : 0u;
        }
    }
    // per-thread serial scan inside each chunk, warp scan across lanes, carry across chunks
    uint32_t carry = 0;
    uint32_t excl[kScanItems / 4];
#pragma unroll
    for (int j = 0; j < kScanItems / 4; j++) {
        v[j][1] += v[j][0];
        v[j][2] += v[j][1];
        v[j][3] += v[j][2];
        uint32_t x = v[j][3];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        excl[j] = carry + x - v[j][3];
        carry += __shfl_sync(0xFFFFFFFFu, x, 31);
    }
    if (lane == 0) s_warp_sum[warp] = carry;
    __syncthreads();

    uint32_t warp_prefix = 0, tile_sum = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        const uint32_t x = s_warp_sum[w];
        if (w < warp) warp_prefix += x;
        tile_sum += x;
    }

    if (tid == 0) {
        uint32_t prefix = 0;
        if (tile > 0) {
            st_status(status + tile, kFlagAggregate | tile_sum);
            int p = (int)tile - 1;
            for (;;) {
                unsigned long long st;
                do { st = ld_status(status + p); } while ((st >> 32) == 0ull);
                prefix += (uint32_t)st;
                if ((st >> 32) == 2ull) break;
                --p;
            }
        }
        st_status(status + tile, kFlagInclusive | (unsigned long long)(prefix + tile_sum));
        s_tile_prefix = prefix;
        if (tile == num_tiles - 1 && d_total) *d_total = prefix + tile_sum;
    }
    __syncthreads();
    const uint32_t add = s_tile_prefix + warp_prefix;

#pragma unroll
    for (int j = 0; j < kScanItems / 4; j++) {
        const uint32_t e = base + j * 128 + lane * 4;
        const uint32_t o = add + excl[j];
        if (vec_ok && e + 3 < n) {
            *reinterpret_cast<uint4*>(out + e) = make_uint4(o + v[j][0], o + v[j][1], o + v[j][2], o + v[j][3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (e + k < n) out[e + k] = o + v[j][k];
        }
    }
}

int launch_scan(lcgs_b200_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n, uint32_t* d_total, cudaStream_t s)
{
    if (n == 0) {
        if (d_total) LCGS_CUDA_CHECK(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint32_t), s));
        return LCGS_B200_OK;
    }
    LCGS_REQUIRE(ctx, n < 0xFFFFFFFFull, "scan: more than 2^32-1 items");
    const uint32_t tiles = (uint32_t)((n + kScanTile - 1) / kScanTile);
    int            rc    = ws_reserve(ctx, ctx->scan_ws, (size_t)tiles * sizeof(unsigned long long));
    if (rc) return rc;
    LCGS_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->scan_ws.ptr, 0, (size_t)tiles * sizeof(unsigned long long), s));
    uint32_t* ticket = ctx->d_scalars + LCGS_SCALAR_SCAN_TICKET;
    LCGS_CUDA_CHECK(ctx, cudaMemsetAsync(ticket, 0, sizeof(uint32_t), s));
    const int vec_ok = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    scan_inclusive_kernel<<<tiles, kScanThreads, 0, s>>>(in, out, (uint32_t)n, tiles,
                                                         (unsigned long long*)ctx->scan_ws.ptr, ticket, d_total, vec_ok);
    LCGS_CUDA_CHECK(ctx, cudaGetLastError());
    return LCGS_B200_OK;
}

}  // namespace lcgs_b200
