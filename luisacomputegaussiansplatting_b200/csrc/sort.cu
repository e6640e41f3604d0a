// sort.cu -- stage 4: onesweep least-significant-digit radix sort of (uint64 key, uint32 value)
// pairs.  Replaces lcpp DeviceRadixSort<>::SortPairs<ulong,uint> (call site
// lcgs/src/gs_tile_splatter/impl.cpp:134-144), which the reference's README calls crude.
//
// One histogram kernel reads the keys once and counts every digit of every pass; then one
// "onesweep" kernel per 8-bit digit reads each pair once and writes it once to its final place for
// that pass: tiles (4096 pairs) are ranked in shared memory with warp-match (__match_any_sync)
// histograms, the per-digit global offsets come from a chained-scan decoupled look-back across
// tiles, and pairs are staged through shared memory so that global stores are coalesced runs.
// Stable: ties keep their input order, so equal (tile, depth) keys stay ordered by Gaussian index,
// exactly what the CPU oracle's stable sort yields.
//
// HBM-bound: algorithmic bytes = 8*n (histogram) + 24*n per pass.
#include <stdlib.h>

#include "common.cuh"

namespace lcgs_b200 {

constexpr uint32_t kStatusAggregate = 1u << 30;
constexpr uint32_t kStatusInclusive = 2u << 30;
constexpr uint32_t kStatusValueMask = (1u << 30) - 1u;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ size_t resolve_n(size_t n_host, const uint32_t* d_n, size_t capacity)
{
    if (!d_n) return n_host;
    size_t n = *d_n;
    return n > capacity ? capacity : n;
}

struct SortPassInfo {
    int      num_passes;
    int      shift[kMaxSortPasses];
    uint32_t mask[kMaxSortPasses];
};

// ---- upfront histogram of every pass's digit ---------------------------------------------------
__global__ void __launch_bounds__(256)
    radix_histogram_kernel(const unsigned long long* __restrict__ keys, size_t n_host, const uint32_t* __restrict__ d_n,
                           size_t capacity, uint32_t* __restrict__ hist, const __grid_constant__ SortPassInfo info)
{
    __shared__ uint32_t s_hist[kMaxSortPasses * kRadix];
    for (int k = threadIdx.x; k < info.num_passes * kRadix; k += blockDim.x) s_hist[k] = 0u;
    __syncthreads();
    const size_t n      = resolve_n(n_host, d_n, capacity);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const unsigned long long key = __ldg(keys + k);
#pragma unroll
        for (int p = 0; p < kMaxSortPasses; p++)
            if (p < info.num_passes) atomicAdd(&s_hist[p * kRadix + (uint32_t)((key >> info.shift[p]) & info.mask[p])], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < info.num_passes * kRadix; k += blockDim.x) {
        const uint32_t c = s_hist[k];
        if (c) atomicAdd(hist + k, c);
    }
}

// exclusive scan of one value per thread over a THREADS-thread block
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /* [THREADS/32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t  x    = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t pre = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++)
        if (w < warp) pre += s_warp[w];
    __syncthreads();  // s_warp may be reused
    return pre + x - v;
}

// ---- one onesweep pass -------------------------------------------------------------------------
// Template: THREADS x ITEMS pairs per tile (warp-striped), PREFETCH = software pipelining of the
// next tile's key loads behind the current tile's look-back and write-out.
// Shared memory (dynamic):
//   s_keys      [TILE]       u64   tile's keys in tile-sorted order
//   s_vals      [TILE]       u32   tile's values in tile-sorted order
//   s_warp_hist [WARPS][256] u32   per-warp digit counters, then exclusive prefix over warps
//   s_tile_start[256]        u32   exclusive digit prefix inside the tile
//   s_digit_base[256]        u32   global base of the digit minus s_tile_start
constexpr int kLookbackWindow = 4;

template <int THREADS, int ITEMS>
constexpr size_t sweep_smem_bytes()
{
    return (size_t)THREADS * ITEMS * 12 + (size_t)(THREADS / 32) * kRadix * 4 + (size_t)kRadix * 4 * 2 + (THREADS / 32) * 4 + 64;
}

template <int THREADS, int ITEMS, int MIN_BLOCKS, bool PREFETCH>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
    onesweep_pass_kernel(const unsigned long long* __restrict__ keys_in, unsigned long long* __restrict__ keys_out,
                         const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out, size_t n_host,
                         const uint32_t* __restrict__ d_n, size_t capacity, const uint32_t* __restrict__ hist /* [256] */,
                         uint32_t* status /* [tiles][256] */, uint32_t* ticket, int shift, uint32_t mask)
{
    static_assert(THREADS >= kRadix && THREADS % 32 == 0, "one thread per digit");
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE  = THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* const s_keys       = reinterpret_cast<unsigned long long*>(smem_raw);
    uint32_t* const           s_vals       = reinterpret_cast<uint32_t*>(s_keys + TILE);
    uint32_t* const           s_warp_hist  = s_vals + TILE;
    uint32_t* const           s_tile_start = s_warp_hist + WARPS * kRadix;
    uint32_t* const           s_digit_base = s_tile_start + kRadix;
    uint32_t* const           s_scan       = s_digit_base + kRadix;  // [WARPS]
    uint32_t* const           s_ticket     = s_scan + WARPS;         // [2]

    const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL      = 0xFFFFFFFFu;
    const unsigned lt_mask   = (1u << lane) - 1u;
    const size_t   n         = resolve_n(n_host, d_n, capacity);
    const uint32_t num_tiles = (uint32_t)((n + TILE - 1) / TILE);
    const uint32_t q0        = warp * (ITEMS * 32) + lane;  // warp-striped: item j sits at q0 + 32*j
    const bool     is_digit  = tid < kRadix;

    // global exclusive scan of this pass's digit histogram (same for every tile)
    const uint32_t bin_base = block_exclusive_scan<THREADS>(is_digit ? __ldg(hist + tid) : 0u, s_scan);

    if (tid == 0) s_ticket[0] = atomicAdd(ticket, 1u);
    __syncthreads();
    uint32_t tile = s_ticket[0];

    unsigned long long key[ITEMS];
    auto load_keys = [&](uint32_t t) {
        const size_t   base = (size_t)t * TILE;
        const uint32_t nv   = t < num_tiles ? (uint32_t)((n - base) < (size_t)TILE ? (n - base) : (size_t)TILE) : 0u;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) key[j] = (q0 + 32 * j < nv) ? __ldg(keys_in + base + q0 + 32 * j) : ~0ull;
    };
    if (PREFETCH) load_keys(tile);

    for (uint32_t it = 0; tile < num_tiles; it++) {
        const size_t   tile_base = (size_t)tile * TILE;
        const uint32_t nvalid    = (uint32_t)((n - tile_base) < (size_t)TILE ? (n - tile_base) : (size_t)TILE);
        // the next ticket is taken early so that its keys can be prefetched; tickets are processed
        // in increasing order by every CTA, which keeps the look-back free of deadlock
        if (tid == 0) s_ticket[(it + 1) & 1] = atomicAdd(ticket, 1u);
        for (int k = tid; k < WARPS * kRadix; k += THREADS) s_warp_hist[k] = 0u;
        if (!PREFETCH) load_keys(tile);
        __syncthreads();

        // ---- rank inside the warp with match_any ---------------------------------------------
        uint32_t  rank[ITEMS];
        uint32_t* my_hist = s_warp_hist + warp * kRadix;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const bool     valid = q0 + 32 * j < nvalid;
            const uint32_t d     = valid ? (uint32_t)((key[j] >> shift) & mask) : (uint32_t)kRadix;
            const unsigned peers = __match_any_sync(FULL, d);
            const unsigned lower = peers & lt_mask;
            uint32_t       pre   = 0;
            if (valid) pre = my_hist[d];
            __syncwarp();
            if (valid && lower == 0u) my_hist[d] = pre + __popc(peers);
            __syncwarp();
            rank[j] = pre + __popc(lower);
        }
        __syncthreads();
        const uint32_t next_tile = s_ticket[(it + 1) & 1];

        // ---- per digit (thread d): prefix over warps, tile histogram, early publish ------------
        uint32_t        tile_count = 0;
        uint32_t* const my_status  = status + (size_t)tile * kRadix + tid;
        if (is_digit) {
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                const uint32_t c              = s_warp_hist[w * kRadix + tid];
                s_warp_hist[w * kRadix + tid] = tile_count;
                tile_count += c;
            }
            if (tile > 0) st_relaxed_u32(my_status, kStatusAggregate | tile_count);
        }
        const uint32_t tile_start = block_exclusive_scan<THREADS>(tile_count, s_scan);
        if (is_digit) s_tile_start[tid] = tile_start;
        __syncthreads();

        // ---- scatter keys into shared memory in tile-sorted order --------------------------------
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            if (q0 + 32 * j < nvalid) {
                const uint32_t d = (uint32_t)((key[j] >> shift) & mask);
                rank[j] += s_tile_start[d] + my_hist[d];
                s_keys[rank[j]] = key[j];
            }
        }
        // ---- start the value loads of this tile and the key loads of the next one ----------------
        uint32_t val[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) val[j] = (q0 + 32 * j < nvalid) ? __ldg(vals_in + tile_base + q0 + 32 * j) : 0u;
        if (PREFETCH) load_keys(next_tile);

        // ---- decoupled look-back, kLookbackWindow predecessors in flight at a time --------------
        if (is_digit) {
            uint32_t prefix = 0;
            if (tile > 0) {
                int  p    = (int)tile - 1;
                bool more = true;
                while (more) {
                    uint32_t st[kLookbackWindow];
#pragma unroll
                    for (int k = 0; k < kLookbackWindow; k++)
                        st[k] = (p - k >= 0) ? ld_relaxed_u32(status + (size_t)(p - k) * kRadix + tid) : kStatusInclusive;
#pragma unroll
                    for (int k = 0; k < kLookbackWindow; k++) {
                        if (!more) break;
                        if ((st[k] >> 30) == 0u) {  // not published yet: poll again from this tile
                            p -= k;
                            break;
                        }
                        prefix += st[k] & kStatusValueMask;
                        if ((st[k] >> 30) == 2u) more = false;
                        else if (k == kLookbackWindow - 1) p -= kLookbackWindow;
                    }
                }
            }
            st_relaxed_u32(my_status, kStatusInclusive | ((prefix + tile_count) & kStatusValueMask));
            s_digit_base[tid] = bin_base + prefix - tile_start;
        }
        __syncthreads();

        // ---- coalesced key write-out; scatter the values -----------------------------------------
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const uint32_t q = tid + j * THREADS;
            if (q < nvalid) {
                const unsigned long long k = s_keys[q];
                keys_out[s_digit_base[(uint32_t)((k >> shift) & mask)] + q] = k;
            }
        }
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
            if (q0 + 32 * j < nvalid) s_vals[rank[j]] = val[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const uint32_t q = tid + j * THREADS;
            if (q < nvalid) vals_out[s_digit_base[(uint32_t)((s_keys[q] >> shift) & mask)] + q] = s_vals[q];
        }
        __syncthreads();  // shared memory is reused by the next tile
        tile = next_tile;
    }
}

typedef void (*SweepKernel)(const unsigned long long*, unsigned long long*, const uint32_t*, uint32_t*, size_t, const uint32_t*,
                            size_t, const uint32_t*, uint32_t*, uint32_t*, int, uint32_t);
struct SweepVariant {
    SweepKernel kernel;
    int         threads, tile, blocks_per_sm;
    size_t      smem;
    const char* name;
};
static const SweepVariant kSweepVariants[] = {
    { onesweep_pass_kernel<256, 16, 3, false>, 256, 4096, 3, sweep_smem_bytes<256, 16>(), "256x16 3/SM" },
    { onesweep_pass_kernel<256, 16, 2, true>, 256, 4096, 2, sweep_smem_bytes<256, 16>(), "256x16 2/SM prefetch" },
    { onesweep_pass_kernel<512, 8, 2, true>, 512, 4096, 2, sweep_smem_bytes<512, 8>(), "512x8 2/SM prefetch" },
    { onesweep_pass_kernel<512, 8, 2, false>, 512, 4096, 2, sweep_smem_bytes<512, 8>(), "512x8 2/SM" },
    { onesweep_pass_kernel<512, 12, 1, true>, 512, 6144, 1, sweep_smem_bytes<512, 12>(), "512x12 1/SM prefetch" },
    { onesweep_pass_kernel<256, 8, 5, false>, 256, 2048, 5, sweep_smem_bytes<256, 8>(), "256x8 5/SM" },
    { onesweep_pass_kernel<256, 12, 3, true>, 256, 3072, 3, sweep_smem_bytes<256, 12>(), "256x12 3/SM prefetch" },
};
constexpr int kNumSweepVariants = (int)(sizeof(kSweepVariants) / sizeof(kSweepVariants[0]));
constexpr int kMaxSweepTile     = 6144;
constexpr int kMinSweepTile     = 2048;

static int sweep_variant_index()
{
    static int idx = -1;
    if (idx < 0) {
        idx           = 0;  // default
        const char* e = getenv("LCGS_SORT_VARIANT");
        if (e && atoi(e) >= 0 && atoi(e) < kNumSweepVariants) idx = atoi(e);
    }
    return idx;
}

__global__ void __launch_bounds__(256)
    copy_pairs_kernel(const unsigned long long* __restrict__ keys_in, unsigned long long* __restrict__ keys_out,
                      const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out, size_t n_host,
                      const uint32_t* __restrict__ d_n, size_t capacity)
{
    const size_t n      = resolve_n(n_host, d_n, capacity);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        keys_out[k] = keys_in[k];
        vals_out[k] = vals_in[k];
    }
}

// workspace layout: [hist: passes*256 u32][tickets are ctx scalars][status: passes*tiles*256 u32][tmp keys][tmp vals]
static size_t sort_ws_layout(size_t n, size_t* off_status, size_t* off_keys, size_t* off_vals)
{
    const size_t tiles = (n + kMinSweepTile - 1) / kMinSweepTile;
    size_t       off   = 0;
    off += (size_t)kMaxSortPasses * kRadix * sizeof(uint32_t);
    off = (off + 255) & ~(size_t)255;
    if (off_status) *off_status = off;
    off += (size_t)kMaxSortPasses * tiles * kRadix * sizeof(uint32_t);
    off = (off + 255) & ~(size_t)255;
    if (off_keys) *off_keys = off;
    off += n * sizeof(uint64_t);
    off = (off + 255) & ~(size_t)255;
    if (off_vals) *off_vals = off;
    off += n * sizeof(uint32_t);
    off = (off + 255) & ~(size_t)255;
    return off;
}

size_t sort_temp_bytes(size_t n) { return sort_ws_layout(n, nullptr, nullptr, nullptr); }

int launch_sort(lcgs_b200_ctx* ctx, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in,
                uint32_t* vals_out, size_t n_host, const uint32_t* d_n, size_t capacity, int begin_bit, int end_bit,
                cudaStream_t s)
{
    LCGS_REQUIRE(ctx, begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit, "sort: bad bit range");
    const size_t bound = d_n ? capacity : n_host;  // upper bound on the number of pairs
    if (bound == 0) return LCGS_B200_OK;
    LCGS_REQUIRE(ctx, bound <= (size_t)kStatusValueMask, "sort: more than 2^30-1 pairs");
    if (!d_n) capacity = n_host;

    SortPassInfo info;
    const int    bits = end_bit - begin_bit;
    info.num_passes   = (bits + kRadixBits - 1) / kRadixBits;
    for (int p = 0; p < kMaxSortPasses; p++) {
        const int lo   = begin_bit + p * kRadixBits;
        const int w    = (p < info.num_passes) ? ((end_bit - lo) < kRadixBits ? (end_bit - lo) : kRadixBits) : 0;
        info.shift[p]  = lo < 64 ? lo : 0;
        info.mask[p]   = w > 0 ? ((1u << w) - 1u) : 0u;
    }

    const auto* kin  = reinterpret_cast<const unsigned long long*>(keys_in);
    auto*       kout = reinterpret_cast<unsigned long long*>(keys_out);
    const unsigned grid_stride_blocks = (unsigned)(((bound + 1023) / 1024) < (size_t)ctx->num_sms * 8
                                                       ? ((bound + 1023) / 1024)
                                                       : (size_t)ctx->num_sms * 8);
    if (info.num_passes == 0) {
        copy_pairs_kernel<<<grid_stride_blocks, 256, 0, s>>>(kin, kout, vals_in, vals_out, n_host, d_n, capacity);
        LCGS_CUDA_CHECK(ctx, cudaGetLastError());
        return LCGS_B200_OK;
    }

    size_t       off_status, off_keys, off_vals;
    const size_t bytes = sort_ws_layout(bound, &off_status, &off_keys, &off_vals);
    int          rc    = ws_reserve(ctx, ctx->sort_ws, bytes);
    if (rc) return rc;
    char*     ws        = (char*)ctx->sort_ws.ptr;
    uint32_t* hist      = (uint32_t*)ws;
    uint32_t* status    = (uint32_t*)(ws + off_status);
    auto*     tmp_keys  = (unsigned long long*)(ws + off_keys);
    uint32_t* tmp_vals  = (uint32_t*)(ws + off_vals);
    const SweepVariant& var   = kSweepVariants[sweep_variant_index()];
    const size_t        tiles = (bound + var.tile - 1) / var.tile;
    uint32_t*           ticket = ctx->d_scalars + LCGS_SCALAR_SORT_TICKET;

    // zero histograms + look-back status (contiguous) and the tickets
    LCGS_CUDA_CHECK(ctx, cudaMemsetAsync(ws, 0, off_status + (size_t)info.num_passes * tiles * kRadix * sizeof(uint32_t), s));
    LCGS_CUDA_CHECK(ctx, cudaMemsetAsync(ticket, 0, kMaxSortPasses * sizeof(uint32_t), s));

    const bool prof = ctx->profiling && ctx->ev_sort[0];
    if (prof) cudaEventRecord(ctx->ev_sort[0], s);
    radix_histogram_kernel<<<grid_stride_blocks, 256, 0, s>>>(kin, n_host, d_n, capacity, hist, info);
    LCGS_CUDA_CHECK(ctx, cudaGetLastError());
    if (prof) cudaEventRecord(ctx->ev_sort[1], s);

    static bool smem_attr_set[kNumSweepVariants] = {};  // opt in to > 48 KB of dynamic shared memory once per process
    if (!smem_attr_set[sweep_variant_index()]) {
        LCGS_CUDA_CHECK(ctx, cudaFuncSetAttribute(var.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)var.smem));
        smem_attr_set[sweep_variant_index()] = true;
    }
    const size_t   max_ctas     = (size_t)ctx->num_sms * var.blocks_per_sm;
    const unsigned sweep_blocks = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    const unsigned long long* src_k = kin;
    const uint32_t*           src_v = vals_in;
    for (int p = 0; p < info.num_passes; p++) {
        const bool          to_out = ((info.num_passes - 1 - p) % 2) == 0;
        unsigned long long* dst_k  = to_out ? kout : tmp_keys;
        uint32_t*           dst_v  = to_out ? vals_out : tmp_vals;
        var.kernel<<<sweep_blocks, var.threads, var.smem, s>>>(src_k, dst_k, src_v, dst_v, n_host, d_n, capacity,
                                                              hist + p * kRadix, status + (size_t)p * tiles * kRadix, ticket + p,
                                                              info.shift[p], info.mask[p]);
        LCGS_CUDA_CHECK(ctx, cudaGetLastError());
        src_k = dst_k;
        src_v = dst_v;
    }
    if (prof) {
        cudaEventRecord(ctx->ev_sort[2], s);
        ctx->sort_passes   = info.num_passes;
        ctx->ev_sort_valid = 1;
    }
    return LCGS_B200_OK;
}

}  // namespace lcgs_b200
