// lookback.cuh -- decoupled look-back of a chained 32-bit sum across tiles (single-pass scans).
// Shared by the scan kernels (scan.cu) and the instance emission (binning.cu).
#pragma once

#include <stdint.h>

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

// Decoupled look-back by one full warp, 32 predecessors per step (every lane returns the exclusive
// prefix of `tile`).  Lane l inspects tile p - l.  The window is consumed up to the nearest inclusive
// prefix; if an unpublished tile comes first, everything in front of it is consumed and the window
// slides so that this tile becomes lane 0 of the next poll.  A single thread walking one predecessor
// per L2 round trip made the first wave of tiles (all started together, only aggregates published)
// a chain of hundreds of serial round trips.
__device__ __forceinline__ uint32_t lookback_warp_u32(unsigned long long* status, uint32_t tile, uint32_t tile_sum)
{
    const int      lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    uint32_t       prefix = 0;
    if (tile > 0) {
        if (lane == 0) st_status(status + tile, kFlagAggregate | tile_sum);
        int p = (int)tile - 1;
        for (;;) {
            const int                q  = p - lane;
            const unsigned long long st = q >= 0 ? ld_status(status + q) : kFlagInclusive;  // before tile 0: inclusive 0
            const uint32_t           fl = (uint32_t)(st >> 32);
            const unsigned           inc = __ballot_sync(FULL, fl == 2u), emp = __ballot_sync(FULL, fl == 0u);
            const int first_inc = inc ? __ffs(inc) - 1 : 32, first_emp = emp ? __ffs(emp) - 1 : 32;
            const int take      = first_inc < first_emp ? first_inc + 1 : first_emp;  // lanes [0, take) are consumed
            prefix += __reduce_add_sync(FULL, lane < take ? (uint32_t)st : 0u);
            if (first_inc < first_emp) break;
            p -= take;
        }
    }
    if (lane == 0) st_status(status + tile, kFlagInclusive | (unsigned long long)(prefix + tile_sum));
    return prefix;
}

}  // namespace lcgs_b200
