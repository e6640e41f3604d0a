// gaussians.h -- host-side Gaussian set and INRIA-3DGS .ply I/O.
// Same role and field names as the reference's app/gaussians.h:15-36 (GaussiansData, read_gs_ply);
// the parser itself is a small purpose-built binary-PLY reader instead of happly.
#pragma once

#include <cstddef>
#include <filesystem>
#include <string>
#include <vector>

namespace lcgs
{

// Page-locked host memory (cudaHostAlloc) so that the five uploads of app/main.cpp:216-222 run at full
// PCIe rate without a staging copy; plain malloc when no CUDA device is present (CPU-side tests).
void* host_alloc_pinned(size_t bytes);
void  host_free_pinned(void* p);

template <typename T>
struct PinnedAllocator {
    using value_type = T;
    PinnedAllocator() = default;
    template <typename U>
    PinnedAllocator(const PinnedAllocator<U>&) noexcept {}
    T*   allocate(size_t n) { return static_cast<T*>(host_alloc_pinned(n * sizeof(T))); }
    void deallocate(T* p, size_t) noexcept { host_free_pinned(p); }
    template <typename U>
    bool operator==(const PinnedAllocator<U>&) const noexcept { return true; }
    template <typename U>
    bool operator!=(const PinnedAllocator<U>&) const noexcept { return false; }
};
using HostArray = std::vector<float, PinnedAllocator<float>>;

struct GaussiansData {
    int       num_gaussians = 0;
    int       sh_deg        = 3;
    HostArray pos;      // [P][3]
    HostArray feature;  // [P][(deg+1)^2][3]  coefficient-major, RGB interleaved
    HostArray opacity;  // [P]     sigmoid(stored logit)
    HostArray scale;    // [P][3]  exp(stored log-scale)
    HostArray rotq;     // [P][4]  (r,x,y,z), normalised

    static float scaling_activation(float x);
    static void  rotation_activation(float& r, float& x, float& y, float& z);
    static float opacity_activation(float x);
    void         resize(int N);
};

// Reads `x y z f_dc_0..2 f_rest_0..44 opacity scale_0..2 rot_0..3` (float32 properties, any order,
// extra properties ignored) from a binary_little_endian or ascii PLY.  Returns false with a message
// in `err` on failure.  A binary body is memory-mapped and de-interleaved + activated by `threads`
// worker threads (0 = hardware concurrency) straight into the pinned arrays: a 6 M-Gaussian checkpoint
// (1.5 GB) is one pass over the page cache instead of 62 whole-file column extractions.
bool read_gs_ply(GaussiansData& gs, const std::filesystem::path& fpath, std::string* err = nullptr, int threads = 0);

// Writes pre-activation values back in the INRIA layout (used by the tests' round trip).
bool write_gs_ply(const std::filesystem::path& fpath, int P, const float* pos, const float* sh, const float* logit_opacity,
                  const float* log_scale, const float* raw_rot);

}  // namespace lcgs
