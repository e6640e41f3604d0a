// gaussians.h -- host-side Gaussian set and INRIA-3DGS .ply I/O.
// Same role and field names as the reference's app/gaussians.h:15-36 (GaussiansData, read_gs_ply);
// the parser itself is a small purpose-built binary-PLY reader instead of happly.
#pragma once

#include <filesystem>
#include <string>
#include <vector>

namespace lcgs
{

struct GaussiansData {
    int                num_gaussians = 0;
    int                sh_deg        = 3;
    std::vector<float> pos;      // [P][3]
    std::vector<float> feature;  // [P][(deg+1)^2][3]  coefficient-major, RGB interleaved
    std::vector<float> opacity;  // [P]     sigmoid(stored logit)
    std::vector<float> scale;    // [P][3]  exp(stored log-scale)
    std::vector<float> rotq;     // [P][4]  (r,x,y,z), normalised

    static float scaling_activation(float x);
    static void  rotation_activation(float& r, float& x, float& y, float& z);
    static float opacity_activation(float x);
    void         resize(int N);
};

// Reads `x y z f_dc_0..2 f_rest_0..44 opacity scale_0..2 rot_0..3` (float32 properties, any order,
// extra properties ignored) from a binary_little_endian or ascii PLY.  Returns false with a message
// in `err` on failure.
bool read_gs_ply(GaussiansData& gs, const std::filesystem::path& fpath, std::string* err = nullptr);

// Writes pre-activation values back in the INRIA layout (used by the tests' round trip).
bool write_gs_ply(const std::filesystem::path& fpath, int P, const float* pos, const float* sh, const float* logit_opacity,
                  const float* log_scale, const float* raw_rot);

}  // namespace lcgs
