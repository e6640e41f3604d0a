// lcgs/util/camera.h -- lcgs::Camera and its matrix helpers, over the C ABI's host functions.
// Same names and semantics as the reference's lcgs/include/lcgs/util/camera.h:15-82.
#pragma once

#include <array>

#include "lcgs/runtime.h"

namespace lcgs
{

struct Camera {
    float3 position{};
    float3 front{};
    float3 up{};
    float3 right{};
    float  fov          = 60.0f;
    float  aspect_ratio = 1.0f;
    int    width        = 512;
    int    height       = 512;
};
static_assert(sizeof(Camera) == sizeof(lcgs_b200_camera), "Camera must be layout-compatible with the C ABI");

using float4x4 = std::array<float, 16>;  // column-major, m[c*4+r]

inline const lcgs_b200_camera* abi(const Camera& c) { return reinterpret_cast<const lcgs_b200_camera*>(&c); }

// Field-by-field conversion of ANY camera with the reference's member names into the ABI struct.  The
// reference's own lcgs::Camera holds four luisa::float3, which are 16-byte aligned 16-byte vectors
// (lcgs/include/lcgs/util/camera.h:15-25): it is 80 bytes, NOT layout-compatible with the packed 64-byte
// lcgs_b200_camera, so it must never be reinterpret_cast -- copy the fields (tests/facade/test_facade.cpp
// checks this against a mock 16-byte float3).
template <typename CameraLike>
inline lcgs_b200_camera to_abi_camera(const CameraLike& c) noexcept
{
    lcgs_b200_camera o;
    o.position[0] = c.position.x; o.position[1] = c.position.y; o.position[2] = c.position.z;
    o.front[0] = c.front.x; o.front[1] = c.front.y; o.front[2] = c.front.z;
    o.up[0] = c.up.x; o.up[1] = c.up.y; o.up[2] = c.up.z;
    o.right[0] = c.right.x; o.right[1] = c.right.y; o.right[2] = c.right.z;
    o.fov = c.fov; o.aspect_ratio = c.aspect_ratio; o.width = c.width; o.height = c.height;
    return o;
}

inline Camera get_lookat_cam(float3 pos, float3 target, float3 world_up)
{
    Camera cam;
    lcgs_b200_get_lookat_cam(&pos.x, &target.x, &world_up.x, reinterpret_cast<lcgs_b200_camera*>(&cam));
    return cam;
}

inline float4x4 local_to_world_matrix(const Camera& cam) noexcept
{
    float4x4 m;
    lcgs_b200_local_to_world_matrix(abi(cam), m.data());
    return m;
}

inline float4x4 world_to_local_matrix(const Camera& cam) noexcept
{
    float4x4 m;
    lcgs_b200_world_to_local_matrix(abi(cam), m.data());
    return m;
}

inline float4x4 projection_matrix(float tanfovx, float tanfovy, float znear = 0.1f, float zfar = 100.0f) noexcept
{
    float4x4 m;
    lcgs_b200_projection_matrix(tanfovx, tanfovy, znear, zfar, m.data());
    return m;
}

}  // namespace lcgs
