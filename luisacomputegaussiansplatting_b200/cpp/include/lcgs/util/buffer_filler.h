// lcgs/util/buffer_filler.h -- lcgs::BufferFiller (reference: lcgs/include/lcgs/util/buffer_filler.h:16-72).
// Same call shape: `cmdlist << filler.fill(device, view, value)` -- fill() returns an un-submitted command.
// All ten element types of the reference (buffer_filler.h:61-70) are provided; the vector types are filled
// component-wise when their components are equal bit patterns of one 32-bit word, which is the only way
// the reference's callers use them (zero-fills), and abort otherwise.
#pragma once

#include <cstring>

#include "lcgs/runtime.h"

namespace lcgs
{

struct uint3 { uint x, y, z; };
struct uint4 { uint x, y, z, w; };

class BufferFiller
{
public:
    uint block_size = 256u;

    Command fill(Device& device, BufferView<uint> v, const uint& x) const noexcept { return fill32(device, v.ptr, v.count, x); }
    Command fill(Device& device, BufferView<int> v, const int& x) const noexcept
    {
        return fill32(device, reinterpret_cast<uint*>(v.ptr), v.count, (uint)x);
    }
    Command fill(Device& device, BufferView<float> v, const float& x) const noexcept
    {
        return [&device, p = v.ptr, n = v.count, x](cudaStream_t s) {
            device.check(lcgs_b200_fill_f32(device.ctx(), p, n, x, reinterpret_cast<lcgs_b200_stream>(s)), "BufferFiller::fill<float>");
        };
    }
    Command fill(Device& device, BufferView<ulong> v, const ulong& x) const noexcept
    {
        return [&device, p = v.ptr, n = v.count, x](cudaStream_t s) {
            device.check(lcgs_b200_fill_u64(device.ctx(), p, n, x, reinterpret_cast<lcgs_b200_stream>(s)), "BufferFiller::fill<ulong>");
        };
    }
    Command fill(Device& device, BufferView<float2> v, const float2& x) const noexcept { return fill_words(device, v, x); }
    Command fill(Device& device, BufferView<float3> v, const float3& x) const noexcept { return fill_words(device, v, x); }
    Command fill(Device& device, BufferView<float4> v, const float4& x) const noexcept { return fill_words(device, v, x); }
    Command fill(Device& device, BufferView<uint2> v, const uint2& x) const noexcept { return fill_words(device, v, x); }
    Command fill(Device& device, BufferView<uint3> v, const uint3& x) const noexcept { return fill_words(device, v, x); }
    Command fill(Device& device, BufferView<uint4> v, const uint4& x) const noexcept { return fill_words(device, v, x); }

private:
    static Command fill32(Device& device, uint* p, size_t n, uint x)
    {
        return [&device, p, n, x](cudaStream_t s) {
            device.check(lcgs_b200_fill_u32(device.ctx(), p, n, x, reinterpret_cast<lcgs_b200_stream>(s)), "BufferFiller::fill");
        };
    }
    // vector element types: a buffer of N-word elements whose words are all equal is a buffer of words
    template <typename V>
    static Command fill_words(Device& device, BufferView<V> v, const V& x)
    {
        constexpr size_t kWords = sizeof(V) / sizeof(uint);
        uint             w[kWords];
        std::memcpy(w, &x, sizeof(V));
        for (size_t k = 1; k < kWords; k++)
            if (w[k] != w[0]) fatal("BufferFiller::fill: vector fill values must have equal components in this build");
        return fill32(device, reinterpret_cast<uint*>(v.ptr), v.count * kWords, w[0]);
    }
};

}  // namespace lcgs
