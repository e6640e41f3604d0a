// lcgs/util/buffer_filler.h -- lcgs::BufferFiller (reference: lcgs/include/lcgs/util/buffer_filler.h:16-72).
// The element types the hot path fills (uint, ulong, float, int) are provided.
#pragma once

#include "lcgs/runtime.h"

namespace lcgs
{

class BufferFiller
{
public:
    uint block_size = 256u;

    void fill(Device& device, Stream& stream, BufferView<uint> v, uint x) const noexcept
    {
        device.check(lcgs_b200_fill_u32(device.ctx(), v.ptr, v.count, x, stream.abi()), "BufferFiller::fill<uint>");
    }
    void fill(Device& device, Stream& stream, BufferView<int> v, int x) const noexcept
    {
        device.check(lcgs_b200_fill_u32(device.ctx(), reinterpret_cast<uint*>(v.ptr), v.count, (uint)x, stream.abi()),
                     "BufferFiller::fill<int>");
    }
    void fill(Device& device, Stream& stream, BufferView<ulong> v, ulong x) const noexcept
    {
        device.check(lcgs_b200_fill_u64(device.ctx(), v.ptr, v.count, x, stream.abi()), "BufferFiller::fill<ulong>");
    }
    void fill(Device& device, Stream& stream, BufferView<float> v, float x) const noexcept
    {
        device.check(lcgs_b200_fill_f32(device.ctx(), v.ptr, v.count, x, stream.abi()), "BufferFiller::fill<float>");
    }
};

}  // namespace lcgs
