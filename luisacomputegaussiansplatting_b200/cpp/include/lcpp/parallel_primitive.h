// lcpp/parallel_primitive.h -- the two lc_parallel_primitive entry points the reference uses
// (DeviceScan<>::InclusiveSum, DeviceRadixSort<>::SortPairs<ulong,uint>; call sites
// app/main.cpp:175-178 and lcgs/src/gs_tile_splatter/impl.cpp:34,50,104,135-143), over the C ABI.
// Temp storage lives in the lcgs_b200 context, so the temp-buffer arguments of the originals are gone.
#pragma once

#include "lcgs/runtime.h"

namespace luisa::parallel_primitive
{

template <typename = void>
class DeviceScan
{
public:
    void create(lcgs::Device& device, lcgs::Stream* stream) noexcept { m_device = &device; m_stream = stream; }
    template <typename T>
    static size_t GetTempStorageBytes(size_t num_items) { return lcgs_b200_scan_temp_bytes(num_items); }
    void InclusiveSum(lcgs::CommandList& cmdlist, lcgs::BufferView<lcgs::uint> d_in, lcgs::BufferView<lcgs::uint> d_out,
                      size_t num_items)
    {
        m_device->check(lcgs_b200_scan_inclusive_u32(m_device->ctx(), d_in.ptr, d_out.ptr, num_items, cmdlist.stream().abi()),
                        "DeviceScan::InclusiveSum");
    }

private:
    lcgs::Device* m_device = nullptr;
    lcgs::Stream* m_stream = nullptr;
};

template <typename = void>
class DeviceRadixSort
{
public:
    void create(lcgs::Device& device, lcgs::Stream* stream) noexcept { m_device = &device; m_stream = stream; }
    template <typename K, typename V>
    static size_t GetSortPairsTempStorageBytes(size_t num_items) { return lcgs_b200_sort_temp_bytes(num_items); }
    template <typename K = lcgs::ulong, typename V = lcgs::uint>
    void SortPairs(lcgs::CommandList& cmdlist, lcgs::BufferView<lcgs::ulong> keys_in, lcgs::BufferView<lcgs::ulong> keys_out,
                   lcgs::BufferView<lcgs::uint> vals_in, lcgs::BufferView<lcgs::uint> vals_out, size_t num_items,
                   int begin_bit = 0, int end_bit = 64)
    {
        m_device->check(lcgs_b200_sort_pairs_u64_u32(m_device->ctx(), keys_in.ptr, keys_out.ptr, vals_in.ptr, vals_out.ptr,
                                                     num_items, begin_bit, end_bit, cmdlist.stream().abi()),
                        "DeviceRadixSort::SortPairs");
    }

private:
    lcgs::Device* m_device = nullptr;
    lcgs::Stream* m_stream = nullptr;
};

}  // namespace luisa::parallel_primitive
