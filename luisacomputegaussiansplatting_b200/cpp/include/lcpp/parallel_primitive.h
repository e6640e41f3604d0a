// lcpp/parallel_primitive.h -- the two lc_parallel_primitive entry points the reference uses
// (DeviceScan<>::InclusiveSum, DeviceRadixSort<>::SortPairs<ulong,uint>; call sites
// app/main.cpp:175-178 and lcgs/src/gs_tile_splatter/impl.cpp:34,50,104,135-143), over the C ABI, with
// the reference's argument lists: the caller-provided temp-storage view is accepted and ignored (temp
// storage lives in the lcgs_b200 context and is grown on demand), and the size queries still answer so
// that ensure_*_temp_buffer (impl.cpp:31-61) keeps working unchanged.
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
    // impl.cpp:104: InclusiveSum(cmdlist, temp, d_in, d_out, num_items) -- enqueue only
    void InclusiveSum(lcgs::CommandList& cmdlist, lcgs::BufferView<lcgs::uint> /*temp*/, lcgs::BufferView<lcgs::uint> d_in,
                      lcgs::BufferView<lcgs::uint> d_out, size_t num_items)
    {
        cmdlist << [dev = m_device, d_in, d_out, num_items](cudaStream_t s) {
            dev->check(lcgs_b200_scan_inclusive_u32(dev->ctx(), d_in.ptr, d_out.ptr, num_items, reinterpret_cast<lcgs_b200_stream>(s)),
                       "DeviceScan::InclusiveSum");
        };
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
    // impl.cpp:135-143: SortPairs<ulong,uint>(cmdlist, temp, keys_in, keys_out, vals_in, vals_out, num_items) -- all
    // 64 key bits, ascending, stable, inputs preserved; begin_bit / end_bit are an extension
    template <typename K = lcgs::ulong, typename V = lcgs::uint>
    void SortPairs(lcgs::CommandList& cmdlist, lcgs::BufferView<lcgs::uint> /*temp*/, lcgs::BufferView<lcgs::ulong> keys_in,
                   lcgs::BufferView<lcgs::ulong> keys_out, lcgs::BufferView<lcgs::uint> vals_in,
                   lcgs::BufferView<lcgs::uint> vals_out, size_t num_items, int begin_bit = 0, int end_bit = 64)
    {
        cmdlist << [dev = m_device, keys_in, keys_out, vals_in, vals_out, num_items, begin_bit, end_bit](cudaStream_t s) {
            dev->check(lcgs_b200_sort_pairs_u64_u32(dev->ctx(), keys_in.ptr, keys_out.ptr, vals_in.ptr, vals_out.ptr, num_items,
                                                    begin_bit, end_bit, reinterpret_cast<lcgs_b200_stream>(s)),
                       "DeviceRadixSort::SortPairs");
        };
    }

private:
    lcgs::Device* m_device = nullptr;
    lcgs::Stream* m_stream = nullptr;
};

}  // namespace luisa::parallel_primitive
