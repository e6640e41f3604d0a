// test_facade.cpp -- compile-and-run checks of the C++ facade (cpp/include) against the reference's call shapes.
//
//   test_facade camera        CPU only: a foreign camera with 16-byte float3 members (what luisa::float3 is,
//                             lcgs/include/lcgs/util/camera.h:15-25) converted with lcgs::to_abi_camera
//   test_facade splat <ply>   GPU: GSTileSplatter::forward's call sequence (lcgs/src/gs_tile_splatter/impl.cpp:63-180),
//                             restated with the reference's statements -- same lcpp / BufferFiller / copy_to /
//                             commit / synchronize calls and argument lists, the private DSL shaders replaced by the
//                             stage entry points -- must give the same buffers, bit for bit, as the facade's own
//                             GSTileSplatter::forward (one C-ABI call)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "gaussians.h"
#include "lcgs/gs_projector.h"
#include "lcgs/gs_tile_splatter.h"
#include "lcgs/sh_preprocessor.h"
#include "lcgs/util/buffer_filler.h"
#include "lcgs/util/camera.h"
#include "lcpp/parallel_primitive.h"

namespace mock_luisa
{
// luisa::float3 is a 16-byte aligned, 16-byte vector
struct alignas(16) float3 {
    float x, y, z;
};
static_assert(sizeof(float3) == 16, "mock of luisa::float3");
struct Camera {  // the reference's lcgs::Camera, member for member
    float3 position, front, up, right;
    float  fov          = 60.0f;
    float  aspect_ratio = 1.0f;
    int    width        = 512;
    int    height       = 512;
};
static_assert(sizeof(Camera) == 80, "four 16-byte vectors + four scalars");
static_assert(sizeof(Camera) != sizeof(lcgs_b200_camera), "which is why it must not be reinterpret_cast");
}  // namespace mock_luisa

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

static int test_camera()
{
    mock_luisa::Camera rc;
    rc.position = { -3.0f, -0.5f, 3.3f };
    rc.front    = { 0.1f, 0.2f, 0.3f };
    rc.up       = { 0.4f, 0.5f, 0.6f };
    rc.right    = { 0.7f, 0.8f, 0.9f };
    rc.fov = 47.0f; rc.aspect_ratio = 1.5f; rc.width = 1920; rc.height = 1080;
    const lcgs_b200_camera a = lcgs::to_abi_camera(rc);
    CHECK(a.position[0] == -3.0f && a.position[1] == -0.5f && a.position[2] == 3.3f);
    CHECK(a.front[0] == 0.1f && a.front[2] == 0.3f && a.up[1] == 0.5f && a.right[0] == 0.7f && a.right[2] == 0.9f);
    CHECK(a.fov == 47.0f && a.aspect_ratio == 1.5f && a.width == 1920 && a.height == 1080);
    // the cast the old documentation suggested reads padding as data
    const auto* wrong = reinterpret_cast<const lcgs_b200_camera*>(&rc);
    CHECK(wrong->front[0] != rc.front.x || wrong->up[0] != rc.up.x);
    // and the facade's own compact Camera agrees with the field copy
    lcgs::Camera fc = lcgs::get_lookat_cam({ -4, -4, 0 }, { 0, 0, 0 }, { 0, 0, 1 });
    const lcgs_b200_camera b = lcgs::to_abi_camera(fc);
    CHECK(std::memcmp(&b, lcgs::abi(fc), sizeof(b)) == 0);
    CHECK(std::fabs(fc.front.x - 0.70710678f) < 1e-6f && std::fabs(fc.front.y - 0.70710678f) < 1e-6f);  // test_camera.cpp:51-67
    std::printf("camera ok\n");
    return 0;
}

namespace lcgs::facade_test  // inside lcgs so that uint / ulong / float3 ... name the facade's types, as in the reference
{
using luisa::parallel_primitive::DeviceRadixSort;
using luisa::parallel_primitive::DeviceScan;

// GSTileSplatter with forward() written out the way the reference does it.  Statements follow impl.cpp:63-180
// one for one; each `(*shad_x)(args).dispatch(n)` is the stage entry point that replaces that private shader.
class ReferenceSequenceSplatter : public GSTileSplatter
{
public:
    static size_t bytes_to_uint_count(size_t bytes) { return (bytes + sizeof(uint) - 1) / sizeof(uint); }

    void ensure_scan_temp_buffer(Device& device, size_t num_items)  // impl.cpp:31-45
    {
        using ScannerT             = DeviceScan<>;
        size_t temp_bytes          = ScannerT::GetTempStorageBytes<uint>(num_items);
        size_t required_uint_count = bytes_to_uint_count(temp_bytes);
        if (m_scan_temp_buffer == nullptr || m_scan_temp_buffer_size < required_uint_count) {
            size_t new_size    = m_scan_temp_buffer_size == 0 ? required_uint_count : std::max(required_uint_count, m_scan_temp_buffer_size * 2);
            m_scan_temp_buffer = std::make_unique<Buffer<uint>>(device.create_buffer<uint>(new_size));
            m_scan_temp_buffer_size = new_size;
        }
    }
    void ensure_radix_sort_temp_buffer(Device& device, size_t num_items)  // impl.cpp:47-61
    {
        using RadixSorterT         = DeviceRadixSort<>;
        size_t temp_bytes          = RadixSorterT::GetSortPairsTempStorageBytes<ulong, uint>(static_cast<uint>(num_items));
        size_t required_uint_count = bytes_to_uint_count(temp_bytes);
        if (m_radix_sort_temp_buffer == nullptr || m_radix_sort_temp_buffer_size < required_uint_count) {
            size_t new_size = m_radix_sort_temp_buffer_size == 0 ? required_uint_count : std::max(required_uint_count, m_radix_sort_temp_buffer_size * 2);
            m_radix_sort_temp_buffer = std::make_unique<Buffer<uint>>(device.create_buffer<uint>(new_size));
            m_radix_sort_temp_buffer_size = new_size;
        }
    }

    int forward(Device& device, Stream& stream, GSTileSplatterAccelProxy accel, GSTileSplatterInputProxy input,
                GSSplatForwardOutputProxy output, bool use_focal = true) noexcept override
    {
        (void)use_focal;
        auto width  = output.width;
        auto height = output.height;
        auto grids  = make_uint2((unsigned int)((width + m_blocks.x - 1u) / m_blocks.x), (unsigned int)((height + m_blocks.y - 1u) / m_blocks.y));

        int  num_gaussians   = input.num_gaussians;
        auto d_point_offsets = accel.point_offsets.subview(0, num_gaussians);
        auto d_tiles_touched = accel.tiles_touched.subview(0, num_gaussians);
        lcgs_b200_ctx* ctx   = device.ctx();
        auto           abi_s = [](cudaStream_t s) { return reinterpret_cast<lcgs_b200_stream>(s); };

        CommandList cmdlist;
        cmdlist << [=, &device](cudaStream_t s) {  // (*shad_allocate_tiles)(...).dispatch(num_gaussians)
            device.check(lcgs_b200_allocate_tiles(ctx, num_gaussians, width, height, input.depth_features.ptr, input.means_2d.ptr,
                                                  input.conic.ptr, d_tiles_touched.ptr, output.radii.ptr, 0, -1, abi_s(s)),
                         "allocate_tiles");
        };
        stream << cmdlist.commit() << synchronize();

        ensure_scan_temp_buffer(device, num_gaussians);
        mp_device_scan->InclusiveSum(cmdlist, m_scan_temp_buffer->view(), d_tiles_touched, d_point_offsets, num_gaussians);

        cmdlist << accel.point_offsets.subview(input.num_gaussians - 1, 1).copy_to(&num_rendered);
        stream << cmdlist.commit() << synchronize();

        if (num_rendered <= 0) { return 0; }

        auto d_point_list_unsorted      = accel.point_list_unsorted.subview(0, num_rendered);
        auto d_point_list_keys_unsorted = accel.point_list_keys_unsorted.subview(0, num_rendered);
        auto d_point_list               = accel.point_list.subview(0, num_rendered);
        auto d_point_list_keys          = accel.point_list_keys.subview(0, num_rendered);

        cmdlist << mp_buffer_filler->fill(device, d_point_list_unsorted, 0u);
        cmdlist << mp_buffer_filler->fill(device, d_point_list_keys_unsorted, 0ull);

        cmdlist << [=, &device](cudaStream_t s) {  // (*shad_copy_with_keys)(...).dispatch(num_gaussians)
            device.check(lcgs_b200_duplicate_keys(ctx, num_gaussians, width, height, input.means_2d.ptr, d_point_offsets.ptr,
                                                  output.radii.ptr, input.depth_features.ptr, d_point_list_keys_unsorted.ptr,
                                                  d_point_list_unsorted.ptr, d_point_list_unsorted.size(), 0, -1, abi_s(s)),
                         "copy_with_keys");
        };
        stream << cmdlist.commit() << synchronize();

        ensure_radix_sort_temp_buffer(device, num_rendered);
        mp_device_radix_sort->SortPairs<ulong, uint>(cmdlist, m_radix_sort_temp_buffer->view(), d_point_list_keys_unsorted,
                                                     d_point_list_keys, d_point_list_unsorted, d_point_list, num_rendered);
        stream << cmdlist.commit() << synchronize();
        auto d_ranges = accel.ranges.subview(0, grids.x * grids.y * 2);
        stream << cmdlist.commit() << synchronize();
        cmdlist << mp_buffer_filler->fill(device, d_ranges, 0u);

        const int n = num_rendered;
        cmdlist << [=, &device](cudaStream_t s) {  // (*shad_get_ranges)(...).dispatch(num_rendered)
            device.check(lcgs_b200_tile_ranges(ctx, d_point_list_keys.ptr, (size_t)n, d_ranges.ptr, (int)(grids.x * grids.y), abi_s(s)),
                         "get_ranges");
        };
        cmdlist << [=, &device](cudaStream_t s) {  // (*m_forward_render_shader)(...).dispatch(resolution)
            device.check(lcgs_b200_blend(ctx, num_gaussians, width, height, &input.bg_color.x, d_ranges.ptr, d_point_list.ptr,
                                         input.means_2d.ptr, input.conic.ptr, input.opacity_features.ptr, input.color_features.ptr,
                                         d_tiles_touched.ptr, output.target_img.ptr, 0, -1, abi_s(s)),
                         "forward_render");
        };
        stream << cmdlist.commit();
        return num_rendered;
    }

private:
    std::unique_ptr<Buffer<uint>> m_scan_temp_buffer, m_radix_sort_temp_buffer;
    size_t                        m_scan_temp_buffer_size = 0, m_radix_sort_temp_buffer_size = 0;
};

template <typename T>
std::vector<T> fetch(Stream& stream, BufferView<T> v, size_t n)
{
    std::vector<T> h(n);
    stream << v.subview(0, n).copy_to(h.data()) << synchronize();
    return h;
}

static int test_splat(const char* ply, int w, int h)
{
    Device device(0);
    Stream stream;
    GaussiansData data;
    std::string   err;
    if (!read_gs_ply(data, ply, &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const int    P = data.num_gaussians;
    const size_t L = 4000000;

    BufferFiller      bf;
    DeviceScan<>      device_scan;
    DeviceRadixSort<> device_radix_sort;
    device_scan.create(device, &stream);
    device_radix_sort.create(device, &stream);
    SHProcessor sh_processor;
    sh_processor.create(device);
    GSProjector projector;
    projector.create(device);

    auto d_pos = device.create_buffer<float>((size_t)P * 3), d_scale = device.create_buffer<float>((size_t)P * 3);
    auto d_rotq = device.create_buffer<float>((size_t)P * 4), d_sh = device.create_buffer<float>((size_t)P * 48);
    auto d_opacity = device.create_buffer<float>((size_t)P);
    Camera cam = get_lookat_cam({ -3.0f, -0.5f, 3.3f }, { 0.0f, 3.0f, 0.5f }, { 0.0f, -1.0f, -1.0f });
    cam.aspect_ratio = (float)w / (float)h; cam.width = w; cam.height = h;
    CommandList up;
    up << d_pos.copy_from(data.pos.data()) << d_scale.copy_from(data.scale.data()) << d_rotq.copy_from(data.rotq.data())
       << d_sh.copy_from(data.feature.data()) << d_opacity.copy_from(data.opacity.data());
    stream << up.commit() << synchronize();  // main.cpp:216-223

    struct Result {
        int                num_rendered;
        std::vector<ulong> keys;
        std::vector<uint>  vals, ranges, tiles;
        std::vector<int>   radii;
        std::vector<float> img, means, conic;
    };
    auto run = [&](GSTileSplatter& splatter) {
        splatter.create(device);
        splatter.set_buffer_filler(&bf);
        splatter.set_device_scan(&device_scan);
        splatter.set_device_radix_sort(&device_radix_sort);
        const auto tw = (w + splatter.m_blocks.x - 1u) / splatter.m_blocks.x, th = (h + splatter.m_blocks.y - 1u) / splatter.m_blocks.y;
        auto d_color = device.create_buffer<float>((size_t)P * 3), d_means_2d = device.create_buffer<float>((size_t)P * 2);
        auto d_depth = device.create_buffer<float>((size_t)P), d_covs_2d = device.create_buffer<float>((size_t)P * 3);
        auto d_tiles = device.create_buffer<uint>((size_t)P), d_offsets = device.create_buffer<uint>((size_t)P);
        auto d_ku = device.create_buffer<ulong>(L), d_k = device.create_buffer<ulong>(L);
        auto d_vu = device.create_buffer<uint>(L), d_v = device.create_buffer<uint>(L);
        auto d_ranges = device.create_buffer<uint>((size_t)tw * th * 2);
        auto d_img = device.create_buffer<float>((size_t)w * h * 3);
        auto d_radii = device.create_buffer<int>((size_t)P);
        CommandList cmd_list;
        sh_processor.process(cmd_list, { P, 3, d_pos }, cam, d_sh, d_color, 3, 3);                                          // main.cpp:268
        projector.forward(cmd_list, { P, d_pos, d_scale, d_rotq, 1.0f }, { d_means_2d, d_covs_2d, d_depth }, cam);          // :269
        stream << cmd_list.commit();                                                                                        // :270
        GSSplatForwardOutputProxy output{ h, w, d_img, d_radii };
        GSTileSplatterAccelProxy  accel{ d_tiles, d_offsets, d_ku, d_vu, d_k, d_v, d_ranges };
        GSTileSplatterInputProxy  input{ P, make_float3(0.f), d_means_2d, d_depth, d_covs_2d, d_color, d_opacity };
        Result r;
        r.num_rendered = splatter.forward(device, stream, accel, input, output);                                            // :299
        stream << synchronize();
        const size_t n = (size_t)std::max(r.num_rendered, 0);
        r.keys = fetch(stream, d_k.view(), n); r.vals = fetch(stream, d_v.view(), n);
        r.ranges = fetch(stream, d_ranges.view(), (size_t)tw * th * 2); r.tiles = fetch(stream, d_tiles.view(), (size_t)P);
        r.radii = fetch(stream, d_radii.view(), (size_t)P); r.img = fetch(stream, d_img.view(), (size_t)w * h * 3);
        r.means = fetch(stream, d_means_2d.view(), (size_t)P * 2); r.conic = fetch(stream, d_covs_2d.view(), (size_t)P * 3);
        return r;
    };
    GSTileSplatter            one_call;
    ReferenceSequenceSplatter by_the_book;
    const Result a = run(one_call), b = run(by_the_book);
    CHECK(a.num_rendered > 0 && a.num_rendered == b.num_rendered);
    CHECK(a.keys == b.keys && a.vals == b.vals && a.ranges == b.ranges && a.tiles == b.tiles && a.radii == b.radii);
    CHECK(a.img.size() == b.img.size() && std::memcmp(a.img.data(), b.img.data(), a.img.size() * 4) == 0);
    CHECK(std::memcmp(a.means.data(), b.means.data(), a.means.size() * 4) == 0);
    CHECK(std::memcmp(a.conic.data(), b.conic.data(), a.conic.size() * 4) == 0);
    std::printf("splat ok: num_rendered %d\n", a.num_rendered);
    return 0;
}
}  // namespace lcgs::facade_test

int main(int argc, char** argv)
{
    if (argc >= 2 && std::string(argv[1]) == "camera") return test_camera();
    if (argc >= 3 && std::string(argv[1]) == "splat") return lcgs::facade_test::test_splat(argv[2], argc >= 4 ? std::atoi(argv[3]) : 512, argc >= 5 ? std::atoi(argv[4]) : 288);
    std::fprintf(stderr, "usage: test_facade camera | splat <ply> [W H]\n");
    return 2;
}
