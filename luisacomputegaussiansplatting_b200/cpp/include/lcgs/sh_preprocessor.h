// lcgs/sh_preprocessor.h -- lcgs::SHProcessor (reference: lcgs/include/lcgs/sh_preprocessor.h:16-57).
#pragma once

#include "lcgs/runtime.h"
#include "lcgs/util/camera.h"

namespace lcgs
{

struct GPUPointsProxy {
    int               N      = 0;
    int               stride = 3;
    BufferView<float> pos;
};

class SHProcessor
{
public:
    void create(Device& device) noexcept { m_device = &device; }
    // Enqueue only.  NB the reference's implementation names the last two parameters (level, channel)
    // and forwards them swapped (sh_preprocessor.cpp:174-181), so the header's `channel` is what the
    // kernel uses as the SH degree; that positional behaviour is kept.
    void process(CommandList& cmdlist, GPUPointsProxy proxy, Camera& camera, BufferView<float> sh, BufferView<float> color,
                 int channel = 3, int level = 3) noexcept
    {
        (void)level;
        cmdlist << [dev = m_device, proxy, cam_pos = camera.position, sh, color, channel](cudaStream_t s) {
            dev->check(lcgs_b200_sh_process(dev->ctx(), proxy.N, channel, &cam_pos.x, proxy.pos.ptr, sh.ptr, color.ptr,
                                            reinterpret_cast<lcgs_b200_stream>(s)),
                       "SHProcessor::process");
        };
    }

private:
    Device* m_device = nullptr;
};

}  // namespace lcgs
