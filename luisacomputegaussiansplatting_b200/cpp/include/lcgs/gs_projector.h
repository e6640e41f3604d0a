// lcgs/gs_projector.h -- lcgs::GSProjector (reference: lcgs/include/lcgs/gs_projector.h:16-43).
#pragma once

#include "lcgs/runtime.h"
#include "lcgs/util/camera.h"

namespace lcgs
{

struct GSProjectorInputProxy {
    int               num_gaussians;
    BufferView<float> pos;
    BufferView<float> scale;
    BufferView<float> rotq;
    float             scale_modifier;
};

struct GSProjectorOutputProxy {
    BufferView<float> means_2d;
    BufferView<float> covs_2d;
    BufferView<float> depth;
};

class GSProjector : public GSModule
{
public:
    void create(Device& device) noexcept { m_device = &device; }
    // Enqueue only; the tan/focal/matrix prologue of gs_projector/impl.cpp:34-42 runs on the host
    // inside lcgs_b200_view_params_from_camera.
    void forward(CommandList& cmdlist, GSProjectorInputProxy input, GSProjectorOutputProxy output, Camera& cam,
                 bool use_focal = true) noexcept
    {
        if (!use_focal) fatal("GSProjector::forward(use_focal=false) is unreachable from the app and not implemented");
        lcgs_b200_view_params vp;
        lcgs_b200_view_params_from_camera(abi(cam), &vp);
        cmdlist << [dev = m_device, input, output, vp](cudaStream_t s) {
            dev->check(lcgs_b200_project(dev->ctx(), input.num_gaussians, input.pos.ptr, input.scale.ptr, input.rotq.ptr,
                                         input.scale_modifier, &vp, output.means_2d.ptr, output.depth.ptr, output.covs_2d.ptr,
                                         reinterpret_cast<lcgs_b200_stream>(s)),
                       "GSProjector::forward");
        };
    }
};

}  // namespace lcgs
