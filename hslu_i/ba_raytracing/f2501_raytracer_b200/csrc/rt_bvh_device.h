// rt_bvh_device.h -- device-side BVH build (rt_bvh_device.cu): same device layout as the host builder (rt_bvh.h).
#pragma once
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "rt_api.h"

namespace rtb {

struct DeviceBvh {  // cudaMalloc'ed arrays on the current device; the caller owns them
    float4 *nodes = nullptr, *prims = nullptr, *shade_n = nullptr;
    uint32_t n_nodes = 0, n_slots = 0, max_depth = 0;
    int32_t nan_hit_slot = -1, nan_hit_slot_culled = -1;
    float box_pad = 0;
};

// Builds nodes / primitive records / stored normals of `desc` on the CURRENT device.  `transmissive[m]` = material m
// is transmissive (TransmissionProperties::mask).  Validates coordinates and material indices on the device.
int build_scene_on_device(const rt_scene_desc *desc, const std::vector<unsigned char> &transmissive, DeviceBvh &out);

}  // namespace rtb
