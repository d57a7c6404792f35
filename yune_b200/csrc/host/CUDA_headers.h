// CUDA_headers.h -- the product's counterpart of the reference's include/CL_headers.h:57-144.
// Same record names the reference's host code uses (Cam, TriangleGPU, AABB, BVHNodeGPU, Material), defined
// as aliases of the C-ABI PODs in include/yune_types.h so that a std::vector<TriangleGPU> can be handed to
// the C ABI without conversion.  No OpenCL typedefs are needed any more.
#ifndef YUNE_CUDA_HEADERS_H
#define YUNE_CUDA_HEADERS_H

#include "yune_types.h"

namespace yune
{
    typedef yune_float4     Float4;
    typedef yune_cam        Cam;
    typedef yune_triangle   TriangleGPU;
    typedef yune_aabb       AABB;
    typedef yune_bvh_node   BVHNodeGPU;
    typedef yune_material   Material;
    typedef yune_quad_light Quad;

    /** Default material, same values as newMaterial() (include/CL_headers.h:109-124). */
    inline Material newMaterial()
    {
        Material m;
        m.ke = {{0.f, 0.f, 0.f, 1.f}};
        m.kd = {{0.3f, 0.3f, 0.3f, 1.f}};
        m.ks = {{0.f, 0.f, 0.f, 1.f}};
        m.n = 1.f; m.k = 1.f; m.px = 1.f; m.py = 1.f;
        m.alpha_x = 100.f; m.alpha_y = 100.f;
        m.is_specular = 0; m.is_transmissive = 0;
        return m;
    }
}
#endif
