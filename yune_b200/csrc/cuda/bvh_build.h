// bvh_build.h -- device-side BVH construction behind the BVHNodeGPU contract (bvh_build.cu).
#ifndef YUNE_BVH_BUILD_H
#define YUNE_BVH_BUILD_H

#include <cuda_runtime.h>
#include <string>
#include "yune_types.h"

namespace yune {

// Everything lives in device memory and is owned by this struct's holder (free_all()).
struct GpuBvh {
    yune_bvh_node* nodes = nullptr; int n_nodes = 0;        // the reference-format array (breadth-first, siblings adjacent)
    float4* pairs = nullptr; int n_inner = 0;               // trav_layout.h: 4 x float4 per inner node, breadth-first
    float4* tris = nullptr; int n_tris = 0;                 // 3 x float4 per triangle, grouped by leaf
    float4* leaf_boxes = nullptr; int n_leaves = 0;         // 2 x float4 per leaf: its box in `nodes`
    float4* shade = nullptr; unsigned char* tri_class = nullptr;   // by original triangle index
    int root_ref = 0, depth = 0, leaf_max = 2;
    int builder = 1, rounds = 0;                            // 0 = LBVH (Karras), 1 = PLOC; PLOC merge rounds
    float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
    float build_ms = 0;                                     // device time from the end of the triangle upload to the last kernel
    void free_all();
};

// The leaves of an UPLOADED reference tree (host arrays, relayout.cpp: referenceLeavesForDevice).  With them the device-built tree
// is only the walk's own tree (accel 1): triangle records carry the reference leaf / visiting rank, the leaf-box filter tests the
// uploaded boxes, hit records are those of the uploaded tree bit for bit; no BVHNodeGPU array is emitted.
struct RefLeaves { const int* leaf_of_tri = nullptr; const int* rank_of_tri = nullptr; const float* leaf_boxes = nullptr; int n_leaves = 0; };

// h_tris: the uploaded TriangleGPU records (host); d_mats: the uploaded Material records (device).
bool buildBvhOnDevice(const yune_triangle* h_tris, int n_tris, const yune_material* d_mats, int n_mats, int leaf_max,
                      cudaStream_t stream, GpuBvh& out, std::string& err, const RefLeaves* ref = nullptr, int builder = 1, int ploc_radius = 32);

} // namespace yune
#endif
