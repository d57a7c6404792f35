// BVH.h -- CPU binned-SAH / spatial-median BVH builder with the reference's exact output contract.
//
// Behavioural restatement of yune::BVH (include/BVH.h:37-68, src/BVH.cpp:56-320): same breadth-first node
// order, same split decisions in float32, same bottom-up refit, so gpu_node_list is byte-identical to the
// reference's for the same triangles (the bytes the reference leaves uninitialised -- vert_list slots past
// vert_len, src/BVHNodeCPU.cpp:30-40 -- are defined here as -1).  The implementation is different: one
// shared primitive-index array that is stable-partitioned in place of per-node std::vectors, and the 19
// candidate planes of a node are scored from one histogram pass instead of 19 re-partitions.
#ifndef YUNE_BVH_H
#define YUNE_BVH_H

#include "CUDA_headers.h"
#include <vector>

namespace yune
{
    /** Per-triangle build input: centroid and padded box (src/TriangleCPU.cpp:41-71). */
    struct TriangleCPU
    {
        TriangleGPU props;
        Float4 centroid;
        AABB aabb;
        void computeCentroid();   /**< (v1+v2)+v3, /3, then computeAABB(). */
        void computeAABB();       /**< min/max of the corners; zero-extent axes get p_max += 0.2. */
    };

    class BVH
    {
        public:
            BVH();
            std::vector<BVHNodeGPU> gpu_node_list;
            /** Build over cpu_tri_list inside the scene box `root` (src/BVH.cpp:56-173). Throws std::runtime_error
             *  when the split recursion cannot terminate (coincident centroids). */
            void createBVH(AABB root, const std::vector<TriangleCPU>& cpu_tri_list, int bvh_bins = 20);
            int bins;
            float bvh_size_kb, bvh_size_mb;

        private:
            int leaf_primitives;
            float cost_isect, cost_trav;
    };
}
#endif
