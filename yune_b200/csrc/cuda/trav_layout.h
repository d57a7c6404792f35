// trav_layout.h -- the traversal-side layout of the reference BVH and triangles in HBM.
//
// The drop-in boundary receives the reference's AoS records (BVHNodeGPU 80 B, TriangleGPU 112 B,
// include/CL_headers.h:67-92).  Traversal never touches those: at upload they are re-laid-out ONCE into
//
//   pair records  (64 B, one per INNER node, breadth-first order so the hot top of the tree is a prefix):
//       q0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)      c0/c1 = the two (adjacent) children
//       q1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//       q2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//       q3 = (ref0, ref1, -, -) as int bits
//     one fetch of four 16-byte vectors tests both children; the 40 dead bytes of vert_list are gone and no
//     node is read twice (the reference reads each node as a child and again as a parent, udpt.cl:300-316).
//   triangle records (48 B, grouped by leaf, leaves in node-index order):
//       t0 = (v1.xyz, original triangle index as int bits), t1 = (v2 - v1, RANK as int bits), t2 = (v3 - v1, 0)
//     RANK = the reference's breadth-first visiting order of that triangle slot; it breaks exact ties in t the way the
//     reference's first-come-first-kept rule does (udpt.cl:373).
//   shading records (64 B, by ORIGINAL triangle index): normalised vertex normals + matID
//       s0 = (N1.xyz, matID bits), s1 = (N2.xyz, 0), s2 = (N3.xyz, 0), s3 spare
//
// child ref: >= 0 inner (pair index); < 0 leaf with ~ref = (first << 4) | count; YUNE_REF_EMPTY = never visit
// (the reference's 'vert_len > 0 || child_idx > 0' filter, udpt.cl:316).
#ifndef YUNE_TRAV_LAYOUT_H
#define YUNE_TRAV_LAYOUT_H

#include "yune_types.h"
#include <vector>
#include <string>

#define YUNE_REF_EMPTY ((int)0x80000000)
#define YUNE_STACK_SIZE 64

namespace yune {

struct F4 { float x, y, z, w; };

struct TravLayoutHost {
    std::vector<F4> pairs;       // 4 per inner node
    std::vector<F4> tris;        // 3 per leaf-ordered triangle
    std::vector<F4> shade;       // 4 per original triangle
    std::vector<F4> leaf_boxes;  // accel 1: 2 per REFERENCE leaf (p_min, p_max exactly as uploaded), indexed by the id in triangle record t2.w
    std::vector<F4> quads;       // accel 2: 7 per wide node: lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4] refs[4]; unused slot = inverted box + YUNE_REF_EMPTY
    int   n_wide = 0, root_wide_ref = YUNE_REF_EMPTY, wide_depth = 0;
    int   accel = 0;
    int   isect = 0;             // 1: `tris` holds the three RAW vertices (t0 = v1, t1 = v2, t2 = v3; same w fields) for the watertight test
    int   n_inner = 0, n_inner_ref = 0, n_leaf_tris = 0, n_tris = 0;   // n_inner counts refinement pairs too; n_inner_ref = reference inner nodes
    int   root_ref = YUNE_REF_EMPTY;
    float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
    int   max_depth = 0;         // stack entries a depth-first walk can need
};

// Builds the layout; returns false and sets `err` when the input is malformed (child index out of range,
// triangle index out of range, leaf with more than 10 slots, tree deeper than YUNE_STACK_SIZE).
// leaf_split: 0 = keep the reference's leaves (up to 10 triangles); N > 0 = refine every leaf holding more than N triangles
// with a private, padded subtree (see relayout.cpp) -- same hits, fewer triangle tests.
//
// accel: 0 = walk the reference tree itself (pair records = the reference's boxes, entered under the reference's predicate).
//        1 = walk OUR OWN binned-SAH tree over the same triangles (tight boxes, conservatively padded, leaves of <= 2), and
//            decide per candidate triangle whether the REFERENCE would have reached it by testing the box of its reference
//            leaf with the reference's exact predicate.  Sound because every reference box contains its children's boxes
//            exactly (getExtent / refit, src/BVH.cpp:218-278) and float subtraction / multiplication are monotone, so
//            "the leaf's box passes" implies "every ancestor's box passes" (rays with a non-finite 1/d excepted).
//            Same hit records, roughly half the box tests and a third of the triangle tests.
//
//        2 = accel 1's tree collapsed into nodes of up to FOUR children (`quads`; half the steps per ray at the same number of
//            box tests).  Layout and walk are pinned against the oracle on the host (tests/test_traversal_hostcheck.py) and on the
//            B200 (tests/test_gpu_parity.py); kernel variant k_trace<., 5 / 7>.  Measured slower than accel 1.
//
// n_nodes == 0: the reference's brute-force mode (udpt.cl:280-284).  Always built like accel 1, with ONE pseudo-leaf whose box
//        every ray passes and visiting rank = triangle index: the hit records of the reference's loop over all triangles.
//
// own_opt_passes: reinsertion passes over the finished own tree (relayout.cpp: OptTree; accel 1 / 2); -1 = 2 up to 2^18 triangles, else 0.
//
// isect: 0 = the reference's Moller-Trumbore (parity mode: hit records bit-identical to the reference's);
//        1 = watertight signed-volume test on the raw vertices, no leaf-box filter (perf mode, own tree only: accel 1).  Differs from
//            mode 0 only for rays within rounding distance of an edge / vertex or of a reference box face (tests pin the rate).
bool buildTravLayout(const yune_triangle* tris, int n_tris, const yune_bvh_node* nodes, int n_nodes,
                     TravLayoutHost& out, std::string& err, int leaf_split = 0, int accel = 0, int isect = 0, int own_opt_passes = -1);

// Reference leaves of an uploaded tree in the form the device layout path (bvh_build.cu) takes them; see relayout.cpp.
bool referenceLeavesForDevice(const yune_triangle* tris, int n_tris, const yune_bvh_node* nodes, int n_nodes,
                              std::vector<int>& leaf_of_tri, std::vector<int>& rank_of_tri, std::vector<F4>& leaf_boxes, bool& usable, std::string& err);

} // namespace yune
#endif
