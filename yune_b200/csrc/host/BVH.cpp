// BVH.cpp -- see BVH.h.  Follows the decisions of src/BVH.cpp:56-320 (reference) step by step; comments
// cite the line that fixes each float32 expression, because the output must match bit for bit.
#include "BVH.h"

#include <algorithm>
#include <limits>
#include <stdexcept>
#include <cstdint>
#include <functional>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>

namespace yune
{
    void TriangleCPU::computeCentroid()
    {
        // src/TriangleCPU.cpp:41-49 -- (v1 + v2) + v3, then / 3.0f, w forced to 1
        for (int k = 0; k < 3; k++) {
            float s = props.v1.s[k] + props.v2.s[k];
            s = s + props.v3.s[k];
            centroid.s[k] = s / 3.0f;
        }
        centroid.s[3] = 1.0f;
        computeAABB();
    }

    void TriangleCPU::computeAABB()
    {
        // src/TriangleCPU.cpp:51-71
        for (int k = 0; k < 3; k++) {
            float lo = std::min(props.v1.s[k], std::min(props.v2.s[k], props.v3.s[k]));
            float hi = std::max(props.v1.s[k], std::max(props.v2.s[k], props.v3.s[k]));
            if (hi - lo == 0.0f) hi += 0.2f;          // flat triangles get a 0.2 slab (:63-67)
            aabb.p_min.s[k] = lo;
            aabb.p_max.s[k] = hi;
        }
        aabb.p_min.s[3] = 1.0f;
        aabb.p_max.s[3] = 1.0f;
    }

    BVH::BVH() : bins(20), bvh_size_kb(0), bvh_size_mb(0), leaf_primitives(10), cost_isect(1.0f), cost_trav(1 / 8.0f) {}

    namespace
    {
        // src/BVH.cpp:316-320
        inline float surfaceArea(const AABB& b)
        {
            float dx = b.p_max.s[0] - b.p_min.s[0], dy = b.p_max.s[1] - b.p_min.s[1], dz = b.p_max.s[2] - b.p_min.s[2];
            return 2 * (dx * dy + dx * dz + dy * dz);
        }
        // src/BVH.cpp:268-278: min/max with bb2 as std::min's first argument
        inline void extend(AABB& acc, const AABB& bb2)
        {
            for (int k = 0; k < 3; k++) {
                acc.p_min.s[k] = std::min(bb2.p_min.s[k], acc.p_min.s[k]);
                acc.p_max.s[k] = std::max(bb2.p_max.s[k], acc.p_max.s[k]);
            }
        }
        inline AABB emptyBox()
        {
            const float big = std::numeric_limits<float>::max();
            AABB b; b.p_min = {{big, big, big, 1.0f}}; b.p_max = {{-big, -big, -big, 1.0f}};
            return b;
        }
        struct BuildNode { int begin, end; };   // range in the shared primitive array

        inline BVHNodeGPU blankNode(const AABB& box)
        {   // BVHNodeCPU(AABB) (src/BVHNodeCPU.cpp:42-52): empty until proven otherwise
            BVHNodeGPU n; n.aabb = box;
            for (int j = 0; j < 10; j++) n.vert_list[j] = -1;
            n.child_idx = -2; n.vert_len = -1;
            return n;
        }
    }

    void BVH::createBVH(AABB root, const std::vector<TriangleCPU>& tris, int bvh_bins)
    {
        const auto t_begin = std::chrono::steady_clock::now();
        gpu_node_list.clear();
        bvh_size_kb = bvh_size_mb = 0;
        bins = bvh_bins;

        // Build records travel WITH the partition (centroid, padded box and index, 64 bytes): every pass over a node's range
        // is then a sequential read, where `tris[prims[j]]` was a random 176-byte gather -- 10.5 M triangles built in 17.8 s,
        // almost all of it cache misses.  Same decisions, same output.
        struct Prim { Float4 c; AABB b; int id; int pad[3]; };
        const int n_tris = (int)tris.size();
        std::vector<Prim> prims(n_tris), scratch(n_tris);
        for (int i = 0; i < n_tris; i++) { prims[i].c = tris[i].centroid; prims[i].b = tris[i].aabb; prims[i].id = i; }

        std::vector<BVHNodeGPU>& nodes = gpu_node_list;
        std::vector<BuildNode> ranges;
        nodes.push_back(blankNode(root));
        ranges.push_back({0, n_tris});
        const size_t node_cap = (size_t)n_tris * 64 + 1024;    // the reference has no guard (appendix B#20)

        // The reference appends children while it walks the node array (:56-173), i.e. level by level.  The nodes of one level
        // own disjoint primitive ranges, so their decisions (leaf / split plane / partition) are taken in parallel; the children
        // are then appended in node order, which reproduces the reference's indices exactly.
        struct Decision { bool split; int n1; AABB c1_box, c2_box; };
        const unsigned hw = std::thread::hardware_concurrency();
        int max_threads = (int)std::min<unsigned>(hw ? hw : 1u, 32u);
        if (const char* e = std::getenv("YUNE_BVH_THREADS")) max_threads = std::max(1, std::min(64, std::atoi(e)));   // 1 = the sequential build
        const int big_node = 1 << 17;                           // nodes with at least this many primitives split their own passes over threads
        // fn(segment, lo, hi) on `n_seg` contiguous segments of [begin, end), one thread each (segment 0 on the caller)
        auto segments = [&](int begin, int end, int n_seg, const std::function<void(int, int, int)>& fn) {
            const long long count = end - begin;
            std::vector<std::thread> pool;
            for (int t = 1; t < n_seg; t++)
                pool.emplace_back(fn, t, begin + (int)(count * t / n_seg), begin + (int)(count * (t + 1) / n_seg));
            fn(0, begin, begin + (int)(count / n_seg));
            for (auto& th : pool) th.join();
        };
        auto decide = [&](const size_t i, Decision& out, const bool inner_parallel)
        {
            out.split = false;
            const int begin = ranges[i].begin, end = ranges[i].end, count = end - begin;
            if (count == 0) return;                             // empty child (:69-70)
            const AABB parent_box = nodes[i].aabb;              // the SPATIAL box handed down by the split, not yet refit

            if (count <= leaf_primitives) {                     // :72-79
                nodes[i].vert_len = count; nodes[i].child_idx = -1;
                for (int j = 0; j < count; j++) nodes[i].vert_list[j] = prims[begin + j].id;
                return;
            }

            // split axis = longest side of the union of the primitives' padded boxes; first maximum wins (:280-314)
            const int n_seg = (inner_parallel && count >= big_node) ? max_threads : 1;
            AABB ext = emptyBox();
            if (n_seg == 1) { for (int j = begin; j < end; j++) extend(ext, prims[j].b); }
            else {                                              // min / max do not depend on the order
                std::vector<AABB> part(n_seg, emptyBox());
                segments(begin, end, n_seg, [&](int t, int lo, int hi) { AABB a = emptyBox(); for (int j = lo; j < hi; j++) extend(a, prims[j].b); part[t] = a; });
                for (int t = 0; t < n_seg; t++) extend(ext, part[t]);
            }
            int axis = 0; float best_len = -std::numeric_limits<float>::max();
            for (int k = 0; k < 3; k++) { float d = ext.p_max.s[k] - ext.p_min.s[k]; if (d > best_len) { best_len = d; axis = k; } }

            // A primitive goes to child 1 iff its centroid is inside child 1's box, borders included (:197-211).
            // Child 1 is the parent box with p_max[axis] lowered to the plane, so the test factors into a
            // plane-independent part and 'centroid[axis] <= plane'.
            auto inside_rest = [&](const Float4& c) {
                for (int k = 0; k < 3; k++) {
                    if (c.s[k] < parent_box.p_min.s[k]) return false;
                    if (k != axis && c.s[k] > parent_box.p_max.s[k]) return false;
                }
                return true;
            };

            float plane = 0.0f; bool have_split = false;
            AABB c1_box = parent_box, c2_box = parent_box;
            if (bins > 2 && count > 20)                         // binned SAH (:85-142)
            {
                const float parent_sa = surfaceArea(parent_box);
                float cost = parent_sa * cost_isect * (float)(size_t)count;     // :91 (not normalised, sic)
                const float initial_cost = cost;
                float increment = parent_box.p_max.s[axis] - parent_box.p_min.s[axis];
                increment /= bins;                                               // :93-94
                const int n_planes = bins - 1;
                std::vector<float> planes(n_planes);
                for (int p = 0; p < n_planes; p++) planes[p] = parent_box.p_min.s[axis] + (p + 1) * increment;   // :103

                // histogram: hist[p] = primitives whose first "inside child 1" plane is p
                std::vector<int> hist(n_planes + 1, 0);
                auto count_range = [&](int lo, int hi, int* h) {
                    for (int j = lo; j < hi; j++) {
                        const Float4& c = prims[j].c;
                        int p = n_planes;
                        if (inside_rest(c)) { p = 0; while (p < n_planes && c.s[axis] > planes[p]) p++; }
                        h[p]++;
                    }
                };
                if (n_seg == 1) count_range(begin, end, hist.data());
                else {
                    std::vector<int> part((size_t)n_seg * (n_planes + 1), 0);
                    segments(begin, end, n_seg, [&](int t, int lo, int hi) { count_range(lo, hi, part.data() + (size_t)t * (n_planes + 1)); });
                    for (int t = 0; t < n_seg; t++) for (int p = 0; p <= n_planes; p++) hist[p] += part[(size_t)t * (n_planes + 1) + p];
                }
                int in_c1 = 0, best_plane = -1;
                for (int p = 0; p < n_planes; p++) {
                    in_c1 += hist[p];
                    AABB b1 = parent_box, b2 = parent_box;
                    b1.p_max.s[axis] = planes[p]; b2.p_min.s[axis] = planes[p];
                    const float sa1 = surfaceArea(b1), sa2 = surfaceArea(b2);
                    size_t n1, n2;                              // populateChildNodes' degenerate-box rules (:177-192)
                    if (sa1 == 0.0f)      { n1 = 0; n2 = (size_t)count; }
                    else if (sa2 == 0.0f) { n1 = (size_t)count; n2 = 0; }
                    else                  { n1 = (size_t)in_c1; n2 = (size_t)(count - in_c1); }
                    const float new_cost = cost_trav + (sa1 / parent_sa) * (cost_isect * n1) + sa2 / parent_sa * (cost_isect * n2);   // :113-115
                    if (new_cost < cost) { cost = new_cost; best_plane = p; }
                }
                if (cost != initial_cost) { plane = planes[best_plane]; have_split = true; }    // else: fall back to the median (:127-128)
            }
            if (!have_split) {                                  // spatial median (:145-160)
                float center = parent_box.p_min.s[axis] + parent_box.p_max.s[axis];
                plane = center / 2.0f;
            }
            c1_box.p_max.s[axis] = plane;
            c2_box.p_min.s[axis] = plane;

            // stable partition, with the same degenerate-box rules
            int n1 = 0, n2 = 0;
            const float sa1 = surfaceArea(c1_box), sa2 = surfaceArea(c2_box);
            if (sa1 == 0.0f) { n2 = count; }
            else if (sa2 == 0.0f) { n1 = count; }
            else if (n_seg == 1) {
                for (int j = begin; j < end; j++) {
                    const Float4& c = prims[j].c;
                    const bool in1 = inside_rest(c) && !(c.s[axis] > plane);
                    if (in1) prims[begin + n1++] = prims[j];     // safe: n1 <= j - begin
                    else scratch[begin + n2++] = prims[j];
                }
                for (int j = 0; j < n2; j++) prims[begin + n1 + j] = scratch[begin + j];
            } else {
                // the same stable partition in two sweeps: count per segment, then every segment writes its two runs where the
                // sequential sweep would have put them (segments are in order, so the order inside both children is kept)
                std::vector<int> c1(n_seg, 0), len(n_seg, 0);
                segments(begin, end, n_seg, [&](int t, int lo, int hi) {
                    int k = 0;
                    for (int j = lo; j < hi; j++) { const Float4& c = prims[j].c; k += (inside_rest(c) && !(c.s[axis] > plane)) ? 1 : 0; }
                    c1[t] = k; len[t] = hi - lo;
                });
                std::vector<int> off1(n_seg, 0), off2(n_seg, 0);
                for (int t = 0; t < n_seg; t++) { off1[t] = n1; n1 += c1[t]; }
                for (int t = 0; t < n_seg; t++) { off2[t] = n1 + n2; n2 += len[t] - c1[t]; }
                segments(begin, end, n_seg, [&](int t, int lo, int hi) {
                    int a = begin + off1[t], b = begin + off2[t];
                    for (int j = lo; j < hi; j++) {
                        const Float4& c = prims[j].c;
                        if (inside_rest(c) && !(c.s[axis] > plane)) scratch[a++] = prims[j]; else scratch[b++] = prims[j];
                    }
                });
                segments(begin, end, n_seg, [&](int, int lo, int hi) { std::copy(scratch.begin() + lo, scratch.begin() + hi, prims.begin() + lo); });
            }

            out.split = true; out.n1 = n1; out.c1_box = c1_box; out.c2_box = c2_box;
        };
        size_t level_begin = 0, level_end = 1;
        std::vector<Decision> dec;
        while (level_begin < level_end)
        {
            const size_t width = level_end - level_begin;
            dec.assign(width, Decision());
            const int n_threads = (n_tris >= (1 << 16) && width >= 8) ? (int)std::min<size_t>((size_t)max_threads, width / 4) : 1;
            auto is_big = [&](size_t k) { return max_threads > 1 && ranges[level_begin + k].end - ranges[level_begin + k].begin >= big_node; };
            for (size_t k = 0; k < width; k++) if (is_big(k)) decide(level_begin + k, dec[k], true);      // one at a time, threads inside
            if (n_threads <= 1) {
                for (size_t k = 0; k < width; k++) if (!is_big(k)) decide(level_begin + k, dec[k], false);
            } else {
                std::atomic<size_t> next(0);
                std::vector<std::thread> pool;
                for (int t = 0; t < n_threads; t++)
                    pool.emplace_back([&]() { for (size_t k; (k = next.fetch_add(1)) < width; ) if (!is_big(k)) decide(level_begin + k, dec[k], false); });
                for (auto& th : pool) th.join();
            }
            for (size_t k = 0; k < width; k++) {
                if (!dec[k].split) continue;
                const size_t i = level_begin + k;
                const int begin = ranges[i].begin, end = ranges[i].end;
                if (nodes.size() + 2 > node_cap)
                    throw std::runtime_error("BVH build does not terminate: more than 10 triangles share a centroid (duplicate geometry)");
                nodes[i].child_idx = (int)nodes.size();             // :161-163
                nodes.push_back(blankNode(dec[k].c1_box)); ranges.push_back({begin, begin + dec[k].n1});
                nodes.push_back(blankNode(dec[k].c2_box)); ranges.push_back({begin + dec[k].n1, end});
            }
            level_begin = level_end; level_end = nodes.size();
        }

        const auto t_split = std::chrono::steady_clock::now();
        // bottom-up refit (resizeBvh, :218-247): leaves take the union of their triangles' padded boxes, inner nodes the
        // union of their two children -- including an empty child's spatial box, which is never refit.
        for (int j = (int)nodes.size() - 1; j >= 0; j--) {
            if (nodes[j].vert_len > 0) {
                AABB acc = emptyBox();
                for (int p = ranges[j].begin; p < ranges[j].end; p++) extend(acc, prims[p].b);
                nodes[j].aabb.p_min = acc.p_min; nodes[j].aabb.p_max = acc.p_max;
            } else if (nodes[j].child_idx > 0) {
                AABB acc = nodes[nodes[j].child_idx].aabb;
                extend(acc, nodes[nodes[j].child_idx + 1].aabb);
                nodes[j].aabb = acc;
            }
        }
        bvh_size_kb = (float)nodes.size() * sizeof(BVHNodeGPU) / 1024;
        bvh_size_mb = bvh_size_kb / 1024;
        if (std::getenv("YUNE_BVH_TIMING")) {
            const auto t_end = std::chrono::steady_clock::now();
            std::fprintf(stderr, "createBVH: %d triangles, %zu nodes: split %.2f s, refit %.2f s\n", n_tris, nodes.size(),
                         std::chrono::duration<double>(t_split - t_begin).count(), std::chrono::duration<double>(t_end - t_split).count());
        }
    }
}
