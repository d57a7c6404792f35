// relayout.cpp -- host-side, once-per-upload conversion of the reference AoS buffers into the traversal
// layout described in trav_layout.h.  Arithmetic here (edge vectors, normalised normals) uses the same
// IEEE float32 expressions the reference evaluates per intersection (udpt.cl:328-329, 377-379), so moving
// them to upload time does not change a single bit of any hit record.
#include "trav_layout.h"
#include "strict_math.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <chrono>
#include <cstdio>

namespace yune {

static inline float bits(int i) { float f; std::memcpy(&f, &i, 4); return f; }

// fn(i) for i in [0, n) on the host's threads (independent iterations only; YUNE_BVH_THREADS=1 runs it in place)
template <class F> static void parallel_for(size_t n, const F& fn)
{
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("YUNE_BVH_THREADS")) hw = (unsigned)std::max(1, std::atoi(e));
    const size_t n_thr = n < (1u << 16) ? 1 : std::min<size_t>(hw ? hw : 1, 32);
    if (n_thr <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_thr; t++)
        pool.emplace_back([&fn, n, n_thr, t]() { for (size_t i = n * t / n_thr, e = n * (t + 1) / n_thr; i < e; i++) fn(i); });
    for (auto& th : pool) th.join();
}

namespace {

// ---- leaf refinement -------------------------------------------------------------------------------------------------
// The reference builder stops at 10 triangles per leaf (src/BVH.cpp:41, 72) and its leaves mix huge wall triangles with
// tiny teapot ones, so a ray that enters a leaf pays up to 10 Moller-Trumbore tests (probe: 22 per extension ray on the
// teapot scene).  At upload each large leaf therefore gets a small private subtree of pair records over its own triangles.
// The reference's boxes and predicates are untouched: a subtree is only entered when the reference would have entered the
// leaf.  Subtree boxes are OURS, built from the true vertices and PADDED, so every triangle that Moller-Trumbore can accept
// is still reached and the winner (with the reference's rank for exact ties) is unchanged; the subtree only skips tests
// that were going to fail.
struct SubTri { int tri, rank; float lo[3], hi[3], c[3]; };

struct Refiner {
    const yune_triangle* tris; TravLayoutHost& out; int leaf_max; int depth_max = 0;
    Refiner(const yune_triangle* t, TravLayoutHost& o, int lm) : tris(t), out(o), leaf_max(lm) {}

    static void bounds(const std::vector<SubTri>& v, int b, int e, float* lo, float* hi)
    {
        for (int k = 0; k < 3; k++) { lo[k] = 3.0e38f; hi[k] = -3.0e38f; }
        for (int i = b; i < e; i++) for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], v[i].lo[k]); hi[k] = std::max(hi[k], v[i].hi[k]); }
    }
    static void padded(const float* lo, const float* hi, float* plo, float* phi)
    {
        float ext = 0.0f, mag = 0.0f;
        for (int k = 0; k < 3; k++) { ext = std::max(ext, hi[k] - lo[k]); mag = std::max(mag, std::max(std::fabs(lo[k]), std::fabs(hi[k]))); }
        const float pad = 2.0e-3f * ext + 4.0e-6f * mag + 1.0e-30f;
        for (int k = 0; k < 3; k++) { plo[k] = lo[k] - pad; phi[k] = hi[k] + pad; }
    }
    int emitLeaf(const std::vector<SubTri>& v, int b, int e)
    {
        const int first = (int)(out.tris.size() / 3);
        for (int i = b; i < e; i++) {
            const yune_triangle& t = tris[v[i].tri];
            V3 v1 = v3(t.v1.s[0], t.v1.s[1], t.v1.s[2]);
            V3 e1 = vsub(v3(t.v2.s[0], t.v2.s[1], t.v2.s[2]), v1);    // v1v2 (udpt.cl:328)
            V3 e2 = vsub(v3(t.v3.s[0], t.v3.s[1], t.v3.s[2]), v1);    // v1v3 (udpt.cl:329)
            out.tris.push_back({v1.x, v1.y, v1.z, bits(v[i].tri)});
            out.tris.push_back({e1.x, e1.y, e1.z, bits(v[i].rank)});
            out.tris.push_back({e2.x, e2.y, e2.z, 0.0f});
        }
        return ~((first << 4) | (e - b));
    }
    // returns the ref of the subtree over v[b, e)
    int build(std::vector<SubTri>& v, int b, int e, int depth)
    {
        if (depth > depth_max) depth_max = depth;
        if (e - b <= leaf_max) return emitLeaf(v, b, e);
        // split: longest axis of the centroid bounds, then the position with the least SA(L)*nL + SA(R)*nR
        float clo[3] = {3e38f, 3e38f, 3e38f}, chi[3] = {-3e38f, -3e38f, -3e38f};
        for (int i = b; i < e; i++) for (int k = 0; k < 3; k++) { clo[k] = std::min(clo[k], v[i].c[k]); chi[k] = std::max(chi[k], v[i].c[k]); }
        int axis = 0; for (int k = 1; k < 3; k++) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
        std::stable_sort(v.begin() + b, v.begin() + e, [axis](const SubTri& x, const SubTri& y) { return x.c[axis] < y.c[axis]; });
        int best = b + (e - b) / 2; float best_cost = 3e38f;
        for (int m = b + 1; m < e; m++) {
            float lo[3], hi[3]; float cost = 0.0f;
            bounds(v, b, m, lo, hi); { float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; cost += (dx * dy + dx * dz + dy * dz) * (m - b); }
            bounds(v, m, e, lo, hi); { float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; cost += (dx * dy + dx * dz + dy * dz) * (e - m); }
            if (cost < best_cost) { best_cost = cost; best = m; }
        }
        const int idx = (int)(out.pairs.size() / 4);
        out.pairs.resize(out.pairs.size() + 4);
        float lo0[3], hi0[3], lo1[3], hi1[3], p0l[3], p0h[3], p1l[3], p1h[3];
        bounds(v, b, best, lo0, hi0); padded(lo0, hi0, p0l, p0h);
        bounds(v, best, e, lo1, hi1); padded(lo1, hi1, p1l, p1h);
        const int r0 = build(v, b, best, depth + 1);
        const int r1 = build(v, best, e, depth + 1);
        F4* q = &out.pairs[(size_t)idx * 4];
        q[0] = {p0l[0], p0h[0], p0l[1], p0h[1]};
        q[1] = {p1l[0], p1h[0], p1l[1], p1h[1]};
        q[2] = {p0l[2], p0h[2], p1l[2], p1h[2]};
        q[3] = {bits(r0), bits(r1), 0.0f, 0.0f};
        return idx;
    }
};

// ---- accel 1: our own tree ---------------------------------------------------------------------------------------------
struct OwnTri { float lo[3], hi[3], c[3]; int tri, rank, leaf; };
struct OwnNode { float lo[3], hi[3]; int left, right, first, count; };   // left < 0: leaf [first, first + count)

struct OwnBuilder {
    std::vector<OwnTri>& t; std::vector<OwnNode> nodes; int depth_max = 0, leaf_max = 2;
    OwnBuilder(std::vector<OwnTri>& tt, int lm) : t(tt), leaf_max(lm) {}
    static float area(const float* lo, const float* hi) { float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dx * dz + dy * dz; }
    // `par_levels` > 0: the two subtrees of a big node are built by two sub-builders on two threads and their node arrays are
    // appended (indices shifted).  The TREE is the same as in a sequential build -- every split only looks at its own range of
    // `t` -- and the layout derived from it (breadth-first pair records, triangles in `t` order) does not depend on how the
    // builder numbered its nodes.
    int build(int b, int e, int depth, int par_levels = 0)
    {
        if (depth > depth_max) depth_max = depth;
        OwnNode nd; nd.left = nd.right = -1; nd.first = b; nd.count = e - b;
        float clo[3] = {3e38f, 3e38f, 3e38f}, chi[3] = {-3e38f, -3e38f, -3e38f};
        for (int k = 0; k < 3; k++) { nd.lo[k] = 3e38f; nd.hi[k] = -3e38f; }
        for (int i = b; i < e; i++) for (int k = 0; k < 3; k++) {
            nd.lo[k] = std::min(nd.lo[k], t[i].lo[k]); nd.hi[k] = std::max(nd.hi[k], t[i].hi[k]);
            clo[k] = std::min(clo[k], t[i].c[k]); chi[k] = std::max(chi[k], t[i].c[k]);
        }
        const int me = (int)nodes.size(); nodes.push_back(nd);
        if (e - b <= leaf_max) return me;
        int axis = 0; for (int k = 1; k < 3; k++) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
        int mid = (b + e) / 2;
        const float ext = chi[axis] - clo[axis];
        if (ext > 0.0f) {
            const int NB = 16;
            int cnt[NB] = {0}; float blo[NB][3], bhi[NB][3];
            for (int q = 0; q < NB; q++) for (int k = 0; k < 3; k++) { blo[q][k] = 3e38f; bhi[q][k] = -3e38f; }
            const float scale = NB * (1.0f - 1e-6f) / ext;
            auto bin_of = [&](const OwnTri& x) { int q = (int)((x.c[axis] - clo[axis]) * scale); return q < 0 ? 0 : (q >= NB ? NB - 1 : q); };
            for (int i = b; i < e; i++) { const int q = bin_of(t[i]); cnt[q]++; for (int k = 0; k < 3; k++) { blo[q][k] = std::min(blo[q][k], t[i].lo[k]); bhi[q][k] = std::max(bhi[q][k], t[i].hi[k]); } }
            float rl[NB][3], rh[NB][3]; int rc[NB];
            { float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f}; int c = 0;
              for (int q = NB - 1; q >= 0; q--) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], blo[q][k]); hi[k] = std::max(hi[k], bhi[q][k]); } c += cnt[q]; for (int k = 0; k < 3; k++) { rl[q][k] = lo[k]; rh[q][k] = hi[k]; } rc[q] = c; } }
            float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f}; int c = 0, best = -1; float best_cost = 3e38f;
            for (int q = 0; q < NB - 1; q++) {
                for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], blo[q][k]); hi[k] = std::max(hi[k], bhi[q][k]); }
                c += cnt[q];
                if (c == 0 || rc[q + 1] == 0) continue;
                const float cost = area(lo, hi) * c + area(rl[q + 1], rh[q + 1]) * rc[q + 1];
                if (cost < best_cost) { best_cost = cost; best = q; }
            }
            if (best >= 0) {
                mid = (int)(std::partition(t.begin() + b, t.begin() + e, [&](const OwnTri& x) { return bin_of(x) <= best; }) - t.begin());
                if (mid == b || mid == e) mid = (b + e) / 2;
            }
        }
        if (par_levels > 0 && e - b >= (1 << 16)) {
            OwnBuilder LB(t, leaf_max), RB(t, leaf_max);
            std::thread th([&]() { LB.build(b, mid, depth + 1, par_levels - 1); });
            RB.build(mid, e, depth + 1, par_levels - 1);
            th.join();
            for (OwnBuilder* sub : {&LB, &RB}) {
                const int off = (int)nodes.size();
                (sub == &LB ? nodes[me].left : nodes[me].right) = off;       // a sub-builder's root is its node 0
                for (OwnNode n : sub->nodes) { if (n.left >= 0) { n.left += off; n.right += off; } nodes.push_back(n); }
                depth_max = std::max(depth_max, sub->depth_max);
            }
            return me;
        }
        const int l = build(b, mid, depth + 1);
        const int r = build(mid, e, depth + 1);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
};

// ---- own-tree optimisation by reinsertion (Bittner, Hapala, Havran: "Fast insertion-based optimization of bounding volume
// hierarchies", CGF 2013; simplified) ---------------------------------------------------------------------------------------
// The walk's cost per ray is one node step per visited inner node, and a random ray visits a node with probability ~ its surface
// area.  A top-down binned build fixes its early splits before it has seen the detail below them; here every subtree is taken out
// of the finished tree in turn (largest boxes first) and put back where it adds the least surface area, found by a
// branch-and-bound search from the root (the place it came from is among the candidates, so a move never makes the sum worse).
// Works on single-triangle leaves; afterwards subtrees of <= leaf_max triangles become the leaves again.  The own tree's topology
// is free -- hit records are decided by the exact leaf-box filter -- so this changes speed only.
struct OptTree {
    struct N { float lo[3], hi[3]; int left, right, parent, tri; };       // tri >= 0: leaf holding t[tri]
    std::vector<N> n; int root = -1;
    static float area(const float* lo, const float* hi) { const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dx * dz + dy * dz; }
    float area(int i) const { return area(n[i].lo, n[i].hi); }
    float union_area(int a, int b) const
    {
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) { lo[k] = std::min(n[a].lo[k], n[b].lo[k]); hi[k] = std::max(n[a].hi[k], n[b].hi[k]); }
        return area(lo, hi);
    }
    void refit_up(int i)
    {
        for (; i >= 0; i = n[i].parent) {
            const N& a = n[n[i].left]; const N& b = n[n[i].right];
            for (int k = 0; k < 3; k++) { n[i].lo[k] = std::min(a.lo[k], b.lo[k]); n[i].hi[k] = std::max(a.hi[k], b.hi[k]); }
        }
    }
    int from_builder(const std::vector<OwnNode>& src, int s, const std::vector<OwnTri>& t, int parent)
    {
        const OwnNode& nd = src[s];
        if (nd.left >= 0) {
            const int me = (int)n.size(); n.push_back(N());
            n[me].parent = parent; n[me].tri = -1;
            const int l = from_builder(src, nd.left, t, me), r = from_builder(src, nd.right, t, me);
            n[me].left = l; n[me].right = r;
            for (int k = 0; k < 3; k++) { n[me].lo[k] = std::min(n[l].lo[k], n[r].lo[k]); n[me].hi[k] = std::max(n[l].hi[k], n[r].hi[k]); }
            return me;
        }
        return leaves(t, nd.first, nd.first + nd.count, parent);
    }
    int leaves(const std::vector<OwnTri>& t, int b, int e, int parent)      // a builder leaf of several triangles: a small balanced subtree
    {
        const int me = (int)n.size(); n.push_back(N());
        n[me].parent = parent;
        if (e - b == 1) {
            n[me].left = n[me].right = -1; n[me].tri = b;
            for (int k = 0; k < 3; k++) { n[me].lo[k] = t[b].lo[k]; n[me].hi[k] = t[b].hi[k]; }
            return me;
        }
        n[me].tri = -1;
        const int m = (b + e) / 2;
        const int l = leaves(t, b, m, me), r = leaves(t, m, e, me);
        n[me].left = l; n[me].right = r;
        for (int k = 0; k < 3; k++) { n[me].lo[k] = std::min(n[l].lo[k], n[r].lo[k]); n[me].hi[k] = std::max(n[l].hi[k], n[r].hi[k]); }
        return me;
    }
    // take subtree x out (its parent p disappears, the sibling moves up) and put it back at the cheapest place
    void reinsert(int x, std::vector<std::pair<float, int>>& heap)
    {
        const int p = n[x].parent;
        if (p < 0 || n[p].parent < 0) return;                  // the root's children stay (the root slot is never recycled)
        const int g = n[p].parent, sib = n[p].left == x ? n[p].right : n[p].left;
        (n[g].left == p ? n[g].left : n[g].right) = sib; n[sib].parent = g;
        refit_up(g);
        const float ax = area(x);
        float best = 3.0e38f; int best_at = sib;
        heap.clear(); heap.push_back({0.0f, root});
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
            const float induced = heap.back().first; const int y = heap.back().second; heap.pop_back();
            if (induced + ax >= best) break;                   // every remaining candidate costs at least this much
            const float direct = union_area(x, y), total = induced + direct;
            if (total < best) { best = total; best_at = y; }
            if (n[y].tri < 0) {
                const float down = total - area(y);            // what the ancestors of a place below y gain
                if (down + ax < best) {
                    for (int c : {n[y].left, n[y].right}) {
                        heap.push_back({down, c});
                        std::push_heap(heap.begin(), heap.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
                    }
                }
            }
        }
        // p becomes the new parent of (best_at, x) in best_at's place
        const int q = n[best_at].parent;
        if (q < 0) {                                            // above the root: p becomes the new root
            n[p].parent = -1; root = p;
        } else { (n[q].left == best_at ? n[q].left : n[q].right) = p; n[p].parent = q; }
        n[p].left = best_at; n[p].right = x; n[best_at].parent = p; n[x].parent = p;
        refit_up(p);
    }
    double cost() const { double c = 0; for (const N& a : n) if (a.tri < 0) c += area(a.lo, a.hi); return c / area(root); }
    void optimise(int passes)      // (see the loop body for what a pass covers)
    {
        std::vector<std::pair<float, int>> heap;
        std::vector<int> order(n.size());
        for (int pass = 0; pass < passes; pass++) {
            for (size_t i = 0; i < n.size(); i++) order[i] = (int)i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return area(a) > area(b); });
            // the largest tenth of the nodes per pass: measured on the teapot scene, box tests per ray 25.2 -> 23.6 / 23.4 / 23.35 after
            // 1 / 2 / 3 such passes against 23.4 / 23.15 after 1 / 2 passes over ALL nodes at ten times the cost
            size_t limit = order.size() / 10 + 1;
            if (const char* e = std::getenv("YUNE_OWN_OPT_FRAC")) limit = (size_t)(order.size() * std::atof(e));
            if (limit > order.size()) limit = order.size();
            for (size_t k = 0; k < limit; k++) { const int x = order[k]; if (x != root) reinsert(x, heap); }
        }
    }
    // back to the builder's form: leaves of <= leaf_max triangles, triangles in depth-first order
    int to_builder(int i, std::vector<OwnNode>& dst, const std::vector<OwnTri>& t, std::vector<OwnTri>& t_out, int leaf_max, int depth, int& depth_max, std::vector<int>& count)
    {
        if (depth > depth_max) depth_max = depth;
        const int me = (int)dst.size(); dst.push_back(OwnNode());
        for (int k = 0; k < 3; k++) { dst[me].lo[k] = n[i].lo[k]; dst[me].hi[k] = n[i].hi[k]; }
        dst[me].first = (int)t_out.size(); dst[me].count = count[i];
        if (count[i] <= leaf_max) {
            dst[me].left = dst[me].right = -1;
            std::vector<int> st{i};
            while (!st.empty()) { const int x = st.back(); st.pop_back(); if (n[x].tri >= 0) t_out.push_back(t[n[x].tri]); else { st.push_back(n[x].right); st.push_back(n[x].left); } }
            return me;
        }
        const int l = to_builder(n[i].left, dst, t, t_out, leaf_max, depth + 1, depth_max, count);
        const int r = to_builder(n[i].right, dst, t, t_out, leaf_max, depth + 1, depth_max, count);
        dst[me].left = l; dst[me].right = r;
        return me;
    }
    int count_tris(int i, std::vector<int>& count) { return count[i] = n[i].tri >= 0 ? 1 : count_tris(n[i].left, count) + count_tris(n[i].right, count); }
};

// Structural checks shared by both layouts: child indices in range and after their parent, leaves sane, every leaf linked from the
// root (accel 1 decides reachability from the leaf boxes alone, so an orphan leaf must be an error, not silently visible).
// Also reports
//   bfs_leaves  the linked leaves in the order the reference's breadth-first queue (udpt.cl:288-320) can reach them: a leaf's
//               position in this list is the visiting rank that breaks exact ties in t.  For an array in breadth-first order (what
//               the reference builder emits) this is node-index order; for a hand-made or deserialised array whose child indices
//               do not increase with the parent index it is not, and the ranks must follow the queue, not the index.
//   nested      every linked inner node's box contains the boxes of its non-empty children exactly.  The reference builder
//               guarantees it (getExtent / resizeBvh, src/BVH.cpp:218-278); accel 1 / 2 RELY on it ("the leaf's box passes" implies
//               "every ancestor's box passes"), so the caller walks the reference tree itself (accel 0) when it does not hold.
static bool validateReferenceTree(const yune_bvh_node* nodes, int n_nodes, int n_tris, std::string& err, std::vector<int>* bfs_leaves = nullptr, bool* nested = nullptr)
{
    std::vector<char> linked(n_nodes, 0);
    linked[0] = 1;
    bool nest = true;
    for (int i = 0; i < n_nodes; i++) {
        const yune_bvh_node& nd = nodes[i];
        const bool is_leaf = nd.child_idx == -1 && nd.vert_len > 0;
        const bool is_inner = !is_leaf && nd.child_idx > 0;
        if (is_leaf) {
            if (nd.vert_len > 10) { err = "leaf with more than 10 triangles"; return false; }
            for (int j = 0; j < nd.vert_len; j++)
                if (nd.vert_list[j] < 0 || nd.vert_list[j] >= n_tris) { err = "leaf references a triangle out of range"; return false; }
            if (!linked[i]) { err = "leaf is not linked from the root"; return false; }
        } else if (is_inner) {
            if (nd.child_idx + 1 >= n_nodes) { err = "child index out of range"; return false; }
            if (nd.child_idx <= i) { err = "BVH is not in breadth-first order (child index <= parent index)"; return false; }
            if (!linked[i]) continue;
            for (int c = nd.child_idx; c <= nd.child_idx + 1; c++) {
                if (linked[c]) { err = "node has two parents"; return false; }
                linked[c] = 1;
                const yune_bvh_node& ch = nodes[c];
                if (!(ch.vert_len > 0 || ch.child_idx > 0)) continue;          // the reference's "empty" node: never visited (udpt.cl:316)
                for (int k = 0; k < 3; k++)
                    if (!(ch.aabb.p_min.s[k] >= nd.aabb.p_min.s[k] && ch.aabb.p_max.s[k] <= nd.aabb.p_max.s[k])) nest = false;
            }
        }
    }
    if (nested) *nested = nest;
    if (bfs_leaves) {
        bfs_leaves->clear();
        std::vector<int> queue; queue.reserve(n_nodes);
        if (n_nodes > 0) queue.push_back(0);
        for (size_t h = 0; h < queue.size(); h++) {
            const yune_bvh_node& nd = nodes[queue[h]];
            if (nd.child_idx == -1 && nd.vert_len > 0) { bfs_leaves->push_back(queue[h]); continue; }
            if (nd.child_idx > 0) { queue.push_back(nd.child_idx); queue.push_back(nd.child_idx + 1); }
        }
    }
    return true;
}

static bool buildOwnLayout(const yune_triangle* tris, int n_tris, const yune_bvh_node* nodes, int n_nodes, TravLayoutHost& out, std::string& err, int leaf_max, int own_opt_passes)
{
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (std::getenv("YUNE_BVH_TIMING")) { auto T1 = std::chrono::steady_clock::now(); std::fprintf(stderr, "  relayout %s: %.2f s\n", what, std::chrono::duration<double>(T1 - T0).count()); T0 = T1; } };
    const bool brute = n_nodes == 0;       // no reference tree: the reference tests every triangle in index order (udpt.cl:280-284)
    std::vector<int> bfs_leaves;
    if (!brute && !validateReferenceTree(nodes, n_nodes, n_tris, err, &bfs_leaves)) return false;
    lap("validate");
    // reference leaves: id, box, visiting rank of every triangle slot
    std::vector<OwnTri> t; t.reserve(n_tris);
    int rank = 0, n_leaves = 0;
    auto padded_bounds = [&](OwnTri& x) {
        const yune_triangle& T = tris[x.tri];
        float ext = 0.0f, mag = 0.0f;
        for (int k = 0; k < 3; k++) {
            x.lo[k] = std::min(T.v1.s[k], std::min(T.v2.s[k], T.v3.s[k]));
            x.hi[k] = std::max(T.v1.s[k], std::max(T.v2.s[k], T.v3.s[k]));
            x.c[k] = 0.5f * (x.lo[k] + x.hi[k]);
            ext = std::max(ext, x.hi[k] - x.lo[k]); mag = std::max(mag, std::max(std::fabs(x.lo[k]), std::fabs(x.hi[k])));
        }
        const float pad = 2.0e-3f * ext + 4.0e-6f * mag + 1.0e-30f;      // same conservative margin as the leaf refinement
        for (int k = 0; k < 3; k++) { x.lo[k] -= pad; x.hi[k] += pad; }
    };
    if (brute) {
        // One pseudo-leaf whose box every ray passes (the filter 'would the reference have reached this triangle?' is always
        // yes), visiting rank = triangle index (exact ties in t go to the lower index, as in the reference's loop).
        const float big = 3.4028234e38f;
        out.leaf_boxes.push_back({-big, -big, -big, 0.0f});
        out.leaf_boxes.push_back({big, big, big, 0.0f});
        t.resize(n_tris);
        parallel_for((size_t)n_tris, [&](size_t i) { OwnTri& x = t[i]; x.tri = (int)i; x.rank = (int)i; x.leaf = 0; padded_bounds(x); });
        n_leaves = 1;
    }
    for (const int i : bfs_leaves) {
        const yune_bvh_node& nd = nodes[i];
        out.leaf_boxes.push_back({nd.aabb.p_min.s[0], nd.aabb.p_min.s[1], nd.aabb.p_min.s[2], 0.0f});
        out.leaf_boxes.push_back({nd.aabb.p_max.s[0], nd.aabb.p_max.s[1], nd.aabb.p_max.s[2], 0.0f});
        for (int j = 0; j < nd.vert_len; j++) {
            const int id = nd.vert_list[j];
            if (id < 0 || id >= n_tris) { err = "leaf references a triangle out of range"; return false; }
            OwnTri x; x.tri = id; x.rank = rank++; x.leaf = n_leaves;
            padded_bounds(x);
            t.push_back(x);
        }
        n_leaves++;
    }
    lap("gather leaves");
    if (t.size() >= ((size_t)1 << 27)) { err = "more than 2^27 triangle slots in the leaves (a leaf reference holds 27 bits)"; return false; }
    out.n_leaf_tris = (int)t.size();
    if (t.empty()) { out.root_ref = YUNE_REF_EMPTY; out.n_inner = out.n_inner_ref = 0; out.max_depth = 0; return true; }
    OwnBuilder B(t, leaf_max < 1 ? 2 : (leaf_max > 8 ? 8 : leaf_max));
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("YUNE_BVH_THREADS")) hw = (unsigned)std::max(1, std::atoi(e));      // 1 = sequential (tests compare both)
    int par_levels = 0;
    while (par_levels < 5 && (2u << par_levels) <= hw) par_levels++;
    int root = B.build(0, (int)t.size(), 0, par_levels);
    lap("own tree");
    // reinsertion passes over the finished tree (small scenes; big ones take the device builder, bvh_build.cu)
    int opt_passes = own_opt_passes >= 0 ? own_opt_passes : (t.size() <= ((size_t)1 << 18) ? 2 : 0);
    if (const char* e = std::getenv("YUNE_OWN_OPT")) opt_passes = std::atoi(e);
    if (opt_passes > 0 && B.nodes[root].left >= 0) {
        OptTree O; O.n.reserve(2 * t.size());
        O.root = O.from_builder(B.nodes, root, t, -1);
        const double c0 = O.cost();
        O.optimise(opt_passes);
        const double c1 = O.cost();
        std::vector<int> count(O.n.size(), 0); O.count_tris(O.root, count);
        std::vector<OwnNode> nodes2; std::vector<OwnTri> t2; nodes2.reserve(B.nodes.size()); t2.reserve(t.size());
        int depth2 = 0;
        const int root2 = O.to_builder(O.root, nodes2, t, t2, B.leaf_max, 0, depth2, count);
        if (std::getenv("YUNE_BVH_TIMING")) std::fprintf(stderr, "  own tree: inner-node area sum %.3f -> %.3f after %d reinsertion passes, depth %d -> %d\n", c0, c1, opt_passes, B.depth_max, depth2);
        if (depth2 + 2 <= YUNE_STACK_SIZE && t2.size() == t.size()) { B.nodes.swap(nodes2); t.swap(t2); root = root2; B.depth_max = depth2; }
        lap("reinsertion");
    }
    if (B.depth_max + 2 > YUNE_STACK_SIZE) { err = "BVH deeper than the traversal stack (YUNE_STACK_SIZE)"; return false; }
    out.max_depth = B.depth_max;
    for (int k = 0; k < 3; k++) { out.root_lo[k] = B.nodes[root].lo[k]; out.root_hi[k] = B.nodes[root].hi[k]; }
    // triangles in builder order (each leaf contiguous)
    out.tris.resize((size_t)t.size() * 3);
    parallel_for(t.size(), [&](size_t i) {
        const yune_triangle& T = tris[t[i].tri];
        V3 v1 = v3(T.v1.s[0], T.v1.s[1], T.v1.s[2]);
        V3 e1 = vsub(v3(T.v2.s[0], T.v2.s[1], T.v2.s[2]), v1);
        V3 e2 = vsub(v3(T.v3.s[0], T.v3.s[1], T.v3.s[2]), v1);
        if (out.isect == 1) { e1 = v3(T.v2.s[0], T.v2.s[1], T.v2.s[2]); e2 = v3(T.v3.s[0], T.v3.s[1], T.v3.s[2]); }   // raw vertices
        out.tris[3 * i + 0] = {v1.x, v1.y, v1.z, bits(t[i].tri)};
        out.tris[3 * i + 1] = {e1.x, e1.y, e1.z, bits(t[i].rank)};
        out.tris[3 * i + 2] = {e2.x, e2.y, e2.z, bits(t[i].leaf)};
    });
    lap("triangle records");
    // pair records in breadth-first order of the inner nodes
    std::vector<int> pair_of(B.nodes.size(), -1), order;
    auto ref_of = [&](int n) { const OwnNode& nd = B.nodes[n]; return nd.left < 0 ? ~((nd.first << 4) | nd.count) : pair_of[n]; };
    if (B.nodes[root].left >= 0) { order.push_back(root); pair_of[root] = 0; }
    for (size_t h = 0; h < order.size(); h++) {
        const OwnNode& nd = B.nodes[order[h]];
        for (int c : {nd.left, nd.right}) if (B.nodes[c].left >= 0) { pair_of[c] = (int)order.size(); order.push_back(c); }
    }
    out.pairs.resize(order.size() * 4);
    for (size_t h = 0; h < order.size(); h++) {
        const OwnNode& nd = B.nodes[order[h]];
        const OwnNode& a = B.nodes[nd.left]; const OwnNode& b = B.nodes[nd.right];
        F4* q = &out.pairs[h * 4];
        q[0] = {a.lo[0], a.hi[0], a.lo[1], a.hi[1]};
        q[1] = {b.lo[0], b.hi[0], b.lo[1], b.hi[1]};
        q[2] = {a.lo[2], a.hi[2], b.lo[2], b.hi[2]};
        q[3] = {bits(ref_of(nd.left)), bits(ref_of(nd.right)), 0.0f, 0.0f};
    }
    out.n_inner = out.n_inner_ref = (int)order.size();
    out.root_ref = ref_of(root);
    lap("pair records");
    if (out.accel != 2) return true;

    // ---- accel 2 (k_trace<., 5 / 7>; bit-exact on the B200, measured slower than accel 1): the same tree collapsed into nodes of up to 4 children ----
    // A wide node starts from the two children of a binary node; the inner child with the largest box is replaced by its own two
    // children until there are four (or only leaves are left).  Breadth-first order again, so the top of the tree is a prefix.
    std::vector<int> wide_of(B.nodes.size(), -1), worder;
    auto wref_of = [&](int n) { const OwnNode& nd = B.nodes[n]; return nd.left < 0 ? ~((nd.first << 4) | nd.count) : wide_of[n]; };
    std::vector<int> wdepth;
    if (B.nodes[root].left >= 0) { worder.push_back(root); wide_of[root] = 0; wdepth.push_back(1); }
    std::vector<int> kids;
    int wide_depth = 0;
    for (size_t h = 0; h < worder.size(); h++) {
        const OwnNode& nd = B.nodes[worder[h]];
        kids.clear(); kids.push_back(nd.left); kids.push_back(nd.right);
        while (kids.size() < 4) {
            int pick = -1; float best_area = -1.0f;
            for (size_t i = 0; i < kids.size(); i++) {
                const OwnNode& c = B.nodes[kids[i]];
                if (c.left < 0) continue;
                const float a = OwnBuilder::area(c.lo, c.hi);
                if (a > best_area) { best_area = a; pick = (int)i; }
            }
            if (pick < 0) break;
            const OwnNode c = B.nodes[kids[pick]];
            kids[pick] = c.left; kids.push_back(c.right);
        }
        for (int c : kids) if (B.nodes[c].left >= 0) { wide_of[c] = (int)worder.size(); worder.push_back(c); wdepth.push_back(wdepth[h] + 1); }
        wide_depth = std::max(wide_depth, wdepth[h]);
        F4 q[7];
        float* f = &q[0].x;                                    // 24 floats: lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4]
        int refs[4];
        for (int i = 0; i < 4; i++) {
            const bool used = i < (int)kids.size();
            for (int a = 0; a < 3; a++) {
                f[8 * a + i]     = used ? B.nodes[kids[i]].lo[a] :  3.0e38f;      // an unused slot is an inverted box: never hit
                f[8 * a + 4 + i] = used ? B.nodes[kids[i]].hi[a] : -3.0e38f;
            }
            refs[i] = used ? wref_of(kids[i]) : YUNE_REF_EMPTY;
        }
        q[6] = {bits(refs[0]), bits(refs[1]), bits(refs[2]), bits(refs[3])};
        out.quads.insert(out.quads.end(), q, q + 7);
    }
    out.n_wide = (int)worder.size();
    out.root_wide_ref = wref_of(root);
    out.wide_depth = wide_depth;
    if (3 * wide_depth + 2 > YUNE_STACK_SIZE) { err = "wide tree deeper than the traversal stack (YUNE_STACK_SIZE)"; return false; }
    lap("wide records");
    return true;
}

} // namespace

// What the DEVICE layout path (bvh_build.cu) needs from an uploaded reference tree: per triangle the id of its reference leaf and
// its visiting rank, and the leaves' uploaded boxes.  `usable` = false when the own-tree argument does not hold for this array
// (boxes that do not nest, a triangle listed by several leaves or by none): the caller then takes the host path.
bool referenceLeavesForDevice(const yune_triangle* tris, int n_tris, const yune_bvh_node* nodes, int n_nodes,
                              std::vector<int>& leaf_of_tri, std::vector<int>& rank_of_tri, std::vector<F4>& leaf_boxes, bool& usable, std::string& err)
{
    usable = false;
    if (n_nodes <= 0 || n_tris <= 0 || !nodes || !tris) return true;
    std::vector<int> bfs_leaves; bool nested = true;
    if (!validateReferenceTree(nodes, n_nodes, n_tris, err, &bfs_leaves, &nested)) return false;
    if (!nested) return true;
    leaf_of_tri.assign((size_t)n_tris, -1); rank_of_tri.assign((size_t)n_tris, 0);
    leaf_boxes.clear(); leaf_boxes.reserve(bfs_leaves.size() * 2);
    int rank = 0, leaf = 0;
    for (const int i : bfs_leaves) {
        const yune_bvh_node& nd = nodes[i];
        leaf_boxes.push_back({nd.aabb.p_min.s[0], nd.aabb.p_min.s[1], nd.aabb.p_min.s[2], 0.0f});
        leaf_boxes.push_back({nd.aabb.p_max.s[0], nd.aabb.p_max.s[1], nd.aabb.p_max.s[2], 0.0f});
        for (int j = 0; j < nd.vert_len; j++) {
            const int id = nd.vert_list[j];
            if (leaf_of_tri[id] >= 0) return true;                  // listed twice: slots != triangles, host path
            leaf_of_tri[id] = leaf; rank_of_tri[id] = rank++;
        }
        leaf++;
    }
    for (int i = 0; i < n_tris; i++) if (leaf_of_tri[i] < 0) return true;      // a triangle the reference never reaches: host path keeps it out of the tree
    usable = true;
    return true;
}

bool buildTravLayout(const yune_triangle* tris, int n_tris, const yune_bvh_node* nodes, int n_nodes,
                     TravLayoutHost& out, std::string& err, int leaf_split, int accel, int isect, int own_opt_passes)
{
    out = TravLayoutHost();
    if (isect != 0 && isect != 1) { err = "isect must be 0 (reference Moller-Trumbore) or 1 (watertight)"; return false; }
    if (n_nodes < 0 || (n_nodes > 0 && !nodes)) { err = "bad BVH buffer"; return false; }
    if (n_tris < 0 || (n_tris > 0 && !tris)) { err = "bad triangle buffer"; return false; }
    if (n_tris >= (1 << 27)) { err = "more than 2^27 triangles"; return false; }
    // bvh_size == 0 is the reference's brute-force mode (udpt.cl:280-284: every triangle, in index order, no box tests).  The
    // result is reproduced by the own-tree walk with a filter that always passes; only the cost differs (a tree walk here).
    if (n_nodes == 0 && accel == 0) accel = 1;
    if (accel < 0 || accel > 2) { err = "accel must be 0, 1 or 2"; return false; }
    out.n_tris = n_tris; out.accel = accel;

    std::vector<int> bfs_leaves; bool nested = true;
    if (n_nodes > 0 && !validateReferenceTree(nodes, n_nodes, n_tris, err, &bfs_leaves, &nested)) return false;
    // accel 1 / 2 answer "would the reference have reached this triangle?" from the leaf box alone, which is only right when every
    // box contains its children's boxes.  A tree that was not built by the reference builder (hand-made, refitted, deserialised)
    // may not nest: walk that tree itself, under the reference's own predicates.
    if (accel >= 1 && n_nodes > 0 && !nested) { accel = 0; out.accel = 0; }
    if (isect == 1 && accel != 1) { err = "isect 1 (watertight) walks the own tree: it needs accel 1 and a tree whose boxes nest"; return false; }
    out.isect = isect;
    if (accel >= 1) {
        if (!buildOwnLayout(tris, n_tris, nodes, n_nodes, out, err, leaf_split, own_opt_passes)) return false;
        goto shade_records;
    }
    {
    // classify nodes; inner nodes get their pair index in node-index (= breadth-first) order
    std::vector<int> ref(n_nodes, YUNE_REF_EMPTY);
    std::vector<char> kind(n_nodes, 0);      // 0 empty, 1 inner, 2 leaf
    int n_inner = 0; long long n_slots = 0;
    for (int i = 0; i < n_nodes; i++) {
        const yune_bvh_node& nd = nodes[i];
        const bool is_leaf = nd.child_idx == -1 && nd.vert_len > 0;      // udpt.cl:301
        const bool is_inner = !is_leaf && nd.child_idx > 0;
        if (is_leaf) {
            if (nd.vert_len > 10) { err = "leaf with more than 10 triangles"; return false; }
            for (int j = 0; j < nd.vert_len; j++)
                if (nd.vert_list[j] < 0 || nd.vert_list[j] >= n_tris) { err = "leaf references a triangle out of range"; return false; }
            kind[i] = 2; n_slots += nd.vert_len;
        } else if (is_inner) {
            if (nd.child_idx + 1 >= n_nodes) { err = "child index out of range"; return false; }
            if (nd.child_idx <= i) { err = "BVH is not in breadth-first order (child index <= parent index)"; return false; }
            kind[i] = 1; ref[i] = n_inner++;
        }
        // anything else is the reference's "empty" node: vert_len <= 0 and child_idx <= 0 -> never visited (udpt.cl:316)
    }
    if (n_slots >= (1ll << 27)) { err = "more than 2^27 triangle slots in the leaves (a leaf reference holds 27 bits)"; return false; }
    out.n_inner_ref = n_inner; out.n_leaf_tris = (int)n_slots;
    for (int k = 0; k < 3; k++) { out.root_lo[k] = nodes[0].aabb.p_min.s[k]; out.root_hi[k] = nodes[0].aabb.p_max.s[k]; }

    // leaves, in the order of the reference's breadth-first queue (= node-index order for a reference-built array): rank = the
    // reference's visiting order of the triangle slots
    out.pairs.resize((size_t)n_inner * 4);
    out.tris.reserve((size_t)n_slots * 3);
    Refiner refiner(tris, out, leaf_split > 0 ? leaf_split : 16);
    int rank = 0, sub_depth = 0;
    std::vector<SubTri> work;
    for (const int i : bfs_leaves) {
        const yune_bvh_node& nd = nodes[i];
        work.clear();
        for (int j = 0; j < nd.vert_len; j++) {
            SubTri st; st.tri = nd.vert_list[j]; st.rank = rank++;
            const yune_triangle& t = tris[st.tri];
            for (int k = 0; k < 3; k++) {
                st.lo[k] = std::min(t.v1.s[k], std::min(t.v2.s[k], t.v3.s[k]));
                st.hi[k] = std::max(t.v1.s[k], std::max(t.v2.s[k], t.v3.s[k]));
                st.c[k] = 0.5f * (st.lo[k] + st.hi[k]);
            }
            work.push_back(st);
        }
        refiner.depth_max = 0;
        ref[i] = refiner.build(work, 0, (int)work.size(), 0);
        if (refiner.depth_max > sub_depth) sub_depth = refiner.depth_max;
    }
    out.root_ref = ref[0];
    out.n_inner = (int)(out.pairs.size() / 4);

    // pair records of the reference's inner nodes
    std::vector<int> depth(n_nodes, 0);      // nodes are breadth-first, so parents precede children
    int max_depth = 0;
    for (int i = 0; i < n_nodes; i++) {
        if (kind[i] != 1) continue;
        const yune_bvh_node& nd = nodes[i];
        const yune_bvh_node& a = nodes[nd.child_idx]; const yune_bvh_node& b = nodes[nd.child_idx + 1];
        F4* q = &out.pairs[(size_t)ref[i] * 4];
        q[0] = {a.aabb.p_min.s[0], a.aabb.p_max.s[0], a.aabb.p_min.s[1], a.aabb.p_max.s[1]};
        q[1] = {b.aabb.p_min.s[0], b.aabb.p_max.s[0], b.aabb.p_min.s[1], b.aabb.p_max.s[1]};
        q[2] = {a.aabb.p_min.s[2], a.aabb.p_max.s[2], b.aabb.p_min.s[2], b.aabb.p_max.s[2]};
        q[3] = {bits(ref[nd.child_idx]), bits(ref[nd.child_idx + 1]), 0.0f, 0.0f};
        depth[nd.child_idx] = depth[nd.child_idx + 1] = depth[i] + 1;
        if (depth[i] + 1 > max_depth) max_depth = depth[i] + 1;
    }
    out.max_depth = max_depth + sub_depth;
    if (out.max_depth + 2 > YUNE_STACK_SIZE) { err = "BVH deeper than the traversal stack (YUNE_STACK_SIZE)"; return false; }

    }
shade_records:
    out.shade.resize((size_t)n_tris * 4);
    parallel_for((size_t)n_tris, [&](size_t i) {
        const yune_triangle& t = tris[i];
        V3 n1 = vnormalize(v3(t.vn1.s[0], t.vn1.s[1], t.vn1.s[2]));       // udpt.cl:377-379
        V3 n2 = vnormalize(v3(t.vn2.s[0], t.vn2.s[1], t.vn2.s[2]));
        V3 n3 = vnormalize(v3(t.vn3.s[0], t.vn3.s[1], t.vn3.s[2]));
        F4* s = &out.shade[(size_t)i * 4];
        s[0] = {n1.x, n1.y, n1.z, bits(t.matID)};
        s[1] = {n2.x, n2.y, n2.z, 0.0f};
        s[2] = {n3.x, n3.y, n3.z, 0.0f};
        s[3] = {0.0f, 0.0f, 0.0f, 0.0f};
    });
    return true;
}

} // namespace yune
