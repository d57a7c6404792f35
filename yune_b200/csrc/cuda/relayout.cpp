// relayout.cpp -- host-side, once-per-upload conversion of the reference AoS buffers into the traversal
// layout described in trav_layout.h.  Arithmetic here (edge vectors, normalised normals) uses the same
// IEEE float32 expressions the reference evaluates per intersection (udpt.cl:328-329, 377-379), so moving
// them to upload time does not change a single bit of any hit record.
#include "trav_layout.h"
#include "strict_math.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace yune {

static inline float bits(int i) { float f; std::memcpy(&f, &i, 4); return f; }

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

} // namespace

bool buildTravLayout(const yune_triangle* tris, int n_tris, const yune_bvh_node* nodes, int n_nodes,
                     TravLayoutHost& out, std::string& err, int leaf_split)
{
    out = TravLayoutHost();
    if (n_nodes <= 0 || !nodes) { err = "no BVH nodes (brute-force mode, bvh_size == 0, is not supported)"; return false; }
    if (n_tris < 0 || (n_tris > 0 && !tris)) { err = "bad triangle buffer"; return false; }
    if (n_tris >= (1 << 27)) { err = "more than 2^27 triangles"; return false; }
    out.n_tris = n_tris;

    // classify nodes; inner nodes get their pair index in node-index (= breadth-first) order
    std::vector<int> ref(n_nodes, YUNE_REF_EMPTY);
    std::vector<char> kind(n_nodes, 0);      // 0 empty, 1 inner, 2 leaf
    int n_inner = 0, n_slots = 0;
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
    out.n_inner_ref = n_inner; out.n_leaf_tris = n_slots;
    for (int k = 0; k < 3; k++) { out.root_lo[k] = nodes[0].aabb.p_min.s[k]; out.root_hi[k] = nodes[0].aabb.p_max.s[k]; }

    // leaves, in node-index order: rank = the reference's breadth-first visiting order of the triangle slots
    out.pairs.resize((size_t)n_inner * 4);
    out.tris.reserve((size_t)n_slots * 3);
    Refiner refiner(tris, out, leaf_split > 0 ? leaf_split : 16);
    int rank = 0, sub_depth = 0;
    std::vector<SubTri> work;
    for (int i = 0; i < n_nodes; i++) {
        if (kind[i] != 2) continue;
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

    out.shade.resize((size_t)n_tris * 4);
    for (int i = 0; i < n_tris; i++) {
        const yune_triangle& t = tris[i];
        V3 n1 = vnormalize(v3(t.vn1.s[0], t.vn1.s[1], t.vn1.s[2]));       // udpt.cl:377-379
        V3 n2 = vnormalize(v3(t.vn2.s[0], t.vn2.s[1], t.vn2.s[2]));
        V3 n3 = vnormalize(v3(t.vn3.s[0], t.vn3.s[1], t.vn3.s[2]));
        F4* s = &out.shade[(size_t)i * 4];
        s[0] = {n1.x, n1.y, n1.z, bits(t.matID)};
        s[1] = {n2.x, n2.y, n2.z, 0.0f};
        s[2] = {n3.x, n3.y, n3.z, 0.0f};
        s[3] = {0.0f, 0.0f, 0.0f, 0.0f};
    }
    return true;
}

} // namespace yune
