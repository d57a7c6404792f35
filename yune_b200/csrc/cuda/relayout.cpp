// relayout.cpp -- host-side, once-per-upload conversion of the reference AoS buffers into the traversal
// layout described in trav_layout.h.  Arithmetic here (edge vectors, normalised normals) uses the same
// IEEE float32 expressions the reference evaluates per intersection (udpt.cl:328-329, 377-379), so moving
// them to upload time does not change a single bit of any hit record.
#include "trav_layout.h"
#include "strict_math.h"

#include <cstring>

namespace yune {

static inline float bits(int i) { float f; std::memcpy(&f, &i, 4); return f; }

bool buildTravLayout(const yune_triangle* tris, int n_tris, const yune_bvh_node* nodes, int n_nodes,
                     TravLayoutHost& out, std::string& err)
{
    out = TravLayoutHost();
    if (n_nodes <= 0 || !nodes) { err = "no BVH nodes (brute-force mode, bvh_size == 0, is not supported)"; return false; }
    if (n_tris < 0 || (n_tris > 0 && !tris)) { err = "bad triangle buffer"; return false; }
    if (n_tris >= (1 << 27)) { err = "more than 2^27 triangles"; return false; }
    out.n_tris = n_tris;

    // classify nodes; give inner nodes their pair index and leaves their first slot, both in node-index order
    std::vector<int> ref(n_nodes, YUNE_REF_EMPTY);
    int n_inner = 0, n_slots = 0;
    for (int i = 0; i < n_nodes; i++) {
        const yune_bvh_node& nd = nodes[i];
        const bool is_leaf = nd.child_idx == -1 && nd.vert_len > 0;      // udpt.cl:301
        const bool is_inner = !is_leaf && nd.child_idx > 0;
        if (is_leaf) {
            if (nd.vert_len > 10) { err = "leaf with more than 10 triangles"; return false; }
            for (int j = 0; j < nd.vert_len; j++)
                if (nd.vert_list[j] < 0 || nd.vert_list[j] >= n_tris) { err = "leaf references a triangle out of range"; return false; }
            ref[i] = ~((n_slots << 4) | nd.vert_len);
            n_slots += nd.vert_len;
        } else if (is_inner) {
            if (nd.child_idx + 1 >= n_nodes) { err = "child index out of range"; return false; }
            if (nd.child_idx <= i) { err = "BVH is not in breadth-first order (child index <= parent index)"; return false; }
            ref[i] = n_inner++;
        }
        // anything else is the reference's "empty" node: vert_len <= 0 and child_idx <= 0 -> never visited (udpt.cl:316)
    }
    out.n_inner = n_inner; out.n_leaf_tris = n_slots; out.root_ref = ref[0];
    for (int k = 0; k < 3; k++) { out.root_lo[k] = nodes[0].aabb.p_min.s[k]; out.root_hi[k] = nodes[0].aabb.p_max.s[k]; }

    out.pairs.resize((size_t)n_inner * 4);
    out.tris.resize((size_t)n_slots * 3);
    std::vector<int> depth(n_nodes, 0);      // nodes are breadth-first, so parents precede children
    int max_depth = 0;
    for (int i = 0; i < n_nodes; i++) {
        const yune_bvh_node& nd = nodes[i];
        if (ref[i] == YUNE_REF_EMPTY) continue;
        if (ref[i] >= 0) {
            const yune_bvh_node& a = nodes[nd.child_idx]; const yune_bvh_node& b = nodes[nd.child_idx + 1];
            F4* q = &out.pairs[(size_t)ref[i] * 4];
            q[0] = {a.aabb.p_min.s[0], a.aabb.p_max.s[0], a.aabb.p_min.s[1], a.aabb.p_max.s[1]};
            q[1] = {b.aabb.p_min.s[0], b.aabb.p_max.s[0], b.aabb.p_min.s[1], b.aabb.p_max.s[1]};
            q[2] = {a.aabb.p_min.s[2], a.aabb.p_max.s[2], b.aabb.p_min.s[2], b.aabb.p_max.s[2]};
            q[3] = {bits(ref[nd.child_idx]), bits(ref[nd.child_idx + 1]), 0.0f, 0.0f};
            depth[nd.child_idx] = depth[nd.child_idx + 1] = depth[i] + 1;
            if (depth[i] + 1 > max_depth) max_depth = depth[i] + 1;
        } else {
            const int first = (~ref[i]) >> 4;
            for (int j = 0; j < nd.vert_len; j++) {
                const yune_triangle& t = tris[nd.vert_list[j]];
                F4* r = &out.tris[(size_t)(first + j) * 3];
                V3 v1 = v3(t.v1.s[0], t.v1.s[1], t.v1.s[2]);
                V3 e1 = vsub(v3(t.v2.s[0], t.v2.s[1], t.v2.s[2]), v1);    // v1v2 (udpt.cl:328)
                V3 e2 = vsub(v3(t.v3.s[0], t.v3.s[1], t.v3.s[2]), v1);    // v1v3 (udpt.cl:329)
                r[0] = {v1.x, v1.y, v1.z, bits(nd.vert_list[j])};
                r[1] = {e1.x, e1.y, e1.z, 0.0f};
                r[2] = {e2.x, e2.y, e2.z, 0.0f};
            }
        }
    }
    out.max_depth = max_depth;
    if (max_depth + 2 > YUNE_STACK_SIZE) { err = "BVH deeper than the traversal stack (YUNE_STACK_SIZE)"; return false; }

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
