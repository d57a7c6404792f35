// tests/hostcheck/hostcheck.cpp -- TEST-ONLY host compile of the product's host/device headers
// (trace_core.h, lights.h, relayout.cpp).  It lets the CPU test tier (-m "not gpu") check the traversal LOGIC
// against the oracle without a GPU.  It is never loaded by the yune_b200 package and is not a fallback: the
// product's only execution path is the CUDA library.
#include "trace_core.h"
#include "lights.h"
#include <cmath>
#include <cstring>

using namespace yune;

namespace {
struct HostPairFetch { const F4* p; void operator()(int i, F4& a, F4& b, F4& c, F4& d) const { const F4* q = p + (size_t)i * 4; a = q[0]; b = q[1]; c = q[2]; d = q[3]; } };
struct HostTriFetch { const F4* p; void operator()(int i, F4& a, F4& b, F4& c) const { const F4* q = p + (size_t)i * 3; a = q[0]; b = q[1]; c = q[2]; } };
struct HostLeafFetch { const F4* p; void operator()(int i, F4& lo, F4& hi) const { lo = p[2 * (size_t)i]; hi = p[2 * (size_t)i + 1]; } };
LightDev unpack(const yune_quad_light& q)
{
    LightDev L;
    L.pos = v3(q.pos.s[0], q.pos.s[1], q.pos.s[2]); L.normal = v3(q.normal.s[0], q.normal.s[1], q.normal.s[2]);
    L.ke = v3(q.ke.s[0], q.ke.s[1], q.ke.s[2]);
    L.edge_l = v3(q.edge_l.s[0], q.edge_l.s[1], q.edge_l.s[2]); L.edge_w = v3(q.edge_w.s[0], q.edge_w.s[1], q.edge_w.s[2]);
    // length(float4) with w = 0
    L.la = sqrtf(L.edge_l.x * L.edge_l.x + L.edge_l.y * L.edge_l.y + L.edge_l.z * L.edge_l.z);
    L.lb = sqrtf(L.edge_w.x * L.edge_w.x + L.edge_w.y * L.edge_w.y + L.edge_w.z * L.edge_w.z);
    return L;
}
}

extern "C" int hc_trace(int n, const float* od6, const float* tmax, int any, const yune_triangle* tris, int ntri,
                        const yune_bvh_node* nodes, int nnodes, const yune_quad_light* lights, int nlights,
                        int* tri_id, int* light_id, float* t_hit, unsigned long long* work2, int leaf_split, int accel)
{
    TravLayoutHost lay; std::string err;
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, leaf_split, accel)) return -1;
    HostLeafFetch lf{lay.leaf_boxes.data()};
    HostPairFetch pf{lay.pairs.data()}; HostTriFetch tf{lay.tris.data()};
    LightDev L[YUNE_MAX_LIGHTS];
    for (int i = 0; i < nlights; i++) L[i] = unpack(lights[i]);
    unsigned long long nb = 0, nt = 0;
    for (int i = 0; i < n; i++) {
        const float* r = od6 + 6 * (size_t)i;
        V3 o = v3(r[0], r[1], r[2]), d = v3(r[3], r[4], r[5]);
        float t = tmax ? tmax[i] : INFINITY;
        int lid = light_loop(L, nlights, o, d, t);
        WorkCount wc = {0, 0};
        if (any) {
            bool occ = lid >= 0;
            if (!occ) {
                if (accel == 1) { HitRec h; trace_own<HostPairFetch, HostTriFetch, HostLeafFetch, true, true>(pf, tf, lf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc); occ = h.tri >= 0; }
                else occ = any_hit<HostPairFetch, HostTriFetch, true>(pf, tf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, &wc);
            }
            tri_id[i] = occ ? 0 : -1; light_id[i] = lid; if (t_hit) t_hit[i] = t;
        } else {
            HitRec h;
            if (accel == 1) trace_own<HostPairFetch, HostTriFetch, HostLeafFetch, false, true>(pf, tf, lf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc);
            else closest_hit<HostPairFetch, HostTriFetch, true>(pf, tf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc);
            tri_id[i] = h.tri; light_id[i] = h.tri >= 0 ? -1 : lid; if (t_hit) t_hit[i] = h.t;
        }
        nb += wc.box; nt += wc.tri;
    }
    if (work2) { work2[0] = nb; work2[1] = nt; }
    return 0;
}

// layout statistics (development aid / test): out[0..3] = n_inner, n_leaf_tris, max_depth, n_inner_ref
extern "C" int hc_layout_info(const yune_triangle* tris, int ntri, const yune_bvh_node* nodes, int nnodes, int leaf_split, int accel, int* out)
{
    TravLayoutHost lay; std::string err;
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, leaf_split, accel)) return -1;
    out[0] = lay.n_inner; out[1] = lay.n_leaf_tris; out[2] = lay.max_depth; out[3] = lay.n_inner_ref;
    return 0;
}
