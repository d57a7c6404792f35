// tests/hostcheck/hostcheck.cpp -- TEST-ONLY host compile of the product's host/device headers
// (trace_core.h, lights.h, relayout.cpp).  It lets the CPU test tier (-m "not gpu") check the traversal LOGIC
// against the oracle without a GPU.  It is never loaded by the yune_b200 package and is not a fallback: the
// product's only execution path is the CUDA library.
#include "trace_core.h"
#include "lights.h"
#include <cmath>
#include <cstring>
#include <cstdlib>

using namespace yune;

namespace {
struct HostPairFetch { const F4* p; void operator()(int i, F4& a, F4& b, F4& c, F4& d) const { const F4* q = p + (size_t)i * 4; a = q[0]; b = q[1]; c = q[2]; d = q[3]; } };
struct HostTriFetch { const F4* p; void operator()(int i, F4& a, F4& b, F4& c) const { const F4* q = p + (size_t)i * 3; a = q[0]; b = q[1]; c = q[2]; } };
struct HostQuadFetch { const F4* p; void operator()(int i, F4* q) const { const F4* s = p + (size_t)i * 7; for (int k = 0; k < 7; k++) q[k] = s[k]; } };
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
    const int isect = (accel >> 8) & 1;              // bit 8 of `accel`: perf-mode (watertight) intersection, own tree only
    accel &= 0xff;
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, leaf_split, accel, isect)) return -1;
    accel = lay.accel;                               // bvh_size == 0 (brute-force mode) always walks the own tree
    HostLeafFetch lf{lay.leaf_boxes.data()};
    HostPairFetch pf{lay.pairs.data()}; HostTriFetch tf{lay.tris.data()}; HostQuadFetch qf{lay.quads.data()};
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
                if (accel == 2) { HitRec h; trace_wide<HostQuadFetch, HostTriFetch, HostLeafFetch, true, true>(qf, tf, lf, lay.root_wide_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc); occ = h.tri >= 0; }
                else if (accel == 1 && isect) { HitRec h; trace_own<HostPairFetch, HostTriFetch, HostLeafFetch, true, true, true>(pf, tf, lf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc); occ = h.tri >= 0; }
                else if (accel == 1) { HitRec h; trace_own<HostPairFetch, HostTriFetch, HostLeafFetch, true, true>(pf, tf, lf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc); occ = h.tri >= 0; }
                else occ = any_hit<HostPairFetch, HostTriFetch, true>(pf, tf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, &wc);
            }
            tri_id[i] = occ ? 0 : -1; light_id[i] = lid; if (t_hit) t_hit[i] = t;
        } else {
            HitRec h;
            if (accel == 2) trace_wide<HostQuadFetch, HostTriFetch, HostLeafFetch, false, true>(qf, tf, lf, lay.root_wide_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc);
            else if (accel == 1 && isect) trace_own<HostPairFetch, HostTriFetch, HostLeafFetch, false, true, true>(pf, tf, lf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc);
            else if (accel == 1) trace_own<HostPairFetch, HostTriFetch, HostLeafFetch, false, true>(pf, tf, lf, lay.root_ref, lay.root_lo, lay.root_hi, o, d, t, h, &wc);
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
    out[0] = lay.n_inner; out[1] = lay.n_leaf_tris; out[2] = accel == 2 ? lay.wide_depth : lay.max_depth; out[3] = accel == 2 ? lay.n_wide : lay.n_inner_ref;
    return 0;
}

// FNV-1a hashes of the four layout arrays (pairs, tris, shade, leaf_boxes): lets a test compare two builds of the same scene
extern "C" int hc_layout_hash(const yune_triangle* tris, int ntri, const yune_bvh_node* nodes, int nnodes, int leaf_split, int accel, unsigned long long* out)
{
    TravLayoutHost lay; std::string err;
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, leaf_split, accel)) return -1;
    auto fnv = [](const std::vector<F4>& v) {
        unsigned long long h = 1469598103934665603ull;
        const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data());
        for (size_t i = 0; i < v.size() * sizeof(F4); i++) { h ^= p[i]; h *= 1099511628211ull; }
        return h;
    };
    out[0] = fnv(lay.pairs); out[1] = fnv(lay.tris); out[2] = fnv(lay.shade); out[3] = fnv(lay.leaf_boxes);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Host restatement of the DEVICE scheduling of k_trace (kernels.cu: postponed leaves, sentinel stack with the top entries
// read up front, per-step vote between a node step and a triangle step, chained node steps, refill of idle lanes), driven for
// simulated 32-lane warps.  The arithmetic is trace_core.h's; what this checks without a GPU is that WHEN a triangle is
// tested relative to the rest of the walk -- which depends on the knobs and on which rays share a warp -- cannot change a
// hit record.  It also reports the lane utilisation of the two step kinds (development aid for the knobs).
// ------------------------------------------------------------------------------------------------------------------
namespace {
const int REF_DONE = -1, STACK_BASE = 2;
struct SimLane {
    RayPre r; float t_best, t_prune, u, v;
    int tri, best_pos, cur, pend_pos, pend_end, sp, where; bool have;
    int stack[YUNE_STACK_SIZE + STACK_BASE];
};
struct SimScene { const TravLayoutHost* lay; HostPairFetch pf; HostTriFetch tf; HostLeafFetch lf; int accel; HostQuadFetch qf; };

void sim_init(SimLane& L, const SimScene& S, V3 o, V3 d, float t_in)
{
    L.r = make_ray(o, d);
    L.t_best = t_in; L.t_prune = t_in * 1.00001f;
    L.u = L.v = 0.0f; L.tri = -1; L.best_pos = -1; L.sp = STACK_BASE;
    L.stack[0] = L.stack[1] = REF_DONE;
    bool hit = S.lay->root_ref != YUNE_REF_EMPTY;
    if (S.accel == 0 && hit) { float e; hit = box_hit(L.r, S.lay->root_lo[0], S.lay->root_hi[0], S.lay->root_lo[1], S.lay->root_hi[1], S.lay->root_lo[2], S.lay->root_hi[2], e); }
    const int ref = hit ? (S.accel == 2 ? S.lay->root_wide_ref : S.lay->root_ref) : REF_DONE;
    const int x = ~ref;
    L.cur = ref >= 0 ? ref : REF_DONE;
    L.pend_pos = ref >= 0 ? 0 : (x >> 4);
    L.pend_end = ref >= 0 ? 0 : (x >> 4) + (x & 15);
}
void sim_inner(SimLane& L, const SimScene& S, bool any_q)
{
    const int top1 = L.stack[L.sp - 1], top2 = L.stack[L.sp - 2];
    F4 q0, q1, q2, q3; S.pf(L.cur, q0, q1, q2, q3);
    const int ref0 = YF_ASINT(q3.x), ref1 = YF_ASINT(q3.y);
    float e0, e1; bool h0, h1;
    if (S.accel == 1) {
        h0 = box_hit_own(L.r, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, L.t_prune, e0);
        h1 = box_hit_own(L.r, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, L.t_prune, e1);
    } else {
        h0 = (ref0 != YUNE_REF_EMPTY) && box_hit(L.r, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, e0) && !(e0 > L.t_prune);
        h1 = (ref1 != YUNE_REF_EMPTY) && box_hit(L.r, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, e1) && !(e1 > L.t_prune);
    }
    const bool both = h0 && h1, any = h0 || h1;
    const bool swap = !any_q && both && (e1 < e0);
    const int near = (h0 && !swap) ? ref0 : ref1, far = swap ? ref0 : ref1;
    const int c0 = any ? near : top1;
    const int c1 = both ? far : (any ? top1 : top2);
    const bool park = c0 < 0 && !(L.pend_pos < L.pend_end);
    const int x = ~c0;
    if (both && !park) L.stack[L.sp] = far;
    L.sp += (both ? 1 : (any ? 0 : -1)) - (park ? 1 : 0);
    L.cur = park ? c1 : c0;
    if (park) { L.pend_pos = x >> 4; L.pend_end = (x >> 4) + (x & 15); }
}
// accel 2: one step over a 4-wide record.  Same rules as sim_inner, stated as "push every hit child far to near, then take
// the next place of the walk": a leaf on top is parked when nothing is parked (and the walk moves on to the place after it),
// otherwise the lane stands on it until its parked triangles have been tested.
void sim_inner_wide(SimLane& L, const SimScene& S, bool any_q)
{
    F4 q[7]; S.qf(L.cur, q);
    auto box = [&](float lox, float hix, float loy, float hiy, float loz, float hiz, float& e) { return box_hit_own(L.r, lox, hix, loy, hiy, loz, hiz, L.t_prune, e); };
    if (std::getenv("YUNE_SIM_WIDE_V1")) {          // the first form of the step (pushes through memory, record order for shadow queries)
        if (any_q) lane_wide_step<SimLane, decltype(box), true>(L, L.stack, q, box, STACK_BASE);
        else lane_wide_step<SimLane, decltype(box), false>(L, L.stack, q, box, STACK_BASE);
        return;
    }
    // the device form (kernels.cu lane_inner_step_wide): keys + lane_wide_finish, distance order for both query kinds
    const int top1 = L.stack[L.sp - 1], top2 = L.stack[L.sp - 2];
    const float* f = &q[0].x;
    int k[4], r[4] = {YF_ASINT(q[6].x), YF_ASINT(q[6].y), YF_ASINT(q[6].z), YF_ASINT(q[6].w)};
    for (int i = 0; i < 4; i++) {
        float e = 0.0f;
        const bool hit = r[i] != YUNE_REF_EMPTY && box(f[i], f[4 + i], f[8 + i], f[12 + i], f[16 + i], f[20 + i], e);
        k[i] = hit ? YF_ASINT(e) : YUNE_KEY_MISS;
    }
    lane_wide_finish(L, L.stack, top1, top2, k[0], k[1], k[2], k[3], r[0], r[1], r[2], r[3]);
}
void sim_tri(SimLane& L, const SimScene& S, bool any_q)
{
    const int top1 = L.stack[L.sp - 1 > 0 ? L.sp - 1 : 0];
    const int pos = L.pend_pos++;
    F4 a, b, c; S.tf(pos, a, b, c);
    float t, u, v;
    bool inside = tri_test(L.r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v);
    if (S.accel >= 1 && inside && t > 0.0f && !(t > L.t_best)) {
        F4 lo, hi; S.lf(YF_ASINT(c.w), lo, hi); float e;
        inside = box_hit(L.r, lo.x, hi.x, lo.y, hi.y, lo.z, hi.z, e);
    }
    bool stop = false;
    if (any_q) { stop = inside && t > 0.0f && t < L.t_best; if (stop) L.tri = 0; }
    else {
        const int rank = YF_ASINT(b.w);
        if (inside && t > 0.0f && (t < L.t_best || (t == L.t_best && L.best_pos >= 0 && rank < L.best_pos))) {
            L.t_best = t; L.u = u; L.v = v; L.tri = YF_ASINT(a.w); L.best_pos = rank; L.t_prune = t * 1.00001f;
        }
    }
    const bool park = !(L.pend_pos < L.pend_end) && L.cur < 0;
    if (park) { const int x = ~L.cur; L.pend_pos = x >> 4; L.pend_end = (x >> 4) + (x & 15); L.sp = L.sp - 1 > 0 ? L.sp - 1 : 0; L.cur = top1; }
    if (stop) { L.cur = REF_DONE; L.pend_end = L.pend_pos; }
}
}

namespace { inline bool any_q_of(int any) { return any != 0; } }
// knobs[4] = refill_idle, phase_min (tri_min), inner_min, inner_chain;  util[4] = inner steps, lanes in them, tri steps, lanes in them
extern "C" int hc_trace_warp(int n, const float* od6, const float* tmax, int any, const yune_triangle* tris, int ntri,
                             const yune_bvh_node* nodes, int nnodes, int* tri_id, float* t_hit, int accel, const int* knobs,
                             unsigned long long* util)
{
    TravLayoutHost lay; std::string err;
    const char* leaf_env = std::getenv("YUNE_SIM_LEAF");             // development: own-tree leaf size for the step model
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, accel == 0 ? 2 : (leaf_env ? std::atoi(leaf_env) : 0), accel)) return -1;
    accel = lay.accel;
    SimScene S{&lay, {lay.pairs.data()}, {lay.tris.data()}, {lay.leaf_boxes.data()}, accel, {lay.quads.data()}};
    auto sim_step = [&](SimLane& L) { if (accel == 2) sim_inner_wide(L, S, any_q_of(any)); else sim_inner(L, S, any_q_of(any)); };
    const int refill_idle = knobs[0], tri_min = knobs[1], inner_min = knobs[2], inner_chain = knobs[3];
    const bool any_q = any != 0;
    std::vector<SimLane> W(32);
    for (auto& L : W) { L.have = false; L.cur = REF_DONE; L.pend_pos = L.pend_end = 0; L.sp = STACK_BASE; }
    int next = 0;                                   // queue head (one warp drains the whole queue here)
    unsigned long long st[4] = {0, 0, 0, 0};
    auto busy = [&](const SimLane& L) { return L.cur >= 0 || L.pend_pos < L.pend_end; };
    for (;;) {
        for (auto& L : W) if (L.have && !busy(L)) { tri_id[L.where] = L.tri; if (t_hit) t_hit[L.where] = L.t_best; L.have = false; }
        int idle = 0; for (auto& L : W) idle += !L.have;
        bool exhausted = next >= n;
        if (idle == 32 && exhausted) break;
        for (auto& L : W) if (!L.have && next < n) {                  // refill: idle lanes take the next queue entries
            const int q = next++; const float* r = od6 + 6 * (size_t)q;
            sim_init(L, S, v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), tmax ? tmax[q] : INFINITY);
            L.where = q; L.have = true;
        }
        exhausted = next >= n;
        const int busy_min = exhausted ? 1 : 33 - refill_idle;
        for (;;) {
            int ni = 0, nt = 0, nb = 0;
            for (auto& L : W) { ni += L.cur >= 0; nt += L.pend_pos < L.pend_end; nb += busy(L); }
            if (nb < busy_min) break;
            static const int rule = [] { const char* e = std::getenv("YUNE_SIM_TRI_RULE"); return e ? std::atoi(e) : 0; }();      // development: other vote rules
            const bool tri_now = rule == 0 ? (nt >= tri_min || nt > ni) : rule == 1 ? (nt >= tri_min || ni == 0) : rule == 2 ? (nt >= tri_min || nt > 2 * ni)
                                : (nt >= tri_min || (nt > ni && ni < 8));
            if (tri_now) { st[2]++; st[3] += nt; for (auto& L : W) if (L.pend_pos < L.pend_end) sim_tri(L, S, any_q); }
            else {
                st[0]++; st[1] += ni; for (auto& L : W) if (L.cur >= 0) sim_step(L);
                for (int k = 0; k < inner_chain; k++) {
                    int c = 0; for (auto& L : W) c += L.cur >= 0;
                    if (c < inner_min) break;
                    st[0]++; st[1] += c; for (auto& L : W) if (L.cur >= 0) sim_step(L);
                }
            }
        }
    }
    if (util) for (int i = 0; i < 4; i++) util[i] = st[i];
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Design aid (no product path): the same scheduling with R ray slots per lane.  A lane takes part in a step when ANY of its
// slots wants the voted operation (one slot advances per lane and step), which is what raises the lanes per step; the price on
// the device would be R times the per-ray registers and stack.  util as in hc_trace_warp (lanes = lanes, not slots).
extern "C" int hc_trace_warp_multi(int n, const float* od6, const float* tmax, int any, const yune_triangle* tris, int ntri,
                                   const yune_bvh_node* nodes, int nnodes, int* tri_id, float* t_hit, int accel, const int* knobs,
                                   int rays_per_lane, unsigned long long* util)
{
    if (rays_per_lane < 1 || rays_per_lane > 4) return -1;
    TravLayoutHost lay; std::string err;
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, accel == 0 ? 2 : 0, accel)) return -1;
    accel = lay.accel;
    const bool any_q = any != 0;
    SimScene S{&lay, {lay.pairs.data()}, {lay.tris.data()}, {lay.leaf_boxes.data()}, accel, {lay.quads.data()}};
    const int R = rays_per_lane, NS = 32 * R;
    const int refill_idle = knobs[0], tri_min = knobs[1], inner_min = knobs[2], inner_chain = knobs[3];
    std::vector<SimLane> W(NS);
    for (auto& L : W) { L.have = false; L.cur = REF_DONE; L.pend_pos = L.pend_end = 0; L.sp = STACK_BASE; }
    int next = 0; unsigned long long st[4] = {0, 0, 0, 0};
    auto busy = [&](const SimLane& L) { return L.cur >= 0 || L.pend_pos < L.pend_end; };
    auto lanes_wanting = [&](bool tri) { int c = 0; for (int l = 0; l < 32; l++) { bool w = false; for (int r = 0; r < R; r++) { const SimLane& L = W[l * R + r]; w = w || (tri ? L.pend_pos < L.pend_end : L.cur >= 0); } c += w; } return c; };
    auto step = [&](bool tri) {
        for (int l = 0; l < 32; l++) for (int r = 0; r < R; r++) {
            SimLane& L = W[l * R + r];
            if (tri ? L.pend_pos < L.pend_end : L.cur >= 0) { if (tri) sim_tri(L, S, any_q); else if (accel == 2) sim_inner_wide(L, S, any_q); else sim_inner(L, S, any_q); break; }
        }
    };
    for (;;) {
        for (auto& L : W) if (L.have && !busy(L)) { tri_id[L.where] = L.tri; if (t_hit) t_hit[L.where] = L.t_best; L.have = false; }
        int idle = 0; for (auto& L : W) idle += !L.have;
        if (idle == NS && next >= n) break;
        for (auto& L : W) if (!L.have && next < n) {
            const int q = next++; const float* r = od6 + 6 * (size_t)q;
            sim_init(L, S, v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), tmax ? tmax[q] : INFINITY);
            L.where = q; L.have = true;
        }
        const int busy_min = next >= n ? 1 : NS + 1 - refill_idle * R;
        for (;;) {
            int nb = 0; for (auto& L : W) nb += busy(L);
            if (nb < busy_min) break;
            const int ni = lanes_wanting(false), nt = lanes_wanting(true);
            if (nt >= tri_min || nt > ni) { st[2]++; st[3] += nt; step(true); }
            else {
                st[0]++; st[1] += ni; step(false);
                for (int k = 0; k < inner_chain; k++) { const int c = lanes_wanting(false); if (c < inner_min) break; st[0]++; st[1] += c; step(false); }
            }
        }
    }
    if (util) for (int i = 0; i < 4; i++) util[i] = st[i];
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Design aid (no product path): how many steps would a WIDER tree take?  The own tree (accel 1) is collapsed into nodes of up
// to `width` children (the inner child with the largest box is replaced by its two children until the node is full) and
// walked front to back with the same conservative box test and pruning; width 2 is the shipped tree.  Returns per-ray
// totals: out[0] = node visits, out[1] = child boxes tested, out[2] = triangle tests, out[3] = rays that hit.
// ------------------------------------------------------------------------------------------------------------------
#include <algorithm>
namespace {
struct WideNode { int n; float lo[8][3], hi[8][3]; int ref[8]; };      // ref >= 0: wide node index; < 0: ~((first << 4) | count)
}
extern "C" int hc_wide_stats(int n, const float* od6, const float* tmax, int any, const yune_triangle* tris, int ntri,
                             const yune_bvh_node* nodes, int nnodes, int width, unsigned long long* out)
{
    if (width < 2 || width > 8) return -1;
    TravLayoutHost lay; std::string err;
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, 0, 1)) return -1;
    const F4* P = lay.pairs.data();
    struct Child { float lo[3], hi[3]; int ref; };
    auto children_of = [&](int pair, Child* c) {
        const F4* q = P + (size_t)pair * 4;
        c[0] = {{q[0].x, q[0].z, q[2].x}, {q[0].y, q[0].w, q[2].y}, YF_ASINT(q[3].x)};
        c[1] = {{q[1].x, q[1].z, q[2].z}, {q[1].y, q[1].w, q[2].w}, YF_ASINT(q[3].y)};
    };
    std::vector<WideNode> wide;
    std::vector<int> wide_of(lay.n_inner, -1), todo;
    auto area = [](const Child& c) { const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2]; return dx * dy + dx * dz + dy * dz; };
    if (lay.root_ref >= 0) { wide_of[lay.root_ref] = 0; wide.push_back(WideNode()); todo.push_back(lay.root_ref); }
    for (size_t k = 0; k < todo.size(); k++) {
        const int pair = todo[k];
        std::vector<Child> cs(2); children_of(pair, cs.data());
        while ((int)cs.size() < width) {
            int pick = -1;
            for (int i = 0; i < (int)cs.size(); i++) if (cs[i].ref >= 0 && (pick < 0 || area(cs[i]) > area(cs[pick]))) pick = i;
            if (pick < 0) break;
            Child two[2]; children_of(cs[pick].ref, two);
            cs[pick] = two[0]; cs.push_back(two[1]);
        }
        WideNode w; w.n = (int)cs.size();
        for (int i = 0; i < w.n; i++) {
            for (int a = 0; a < 3; a++) { w.lo[i][a] = cs[i].lo[a]; w.hi[i][a] = cs[i].hi[a]; }
            if (cs[i].ref >= 0) { wide_of[cs[i].ref] = (int)wide.size() + 0; wide.push_back(WideNode()); todo.push_back(cs[i].ref); w.ref[i] = wide_of[cs[i].ref]; }
            else w.ref[i] = cs[i].ref;
        }
        wide[wide_of[pair]] = w;
    }
    HostTriFetch tf{lay.tris.data()}; HostLeafFetch lf{lay.leaf_boxes.data()};
    unsigned long long visits = 0, boxes = 0, tests = 0, hits = 0;
    std::vector<int> stack;
    for (int i = 0; i < n; i++) {
        const float* r = od6 + 6 * (size_t)i;
        RayPre ray = make_ray(v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]));
        float t_best = tmax ? tmax[i] : INFINITY, t_prune = t_best * 1.00001f; int tri = -1, best_pos = -1;
        stack.clear();
        if (lay.root_ref != YUNE_REF_EMPTY) stack.push_back(lay.root_ref >= 0 ? 0 : lay.root_ref);
        bool done = false;
        while (!stack.empty() && !done) {
            const int ref = stack.back(); stack.pop_back();
            if (ref < 0) {
                const int x = ~ref;
                for (int pos = x >> 4; pos < (x >> 4) + (x & 15) && !done; pos++) {
                    F4 a, b, c; tf(pos, a, b, c); float t, u, v; tests++;
                    if (!tri_test(ray, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v) || !(t > 0.0f) || t > t_best) continue;
                    F4 lo, hi; lf(YF_ASINT(c.w), lo, hi); float e;
                    if (!box_hit(ray, lo.x, hi.x, lo.y, hi.y, lo.z, hi.z, e)) continue;
                    if (any) { if (t < t_best) { tri = 0; done = true; } }
                    else if (t < t_best || (best_pos >= 0 && YF_ASINT(b.w) < best_pos)) { t_best = t; t_prune = t * 1.00001f; tri = YF_ASINT(a.w); best_pos = YF_ASINT(b.w); }
                }
                continue;
            }
            const WideNode& w = wide[ref]; visits++; boxes += w.n;
            std::pair<float, int> hit[8]; int nh = 0;
            for (int k = 0; k < w.n; k++) { float e; if (box_hit_own(ray, w.lo[k][0], w.hi[k][0], w.lo[k][1], w.hi[k][1], w.lo[k][2], w.hi[k][2], t_prune, e)) hit[nh++] = {e, w.ref[k]}; }
            if (!any) for (int a = 1; a < nh; a++) { const std::pair<float, int> v = hit[a]; int b = a; while (b > 0 && hit[b - 1].first < v.first) { hit[b] = hit[b - 1]; b--; } hit[b] = v; }      // farthest first
            for (int k = 0; k < nh; k++) stack.push_back(hit[k].second);      // nearest on top
        }
        hits += tri >= 0;
    }
    out[0] = visits; out[1] = boxes; out[2] = tests; out[3] = hits;
    return 0;
}

// which walk a layout ended up with (accel 1 / 2 fall back to 0 when the uploaded boxes do not nest; bvh_size == 0 forces 1)
extern "C" int hc_layout_accel(const yune_triangle* tris, int ntri, const yune_bvh_node* nodes, int nnodes, int leaf_split, int accel)
{
    TravLayoutHost lay; std::string err;
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, leaf_split, accel)) return -1;
    return lay.accel;
}

// what the device layout path takes from an uploaded tree (relayout.cpp: referenceLeavesForDevice); returns -1 malformed, 0 not usable, 1 usable
extern "C" int hc_reference_leaves(const yune_triangle* tris, int ntri, const yune_bvh_node* nodes, int nnodes, int* leaf_of_tri, int* rank_of_tri, float* leaf_boxes8, int* n_leaves)
{
    std::vector<int> l, r; std::vector<F4> b; bool usable = false; std::string err;
    if (!referenceLeavesForDevice(tris, ntri, nodes, nnodes, l, r, b, usable, err)) return -1;
    if (!usable) return 0;
    for (int i = 0; i < ntri; i++) { leaf_of_tri[i] = l[i]; rank_of_tri[i] = r[i]; }
    *n_leaves = (int)(b.size() / 2);
    if (leaf_boxes8) std::memcpy(leaf_boxes8, b.data(), b.size() * sizeof(F4));
    return 1;
}
