// tests/hostcheck/simd_sim.cpp -- DEVELOPMENT TOOL (test-only host code): simulates the warp scheduling policies of
// k_trace on real ray sets with the product's own state machine (trace_core.h), to estimate SIMD utilisation
// (active lanes per executed step) before a policy is written in CUDA.  Not part of the product.
#include "trace_core.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
using namespace yune;
namespace {
struct HostPairFetch { const F4* p; void operator()(int i, F4& a, F4& b, F4& c, F4& d) const { const F4* q = p + (size_t)i * 4; a = q[0]; b = q[1]; c = q[2]; d = q[3]; } };
struct HostTriFetch { const F4* p; void operator()(int i, F4& a, F4& b, F4& c) const { const F4* q = p + (size_t)i * 3; a = q[0]; b = q[1]; c = q[2]; } };
struct LaneS { TraceState s; int stack[YUNE_STACK_SIZE]; bool have; int pend_pos[8], pend_end[8], npend; };
}
// policy 0: current (phases with hysteresis phase_min, phase_max, refill at refill_idle)
// policy 1: postponed leaves with capacity cap: INNER phase keeps traversing, leaves are queued per lane
extern "C" void simd_sim(int n, const float* od6, const float* tmax, int any, const yune_triangle* tris, int ntri,
                         const yune_bvh_node* nodes, int nnodes, int leaf_split, int policy, int phase_min, int phase_max, int refill_idle, int cap,
                         int rays_per_warp, double* out)
{
    TravLayoutHost lay; std::string err;
    if (!buildTravLayout(tris, ntri, nodes, nnodes, lay, err, leaf_split)) { out[0] = -1; return; }
    HostPairFetch pf{lay.pairs.data()}; HostTriFetch tf{lay.tris.data()};
    double inner_exec = 0, inner_lanes = 0, tri_exec = 0, tri_lanes = 0, passes = 0;
    WorkCount wc = {0, 0};
    for (int w0 = 0; w0 < n; w0 += rays_per_warp) {
        const int wn = std::min(rays_per_warp, n - w0);
        int next = 0;
        std::vector<LaneS> L(32);
        for (auto& l : L) { l.have = false; l.npend = 0; }
        bool exhausted = false;
        for (;;) {
            // retire + refill
            int idle = 0;
            for (auto& l : L) { if (l.have && l.s.done && l.npend == 0) l.have = false; if (!l.have) idle++; }
            if (idle == 32 && exhausted) break;
            if (!exhausted && (idle == 32 || idle >= refill_idle)) {
                for (auto& l : L) if (!l.have && next < wn) {
                    const float* r = od6 + 6 * (size_t)(w0 + next);
                    const float t = tmax ? tmax[w0 + next] : INFINITY; next++;
                    ts_init<true>(l.s, v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), t, lay.root_ref, lay.root_lo, lay.root_hi, &wc);
                    l.have = true; l.npend = 0;
                }
                if (next >= wn) exhausted = true;
            }
            passes++;
            bool progressed = false;
            if (policy == 0) {
                for (int k = 0; k < phase_max; k++) {
                    int want = 0; for (auto& l : L) if (l.have && !l.s.done && !(l.s.leaf_pos < l.s.leaf_end)) want++;
                    if (want < phase_min) break;
                    for (auto& l : L) if (l.have && !l.s.done && !(l.s.leaf_pos < l.s.leaf_end)) { if (any) ts_inner_step<HostPairFetch, true, true>(l.s, l.stack, pf, &wc); else ts_inner_step<HostPairFetch, false, true>(l.s, l.stack, pf, &wc); }
                    inner_exec++; inner_lanes += want; progressed = true;
                }
                for (int k = 0; k < phase_max; k++) {
                    int want = 0; for (auto& l : L) if (l.have && l.s.leaf_pos < l.s.leaf_end) want++;
                    if (want < phase_min) break;
                    for (auto& l : L) if (l.have && l.s.leaf_pos < l.s.leaf_end) { if (any) ts_tri_step<HostTriFetch, true, true>(l.s, l.stack, tf, &wc); else ts_tri_step<HostTriFetch, false, true>(l.s, l.stack, tf, &wc); }
                    tri_exec++; tri_lanes += want; progressed = true;
                }
                if (!progressed) {
                    int wi = 0, wt = 0;
                    for (auto& l : L) if (l.have && !l.s.done && !(l.s.leaf_pos < l.s.leaf_end)) { wi++; if (any) ts_inner_step<HostPairFetch, true, true>(l.s, l.stack, pf, &wc); else ts_inner_step<HostPairFetch, false, true>(l.s, l.stack, pf, &wc); }
                    for (auto& l : L) if (l.have && l.s.leaf_pos < l.s.leaf_end) { wt++; if (any) ts_tri_step<HostTriFetch, true, true>(l.s, l.stack, tf, &wc); else ts_tri_step<HostTriFetch, false, true>(l.s, l.stack, tf, &wc); }
                    if (wi) { inner_exec++; inner_lanes += wi; } if (wt) { tri_exec++; tri_lanes += wt; }
                }
            } else {
                // postponed leaves: a lane in TRI mode moves its leaf range to the pending list and pops on, until cap leaves are pending
                auto wants_inner = [&](LaneS& l) {
                    if (!l.have || l.s.done) return false;
                    while (l.s.leaf_pos < l.s.leaf_end && l.npend < cap) {      // defer the leaf, continue the walk
                        l.pend_pos[l.npend] = l.s.leaf_pos; l.pend_end[l.npend] = l.s.leaf_end; l.npend++;
                        l.s.leaf_pos = l.s.leaf_end = 0; ts_pop(l.s, l.stack);
                        if (l.s.done) return false;
                    }
                    return !(l.s.leaf_pos < l.s.leaf_end);
                };
                for (int k = 0; k < phase_max; k++) {
                    int want = 0; std::vector<int> idx;
                    for (int i = 0; i < 32; i++) if (wants_inner(L[i])) { want++; idx.push_back(i); }
                    if (want < phase_min) break;
                    for (int i : idx) { if (any) ts_inner_step<HostPairFetch, true, true>(L[i].s, L[i].stack, pf, &wc); else ts_inner_step<HostPairFetch, false, true>(L[i].s, L[i].stack, pf, &wc); }
                    inner_exec++; inner_lanes += want; progressed = true;
                }
                // TRI phase: process pending triangles one per lane per step (the current leaf of the state machine counts too)
                auto tri_one = [&](LaneS& l) {
                    // pending leaves first
                    if (l.npend > 0) {
                        TraceState tmp = l.s; int dummy_stack[1]; tmp.sp = 0; tmp.done = false;
                        tmp.leaf_pos = l.pend_pos[0]; tmp.leaf_end = l.pend_pos[0] + 1;
                        const float tb = l.s.t_best; (void)tb;
                        // run one triangle test against the lane's current best
                        int pos = l.pend_pos[0]++;
                        F4 a, b, c; tf(pos, a, b, c); float t, u, v; wc.tri++;
                        if (tri_test(l.s.r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v)) {
                            if (any) { if (t > 0.0f && t < l.s.t_best) { l.s.tri = 0; l.s.done = true; l.s.sp = 0; l.s.leaf_pos = l.s.leaf_end = 0; l.npend = 0; return; } }
                            else if (t > 0.0f && (t < l.s.t_best || (t == l.s.t_best && l.s.best_pos >= 0 && YF_ASINT(b.w) < l.s.best_pos))) {
                                l.s.t_best = t; l.s.u = u; l.s.v = v; l.s.tri = YF_ASINT(a.w); l.s.best_pos = YF_ASINT(b.w); l.s.t_prune = t * 1.00001f; }
                        }
                        if (l.pend_pos[0] >= l.pend_end[0]) { for (int j = 1; j < l.npend; j++) { l.pend_pos[j - 1] = l.pend_pos[j]; l.pend_end[j - 1] = l.pend_end[j]; } l.npend--; }
                        (void)dummy_stack;
                        return;
                    }
                    if (any) ts_tri_step<HostTriFetch, true, true>(l.s, l.stack, tf, &wc); else ts_tri_step<HostTriFetch, false, true>(l.s, l.stack, tf, &wc);
                };
                for (int k = 0; k < phase_max * 2; k++) {
                    int want = 0; for (auto& l : L) if (l.have && (l.npend > 0 || l.s.leaf_pos < l.s.leaf_end)) want++;
                    if (want < phase_min) break;
                    for (auto& l : L) if (l.have && (l.npend > 0 || l.s.leaf_pos < l.s.leaf_end)) tri_one(l);
                    tri_exec++; tri_lanes += want; progressed = true;
                }
                if (!progressed) {
                    int wi = 0, wt = 0;
                    for (int i = 0; i < 32; i++) if (wants_inner(L[i])) { wi++; if (any) ts_inner_step<HostPairFetch, true, true>(L[i].s, L[i].stack, pf, &wc); else ts_inner_step<HostPairFetch, false, true>(L[i].s, L[i].stack, pf, &wc); }
                    for (auto& l : L) if (l.have && (l.npend > 0 || l.s.leaf_pos < l.s.leaf_end)) { wt++; tri_one(l); }
                    if (wi) { inner_exec++; inner_lanes += wi; } if (wt) { tri_exec++; tri_lanes += wt; }
                }
            }
        }
    }
    out[0] = inner_exec; out[1] = inner_lanes; out[2] = tri_exec; out[3] = tri_lanes; out[4] = passes; out[5] = wc.box; out[6] = wc.tri;
}
