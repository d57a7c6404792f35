// trace_core.h -- ray / box, ray / triangle and the stack traversal, written once for device and host.
//
// What it reproduces (reference: kernels/legacy/udpt.cl):
//   box_hit()   = rayAabbIntersection (:392-431), same slab arithmetic ((p - o) * (1/d)), same NaN guards.
//   tri_test()  = rayTriangleIntersection (:326-390), Moller-Trumbore in the reference's operation order.
//   closest_hit / any_hit = what traverseBVH (:288-324) returns, but walked depth-first, near child first,
//                 skipping boxes that start behind the current hit.  The reference walks breadth-first and
//                 never prunes; the RESULT is the same triangle because (a) a box is entered under exactly the
//                 reference's predicate, (b) exact ties in t go to the lower breadth-first rank, as in the
//                 reference, and (c) pruning keeps a 1e-5 relative margin so float noise between a box's entry
//                 distance and the t of a triangle lying on that box's face cannot drop a winner.
//
// The node/triangle fetchers are template parameters: on the device they read shared memory or the
// read-only path with 16-byte vector loads; the host build (tests/hostcheck) reads plain arrays.
#ifndef YUNE_TRACE_CORE_H
#define YUNE_TRACE_CORE_H

#include "strict_math.h"
#include "trav_layout.h"

namespace yune {

struct RayPre {
    V3 o, d, inv;     // inv = 1 / d, IEEE division (udpt.cl:395)
    bool guard;       // some 1/d is not finite -> NaNs are possible in the slab products
};

YUNE_HD RayPre make_ray(V3 o, V3 d)
{
    RayPre r; r.o = o; r.d = d;
    r.inv = v3(YF_DIV(1.0f, d.x), YF_DIV(1.0f, d.y), YF_DIV(1.0f, d.z));
    // finite <=> |x| < inf (false for NaN too)
    r.guard = !(fabsf(r.inv.x) < INFINITY && fabsf(r.inv.y) < INFINITY && fabsf(r.inv.z) < INFINITY);
    return r;
}

// One slab axis exactly as the reference writes it (udpt.cl:400-416).
YUNE_HD void slab_guarded(float lo, float hi, float o, float inv, float& t_min, float& t_max)
{
    const float a = YF_MUL(YF_SUB(lo, o), inv), b = YF_MUL(YF_SUB(hi, o), inv);
    if (!(a != a)) {
        t_min = fmaxf(cl_min(a, b), t_min);
        t_max = cl_min(fmaxf(a, b), t_max);
    }
}

// Returns the reference's predicate 't_max > fmax(t_min, 0)' and the entry distance fmax(t_min, 0).
// (The reference's early 'if (t_max < t_min) return false' after the y slab cannot change the outcome:
//  the z slab only raises t_min and lowers t_max.)
YUNE_HD bool box_hit(const RayPre& r, float lox, float hix, float loy, float hiy, float loz, float hiz, float& entry)
{
    float t_min, t_max;
    if (!r.guard) {
        // no NaN can occur: plain min/max give the same values as the guarded form
        const float ax = YF_MUL(YF_SUB(lox, r.o.x), r.inv.x), bx = YF_MUL(YF_SUB(hix, r.o.x), r.inv.x);
        const float ay = YF_MUL(YF_SUB(loy, r.o.y), r.inv.y), by = YF_MUL(YF_SUB(hiy, r.o.y), r.inv.y);
        const float az = YF_MUL(YF_SUB(loz, r.o.z), r.inv.z), bz = YF_MUL(YF_SUB(hiz, r.o.z), r.inv.z);
        t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        t_max = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    } else {
        t_min = -INFINITY; t_max = INFINITY;
        slab_guarded(lox, hix, r.o.x, r.inv.x, t_min, t_max);
        slab_guarded(loy, hiy, r.o.y, r.inv.y, t_min, t_max);
        slab_guarded(loz, hiz, r.o.z, r.inv.z, t_min, t_max);
    }
    entry = fmaxf(t_min, 0.0f);
    return t_max > entry;
}

// Moller-Trumbore, two-sided, no determinant epsilon, borders included (udpt.cl:326-350).  `tri` points at the
// 48-byte record (v1, e1 = v2 - v1, e2 = v3 - v1).  Returns true when (u, v) are inside; t may still be rejected
// by the caller's distance test.  NaNs (det == 0) fall through the comparisons exactly as in the reference.
YUNE_HD bool tri_test(const RayPre& r, V3 v1, V3 e1, V3 e2, float& t, float& u, float& v)
{
    const V3 pvec = vcross(r.d, e2);
    const float det = vdot(e1, pvec);
    const float inv_det = YF_DIV(1.0f, det);
    const V3 dist = vsub(r.o, v1);
    u = YF_MUL(vdot(pvec, dist), inv_det);
    if (u < 0.0f || u > 1.0f) return false;
    const V3 qvec = vcross(dist, e1);
    v = YF_MUL(vdot(qvec, r.d), inv_det);
    if (v < 0.0f || YF_ADD(u, v) > 1.0f) return false;
    t = YF_MUL(vdot(e2, qvec), inv_det);
    return true;
}

struct HitRec { float t, u, v; int tri; };   // tri = original triangle index, -1 = nothing closer than t_in

struct WorkCount { unsigned box, tri; };

// Closest hit.  `fetch_pair(idx, q0..q3)` and `fetch_tri(pos, t0, t1, t2)` load one record each.
template <class PairFetch, class TriFetch, bool COUNT>
YUNE_HD void closest_hit(const PairFetch& fetch_pair, const TriFetch& fetch_tri, int root_ref,
                         const float* root_lo, const float* root_hi,
                         const RayPre& r, float t_in, HitRec& out, WorkCount* wc)
{
    out.t = t_in; out.u = 0.0f; out.v = 0.0f; out.tri = -1;
    if (root_ref == YUNE_REF_EMPTY) return;
    float entry;
    if (COUNT) wc->box++;
    if (!box_hit(r, root_lo[0], root_hi[0], root_lo[1], root_hi[1], root_lo[2], root_hi[2], entry)) return;   // udpt.cl:295-296

    int stack[YUNE_STACK_SIZE];
    int sp = 0, cur = root_ref, best_pos = -1;
    float t_prune = t_in * 1.00001f;          // inf stays inf
    for (;;) {
        if (cur >= 0) {
            F4 q0, q1, q2, q3;
            fetch_pair(cur, q0, q1, q2, q3);
            const int ref0 = YF_ASINT(q3.x), ref1 = YF_ASINT(q3.y);
            float e0, e1;
            if (COUNT) wc->box += (ref0 != YUNE_REF_EMPTY) + (ref1 != YUNE_REF_EMPTY);
            bool h0 = (ref0 != YUNE_REF_EMPTY) && box_hit(r, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, e0);
            bool h1 = (ref1 != YUNE_REF_EMPTY) && box_hit(r, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, e1);
            h0 = h0 && !(e0 > t_prune);
            h1 = h1 && !(e1 > t_prune);
            if (h0 && h1) {
                const bool swap = e1 < e0;
                stack[sp++] = swap ? ref0 : ref1;
                cur = swap ? ref1 : ref0;
                continue;
            }
            if (h0) { cur = ref0; continue; }
            if (h1) { cur = ref1; continue; }
        } else {
            const int first = (~cur) >> 4, count = (~cur) & 15;
            for (int k = 0; k < count; k++) {
                F4 a, b, c;
                fetch_tri(first + k, a, b, c);
                float t, u, v;
                if (COUNT) wc->tri++;
                if (!tri_test(r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v)) continue;
                // udpt.cl:373 't > 0 && t < ray->length', with the reference's first-come rule for exact ties
                if (t > 0.0f && (t < out.t || (t == out.t && best_pos >= 0 && first + k < best_pos))) {
                    out.t = t; out.u = u; out.v = v; out.tri = YF_ASINT(a.w); best_pos = first + k;
                    t_prune = t * 1.00001f;
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
}

// Any hit in (0, t_in): the reference's shadow-ray early-out (udpt.cl:306-308).
template <class PairFetch, class TriFetch, bool COUNT>
YUNE_HD bool any_hit(const PairFetch& fetch_pair, const TriFetch& fetch_tri, int root_ref,
                     const float* root_lo, const float* root_hi, const RayPre& r, float t_in, WorkCount* wc)
{
    if (root_ref == YUNE_REF_EMPTY) return false;
    float entry;
    if (COUNT) wc->box++;
    if (!box_hit(r, root_lo[0], root_hi[0], root_lo[1], root_hi[1], root_lo[2], root_hi[2], entry)) return false;
    int stack[YUNE_STACK_SIZE];
    int sp = 0, cur = root_ref;
    const float t_prune = t_in * 1.00001f;
    for (;;) {
        if (cur >= 0) {
            F4 q0, q1, q2, q3;
            fetch_pair(cur, q0, q1, q2, q3);
            const int ref0 = YF_ASINT(q3.x), ref1 = YF_ASINT(q3.y);
            float e0, e1;
            if (COUNT) wc->box += (ref0 != YUNE_REF_EMPTY) + (ref1 != YUNE_REF_EMPTY);
            bool h0 = (ref0 != YUNE_REF_EMPTY) && box_hit(r, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, e0);
            bool h1 = (ref1 != YUNE_REF_EMPTY) && box_hit(r, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, e1);
            h0 = h0 && !(e0 > t_prune);
            h1 = h1 && !(e1 > t_prune);
            if (h0 && h1) { stack[sp++] = ref1; cur = ref0; continue; }
            if (h0) { cur = ref0; continue; }
            if (h1) { cur = ref1; continue; }
        } else {
            const int first = (~cur) >> 4, count = (~cur) & 15;
            for (int k = 0; k < count; k++) {
                F4 a, b, c;
                fetch_tri(first + k, a, b, c);
                float t, u, v;
                if (COUNT) wc->tri++;
                if (tri_test(r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v) && t > 0.0f && t < t_in)
                    return true;
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return false;
}

} // namespace yune
#endif
