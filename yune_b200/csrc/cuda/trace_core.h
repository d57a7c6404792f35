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
    V3 oi;            // o * inv, for the fused slab test of our own tree (accel 1)
    bool guard;       // some 1/d is not finite -> NaNs are possible in the slab products
};

YUNE_HD RayPre make_ray(V3 o, V3 d)
{
    RayPre r; r.o = o; r.d = d;
    r.inv = v3(YF_DIV(1.0f, d.x), YF_DIV(1.0f, d.y), YF_DIV(1.0f, d.z));
    // finite <=> |x| < inf (false for NaN too)
    r.guard = !(fabsf(r.inv.x) < INFINITY && fabsf(r.inv.y) < INFINITY && fabsf(r.inv.z) < INFINITY);
    r.oi = v3(YF_MUL(o.x, r.inv.x), YF_MUL(o.y, r.inv.y), YF_MUL(o.z, r.inv.z));
    // o * (1/d) can overflow although 1/d is finite (|d| ~ 1e-37 with |o| ~ 1e2): lo * inv - oi would then be -inf for both planes of
    // a box and box_hit_own would lose it.  Such rays take the guarded subtract-then-multiply form too (kernels.cu: lane_init).
    r.guard = r.guard || !(fabsf(r.oi.x) < INFINITY && fabsf(r.oi.y) < INFINITY && fabsf(r.oi.z) < INFINITY);
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

// Slab test for OUR boxes (accel 1): one fused multiply-add per plane, result widened by a few ulps.  It only has to be
// CONSERVATIVE (never reject a box that holds a triangle the ray hits); which triangles the reference would have reached is
// decided separately, with box_hit() on the reference leaf's box.  A NaN (0 * inf on an axis the ray does not move along) is
// dropped by fminf/fmaxf, i.e. that axis does not constrain -- conservative again.
#if defined(__CUDA_ARCH__)
  #define YF_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
  #define YF_FMA(a, b, c) fmaf((a), (b), (c))
#endif
YUNE_HD bool box_hit_own(const RayPre& r, float lox, float hix, float loy, float hiy, float loz, float hiz, float t_prune, float& entry)
{
    if (r.guard) {      // 1/d not finite: lo * inf - o * inf is inf - inf; use the subtract-then-multiply form with its NaN guards
        const bool h = box_hit(r, lox, hix, loy, hiy, loz, hiz, entry);
        return h && !(entry > t_prune);
    }
    const float ax = YF_FMA(lox, r.inv.x, -r.oi.x), bx = YF_FMA(hix, r.inv.x, -r.oi.x);
    const float ay = YF_FMA(loy, r.inv.y, -r.oi.y), by = YF_FMA(hiy, r.inv.y, -r.oi.y);
    const float az = YF_FMA(loz, r.inv.z, -r.oi.z), bz = YF_FMA(hiz, r.inv.z, -r.oi.z);
    const float t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    const float t_max = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t_prune));
    entry = t_min;
    return YF_MUL(t_max, 1.0000021f) >= t_min;        // (t_min >= 0) at least as wide as t_max (1 + 1e-6) >= t_min (1 - 1e-6)
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

// PERF-MODE intersection (option "isect" = 1; north star: "a watertight ray-triangle test"; replaces rayTriangleIntersection,
// udpt.cl:326-390, whose edge decisions are not consistent between the two triangles that share an edge).  Woop / Benthin / Wald's
// formulation: vertices relative to the ray origin, axes permuted so that z is the ray's major axis, sheared so that the ray
// becomes the z axis; then 2-D edge functions
//     U = C'x B'y - C'y B'x    V = A'x C'y - A'y C'x    W = B'x A'y - B'y A'x        inside <=> no two of them have opposite signs
//     det = U + V + W,   t = (U A'z + V B'z + W C'z) / det,   (u, v) = (V, W) / det   (the reference's barycentrics of v2 / v3)
// The sheared coordinates are offsets from the RAY (small near the hit), so the products do not cancel the way origin-relative
// triple products do.  Watertight along edges WITHOUT the paper's double-precision fallback because nothing here is contracted
// into an FMA: a sheared vertex is a function of (ray, vertex bits) only -- the same for every triangle that uses the vertex --
// and (a*b) - (c*d) is exactly antisymmetric, so the two triangles of an edge (P, Q) see edge values that are exact negatives of
// each other: a ray cannot be outside of both, and a ray on the edge (value 0) is inside both.  Borders included, two-sided, like
// the reference's test.  Needs the three RAW vertices (trav_layout.h, isect 1): v1 + (v2 - v1) would not reproduce v2's bits.
struct WtRay { V3 S; int kz; };        // S = (d[kx] / d[kz], d[ky] / d[kz], 1 / d[kz]); (kx, ky, kz) cyclic, kz = major axis of d
YUNE_HD V3 wt_perm(V3 p, int kz) { return kz == 2 ? p : (kz == 0 ? v3(p.y, p.z, p.x) : v3(p.z, p.x, p.y)); }
YUNE_HD WtRay wt_setup(V3 d)
{
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    WtRay w;
    w.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    const V3 pd = wt_perm(d, w.kz);
    const float sz = YF_DIV(1.0f, pd.z);
    w.S = v3(YF_MUL(pd.x, sz), YF_MUL(pd.y, sz), sz);
    return w;
}
YUNE_HD bool tri_test_watertight(V3 o, const WtRay& w, V3 v1, V3 v2, V3 v3_, float& t, float& u, float& v)
{
    const V3 A = wt_perm(vsub(v1, o), w.kz), B = wt_perm(vsub(v2, o), w.kz), C = wt_perm(vsub(v3_, o), w.kz);
    const float Ax = YF_SUB(A.x, YF_MUL(w.S.x, A.z)), Ay = YF_SUB(A.y, YF_MUL(w.S.y, A.z));
    const float Bx = YF_SUB(B.x, YF_MUL(w.S.x, B.z)), By = YF_SUB(B.y, YF_MUL(w.S.y, B.z));
    const float Cx = YF_SUB(C.x, YF_MUL(w.S.x, C.z)), Cy = YF_SUB(C.y, YF_MUL(w.S.y, C.z));
    const float U = YF_SUB(YF_MUL(Cx, By), YF_MUL(Cy, Bx));
    const float V = YF_SUB(YF_MUL(Ax, Cy), YF_MUL(Ay, Cx));
    const float W = YF_SUB(YF_MUL(Bx, Ay), YF_MUL(By, Ax));
    const bool neg = (U < 0.0f) || (V < 0.0f) || (W < 0.0f), pos = (U > 0.0f) || (V > 0.0f) || (W > 0.0f);
    const float det = YF_ADD(YF_ADD(U, V), W);
    const float inv_det = YF_DIV(1.0f, det);
    const float T = YF_ADD(YF_ADD(YF_MUL(U, YF_MUL(w.S.z, A.z)), YF_MUL(V, YF_MUL(w.S.z, B.z))), YF_MUL(W, YF_MUL(w.S.z, C.z)));
    t = YF_MUL(T, inv_det);
    u = YF_MUL(V, inv_det); v = YF_MUL(W, inv_det);
    return !(neg && pos) && det != 0.0f;           // a NaN passes here and fails the caller's 't > 0'
}

struct HitRec { float t, u, v; int tri; };   // tri = original triangle index, -1 = nothing closer than t_in

struct WorkCount { unsigned box, tri; };

// ------------------------------------------------------------------------------------------------------------------
// The walk as a per-ray STATE MACHINE whose transitions are two fixed-size operations:
//     inner step = fetch one pair record, test both child boxes, descend / push / pop
//     tri step   = fetch one triangle record, Moller-Trumbore, maybe accept
// A ray is always in exactly one of three modes: INNER (cur >= 0), TRI (leaf_pos < leaf_end) or done.  Written this way
// the GPU kernel can schedule a warp by majority vote -- every pass executes the ONE operation most lanes are waiting for,
// with all those lanes converged -- instead of letting 32 rays serialise through divergent loops (the first version of
// this kernel ran with 4.4 of 32 threads active per instruction; profiles/ has the ncu capture).  The host wrappers
// below drive the same machine one ray at a time, which is what tests/hostcheck compares against the oracle.
// ------------------------------------------------------------------------------------------------------------------
#define YUNE_MODE_NONE  0
#define YUNE_MODE_INNER 1
#define YUNE_MODE_TRI   2

struct TraceState {
    RayPre r;
    float t_best, t_prune;      // current ray length; pruning bound = t_best * (1 + 1e-5)
    float u, v;
    int   tri, best_pos;        // accepted triangle (original index) and its reference visiting rank; -1 = none
    int   cur;                  // INNER mode: pair index
    int   leaf_pos, leaf_end;   // TRI mode: triangles [leaf_pos, leaf_end) of the current leaf are still to be tested
    int   sp;
    bool  done;
};

YUNE_HD int ts_mode(const TraceState& s) { return s.done ? YUNE_MODE_NONE : (s.leaf_pos < s.leaf_end ? YUNE_MODE_TRI : YUNE_MODE_INNER); }

YUNE_HD void ts_enter(TraceState& s, int ref)
{
    if (ref >= 0) { s.cur = ref; s.leaf_pos = s.leaf_end = 0; }
    else { const int x = ~ref; s.leaf_pos = x >> 4; s.leaf_end = (x >> 4) + (x & 15); }
}
YUNE_HD void ts_pop(TraceState& s, const int* stack)
{
    if (s.sp == 0) { s.done = true; s.leaf_pos = s.leaf_end = 0; return; }
    ts_enter(s, stack[--s.sp]);
}

template <bool COUNT>
YUNE_HD void ts_init(TraceState& s, V3 o, V3 d, float t_in, int root_ref, const float* root_lo, const float* root_hi, WorkCount* wc)
{
    s.r = make_ray(o, d);
    s.t_best = t_in; s.t_prune = t_in * 1.00001f;      // inf stays inf
    s.u = 0.0f; s.v = 0.0f; s.tri = -1; s.best_pos = -1;
    s.cur = 0; s.leaf_pos = s.leaf_end = 0; s.sp = 0; s.done = false;
    if (root_ref == YUNE_REF_EMPTY) { s.done = true; return; }
    float entry;
    if (COUNT) wc->box++;
    if (!box_hit(s.r, root_lo[0], root_hi[0], root_lo[1], root_hi[1], root_lo[2], root_hi[2], entry)) { s.done = true; return; }   // udpt.cl:295-296
    ts_enter(s, root_ref);
}

// One inner step.  ANY = shadow query (no near/far ordering needed, keeps the reference's child order).
template <class PairFetch, bool ANY, bool COUNT>
YUNE_HD void ts_inner_step(TraceState& s, int* stack, const PairFetch& fetch_pair, WorkCount* wc)
{
    F4 q0, q1, q2, q3;
    fetch_pair(s.cur, q0, q1, q2, q3);
    const int ref0 = YF_ASINT(q3.x), ref1 = YF_ASINT(q3.y);
    float e0, e1;
    if (COUNT) wc->box += (ref0 != YUNE_REF_EMPTY) + (ref1 != YUNE_REF_EMPTY);
    bool h0 = (ref0 != YUNE_REF_EMPTY) && box_hit(s.r, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, e0);
    bool h1 = (ref1 != YUNE_REF_EMPTY) && box_hit(s.r, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, e1);
    h0 = h0 && !(e0 > s.t_prune);
    h1 = h1 && !(e1 > s.t_prune);
    if (h0 && h1) {
        const bool swap = !ANY && (e1 < e0);
        stack[s.sp++] = swap ? ref0 : ref1;
        ts_enter(s, swap ? ref1 : ref0);
    } else if (h0) ts_enter(s, ref0);
    else if (h1) ts_enter(s, ref1);
    else ts_pop(s, stack);
}

// One triangle step.
template <class TriFetch, bool ANY, bool COUNT>
YUNE_HD void ts_tri_step(TraceState& s, const int* stack, const TriFetch& fetch_tri, WorkCount* wc)
{
    const int pos = s.leaf_pos++;
    F4 a, b, c;
    fetch_tri(pos, a, b, c);
    float t, u, v;
    if (COUNT) wc->tri++;
    if (tri_test(s.r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v)) {
        if (ANY) {
            if (t > 0.0f && t < s.t_best) { s.tri = 0; s.done = true; s.leaf_pos = s.leaf_end = 0; return; }   // udpt.cl:306-308
        } else if (t > 0.0f && (t < s.t_best || (t == s.t_best && s.best_pos >= 0 && YF_ASINT(b.w) < s.best_pos))) {
            // udpt.cl:373 't > 0 && t < ray->length', with the reference's first-come rule for exact ties
            s.t_best = t; s.u = u; s.v = v; s.tri = YF_ASINT(a.w); s.best_pos = YF_ASINT(b.w);
            s.t_prune = t * 1.00001f;
        }
    }
    if (s.leaf_pos >= s.leaf_end) ts_pop(s, stack);
}

// ---- accel 1: the same machine over our own tree ----
template <class PairFetch, bool ANY, bool COUNT>
YUNE_HD void ts_inner_step_own(TraceState& s, int* stack, const PairFetch& fetch_pair, WorkCount* wc)
{
    F4 q0, q1, q2, q3;
    fetch_pair(s.cur, q0, q1, q2, q3);
    const int ref0 = YF_ASINT(q3.x), ref1 = YF_ASINT(q3.y);
    float e0, e1;
    if (COUNT) wc->box += 2;
    const bool h0 = box_hit_own(s.r, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, s.t_prune, e0);
    const bool h1 = box_hit_own(s.r, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, s.t_prune, e1);
    if (h0 && h1) {
        const bool swap = !ANY && (e1 < e0);
        stack[s.sp++] = swap ? ref0 : ref1;
        ts_enter(s, swap ? ref1 : ref0);
    } else if (h0) ts_enter(s, ref0);
    else if (h1) ts_enter(s, ref1);
    else ts_pop(s, stack);
}
// fetch_leaf_box(id, lo, hi) returns the uploaded box of reference leaf `id`
// WT: perf-mode intersection (isect 1): raw vertices, tri_test_watertight, no leaf-box filter (that filter reproduces the
// reference's misses at grazing boxes, which is exactly what a watertight walk must not do).
template <class TriFetch, class LeafBoxFetch, bool ANY, bool COUNT, bool WT = false>
YUNE_HD void ts_tri_step_own(TraceState& s, const int* stack, const TriFetch& fetch_tri, const LeafBoxFetch& fetch_leaf_box, WorkCount* wc)
{
    const int pos = s.leaf_pos++;
    F4 a, b, c, lo, hi;
    fetch_tri(pos, a, b, c);
    float t, u, v, entry;
    bool inside;
    if (WT) {
        if (COUNT) wc->tri++;
        inside = tri_test_watertight(s.r.o, wt_setup(s.r.d), v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v);
    } else {
        fetch_leaf_box(YF_ASINT(c.w), lo, hi);
        if (COUNT) { wc->tri++; wc->box++; }
        // would the reference have reached this triangle?  <=> the box of its leaf passes the reference's predicate (udpt.cl:392-431)
        inside = box_hit(s.r, lo.x, hi.x, lo.y, hi.y, lo.z, hi.z, entry) && tri_test(s.r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v);
    }
    if (inside) {
        if (ANY) {
            if (t > 0.0f && t < s.t_best) { s.tri = 0; s.done = true; s.leaf_pos = s.leaf_end = 0; return; }
        } else if (t > 0.0f && (t < s.t_best || (t == s.t_best && s.best_pos >= 0 && YF_ASINT(b.w) < s.best_pos))) {
            s.t_best = t; s.u = u; s.v = v; s.tri = YF_ASINT(a.w); s.best_pos = YF_ASINT(b.w);
            s.t_prune = t * 1.00001f;
        }
    }
    if (s.leaf_pos >= s.leaf_end) ts_pop(s, stack);
}
template <bool COUNT>
YUNE_HD void ts_init_own(TraceState& s, V3 o, V3 d, float t_in, int root_ref, const float* root_lo, const float* root_hi, WorkCount* wc)
{
    s.r = make_ray(o, d);
    s.t_best = t_in; s.t_prune = t_in * 1.00001f;
    s.u = 0.0f; s.v = 0.0f; s.tri = -1; s.best_pos = -1;
    s.cur = 0; s.leaf_pos = s.leaf_end = 0; s.sp = 0; s.done = false;
    if (root_ref == YUNE_REF_EMPTY) { s.done = true; return; }
    float entry;
    if (COUNT) wc->box++;
    if (!box_hit_own(s.r, root_lo[0], root_hi[0], root_lo[1], root_hi[1], root_lo[2], root_hi[2], s.t_prune, entry)) { s.done = true; return; }
    ts_enter(s, root_ref);
}
template <class PairFetch, class TriFetch, class LeafBoxFetch, bool ANY, bool COUNT, bool WT = false>
YUNE_HD void trace_own(const PairFetch& fetch_pair, const TriFetch& fetch_tri, const LeafBoxFetch& fetch_leaf_box, int root_ref,
                       const float* root_lo, const float* root_hi, V3 o, V3 d, float t_in, HitRec& out, WorkCount* wc)
{
    TraceState s; int stack[YUNE_STACK_SIZE];
    ts_init_own<COUNT>(s, o, d, t_in, root_ref, root_lo, root_hi, wc);
    while (!s.done) {
        if (s.leaf_pos < s.leaf_end) ts_tri_step_own<TriFetch, LeafBoxFetch, ANY, COUNT, WT>(s, stack, fetch_tri, fetch_leaf_box, wc);
        else ts_inner_step_own<PairFetch, ANY, COUNT>(s, stack, fetch_pair, wc);
    }
    out.t = s.t_best; out.u = s.u; out.v = s.v; out.tri = s.tri;
}

// ---- accel 2: the own tree with up to four children per node (trav_layout.h `quads`) ----
// One step fetches a 112-byte record, tests the (up to) four child boxes, enters the nearest hit child and pushes the others
// far to near.  Leaves are the same triangle ranges as in accel 1, tested by ts_tri_step_own.
// Ascending sort of four (entry, ref) pairs with the 5-comparator network (0,1)(2,3)(0,2)(1,3)(1,2): branch-free selects, the
// form the device step will use.  Misses carry entry = +inf and sink to the end; the order among equal entries is irrelevant
// (hit records do not depend on the visiting order, tests/test_traversal_hostcheck.py).
YUNE_HD void cmp_swap(float& ea, int& ra, float& eb, int& rb)
{
    const bool sw = eb < ea;
    const float e_lo = sw ? eb : ea, e_hi = sw ? ea : eb; const int r_lo = sw ? rb : ra, r_hi = sw ? ra : rb;
    ea = e_lo; eb = e_hi; ra = r_lo; rb = r_hi;
}
YUNE_HD void sort4(float* e, int* r)
{
    cmp_swap(e[0], r[0], e[1], r[1]); cmp_swap(e[2], r[2], e[3], r[3]);
    cmp_swap(e[0], r[0], e[2], r[2]); cmp_swap(e[1], r[1], e[3], r[3]);
    cmp_swap(e[1], r[1], e[2], r[2]);
}
template <class QuadFetch, bool ANY, bool COUNT>
YUNE_HD void ts_inner_step_wide(TraceState& s, int* stack, const QuadFetch& fetch_quad, WorkCount* wc)
{
    F4 q[7];
    fetch_quad(s.cur, q);
    const float* f = &q[0].x;
    int ref[4] = {YF_ASINT(q[6].x), YF_ASINT(q[6].y), YF_ASINT(q[6].z), YF_ASINT(q[6].w)};
    float e[4]; int n = 0;
    for (int i = 0; i < 4; i++) {
        float ent;
        const bool used = ref[i] != YUNE_REF_EMPTY;
        if (COUNT && used) wc->box++;
        const bool hit = used && box_hit_own(s.r, f[i], f[4 + i], f[8 + i], f[12 + i], f[16 + i], f[20 + i], s.t_prune, ent);
        e[i] = hit ? fminf(ent, 3.0e38f) : INFINITY;      // a hit's key stays finite: +inf marks a miss
        n += hit;
    }
    if (ANY) {       // shadow query: keep the record's order, just move the misses behind the hits (stable compaction)
        int k = 0; float e2[4]; int r2[4];
        for (int i = 0; i < 4; i++) if (e[i] < INFINITY) { e2[k] = e[i]; r2[k] = ref[i]; k++; }
        for (int i = 0; i < k; i++) { e[i] = e2[i]; ref[i] = r2[i]; }
    } else sort4(e, ref);
    if (n == 0) { ts_pop(s, stack); return; }
    for (int k = 3; k >= 1; k--) if (k < n) stack[s.sp++] = ref[k];      // far to near: the nearest pushed child is popped first
    ts_enter(s, ref[0]);
}
// The wide step in the form the DEVICE scheduling needs (postponed leaves, sentinel stack: kernels.cu's Lane and the simulated
// lanes of tests/hostcheck share the fields used here).  `q` = the seven 16-byte vectors of the record, already fetched;
// box_test(lox, hix, loy, hiy, loz, hiz, entry) = the conservative slab test against the lane's ray and pruning distance.
// Every hit child is pushed far to near, then the walk takes its next place: a leaf on top is PARKED when nothing is parked
// (and the walk moves on to the place after it), otherwise the lane stands on it until its parked triangles have been tested.
// stack[0 .. stack_base) hold the sentinel `ref_done` (< 0, an empty triangle range), so popping an empty stack ends the walk.
template <class LaneT, class BoxTest, bool ANY>
YUNE_HD void lane_wide_step(LaneT& L, int* stack, const F4* q, const BoxTest& box_test, int stack_base)
{
    const float* f = &q[0].x;
    int ref[4] = {YF_ASINT(q[6].x), YF_ASINT(q[6].y), YF_ASINT(q[6].z), YF_ASINT(q[6].w)};
    float e[4]; int n = 0;
    for (int i = 0; i < 4; i++) {
        float ent = 0.0f;
        const bool hit = ref[i] != YUNE_REF_EMPTY && box_test(f[i], f[4 + i], f[8 + i], f[12 + i], f[16 + i], f[20 + i], ent);
        e[i] = hit ? fminf(ent, 3.0e38f) : INFINITY;
        n += hit;
    }
    if (ANY) {
        int k = 0; int r2[4];
        for (int i = 0; i < 4; i++) if (e[i] < INFINITY) r2[k++] = ref[i];
        for (int i = 0; i < k; i++) ref[i] = r2[i];
    } else sort4(e, ref);
    for (int k = 3; k >= 0; k--) if (k < n) stack[L.sp++] = ref[k];
    int c0 = stack[L.sp - 1];
    L.sp = L.sp - 1 > stack_base ? L.sp - 1 : stack_base;
    if (c0 < 0 && !(L.pend_pos < L.pend_end)) {
        const int x = ~c0; L.pend_pos = x >> 4; L.pend_end = (x >> 4) + (x & 15);
        c0 = stack[L.sp - 1];
        L.sp = L.sp - 1 > stack_base ? L.sp - 1 : stack_base;
    }
    L.cur = c0;
}

// The wide step as the device kernel runs it (accel 2), in two halves so that host and device share everything but the fetch:
//   keys    k_i = int bits of child i's entry distance (>= 0, so the integer order is the float order), YUNE_KEY_MISS for a miss;
//           the device computes them from near / far planes picked by the ray's direction signs (one FFMA per plane, no
//           per-axis min / max: for lo <= hi and a finite 1/d the fused product-difference is monotone, so near <= far and the
//           values are those of box_hit_own), the host calls box_hit_own;
//   finish  sorts the four (key, ref) pairs with the 5-comparator network and updates the walk with selects only: the two
//           places the walk visits next are c0 / c1 (a hit, or the entries read from the top of the stack up front), at most
//           three refs are stored, each at a position that only depends on the hit count.  Same rules as lane_wide_step.
#define YUNE_KEY_MISS 0x7f800000
YUNE_HD void key_cswap(int& ka, int& ra, int& kb, int& rb)
{
    const bool sw = kb < ka;
    const int k_lo = sw ? kb : ka, k_hi = sw ? ka : kb, r_lo = sw ? rb : ra, r_hi = sw ? ra : rb;
    ka = k_lo; kb = k_hi; ra = r_lo; rb = r_hi;
}
template <class LaneT>
YUNE_HD void lane_wide_finish(LaneT& L, int* stack, int top1, int top2, int k0, int k1, int k2, int k3, int r0, int r1, int r2, int r3)
{
    key_cswap(k0, r0, k1, r1); key_cswap(k2, r2, k3, r3);
    key_cswap(k0, r0, k2, r2); key_cswap(k1, r1, k3, r3);
    key_cswap(k1, r1, k2, r2);
    const bool h1 = k0 < YUNE_KEY_MISS, h2 = k1 < YUNE_KEY_MISS, h3 = k2 < YUNE_KEY_MISS, h4 = k3 < YUNE_KEY_MISS;   // at least 1 / 2 / 3 / 4 hits
    const int n = (h1 ? 1 : 0) + (h2 ? 1 : 0) + (h3 ? 1 : 0) + (h4 ? 1 : 0);
    const int c0 = h1 ? r0 : top1;
    const int c1 = h2 ? r1 : (h1 ? top1 : top2);
    const bool park = c0 < 0 && !(L.pend_pos < L.pend_end);      // c0 is a leaf (or the sentinel: an empty leaf) and nothing is parked
    const int x = ~c0;
    int* s = stack + L.sp;
    if (h4) s[0] = r3;                                            // far to near: r3 (only with four hits) lands lowest
    if (h3) s[n - 3] = r2;
    if (h2 && !park) s[n - 2] = r1;                               // with a parked c0 the walk continues at r1 itself
    L.sp += n - 1 - (park ? 1 : 0);
    L.cur = park ? c1 : c0;
    L.pend_pos = park ? (x >> 4) : L.pend_pos;
    L.pend_end = park ? (x >> 4) + (x & 15) : L.pend_end;
}

template <class QuadFetch, class TriFetch, class LeafBoxFetch, bool ANY, bool COUNT>
YUNE_HD void trace_wide(const QuadFetch& fetch_quad, const TriFetch& fetch_tri, const LeafBoxFetch& fetch_leaf_box, int root_ref,
                        const float* root_lo, const float* root_hi, V3 o, V3 d, float t_in, HitRec& out, WorkCount* wc)
{
    TraceState s; int stack[YUNE_STACK_SIZE];
    ts_init_own<COUNT>(s, o, d, t_in, root_ref, root_lo, root_hi, wc);
    while (!s.done) {
        if (s.leaf_pos < s.leaf_end) ts_tri_step_own<TriFetch, LeafBoxFetch, ANY, COUNT>(s, stack, fetch_tri, fetch_leaf_box, wc);
        else ts_inner_step_wide<QuadFetch, ANY, COUNT>(s, stack, fetch_quad, wc);
    }
    out.t = s.t_best; out.u = s.u; out.v = s.v; out.tri = s.tri;
}

// ---- whole-ray wrappers (host check, hooks): drive the machine until done ----
template <class PairFetch, class TriFetch, bool COUNT>
YUNE_HD void closest_hit(const PairFetch& fetch_pair, const TriFetch& fetch_tri, int root_ref,
                         const float* root_lo, const float* root_hi, V3 o, V3 d, float t_in, HitRec& out, WorkCount* wc)
{
    TraceState s; int stack[YUNE_STACK_SIZE];
    ts_init<COUNT>(s, o, d, t_in, root_ref, root_lo, root_hi, wc);
    while (!s.done) {
        if (s.leaf_pos < s.leaf_end) ts_tri_step<TriFetch, false, COUNT>(s, stack, fetch_tri, wc);
        else ts_inner_step<PairFetch, false, COUNT>(s, stack, fetch_pair, wc);
    }
    out.t = s.t_best; out.u = s.u; out.v = s.v; out.tri = s.tri;
}
template <class PairFetch, class TriFetch, bool COUNT>
YUNE_HD bool any_hit(const PairFetch& fetch_pair, const TriFetch& fetch_tri, int root_ref,
                     const float* root_lo, const float* root_hi, V3 o, V3 d, float t_in, WorkCount* wc)
{
    TraceState s; int stack[YUNE_STACK_SIZE];
    ts_init<COUNT>(s, o, d, t_in, root_ref, root_lo, root_hi, wc);
    while (!s.done) {
        if (s.leaf_pos < s.leaf_end) ts_tri_step<TriFetch, true, COUNT>(s, stack, fetch_tri, wc);
        else ts_inner_step<PairFetch, true, COUNT>(s, stack, fetch_pair, wc);
    }
    return s.tri >= 0;
}

} // namespace yune
#endif
