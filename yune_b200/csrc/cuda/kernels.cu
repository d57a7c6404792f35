// kernels.cu -- the sm_100a wavefront kernels of the path-tracing hot path.
//
//   k_trace        persistent-thread BVH walk: drains the shadow queue (any-hit) then the extension queue
//                  (closest hit).  Replaces traceRay/traverseBVH/rayTriangleIntersection/rayAabbIntersection
//                  (kernels/legacy/udpt.cl:240-431).
//   k_shade_dense  "logic + material" stage as a persistent kernel that sorts path slots inside each block (classify /
//                  diffuse / specular / regenerate rounds): collects last iteration's NEE answers, processes the hit
//                  (udpt.cl:433-533 `shading`), creates the NEE / MIS shadow rays (:535-609 `evaluateDirectLighting`),
//                  samples the next direction, regenerates finished slots with fresh camera rays (:158-238
//                  `pathtracer`/`createRay`) and accumulates finished samples (:193-210).  Queue space: one atomic per
//                  counter per round.
//   k_tonemap      kernels/post-proc/tonemap.cl:14-47 on the mean image.
//   k_hook_*       parity hooks: primary rays / arbitrary rays through the same k_trace.
//
// Tensor cores are not used: no stage is a dense contraction (every ray gathers its own nodes/triangles).
#include <cstdlib>
#include "kernel_common.cuh"
#include "trace_core.h"

namespace yune {

// ---- device-side walk: the same state machine as trace_core.h (tests/hostcheck validates that one against the oracle;
// tests/test_gpu_parity.py validates this one), written branch-free so that a warp's lanes stay converged. ----
//
// Per-lane state:
//   cur            what the walk visits next: >= 0 pair index (an INNER step can run); < 0 a leaf reference that waits because
//                  the lane already holds postponed triangles; YUNE_REF_DONE (-1, "a leaf with no triangles") = stack exhausted.
//   [pend_pos, pend_end)  POSTPONED triangles: a leaf the walk reached is not tested on the spot.  Its range is parked here and
//                  the walk goes on with the next stack entry, so a lane keeps taking part in INNER steps until it reaches a
//                  second leaf.  TRI steps run when enough lanes hold parked triangles.  (Aila & Laine's speculative traversal;
//                  the ncu capture of the phase-alternating version showed 18 of ~29 ray-holding lanes active in the inner body:
//                  the others sat on a leaf waiting for the phase to turn.)  Testing a leaf later than the reference order
//                  cannot change the result: the closest hit is order-independent (exact ties in t go by reference rank) and a
//                  late t_best only costs a few box tests that earlier pruning would have saved.
//   a ray is finished when cur == DONE and nothing is parked.
#define YUNE_REF_DONE (-1)

struct Lane {
    V3 o, d, inv, oi;       // oi = o * inv: fused slab test of our own tree (ACCEL 1)
    float t_best, t_prune, u, v;
    int tri, best_pos, cur, pend_pos, pend_end, sp;
    int sgn;                // ACCEL bit 2: byte k = 16 when d[k] < 0 (which of a wide record's lo / hi vectors holds the NEAR planes)
    bool guard;
};

// The stack lives in local memory above TWO sentinel entries (stack[0] = stack[1] = DONE, sp starts at 2).  Every step loads the
// top entries it MIGHT need up front, unconditionally and next to the node / triangle fetch; what the walk does next is then
// a chain of selects.  (ncu, per-SASS-line view of the previous version: `if (miss) next = stack[--sp]` and the parking code
// compiled to branches that 3 of 32 lanes took -- 15 issue slots per inner step for one pop.)
#define YUNE_STACK_BASE 2

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds128v(uint32_t addr)       // ordered against the stores above (a lane re-reads what it wrote)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int2 lds64(uint32_t addr)
{
    int2 v;
    asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ bool box_fast(V3 o, V3 inv, float lox, float hix, float loy, float hiy, float loz, float hiz, float& entry)
{
    const float ax = YF_MUL(YF_SUB(lox, o.x), inv.x), bx = YF_MUL(YF_SUB(hix, o.x), inv.x);
    const float ay = YF_MUL(YF_SUB(loy, o.y), inv.y), by = YF_MUL(YF_SUB(hiy, o.y), inv.y);
    const float az = YF_MUL(YF_SUB(loz, o.z), inv.z), bz = YF_MUL(YF_SUB(hiz, o.z), inv.z);
    const float t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float t_max = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    entry = fmaxf(t_min, 0.0f);
    return t_max > entry;
}
__device__ __noinline__ float box_guarded(V3 o, V3 inv, float lox, float hix, float loy, float hiy, float loz, float hiz)
{
    float t_min = -INFINITY, t_max = INFINITY;       // the reference's per-axis NaN guards (udpt.cl:400-416), rare path
    slab_guarded(lox, hix, o.x, inv.x, t_min, t_max);
    slab_guarded(loy, hiy, o.y, inv.y, t_min, t_max);
    slab_guarded(loz, hiz, o.z, inv.z, t_min, t_max);
    const float entry = fmaxf(t_min, 0.0f);
    return (t_max > entry) ? entry : -1.0f;          // entry >= 0 on a hit, -1 on a miss
}

// Conservative slab test for OUR boxes (trace_core.h: box_hit_own): one FFMA per plane, pruning distance folded in.
__device__ __forceinline__ bool box_own(const Lane& L, float lox, float hix, float loy, float hiy, float loz, float hiz, float& entry)
{
    const float ax = __fmaf_rn(lox, L.inv.x, -L.oi.x), bx = __fmaf_rn(hix, L.inv.x, -L.oi.x);
    const float ay = __fmaf_rn(loy, L.inv.y, -L.oi.y), by = __fmaf_rn(hiy, L.inv.y, -L.oi.y);
    const float az = __fmaf_rn(loz, L.inv.z, -L.oi.z), bz = __fmaf_rn(hiz, L.inv.z, -L.oi.z);
    const float t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    const float t_max = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), L.t_prune));
    entry = t_min;
    return YF_MUL(t_max, 1.0000021f) >= t_min;        // (t_min >= 0) at least as wide as t_max (1 + 1e-6) >= t_min (1 - 1e-6)
}

// ACCEL bit 2 keeps the part of a ray that node steps never touch -- origin, direction, the barycentrics of the best hit -- in
// SHARED memory next to the staged records (two float4 per thread: (o.xyz, u) at s_ray, (d.xyz, v) at s_ray + 16 * blockDim.x;
// a warp's LDS.128 / STS.128 are conflict-free).  The wide step holds 24 plane values + 4 refs + 4 keys at once; with the whole
// ray in registers ptxas spilled the walk's own state (sp, the parked range, the loop counter) around every step.
template <bool ANY, bool COUNT, int ACCEL>
__device__ __forceinline__ void lane_init(Lane& L, const DevScene& sc, float4 o, float4 d, WorkCount& wc, const uint32_t s_ray)
{
    if (ACCEL & 4) { sts128(s_ray, make_float4(o.x, o.y, o.z, 0.0f)); sts128(s_ray + 16u * blockDim.x, make_float4(d.x, d.y, d.z, 0.0f)); }
    else { L.o = xyz(o); L.d = xyz(d); L.u = 0.0f; L.v = 0.0f; }
    L.inv = v3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z));       // correctly rounded 1/x == '1 / ray->dir' (udpt.cl:395)
    L.guard = !(fabsf(L.inv.x) < INFINITY && fabsf(L.inv.y) < INFINITY && fabsf(L.inv.z) < INFINITY);
    L.t_best = o.w; L.t_prune = o.w * 1.00001f;
    L.tri = -1; L.best_pos = -1; L.sp = YUNE_STACK_BASE;
    L.oi = v3(YF_MUL(o.x, L.inv.x), YF_MUL(o.y, L.inv.y), YF_MUL(o.z, L.inv.z));
    if (ACCEL & 1)      // o * (1/d) can overflow although 1/d is finite: lo * inv - oi would then be -inf for both planes (box lost)
        L.guard = L.guard || !(fabsf(L.oi.x) < INFINITY && fabsf(L.oi.y) < INFINITY && fabsf(L.oi.z) < INFINITY);
    if (ACCEL & 4) L.sgn = (L.inv.x < 0.0f ? 16 : 0) | (L.inv.y < 0.0f ? 16 << 8 : 0) | (L.inv.z < 0.0f ? 16 << 16 : 0);
    if (ACCEL & 8) { const WtRay w = wt_setup(xyz(d)); L.d = w.S; L.sgn = w.kz; }      // isect 1: node steps never read the direction itself
    bool hit = sc.root_ref != YUNE_REF_EMPTY;
    if ((ACCEL & 1) == 0 && hit) {
        // the reference tests the root box first (udpt.cl:295-296).  With ACCEL 1 the boxes of our own tree only prune -- what the
        // reference would have reached is decided per triangle by the leaf-box filter -- so the root test is skipped there.
        float entry;
        if (COUNT) wc.box++;
        if (L.guard) { entry = box_guarded(xyz(o), L.inv, sc.root_lo[0], sc.root_hi[0], sc.root_lo[1], sc.root_hi[1], sc.root_lo[2], sc.root_hi[2]); hit = entry >= 0.0f; }
        else hit = box_fast(xyz(o), L.inv, sc.root_lo[0], sc.root_hi[0], sc.root_lo[1], sc.root_hi[1], sc.root_lo[2], sc.root_hi[2], entry);
    }
    const int ref = hit ? sc.root_ref : YUNE_REF_DONE;
    const int x = ~ref;                                                 // a root that is a leaf is parked at once
    L.cur = ref >= 0 ? ref : YUNE_REF_DONE;
    L.pend_pos = ref >= 0 ? 0 : (x >> 4);
    L.pend_end = ref >= 0 ? 0 : (x >> 4) + (x & 15);
}

template <bool ANY, bool COUNT, int ACCEL>
__device__ __forceinline__ void lane_inner_step(Lane& L, int* stack, const DevScene& sc, const uint32_t s_box, const uint32_t s_ref, WorkCount& wc)
{
    const int top1 = stack[L.sp - 1], top2 = stack[L.sp - 2];       // cur >= 0 implies sp >= YUNE_STACK_BASE
    float4 q0, q1, q2, q3;
    int ref0, ref1;
    if ((ACCEL & 2) || L.cur < sc.n_smem_pairs) {       // ACCEL bit 1: every pair record is staged, no global path compiled in
        // Shared-memory copy: boxes at a 48-byte stride, child refs in a separate int2 array.  With the 64-byte records of
        // the global layout every lane's q_k would fall into the same two 16-byte bank columns (64 * idx mod 128) and an
        // LDS.128 of 18 scattered lanes took ~15 wavefronts (ncu); 48 * idx mod 128 visits all eight columns.
        const uint32_t p = s_box + 48u * (uint32_t)L.cur;
        q0 = lds128(p); q1 = lds128(p + 16); q2 = lds128(p + 32);
        const int2 rr = lds64(s_ref + 8u * (uint32_t)L.cur); ref0 = rr.x; ref1 = rr.y;
    } else {
        const float4* p = sc.pairs + 4 * (size_t)L.cur; q0 = __ldg(p); q1 = __ldg(p + 1); q2 = __ldg(p + 2); q3 = __ldg(p + 3);
        ref0 = __float_as_int(q3.x); ref1 = __float_as_int(q3.y);
    }
    float e0, e1; bool h0, h1;
    if ((ACCEL & 1) == 1 && !L.guard) {
        h0 = box_own(L, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, e0);
        h1 = box_own(L, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, e1);
    } else if (!L.guard) {
        h0 = box_fast(L.o, L.inv, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, e0);
        h1 = box_fast(L.o, L.inv, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, e1);
    } else {
        e0 = box_guarded(L.o, L.inv, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y); h0 = e0 >= 0.0f;
        e1 = box_guarded(L.o, L.inv, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w); h1 = e1 >= 0.0f;
    }
    if (COUNT) wc.box += (ref0 != YUNE_REF_EMPTY) + (ref1 != YUNE_REF_EMPTY);
    if ((ACCEL & 1) == 0 || L.guard) {        // (ACCEL 1 folds the pruning distance into box_own; its guarded fallback does not)
        h0 = h0 && (ref0 != YUNE_REF_EMPTY) && !(e0 > L.t_prune);
        h1 = h1 && (ref1 != YUNE_REF_EMPTY) && !(e1 > L.t_prune);
    }
    const bool both = h0 && h1, any = h0 || h1;
    const bool swap = !ANY && both && (e1 < e0);
    const int near = (h0 && !swap) ? ref0 : ref1;                   // meaningful when any
    const int far  = swap ? ref0 : ref1;                            // meaningful when both
    // the next two places the walk would visit, in order
    const int c0 = any ? near : top1;
    const int c1 = both ? far : (any ? top1 : top2);
    const bool park = c0 < 0 && !(L.pend_pos < L.pend_end);         // c0 is a leaf (or DONE = empty leaf) and nothing is parked
    const int x = ~c0;
    if (both && !park) stack[L.sp] = far;
    L.sp += (both ? 1 : (any ? 0 : -1)) - (park ? 1 : 0);
    L.cur = park ? c1 : c0;
    L.pend_pos = park ? (x >> 4) : L.pend_pos;
    L.pend_end = park ? (x >> 4) + (x & 15) : L.pend_end;
}

// ACCEL bit 2 (value 4; accel option 2): the own tree collapsed into 4-wide records (trav_layout.h `quads`: 7 float4 per record
// = lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4] refs[4]; sc.pairs points at them, sc.n_inner counts them, the first
// sc.n_smem_pairs are staged at a 112-byte stride).  Half the steps of the binary walk for the same number of box tests.
// The NEAR planes of an axis are the lo vector when d > 0 and the hi vector when d < 0, so the fetch address picks them
// (Lane::sgn) and a box costs 6 FFMA + 2 three-input min / max instead of 6 FFMA + 6 min / max + 2; an unused slot (lo = 3e38,
// hi = -3e38) misses by itself.  What the walk does next is trace_core.h's lane_wide_finish (selects only).
struct WideKeys { int k0, k1, k2, k3; };
__device__ __noinline__ WideKeys wide_keys_guarded(V3 o, V3 inv, float t_prune, const float4* q)     // rare: some 1/d (or o/d) is not finite
{
    const float4 lx = q[0], hx = q[1], ly = q[2], hy = q[3], lz = q[4], hz = q[5];
    const int4 rf = *reinterpret_cast<const int4*>(q + 6);
    const float e0 = box_guarded(o, inv, lx.x, hx.x, ly.x, hy.x, lz.x, hz.x), e1 = box_guarded(o, inv, lx.y, hx.y, ly.y, hy.y, lz.y, hz.y);
    const float e2 = box_guarded(o, inv, lx.z, hx.z, ly.z, hy.z, lz.z, hz.z), e3 = box_guarded(o, inv, lx.w, hx.w, ly.w, hy.w, lz.w, hz.w);
    WideKeys K;
    K.k0 = (rf.x != YUNE_REF_EMPTY && e0 >= 0.0f && !(e0 > t_prune)) ? __float_as_int(e0) : YUNE_KEY_MISS;
    K.k1 = (rf.y != YUNE_REF_EMPTY && e1 >= 0.0f && !(e1 > t_prune)) ? __float_as_int(e1) : YUNE_KEY_MISS;
    K.k2 = (rf.z != YUNE_REF_EMPTY && e2 >= 0.0f && !(e2 > t_prune)) ? __float_as_int(e2) : YUNE_KEY_MISS;
    K.k3 = (rf.w != YUNE_REF_EMPTY && e3 >= 0.0f && !(e3 > t_prune)) ? __float_as_int(e3) : YUNE_KEY_MISS;
    return K;
}
// The fetches of the wide step are `volatile` asm so that ptxas keeps them in this order: x and y planes first, their partial
// intervals folded into 8 registers, then the z planes, then the refs and the two stack tops.  Left to itself it hoists all seven
// LDS.128 (28 registers) above the arithmetic and spills the walk's own state around every step (64 registers per thread).
__device__ __forceinline__ float4 ld_plane(const bool smem, const uint32_t sa, const char* ga, const int off)
{
    float4 v;
    if (smem) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa + off));
    else asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ga + off));
    return v;
}
template <bool ANY, bool COUNT, bool ALL_STAGED>
__device__ __forceinline__ void lane_inner_step_wide(Lane& L, int* stack, const DevScene& sc, const uint32_t s_quad, const uint32_t s_ray, WorkCount& wc)
{
    if (COUNT) wc.box += 4;
    const bool smem = ALL_STAGED || L.cur < sc.n_smem_pairs;
    const uint32_t sa = s_quad + 112u * (uint32_t)L.cur;
    const char* ga = ALL_STAGED ? nullptr : reinterpret_cast<const char*>(sc.pairs + 7 * (size_t)L.cur);
    WideKeys K;
    if (!L.guard) {
        const int sx = L.sgn & 0xff, sy = (L.sgn >> 8) & 0xff, sz = L.sgn >> 16;
        const float4 nx = ld_plane(smem, sa, ga, sx), fx = ld_plane(smem, sa, ga, 16 - sx);
        const float4 ny = ld_plane(smem, sa, ga, 32 + sy), fy = ld_plane(smem, sa, ga, 48 - sy);
        // per child: t_min = max(near planes, 0), t_max = min(far planes, pruning distance); one FFMA per plane
        float lo0 = fmaxf(__fmaf_rn(nx.x, L.inv.x, -L.oi.x), __fmaf_rn(ny.x, L.inv.y, -L.oi.y)), hi0 = fminf(__fmaf_rn(fx.x, L.inv.x, -L.oi.x), __fmaf_rn(fy.x, L.inv.y, -L.oi.y));
        float lo1 = fmaxf(__fmaf_rn(nx.y, L.inv.x, -L.oi.x), __fmaf_rn(ny.y, L.inv.y, -L.oi.y)), hi1 = fminf(__fmaf_rn(fx.y, L.inv.x, -L.oi.x), __fmaf_rn(fy.y, L.inv.y, -L.oi.y));
        float lo2 = fmaxf(__fmaf_rn(nx.z, L.inv.x, -L.oi.x), __fmaf_rn(ny.z, L.inv.y, -L.oi.y)), hi2 = fminf(__fmaf_rn(fx.z, L.inv.x, -L.oi.x), __fmaf_rn(fy.z, L.inv.y, -L.oi.y));
        float lo3 = fmaxf(__fmaf_rn(nx.w, L.inv.x, -L.oi.x), __fmaf_rn(ny.w, L.inv.y, -L.oi.y)), hi3 = fminf(__fmaf_rn(fx.w, L.inv.x, -L.oi.x), __fmaf_rn(fy.w, L.inv.y, -L.oi.y));
        const float4 nz = ld_plane(smem, sa, ga, 64 + sz), fz = ld_plane(smem, sa, ga, 80 - sz);
        lo0 = fmaxf(lo0, fmaxf(__fmaf_rn(nz.x, L.inv.z, -L.oi.z), 0.0f)); hi0 = fminf(hi0, fminf(__fmaf_rn(fz.x, L.inv.z, -L.oi.z), L.t_prune));
        lo1 = fmaxf(lo1, fmaxf(__fmaf_rn(nz.y, L.inv.z, -L.oi.z), 0.0f)); hi1 = fminf(hi1, fminf(__fmaf_rn(fz.y, L.inv.z, -L.oi.z), L.t_prune));
        lo2 = fmaxf(lo2, fmaxf(__fmaf_rn(nz.z, L.inv.z, -L.oi.z), 0.0f)); hi2 = fminf(hi2, fminf(__fmaf_rn(fz.z, L.inv.z, -L.oi.z), L.t_prune));
        lo3 = fmaxf(lo3, fmaxf(__fmaf_rn(nz.w, L.inv.z, -L.oi.z), 0.0f)); hi3 = fminf(hi3, fminf(__fmaf_rn(fz.w, L.inv.z, -L.oi.z), L.t_prune));
        // box_own's predicate and entry distance: (t_min >= 0) at least as wide as t_max (1 + 1e-6) >= t_min (1 - 1e-6)
        K.k0 = YF_MUL(hi0, 1.0000021f) >= lo0 ? __float_as_int(lo0) : YUNE_KEY_MISS;
        K.k1 = YF_MUL(hi1, 1.0000021f) >= lo1 ? __float_as_int(lo1) : YUNE_KEY_MISS;
        K.k2 = YF_MUL(hi2, 1.0000021f) >= lo2 ? __float_as_int(lo2) : YUNE_KEY_MISS;
        K.k3 = YF_MUL(hi3, 1.0000021f) >= lo3 ? __float_as_int(lo3) : YUNE_KEY_MISS;
    } else {
        const float4* q = smem ? reinterpret_cast<const float4*>(__cvta_shared_to_generic(sa)) : reinterpret_cast<const float4*>(ga);
        K = wide_keys_guarded(xyz(lds128v(s_ray)), L.inv, L.t_prune, q);
    }
    const float4 r4 = ld_plane(smem, sa, ga, 96);
    const int top1 = stack[L.sp - 1], top2 = stack[L.sp - 2];       // cur >= 0 implies sp >= YUNE_STACK_BASE
    lane_wide_finish(L, stack, top1, top2, K.k0, K.k1, K.k2, K.k3, __float_as_int(r4.x), __float_as_int(r4.y), __float_as_int(r4.z), __float_as_int(r4.w));
}

template <bool ANY, bool COUNT, int ACCEL>
__device__ __forceinline__ void lane_node_step(Lane& L, int* stack, const DevScene& sc, const uint32_t s_box, const uint32_t s_ref, const uint32_t s_ray, WorkCount& wc)
{
    if (ACCEL & 4) lane_inner_step_wide<ANY, COUNT, (ACCEL & 2) != 0>(L, stack, sc, s_box, s_ray, wc);
    else lane_inner_step<ANY, COUNT, ACCEL>(L, stack, sc, s_box, s_ref, wc);
}

template <bool ANY, bool COUNT, int ACCEL>
__device__ __forceinline__ void lane_tri_step(Lane& L, const int* stack, const DevScene& sc, const uint32_t s_ray, WorkCount& wc)
{
    const int top1 = stack[max(L.sp - 1, 0)];
    V3 ro, rd;
    if (ACCEL & 4) { ro = xyz(lds128v(s_ray)); rd = xyz(lds128v(s_ray + 16u * blockDim.x)); }
    else { ro = L.o; rd = L.d; }
    const int pos = L.pend_pos++;
    const float4* p = sc.tris + 3 * (size_t)pos;
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    if (COUNT) wc.tri++;
    // rayTriangleIntersection (udpt.cl:326-373), evaluated without early exits: the rejected lanes would idle anyway, and
    // NaNs (det == 0) fall through the comparisons exactly as in the sequential form.
    float t, u, v; bool inside;
    if (ACCEL & 8) {
        // isect 1 (perf mode): watertight test on the raw vertices (trace_core.h), no leaf-box filter
        WtRay w; w.S = rd; w.kz = L.sgn;
        inside = tri_test_watertight(ro, w, xyz(a), xyz(b), xyz(c), t, u, v);
    } else {
        const V3 e1 = xyz(b), e2 = xyz(c);
        const V3 pvec = vcross(rd, e2);
        const float det = vdot(e1, pvec);
        const float inv_det = __frcp_rn(det);
        const V3 dist = vsub(ro, xyz(a));
        u = YF_MUL(vdot(pvec, dist), inv_det);
        const V3 qvec = vcross(dist, e1);
        v = YF_MUL(vdot(qvec, rd), inv_det);
        t = YF_MUL(vdot(e2, qvec), inv_det);
        inside = !(u < 0.0f || u > 1.0f) && !(v < 0.0f || YF_ADD(u, v) > 1.0f);
    }
    if ((ACCEL & 9) == 1) {
        // Would the reference have reached this triangle?  <=> the uploaded box of its reference leaf passes the reference's
        // own predicate (its ancestors' boxes contain it exactly, so they pass too).  Geometrically a ray that hits the
        // triangle always crosses that box, so the test can only ever reject in last-bit grazing cases; it is evaluated just
        // for the lanes that are about to accept a hit, which keeps its two loads off the L1 pipe almost always.
        const bool candidate = inside && t > 0.0f && !(t > L.t_best);
        if (candidate) {
            const float4* lb = sc.leaf_boxes + 2 * (size_t)__float_as_int(c.w);
            const float4 lo = __ldg(lb), hi = __ldg(lb + 1);
            float entry; bool reach;
            if (COUNT) wc.box++;
            if (!L.guard) reach = box_fast(ro, L.inv, lo.x, hi.x, lo.y, hi.y, lo.z, hi.z, entry);
            else reach = box_guarded(ro, L.inv, lo.x, hi.x, lo.y, hi.y, lo.z, hi.z) >= 0.0f;
            inside = reach;
        }
    }
    bool stop = false;
    if (ANY) {
        stop = inside && t > 0.0f && t < L.t_best;                                  // udpt.cl:306-308
        L.tri = stop ? 0 : L.tri;
    } else {
        const int rank = __float_as_int(b.w);
        const bool accept = inside && t > 0.0f && (t < L.t_best || (t == L.t_best && L.best_pos >= 0 && rank < L.best_pos));
        L.t_best = accept ? t : L.t_best;
        if (ACCEL & 4) { if (accept) { sts32(s_ray + 12u, u); sts32(s_ray + 16u * blockDim.x + 12u, v); } }
        else { L.u = accept ? u : L.u; L.v = accept ? v : L.v; }
        L.tri = accept ? __float_as_int(a.w) : L.tri; L.best_pos = accept ? rank : L.best_pos;
        L.t_prune = accept ? t * 1.00001f : L.t_prune;
    }
    // parked range used up and the walk stands on a leaf: park that one and move on
    const bool park = !(L.pend_pos < L.pend_end) && L.cur < 0;
    const int x = ~L.cur;
    L.pend_pos = park ? (x >> 4) : L.pend_pos;
    L.pend_end = park ? (x >> 4) + (x & 15) : L.pend_end;
    L.sp = park ? max(L.sp - 1, 0) : L.sp;
    L.cur = park ? top1 : L.cur;
    if (ANY) { L.cur = stop ? YUNE_REF_DONE : L.cur; L.pend_end = stop ? L.pend_pos : L.pend_end; }
}

// Drains one ray queue with a persistent warp.  Lanes hold one ray each.  Every pass of the inner loop executes ONE operation
// for all lanes that can take it: a TRI step when at least `tri_min` lanes hold parked triangles (or more lanes than could do
// an INNER step), otherwise an INNER step -- followed at once by up to `inner_chain` more while at least `inner_min` lanes still can.
// Lanes that finish are refilled from the queue head (one atomicAdd per chunk of up to YUNE_FETCH_CHUNK rays, ballot/popc ranks) once
// `refill_idle` of them are idle.
#ifndef YUNE_FETCH_CHUNK
#define YUNE_FETCH_CHUNK 128
#endif

template <bool ANY, bool COUNT, int ACCEL>
__device__ __forceinline__ void trace_queue(const TraceArgs& A, const uint32_t s_box, const uint32_t s_ref, const uint32_t s_ray, WorkCount& wc)
{
    const DevScene& sc = A.sc;
    const int lane = threadIdx.x & 31;
    const unsigned lane_lt = (1u << lane) - 1u;
    const int n = ANY ? *A.n_shadow : *A.n_extend;
    int* fetch = ANY ? A.fetch_shadow : A.fetch_extend;
    int stack[YUNE_STACK_SIZE + YUNE_STACK_BASE];
    stack[0] = stack[1] = YUNE_REF_DONE;
    Lane L; L.cur = YUNE_REF_DONE; L.pend_pos = L.pend_end = 0; L.sp = YUNE_STACK_BASE; L.tri = -1; L.guard = false;
    bool have = false;          // this lane holds a ray
    int  where = 0;             // extension: slot index; shadow: answer target
    bool exhausted = (n == 0), last_chunk = false;
    // Queue entries a warp takes per atomic: YUNE_FETCH_CHUNK when the queue is long; when it is short (the drain of a job: a
    // few thousand rays for 4736 warps) just enough to give every warp one refill, so that the rays run side by side instead
    // of 128 at a time in a few warps (launch list of a 1-spp frame: the trace kernel took 100-200 us for a handful of rays).
    const int n_warps = (int)(gridDim.x * (blockDim.x >> 5));
    const int fetch_chunk = min(YUNE_FETCH_CHUNK, max(8, ((n / n_warps) + 7) & ~7));
    int chunk_next = 0, chunk_end = 0;            // warp-uniform: the private range of queue entries still to hand out
    // ACCEL bit 4: the four scheduling knobs are at their defaults and compiled in as immediates (as kernel parameters they are
    // re-read from the constant bank inside the step loop -- no register is free to hold them: trace -1.7 %)
    const int refill_idle = (ACCEL & 16) ? YUNE_DEF_REFILL_IDLE : A.refill_idle, tri_min = (ACCEL & 16) ? YUNE_DEF_PHASE_MIN : A.phase_min,
              inner_min = (ACCEL & 16) ? YUNE_DEF_INNER_MIN : A.inner_min, inner_chain = (ACCEL & 16) ? YUNE_DEF_INNER_CHAIN : A.inner_chain;

    for (;;) {
        // ---- retire finished rays ----
        if (have && L.cur == YUNE_REF_DONE && !(L.pend_pos < L.pend_end)) {
            if (ANY) {
                const unsigned char vis = L.tri >= 0 ? 0 : 1;
                if (where >= 0) A.vis_a[where] = vis; else A.vis_b[~where] = vis;
            } else if (ACCEL & 4) A.hit[where] = make_float4(L.t_best, lds128v(s_ray).w, lds128v(s_ray + 16u * blockDim.x).w, __int_as_float(L.tri));
            else A.hit[where] = make_float4(L.t_best, L.u, L.v, __int_as_float(L.tri));
            have = false;
        }
        // ---- refill: lanes take rays from the warp's private chunk; a new chunk costs one atomic per YUNE_FETCH_CHUNK rays ----
        const unsigned idle = __ballot_sync(0xffffffffu, !have);
        if (idle == 0xffffffffu && exhausted) break;
        if (!exhausted) {
            if (chunk_next >= chunk_end) {
                int base = 0;
                if (lane == 0) base = atomicAdd(fetch, fetch_chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                chunk_next = base; chunk_end = min(base + fetch_chunk, n);
                last_chunk = base + fetch_chunk >= n;
            }
            const int q = chunk_next + __popc(idle & lane_lt);
            if (!have && q < chunk_end) {
                float4 o, d;
                if (ANY) { const int qq = A.sq_idx ? A.sq_idx[q] : q; o = A.sq_o[qq]; d = A.sq_d[qq]; where = __float_as_int(d.w); }
                else { where = A.eq ? A.eq[q] : q; const size_t k = (size_t)where * A.ray_stride; o = A.ray_o[k]; d = A.ray_d[k]; }
                lane_init<ANY, COUNT, ACCEL>(L, sc, o, d, wc, s_ray);
                have = true;
            }
            chunk_next = min(chunk_next + __popc(idle), chunk_end);
            exhausted = last_chunk && chunk_next >= chunk_end;
        }
        // ---- steps, until enough lanes have nothing left to do ----
        const int busy_min = exhausted ? 1 : 33 - refill_idle;      // keep stepping while at least this many lanes have work
        #pragma unroll 1
        for (;;) {
            const bool wt = L.pend_pos < L.pend_end;
            const unsigned bi = __ballot_sync(0xffffffffu, L.cur >= 0), bt = __ballot_sync(0xffffffffu, wt);
            if (__popc(bi | bt) < busy_min) break;
            const int ni = __popc(bi), nt = __popc(bt);
            if (nt >= tri_min || nt > ni) { if (wt) lane_tri_step<ANY, COUNT, ACCEL>(L, stack, sc, s_ray, wc); }
            else {
                if (L.cur >= 0) lane_node_step<ANY, COUNT, ACCEL>(L, stack, sc, s_box, s_ref, s_ray, wc);
                #pragma unroll 1
                for (int k = 0; k < inner_chain && __popc(__ballot_sync(0xffffffffu, L.cur >= 0)) >= inner_min; k++) {
                    if (L.cur >= 0) lane_node_step<ANY, COUNT, ACCEL>(L, stack, sc, s_box, s_ref, s_ray, wc);
                }
            }
            __syncwarp();
        }
    }
}

template <bool COUNT, int ACCEL>
__global__ void __launch_bounds__(YUNE_TRACE_MAX_BLOCK, 1) k_trace(TraceArgs A)
{
    extern __shared__ float4 s_box[];                    // [3 * n_smem_pairs] boxes, then [n_smem_pairs] int2 child refs
    const DevScene& sc = A.sc;
    int2* s_ref = reinterpret_cast<int2*>(s_box + 3 * sc.n_smem_pairs);
    if (ACCEL & 4) {                                     // wide records are staged as they are (7 float4 = 112 bytes each)
        for (int i = threadIdx.x; i < sc.n_smem_pairs * 7; i += blockDim.x) s_box[i] = __ldg(sc.pairs + i);
    } else
    for (int i = threadIdx.x; i < sc.n_smem_pairs * 4; i += blockDim.x) {
        const float4 v = __ldg(sc.pairs + i);
        const int node = i >> 2, k = i & 3;
        if (k < 3) s_box[3 * node + k] = v;
        else s_ref[node] = make_int2(__float_as_int(v.x), __float_as_int(v.y));
    }
    __syncthreads();

    WorkCount wc; wc.box = 0; wc.tri = 0;
    uint32_t a_box = (uint32_t)__cvta_generic_to_shared(s_box), a_ref = (uint32_t)__cvta_generic_to_shared(s_ref);
    // Opaque to the compiler: otherwise ptxas re-derives this address from SR_CgaCtaId in every node step (S2R + MOV + LEA) instead
    // of keeping it in a (uniform) register: trace -1.1 %.
    asm volatile("mov.u32 %0, %0;" : "+r"(a_box));
    const uint32_t a_ray = a_box + 112u * (uint32_t)sc.n_smem_pairs + 16u * threadIdx.x;      // ACCEL bit 2: this thread's (o, u) record
    trace_queue<true, COUNT, ACCEL>(A, a_box, a_ref, a_ray, wc);      // shadow rays: any hit
    trace_queue<false, COUNT, ACCEL>(A, a_box, a_ref, a_ray, wc);     // extension rays: closest hit
    if (COUNT) {
        const int lane = threadIdx.x & 31;
        unsigned long long b = wc.box, t = wc.tri;
        for (int o = 16; o > 0; o >>= 1) { b += __shfl_down_sync(0xffffffffu, b, o); t += __shfl_down_sync(0xffffffffu, t, o); }
        if (lane == 0 && A.tot) { atomicAdd(&A.tot->box_tests, b); atomicAdd(&A.tot->tri_tests, t); }
    }
}

// One thread: fold this iteration's queue sizes into the running totals and clear the OTHER parity's counters so
// the next shade launch starts from zero.  Runs between k_trace and the next k_shade on the same stream.
__global__ void k_iter_end(IterCounters* ctr, Totals* tot, int parity)
{
    IterCounters& c = ctr[parity];
    tot->extend_rays += (unsigned long long)c.n_extend;
    tot->shadow_rays += (unsigned long long)c.n_shadow;
    tot->live_last = c.live;
    tot->iterations += 1;
    IterCounters z; z.n_extend = z.n_shadow = z.n_events = z.live = 0; z.fetch_extend = z.fetch_shadow = 0;
    ctr[parity ^ 1] = z;
}


// ------------------------------------------------------------------------------------------------------------
// Shade stage with IN-BLOCK SORTING (default).  ncu on the kernel above: 14.8 of 32 threads active per instruction, because
// the lanes of a warp are in different situations (diffuse surface 53 %, specular surface 19 %, miss / drain / free 28 %)
// and take turns through ~10 K SASS instructions.  Here a persistent block walks chunks of YUNE_SHADE_BLOCK slots:
//   A  classify   every lane, cheap: settle last iteration's NEE answers, retire misses and drained paths into the
//                 accumulation buffer, and append the slot to one of three shared-memory lists
//   D  diffuse    runs only when the list holds a full block of diffuse surface hits: NEE (+ MIS rays), lobe sampling,
//                 roulette -- every lane of every warp on the same code
//   S  specular   likewise for mirror / glass hits (no NEE)
//   R  regenerate likewise for slots that need the next (pixel, sample): camera ray in fp64
// Leftovers stay in the lists for the next chunk; the last pass flushes them.  Per-slot results are the same as the
// fused kernel's (same arithmetic, same random-number addressing); only the order of queue entries differs.
// ------------------------------------------------------------------------------------------------------------
enum { YL_DIFFUSE = 0, YL_SPECULAR = 1 };


struct DenseShared {
    // A classify phase adds up to YUNE_CLASSIFY_N blocks of entries to < 1 block of leftovers; the regeneration list is emptied to
    // < 1 block before every surface round (which adds at most one block to it).
    int surf[2][(YUNE_CLASSIFY_N + 1) * YUNE_SHADE_BLOCK];      // diffuse / specular surface hits waiting for a full round
    int regen[(YUNE_CLASSIFY_N + 1) * YUNE_SHADE_BLOCK];        // slots waiting for a fresh sample (fed by A, D and S)
    int n_surf[2], n_regen;
    int cnt[3 * (YUNE_NW + 1)];             // block_alloc scratch
    unsigned long long sample_base; int ext_base;
    int visits[3];                          // rounds' entries by kind (statistics)
};

// phase A for one slot
// Returns whether the slot goes on to a surface round (it may stay in flight).
struct ClassifyIn { uint4 meta; float hit_w; unsigned char vis_l; bool valid, spec; float4 col, thr, pend; };
// the loads of phase A, all independent of each other: issued together (and, with YUNE_CLASSIFY2, for two slots at once)
__device__ __forceinline__ ClassifyIn classify_load(const RenderArgs& A, const int s)
{
    const PathPool& P = A.pool;
    ClassifyIn in;
    in.valid = s < P.n_slots;
    in.meta = in.valid ? P.meta[s] : make_uint4(0, 0, 0, YS_DONE);
    in.hit_w = in.valid ? P.hit[s].w : 0.0f;                                   // unconditional: in flight together with meta
    in.vis_l = in.valid ? P.vis_l[s] : 0;                                      // likewise: saves the pending-NEE path one dependent round trip
    // Speculative: the state a pending NEE answer needs is requested before the flags are known (one dependent round trip less
    // for the ~half of the slots that have one).  Not while the pool drains (A.tail): most slots are DONE then and the scan
    // itself is what an iteration costs.
    in.spec = in.valid && !A.tail;
    in.col = make_float4(0, 0, 0, 0); in.thr = in.col; in.pend = in.col;
    if (in.spec) { in.col = P.col[s]; in.thr = P.thr[s]; in.pend = P.pend_l[s]; }
    return in;
}
__device__ __forceinline__ bool classify_finish(const RenderArgs& A, const int s, const ClassifyIn& in, DenseShared& sh);
__device__ __forceinline__ bool classify_slot(const RenderArgs& A, const int s, DenseShared& sh)
{
    const ClassifyIn in = classify_load(A, s);
    return classify_finish(A, s, in, sh);
}
__device__ __forceinline__ bool classify_finish(const RenderArgs& A, const int s, const ClassifyIn& in, DenseShared& sh)
{
    const PathPool& P = A.pool;
    const bool valid = in.valid, spec = in.spec;
    const uint4 meta = in.meta; const float hit_w = in.hit_w; const unsigned char vis_l = in.vis_l;
    float4 spec_col = in.col, spec_thr = in.thr, spec_pend = in.pend;
    const unsigned state = meta.w & YS_STATE_MASK;
    bool to_regen = valid && state == YS_FREE, to_d = false, to_s = false;
    if (state == YS_TRACE || state == YS_DRAIN) {
        const int tri = state == YS_TRACE ? __float_as_int(hit_w) : -1;
        if (tri >= 0) {
            to_s = A.sc.tri_class[tri] != 0; to_d = !to_s;
#ifdef YUNE_PREFETCH_RAY
            // the surface round of this slot runs a chunk or two later and reads the ray record, which classify does not touch:
            // request it into L2 now so that the round's loads are L2 hits instead of a DRAM round trip
            asm volatile("prefetch.global.L2 [%0];" :: "l"(P.ray_o.p + 2 * (size_t)s));
#endif
        }
        const bool pend = (meta.w & (YF_PEND_EVT | YF_PEND_L)) != 0;
        if (pend || tri < 0) {
            if (!spec) { spec_col = P.col[s]; spec_thr = P.thr[s]; if (meta.w & YF_PEND_L) spec_pend = P.pend_l[s]; }
            V3 col = xyz(spec_col);
            bool dirty = false;
            // resolve the NEE launched at the previous visit (udpt.cl:551-608): nee = light sample [+ BRDF sample]
            if (meta.w & YF_PEND_EVT) {
                const V3 T = xyz(P.thr[s]);
                const int e = P.evt_idx[s];
                const float4 e0 = P.evt[3 * (size_t)e], e1 = P.evt[3 * (size_t)e + 1], e2 = P.evt[3 * (size_t)e + 2];
                const int ef = __float_as_int(e0.w);
                const bool visS = (ef & YE_HAS_S) && P.evt_vis[4 * (size_t)e + 0];
                const bool visMV = (ef & YE_HAS_MV) && P.evt_vis[4 * (size_t)e + 1];
                const bool visMO = (ef & YE_MO_IS_MV) ? visMV : ((ef & YE_HAS_MO) && P.evt_vis[4 * (size_t)e + 2]);
                V3 nee;
                if (visS) nee = vadd(xyz(e0), visMV ? xyz(e1) : v3(0, 0, 0));
                else      nee = ((ef & (YE_HAS_MO | YE_MO_IS_MV)) && visMO) ? xyz(e2) : v3(0, 0, 0);
                col = vadd(col, vmul(T, nee)); dirty = true;
            } else if ((meta.w & YF_PEND_L) && vis_l) { col = vadd(col, vmul(xyz(spec_thr), xyz(spec_pend))); dirty = true; }
            if (tri < 0) {
                if (state == YS_TRACE) {                                            // nothing hit, or a light
                    const int lid = (int)((meta.w >> YF_LID_SHIFT) & YF_LID_MASK) - 1;
                    if (meta.z == 0) {                                              // udpt.cl:437-446
                        if (lid >= 0) col = (vdot(xyz(P.ray_d[s]), A.lights.l[lid].normal) < 0.0f) ? v3(1.0f, 1.0f, 1.0f) : v3(0.1f, 0.1f, 0.1f);
                        else col = v3(0.4f, 0.4f, 0.4f);
                    }
                    else if (lid >= 0 && (meta.w & YF_PREV_SPEC)) col = vadd(col, vmul(xyz(spec_thr), A.lights.l[lid].ke));   // :490-493
                }
                finish_sample(A, meta.x, col);                                      // udpt.cl:193-210
                to_regen = true;
            } else if (dirty) P.col[s] = f4(col, 0.0f);
        }
    }
    list_push(sh.surf[YL_DIFFUSE], &sh.n_surf[YL_DIFFUSE], to_d, s);
    list_push(sh.surf[YL_SPECULAR], &sh.n_surf[YL_SPECULAR], to_s, s);
    list_push(sh.regen, &sh.n_regen, to_regen, s);
    return to_d || to_s;
}

// phases D / S for one list entry per thread (s < 0: idle lane of a flush round); every thread of the block calls it
template <bool MIS, bool SPEC>
__device__ __forceinline__ void surface_round(const RenderArgs& A, const int s, DenseShared& sh, int& live)
{
    const PathPool& P = A.pool;
    IterCounters* C = A.ctr + A.parity;
    const LightDev* lights = A.lights.l;
    const int n_lights = A.lights.n;
    bool has_ext = false, finished = false;
    V3 ext_o = v3(0, 0, 0), ext_d = v3(0, 0, 1); float ext_t = INFINITY; int ext_lid = -1;
    NeeOut N; N.S.has = N.MV.has = N.MO.has = false; N.mo_is_mv = false; N.Lv = N.BV = N.BO = v3(0, 0, 0);
    uint4 meta = make_uint4(0, 0, 0, 0);
    V3 col = v3(0, 0, 0), T = v3(1, 1, 1), Tn = v3(1, 1, 1);
    unsigned new_flags = 0;
    const unsigned act = __ballot_sync(0xffffffffu, s >= 0);

    if (s >= 0) {
        meta = P.meta[s];
        const float4 hit = P.hit[s], ro = P.ray_o[s], rd = P.ray_d[s];
        col = xyz(P.col[s]); T = xyz(P.thr[s]);
        const int tri = __float_as_int(hit.w);
        const V3 o = xyz(ro), d = xyz(rd);
        const unsigned vtx = meta.z;
        bool terminate = false;
        // ---- the surface point (udpt.cl:375-385)
        const float4 s0 = __ldg(A.sc.shade + 4 * (size_t)tri), s1 = __ldg(A.sc.shade + 4 * (size_t)tri + 1), s2 = __ldg(A.sc.shade + 4 * (size_t)tri + 2);
        const MatDev mat = load_material(A.sc.mats, __float_as_int(s0.w));
        const float bw = YF_SUB(YF_SUB(1.0f, hit.y), hit.z);
        const V3 hp = vadd(o, vscale(d, hit.x));
        const V3 n = vnormalize(vmadd3(xyz(s0), bw, xyz(s1), hit.y, xyz(s2), hit.z));
        const V3 w_o = vneg(d);
        U4 u_nee; u_nee.x = u_nee.y = u_nee.z = u_nee.w = 0u;
        if (!SPEC || (vtx > 0 && (int)vtx - 1 > A.rr_threshold)) u_nee = draw4(A.seed, meta.x, meta.y, vtx, YUNE_BLK_NEE);
        if (vtx > 0) {                                                          // arrival of bounce i = vtx - 1
            Tn = xyz(P.thr_next[s]);
            col = vadd(col, vmul(T, mat.ke));                                   // :498
            T = Tn;                                                             // :504-507 (product was formed at sampling time)
            if ((int)vtx - 1 > A.rr_threshold) {                                // :514-523
                const float p = cl_min(luminance(T), 0.95f);
                const float r = u01(u_nee.x);
                if (r >= p) terminate = true;
                else T = vscale(T, YF_DIV(1.0f, p));
            }
        }
        bool nee_pending = false;
        __syncwarp(act);                                                        // see nee_sample: keep the warp together
        const unsigned alive = __ballot_sync(act, !terminate);
        if (!terminate) {
            // ---- next-event estimation at this vertex (evaluateDirectLighting, :535-609)
            LobePrep lobes; lobes.mode = 0; lobes.pd = lobes.ps = 0.0f;
            V3 Nx = v3(1, 0, 0), Ny = v3(0, 1, 0);
            if (!SPEC) {
                lobes = lobe_prepare(mat, false);                               // shared by the NEE and the bounce below
                onb(n, Nx, Ny);
                col = vadd(col, vmul(T, mat.ke));                               // the 'emission' term of every return path
                nee_sample<MIS, false>(alive, lights, n_lights, mat, lobes, hp, n, Nx, Ny, w_o, u_nee, A.seed, meta.x, meta.y, vtx, A.oren_nayar != 0, N);
                nee_pending = N.S.has || N.MV.has || N.MO.has;
            }
            __syncwarp(alive);
            // ---- continue the path (udpt.cl:463-530)
            if (!A.gi_check) terminate = true;
            else {
                const U4 u_b = draw4(A.seed, meta.x, meta.y, vtx, YUNE_BLK_BOUNCE);
                V3 dir = v3(0, 0, 1);
                if (SPEC) {
                    float ior = 1.0f;
                    dir = sample_specular(mat, w_o, n, u01(u_b.w), ior);
                    Tn = vscale(T, ior);
                } else {
                    float prob = 0.0f, pdf = 1.0f;
                    const bool glossy = select_lobe_r(lobes, u01(u_b.x), false, prob);
                    if (prob == 0.0f) terminate = true;                         // absorbed (:475-476)
                    else {
                        dir = glossy ? sample_phong(w_o, n, mat.px, mat.py, u01(u_b.y), u01(u_b.z), true, pdf)
                                     : sample_cosine_onb(n, Nx, Ny, u01(u_b.y), u01(u_b.z), pdf);
                        if (pdf <= 0.0f) terminate = true;                      // :488
                        else Tn = vdivs(vscale(vmul(T, eval_brdf(mat, dir, w_o, n, glossy, prob, true, A.oren_nayar != 0)), fmaxf(vdot(dir, n), 0.0f)), pdf);
                    }
                }
                if (!terminate) {
                    has_ext = true;
                    ext_d = dir; ext_o = vadd(hp, vscale(dir, YUNE_EPS)); ext_t = INFINITY;
                    ext_lid = light_loop(lights, n_lights, ext_o, ext_d, ext_t);
                    new_flags = YS_TRACE | (SPEC ? YF_PREV_SPEC : 0u) | ((unsigned)(ext_lid + 1) << YF_LID_SHIFT);
                    meta.z = vtx + 1;
                }
            }
        }
        __syncwarp(act);
        if (!has_ext) {
            if (nee_pending) new_flags = YS_DRAIN;
            else { finished = true; new_flags = YS_FREE; finish_sample(A, meta.x, col); }
        }
    }
    list_push(sh.regen, &sh.n_regen, finished, s);
    if (s >= 0 && !finished) live++;

    // ---- queue pushes (block-wide compaction: one atomic per counter per round) and state write-back
    if (SPEC) {
        int* const counters[1] = { &C->n_extend };
        const int cnt[1] = { has_ext ? 1 : 0 };
        int first[1];
        block_alloc<1, YUNE_NW>(counters, cnt, first, sh.cnt);
        if (has_ext) P.eq[first[0]] = s;
    } else {
        const bool is_event = N.MV.has || N.MO.has;
        int* const counters[3] = { &C->n_extend, &C->n_shadow, &C->n_events };
        const int cnt[3] = { has_ext ? 1 : 0, (N.S.has ? 1 : 0) + (N.MV.has ? 1 : 0) + (N.MO.has ? 1 : 0), is_event ? 1 : 0 };
        int first[3];
        block_alloc<3, YUNE_NW>(counters, cnt, first, sh.cnt);
        if (has_ext) P.eq[first[0]] = s;
        const int ev = A.parity * P.n_slots + first[2];
        if (is_event) {
            const int ef = (N.S.has ? YE_HAS_S : 0) | (N.MV.has ? YE_HAS_MV : 0) | (N.MO.has ? YE_HAS_MO : 0) | (N.mo_is_mv ? YE_MO_IS_MV : 0);
            P.evt[3 * (size_t)ev] = f4(N.Lv, __int_as_float(ef));
            P.evt[3 * (size_t)ev + 1] = f4(N.BV, 0.0f);
            P.evt[3 * (size_t)ev + 2] = f4(N.BO, 0.0f);
            P.evt_idx[s] = ev;
            new_flags |= YF_PEND_EVT;
        } else if (N.S.has) new_flags |= YF_PEND_L;
        int qs = first[1];
        if (N.S.has) {
            P.sq_o[qs] = f4(N.S.o, N.S.tmax);
            P.sq_d[qs] = f4(N.S.d, __int_as_float(is_event ? ~(4 * ev + 0) : s));
            if (!is_event) P.pend_l[s] = f4(N.Lv, 0.0f);
            qs++;
        }
        if (MIS) {
            if (N.MV.has) { P.sq_o[qs] = f4(N.MV.o, N.MV.tmax); P.sq_d[qs] = f4(N.MV.d, __int_as_float(~(4 * ev + 1))); qs++; }
            if (N.MO.has) { P.sq_o[qs] = f4(N.MO.o, N.MO.tmax); P.sq_d[qs] = f4(N.MO.d, __int_as_float(~(4 * ev + 2))); qs++; }
        }
    }
    if (s >= 0 && !finished) {
        meta.w = new_flags;
        P.meta[s] = meta;
        P.col[s] = f4(col, 0.0f);
        P.thr[s] = f4(T, 0.0f);
        if (has_ext) {
            P.ray_o[s] = f4(ext_o, ext_t);
            P.ray_d[s] = f4(ext_d, __int_as_float(ext_lid));
            P.thr_next[s] = f4(Tn, 0.0f);
        }
    }
}

// phase R: `take` list entries [from, from + take), thread t serves entry from + t
__device__ __forceinline__ void regen_round(const RenderArgs& A, const int from, const int take, DenseShared& sh, int& live)
{
    const PathPool& P = A.pool;
    IterCounters* C = A.ctr + A.parity;
    if (threadIdx.x == 0) {
        const unsigned long long base = atomicAdd(&A.tot->next_sample, (unsigned long long)take);     // next (pixel, sample) in global order
        const unsigned long long total = A.tot->n_samples;
        const int n_fresh = base >= total ? 0 : (int)((total - base) < (unsigned long long)take ? (total - base) : (unsigned long long)take);
        sh.sample_base = base;
        sh.ext_base = n_fresh > 0 ? atomicAdd(&C->n_extend, n_fresh) : 0;
    }
    __syncthreads();
    if ((int)threadIdx.x < take) {
        const int s = sh.regen[from + threadIdx.x];
        const unsigned long long g = sh.sample_base + threadIdx.x;
        if (g < A.tot->n_samples) {
            // udpt.cl:164-189
            const unsigned long long n_pix = (unsigned long long)A.width * A.height;
            const unsigned pixel = (unsigned)(g % n_pix);
            const unsigned sample = (unsigned)(A.spp_begin + (int)(g / n_pix));
            const int px = pixel % A.width, py = pixel / A.width;
            const U4 uj = draw4(A.seed, pixel, sample, YUNE_VERTEX_CAMERA, 0u);
            V3 ro, rd;
            create_ray(A.cam, A.width, A.height, (float)px + u01(uj.x), (float)py + u01(uj.y), ro, rd);
            float rt = INFINITY;
            const int rl = light_loop(A.lights.l, A.lights.n, ro, rd, rt);
            P.eq[sh.ext_base + threadIdx.x] = s;
            P.ray_o[s] = f4(ro, rt);
            P.ray_d[s] = f4(rd, __int_as_float(rl));
            P.meta[s] = make_uint4(pixel, sample, 0u, YS_TRACE | ((unsigned)(rl + 1) << YF_LID_SHIFT));
            P.col[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            P.thr[s] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            live++;
        } else P.meta[s] = make_uint4(0u, 0u, 0u, YS_DONE);
    }
    __syncthreads();
}

template <bool MIS>
__global__ void __launch_bounds__(YUNE_SHADE_BLOCK, YUNE_SHADE_MIN_BLOCKS) k_shade_dense(RenderArgs A)
{
    __shared__ DenseShared sh;
    const int tid = threadIdx.x;
    if (tid == 0) { sh.n_surf[0] = sh.n_surf[1] = 0; sh.n_regen = 0; sh.visits[0] = sh.visits[1] = sh.visits[2] = 0; }
    __syncthreads();
    // A chunk is YUNE_CLASSIFY_N blocks of slots: every thread classifies N slots with all their (independent) loads in flight
    // together.  Phase A is where the kernel waits for DRAM (per-line stall view: the first use of meta / hit and the barrier after
    // it); two slots per thread: 1.162 -> 1.129 ms per launch (profiles/r2_ab_shade_classify2.log).
    const int CH = YUNE_CLASSIFY_N * YUNE_SHADE_BLOCK;
    const int n_chunks = (A.pool.n_slots + CH - 1) / CH;
    int live = 0;                                   // slots this thread left in flight (TRACE or DRAIN)
    for (int chunk = blockIdx.x; ; chunk += gridDim.x) {
        const bool flush = chunk >= n_chunks;       // block-uniform: the pass after the last chunk empties the lists
        // While the pool drains (A.tail: every sample has been handed out) a chunk whose slots had nothing in flight after the
        // previous iteration is skipped by reading one byte.
        if (!flush && A.tail && A.chunk_live[chunk] == 0) continue;
        bool cont = false;
        if (!flush) {
            ClassifyIn in[YUNE_CLASSIFY_N];
            #pragma unroll
            for (int j = 0; j < YUNE_CLASSIFY_N; j++) in[j] = classify_load(A, chunk * CH + j * YUNE_SHADE_BLOCK + tid);
#if YUNE_CLASSIFY_N == 2
            // ONE copy of the finish code, the slot's loaded state picked by selects: the kernel's hot path is instruction-cache
            // bound (no_instruction is its second-largest stall) and the unrolled form costs 300 SASS lines: 1.112 -> 1.081 ms
            #pragma unroll 1
            for (int j = 0; j < 2; j++) {
                ClassifyIn x = in[0];
                if (j) x = in[1];
                cont = classify_finish(A, chunk * CH + j * YUNE_SHADE_BLOCK + tid, x, sh) || cont;
            }
#else
            #pragma unroll
            for (int j = 0; j < YUNE_CLASSIFY_N; j++) cont = classify_finish(A, chunk * CH + j * YUNE_SHADE_BLOCK + tid, in[j], sh) || cont;
#endif
        }
        if (A.tail && !flush) {
            const int any = __syncthreads_or(cont ? 1 : 0);
            if (tid == 0) A.chunk_live[chunk] = any ? 1 : 0;
        } else __syncthreads();
        for (;;) {
            // regeneration rounds first: keeps that list below one block before a surface round adds to it
            for (;;) {
                const int n = sh.n_regen;
                if (!(n >= YUNE_SHADE_BLOCK || (flush && n > 0))) break;
                const int take = n < YUNE_SHADE_BLOCK ? n : YUNE_SHADE_BLOCK;
                __syncthreads();
                if (tid == 0) { sh.n_regen = n - take; sh.visits[2] += take; }
                regen_round(A, n - take, take, sh, live);
            }
            bool did = false;
            YUNE_NO_UNROLL
            for (int k = 0; k < 2; k++) {
                const int n = sh.n_surf[k];
                if (n >= YUNE_SHADE_BLOCK || (flush && n > 0)) {
                    const int take = n < YUNE_SHADE_BLOCK ? n : YUNE_SHADE_BLOCK;
                    const int s = tid < take ? sh.surf[k][n - take + tid] : -1;
                    __syncthreads();
                    if (tid == 0) { sh.n_surf[k] = n - take; sh.visits[k] += take; }
                    if (k == YL_DIFFUSE) surface_round<MIS, false>(A, s, sh, live);
                    else                 surface_round<MIS, true>(A, s, sh, live);
                    __syncthreads();
                    did = true;
                    break;                          // back to the regeneration list before the next surface round
                }
            }
            if (!did) break;
        }
        if (flush) break;
        __syncthreads();                            // every warp has read the list counts before the next chunk's classify bumps them
    }
    // one atomic per block: how many slots are still in flight
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
    if ((tid & 31) == 0) sh.cnt[tid >> 5] = live;
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int w = 0; w < YUNE_NW; w++) total += sh.cnt[w];
        if (total > 0) atomicAdd(&A.ctr[A.parity].live, total);
        if (sh.visits[0]) atomicAdd(&A.tot->visits_d, (unsigned long long)sh.visits[0]);
        if (sh.visits[1]) atomicAdd(&A.tot->visits_s, (unsigned long long)sh.visits[1]);
        if (sh.visits[2]) atomicAdd(&A.tot->visits_r, (unsigned long long)sh.visits[2]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// pool / buffer initialisation, tonemap, hooks
// ------------------------------------------------------------------------------------------------------------
__global__ void k_pool_reset(PathPool P)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.n_slots) P.meta[s] = make_uint4(0, 0, 0, YS_FREE);
}
// A pipelined call (yune_ctx::carry) continues on a pool that the previous call left in its drain phase: slots that found no
// sample to start are DONE.  They take part again.
__global__ void k_pool_revive(PathPool P)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.n_slots && (P.meta[s].w & YS_STATE_MASK) == YS_DONE) P.meta[s].w = YS_FREE;
}
__global__ void k_fill_f4(float4* p, size_t n, float4 v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// option "deterministic": the float sum buffer is a view of the fixed-point one (rgb * 2^-24, a = count), and back
__global__ void k_fix_to_sum(const long long* fix, float4* sum, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const longlong2 a = reinterpret_cast<const longlong2*>(fix)[2 * i], b = reinterpret_cast<const longlong2*>(fix)[2 * i + 1];
    const float inv = 1.0f / YUNE_FIX_SCALE;
    sum[i] = make_float4(__ll2float_rn(a.x) * inv, __ll2float_rn(a.y) * inv, __ll2float_rn(b.x) * inv, __ll2float_rn(b.y));
}
__global__ void k_sum_to_fix(const float4* sum, long long* fix, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = sum[i];
    fix[4 * i + 0] = to_fixed(v.x); fix[4 * i + 1] = to_fixed(v.y); fix[4 * i + 2] = to_fixed(v.z); fix[4 * i + 3] = __float2ll_rn(v.w);
}

// tonemap.cl:14-47 on mean = sum / count.  Output keeps the reference's "gamma on all four channels".
__global__ void k_tonemap(const float4* sum, float4* hdr, float4* ldr, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 sv = sum[i];
    float4 c;
    if (sv.w > 0.0f) { c.x = YF_DIV(sv.x, sv.w); c.y = YF_DIV(sv.y, sv.w); c.z = YF_DIV(sv.z, sv.w); c.w = sv.w; }
    else c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hdr) hdr[i] = c;
    if (!ldr) return;
    const float lum_world = YF_ADD(YF_ADD(YF_ADD(YF_MUL(0.212671f, c.x), YF_MUL(0.715160f, c.y)), YF_MUL(0.072169f, c.z)), 0.001f);
    const float lum_white = 1.0f;
    const float lum_display = YF_DIV(YF_MUL(lum_world, YF_ADD(1.0f, YF_DIV(lum_world, YF_MUL(lum_white, lum_white)))), YF_ADD(1.0f, lum_world));
    float4 l;
    l.x = YF_MUL(lum_display, powf(YF_DIV(c.x, lum_world), 1.0f));
    l.y = YF_MUL(lum_display, powf(YF_DIV(c.y, lum_world), 1.0f));
    l.z = YF_MUL(lum_display, powf(YF_DIV(c.z, lum_world), 1.0f));
    l.w = YF_MUL(lum_display, powf(YF_DIV(c.w, lum_world), 1.0f));
    const float g = YF_DIV(1.0f, 2.2f);
    ldr[i] = make_float4(powf(l.x, g), powf(l.y, g), powf(l.z, g), powf(l.w, g));
}

// Primary rays for the parity hook: same create_ray as the renderer; jitter from the REFERENCE generator
// (udpt.cl:175-187) or pixel centres.
__global__ void k_hook_primary(RenderArgs A, int jitter_mode, uint32_t rand, float4* ray_o, float4* ray_d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.width * A.height) return;
    const int x = i % A.width, y = i / A.width;
    float r1 = 0.5f, r2 = 0.5f;
    if (jitter_mode == 1) {
        uint32_t seed = (uint32_t)((y + 1) * A.width + (x + 1));
        seed = rand * seed;
        seed = wang_hash(seed);
        if (seed == 0) seed = wang_hash(seed);
        seed = xor_shift(seed); r1 = u01(seed);
        seed = xor_shift(seed); r2 = u01(seed);
    }
    V3 o, d; create_ray(A.cam, A.width, A.height, (float)x + r1, (float)y + r2, o, d);
    float t = INFINITY;
    const int lid = light_loop(A.lights.l, A.lights.n, o, d, t);
    ray_o[i] = f4(o, t); ray_d[i] = f4(d, __int_as_float(lid));
}

// Arbitrary rays for the parity hook: run traceRay's light loop, then hand the ray to k_trace.
__global__ void k_hook_prepare(LightSet L, int n, const float* od6, const float* tmax, int any_hit,
                               float4* ray_o, float4* ray_d, int* light_id, unsigned char* vis)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 o = v3(od6[6 * (size_t)i], od6[6 * (size_t)i + 1], od6[6 * (size_t)i + 2]);
    const V3 d = v3(od6[6 * (size_t)i + 3], od6[6 * (size_t)i + 4], od6[6 * (size_t)i + 5]);
    float t = tmax ? tmax[i] : INFINITY;
    const int lid = light_loop(L.l, L.n, o, d, t);
    light_id[i] = lid;
    if (any_hit) {
        // a light inside the segment already occludes (traceRay ORs the two answers, udpt.cl:278-279)
        ray_o[i] = f4(o, lid >= 0 ? -1.0f : t);
        ray_d[i] = f4(d, __int_as_float(i));
        vis[i] = 0;
    } else {
        ray_o[i] = f4(o, t); ray_d[i] = f4(d, __int_as_float(lid));
    }
}
__global__ void k_hook_finish(int n, int any_hit, const float4* ray_o, const float4* hit, const unsigned char* vis,
                              int* tri_id, int* light_id, float* t_hit)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (any_hit) {
        const bool occluded = light_id[i] >= 0 || vis[i] == 0;
        tri_id[i] = occluded ? 0 : -1;
        if (t_hit) t_hit[i] = ray_o[i].w;
    } else {
        const float4 h = hit[i];
        const int tri = __float_as_int(h.w);
        tri_id[i] = tri;
        if (tri >= 0) light_id[i] = -1;
        if (t_hit) t_hit[i] = h.x;
    }
}

// option "sort_rays": the sorting keys of both queues (Morton cell of the ray origin), computed just before the sorts so that the
// shade kernel carries no code for an option that is off by default
__global__ void k_ray_keys(DevScene sc, PathPool P, int n_ext, int n_sh, unsigned* eq_key, unsigned* sq_key)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ext) eq_key[i] = origin_key(sc, xyz(P.ray_o[P.eq[i]]));
    if (i < n_sh) sq_key[i] = origin_key(sc, xyz(P.sq_o[i]));
}

// Copies the first `max_rays` rays of this iteration's two queues into side buffers (measurement aid: the bench hands
// them to the oracle, which counts the box/triangle tests of ITS ordered walk -- the roofline's work model).
__global__ void k_capture(PathPool P, const IterCounters* c, int max_rays, float4* ext_o, float4* ext_d, float4* sh_o, float4* sh_d, int* counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int ne = min(c->n_extend, max_rays), ns = min(c->n_shadow, max_rays);
    if (i == 0) { counts[0] = ne; counts[1] = ns; counts[2] = c->n_extend; counts[3] = c->n_shadow; }
    if (i < ne) { const int s = P.eq[i]; ext_o[i] = P.ray_o[s]; ext_d[i] = P.ray_d[s]; }
    if (i < ns) { sh_o[i] = P.sq_o[i]; sh_d[i] = P.sq_d[i]; }
}

// ------------------------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------------------------
static inline int ceil_div(long long a, int b) { return (int)((a + b - 1) / b); }

// Which instantiation serves a scene: ACCEL bit 0 = own tree + leaf-box filter, bit 1 = every node record is staged in shared
// memory (no global node path compiled in), bit 2 = 4-wide records, bit 3 = watertight intersection instead of the filter + the
// reference's Moller-Trumbore (isect 1, own binary tree only).
typedef void (*TraceKernel)(TraceArgs);
static TraceKernel trace_kernel(const DevScene& sc, bool count, bool default_knobs)
{
    const bool all_staged = sc.n_smem_pairs >= sc.n_inner;
    if (default_knobs && !count && sc.isect == 0 && sc.accel == 1) return all_staged ? k_trace<false, 19> : k_trace<false, 17>;      // the two production variants
    if (sc.isect == 1) return count ? k_trace<true, 9> : (all_staged ? k_trace<false, 11> : k_trace<false, 9>);
    if (sc.accel == 2) return count ? k_trace<true, 5> : (all_staged ? k_trace<false, 7> : k_trace<false, 5>);
    if (sc.accel == 1) return count ? k_trace<true, 1> : (all_staged ? k_trace<false, 3> : k_trace<false, 1>);
    return count ? k_trace<true, 0> : k_trace<false, 0>;
}
static bool knobs_are_default(int refill_idle, int phase_min, int inner_min, int inner_chain)
{
    return refill_idle == YUNE_DEF_REFILL_IDLE && phase_min == YUNE_DEF_PHASE_MIN && inner_min == YUNE_DEF_INNER_MIN && inner_chain == YUNE_DEF_INNER_CHAIN;
}
int trace_variant_id(const DevScene& sc, bool count, bool default_knobs)
{
    return sc.accel | (sc.n_smem_pairs >= sc.n_inner ? 4 : 0) | (count ? 8 : 0) | (sc.isect ? 16 : 0) | (default_knobs ? 32 : 0);
}
cudaError_t launch_trace(const TraceArgs& a, int grid, int block, size_t smem_bytes, bool count, cudaStream_t st)
{
    trace_kernel(a.sc, count, knobs_are_default(a.refill_idle, a.phase_min, a.inner_min, a.inner_chain))<<<grid, block, smem_bytes, st>>>(a);
    return cudaGetLastError();
}
// Shared-memory opt-in and resident blocks per SM of the instantiation that will be launched (the caller caches the answer per
// (variant, block, smem) and per context).
cudaError_t trace_prepare(const DevScene& sc, bool count, bool default_knobs, int block, size_t smem_bytes, int* blocks_per_sm)
{
    TraceKernel k = trace_kernel(sc, count, default_knobs);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, block, smem_bytes);
}
cudaError_t launch_iter_end(IterCounters* ctr, Totals* tot, int parity, cudaStream_t st)
{
    k_iter_end<<<1, 1, 0, st>>>(ctr, tot, parity);
    return cudaGetLastError();
}
cudaError_t launch_shade_dense(const RenderArgs& a, int sm_count, int blocks_per_sm, int* occ, cudaStream_t st)
{      // occ[2]: resident blocks per SM of the two instantiations, cached by the caller per context (0 = not asked yet)
    const int v = a.mis ? 1 : 0;
    if (occ[v] == 0) {
        // The kernel needs 3 x 9.4 KB of shared memory per SM; left alone the driver configures 64 KB (ncu: launch__shared_mem_config_size),
        // 32 KB of it taken from an L1 whose hit rate is 42 %.  Ask for the smallest carve-out that fits.
        int carve = 13;                                  // per cent of 228 KB -> the 32 KB configuration
        if (const char* e = getenv("YUNE_SHADE_CARVEOUT")) carve = atoi(e);
        if (carve >= 0) {
            if (v) cudaFuncSetAttribute(k_shade_dense<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            else   cudaFuncSetAttribute(k_shade_dense<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        }
        cudaError_t e = v ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[v], k_shade_dense<true>, YUNE_SHADE_BLOCK, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[v], k_shade_dense<false>, YUNE_SHADE_BLOCK, 0);
        if (e != cudaSuccess) return e;
        if (occ[v] < 1) occ[v] = 1;
    }
    int grid = sm_count * (blocks_per_sm > 0 ? blocks_per_sm : occ[v]);
    const int n_chunks = ceil_div(a.pool.n_slots, YUNE_SHADE_BLOCK);
    if (grid > n_chunks) grid = n_chunks;
    if (grid < 1) grid = 1;
    if (a.mis) k_shade_dense<true><<<grid, YUNE_SHADE_BLOCK, 0, st>>>(a);
    else       k_shade_dense<false><<<grid, YUNE_SHADE_BLOCK, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_capture(const PathPool& p, const IterCounters* c, int max_rays, float4* ext_o, float4* ext_d, float4* sh_o, float4* sh_d, int* counts, cudaStream_t st)
{
    k_capture<<<ceil_div(max_rays, 256), 256, 0, st>>>(p, c, max_rays, ext_o, ext_d, sh_o, sh_d, counts);
    return cudaGetLastError();
}
cudaError_t launch_ray_keys(const DevScene& sc, const PathPool& p, int n_ext, int n_sh, unsigned* eq_key, unsigned* sq_key, cudaStream_t st)
{
    const int n = n_ext > n_sh ? n_ext : n_sh;
    if (n > 0) k_ray_keys<<<ceil_div(n, 256), 256, 0, st>>>(sc, p, n_ext, n_sh, eq_key, sq_key);
    return cudaGetLastError();
}
cudaError_t launch_pool_reset(const PathPool& p, cudaStream_t st)
{
    k_pool_reset<<<ceil_div(p.n_slots, 256), 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_pool_revive(const PathPool& p, cudaStream_t st)
{
    k_pool_revive<<<ceil_div(p.n_slots, 256), 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_fill_f4(float4* p, size_t n, float4 v, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    k_fill_f4<<<ceil_div((long long)n, 256), 256, 0, st>>>(p, n, v);
    return cudaGetLastError();
}
cudaError_t launch_fix_to_sum(const long long* fix, float4* sum, size_t n, cudaStream_t st)
{
    if (n) k_fix_to_sum<<<ceil_div((long long)n, 256), 256, 0, st>>>(fix, sum, n);
    return cudaGetLastError();
}
cudaError_t launch_sum_to_fix(const float4* sum, long long* fix, size_t n, cudaStream_t st)
{
    if (n) k_sum_to_fix<<<ceil_div((long long)n, 256), 256, 0, st>>>(sum, fix, n);
    return cudaGetLastError();
}
cudaError_t launch_tonemap(const float4* sum, float4* hdr, float4* ldr, int n, cudaStream_t st)
{
    k_tonemap<<<ceil_div(n, 256), 256, 0, st>>>(sum, hdr, ldr, n);
    return cudaGetLastError();
}
cudaError_t launch_hook_primary(const RenderArgs& a, int jitter_mode, uint32_t rand, float4* ray_o, float4* ray_d, cudaStream_t st)
{
    k_hook_primary<<<ceil_div((long long)a.width * a.height, 256), 256, 0, st>>>(a, jitter_mode, rand, ray_o, ray_d);
    return cudaGetLastError();
}
cudaError_t launch_hook_prepare(const LightSet& L, int n, const float* od6, const float* tmax, int any_hit,
                                float4* ray_o, float4* ray_d, int* light_id, unsigned char* vis, cudaStream_t st)
{
    k_hook_prepare<<<ceil_div(n, 256), 256, 0, st>>>(L, n, od6, tmax, any_hit, ray_o, ray_d, light_id, vis);
    return cudaGetLastError();
}
cudaError_t launch_hook_finish(int n, int any_hit, const float4* ray_o, const float4* hit, const unsigned char* vis,
                               int* tri_id, int* light_id, float* t_hit, cudaStream_t st)
{
    k_hook_finish<<<ceil_div(n, 256), 256, 0, st>>>(n, any_hit, ray_o, hit, vis, tri_id, light_id, t_hit);
    return cudaGetLastError();
}

} // namespace yune
