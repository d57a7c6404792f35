// wavefront_types.h -- argument blocks shared by the kernels and the context (engine.cu).
#ifndef YUNE_WAVEFRONT_TYPES_H
#define YUNE_WAVEFRONT_TYPES_H

#include <cuda_runtime.h>
#include <stdint.h>
#include "lights.h"

namespace yune {

// Scene in its traversal/shading layout (trav_layout.h).  Passed to kernels BY VALUE (constant bank).
struct DevScene {
    const float4* pairs;     // 4 x float4 per inner node, breadth-first
    const float4* tris;      // 3 x float4 per leaf-ordered triangle
    const float4* shade;     // 4 x float4 per original triangle (normalised normals, matID)
    const float4* mats;      // 5 x float4 per material (the 80-byte reference record, untouched)
    const unsigned char* tri_class; // 1 byte per original triangle: 1 = specular material
    const float4* leaf_boxes;// accel 1: 2 x float4 per REFERENCE leaf (its uploaded box), indexed by triangle record t2.w
    int   accel;             // 0 = walk the reference tree, 1 = walk our own tree + exact leaf-box filter (trav_layout.h),
                             // 2 = as 1 over 4-wide records (`pairs` then holds 7 x float4 per record, n_inner counts them)
    int   isect;             // 0 = the reference's Moller-Trumbore + leaf-box filter (parity mode), 1 = watertight test on raw vertices (accel 1 only)
    int   n_inner, n_tris, n_mats, root_ref;
    float root_lo[3], root_hi[3];
    int   n_smem_pairs;      // pair records [0, n_smem_pairs) are staged in shared memory by the trace kernel
};

struct LightSet { LightDev l[YUNE_MAX_LIGHTS]; int n; };

// Per-iteration queue heads/counters (double buffered by iteration parity) + running totals.
// Every counter sits in its own 128-byte line: they are hit by one atomic per block (shade) or per ray chunk (trace), and
// atomics on the same line serialise in one L2 slice.
struct IterCounters {
    alignas(128) int n_extend;
    alignas(128) int n_shadow;
    alignas(128) int n_events;
    alignas(128) int live;
    alignas(128) int fetch_extend;
    alignas(128) int fetch_shadow;
};
struct Totals {
    alignas(128) unsigned long long next_sample;     // next global sample index to hand out (own line: atomics)
    alignas(128) unsigned long long n_samples;       // samples to render in this call
    unsigned long long samples_done;
    unsigned long long extend_rays, shadow_rays, box_tests, tri_tests;
    unsigned long long visits_d, visits_s, visits_r;   // shade-stage rounds by kind (one atomic per block per launch)
    int live_last;                      // slots still busy after the last completed iteration
    int iterations;
};

// path-state flags (meta.w)
#define YS_FREE   0u      // needs a new sample
#define YS_TRACE  1u      // an extension ray is in flight; hit[] holds its answer on the next visit
#define YS_DRAIN  2u      // path ended but NEE answers are still in flight
#define YS_DONE   3u      // no samples left for this slot
#define YS_STATE_MASK 3u
#define YF_PEND_L     4u    // pendL/visL hold an unresolved NEE light sample
#define YF_PREV_SPEC  8u    // the vertex that launched the extension ray was specular (udpt.cl:490)
#define YF_PEND_EVT  16u    // evt_idx points at an MIS event record
#define YF_LID_SHIFT  8     // bits 8..11: 1 + index of the light that bounds the in-flight extension ray (0 = none); a copy of
#define YF_LID_MASK  15u    //   ray_d.w so that the classify phase knows it from the first load (k_shade_dense only)

// One 16-byte field of a 32-byte two-field record: element s lives at p[2 s].  Fields that are read and written by the same
// visit of a slot share a record, so that every 32-byte DRAM sector the shade rounds touch belongs to ONE slot: with plain
// 16-byte arrays a sector holds two neighbouring slots, the neighbour is usually in another list (or idle), and ncu showed
// 1.9x the algorithmic traffic (half-used sector reads, read-modify-write of half-written sectors).
struct F4Half {
    float4* p;
    __host__ __device__ __forceinline__ float4& operator[](size_t s) const { return p[2 * s]; }
};

// Path pool: one entry per slot; 32-byte records (ray_o, ray_d) (thr, thr_next) (col, pend_l), 16-byte arrays meta and hit.
struct PathPool {
    int      n_slots;
    F4Half   ray_o;      // xyz origin, w = ray length so far (t of an analytic light hit, or +inf)
    F4Half   ray_d;      // xyz direction, w = int bits: index of the light that owns that length, or -1
    float4*  hit;        // t, u, v, int bits: original triangle index or -1     (written by the trace kernel)
    F4Half   thr;        // xyz throughput at the current vertex (after Russian roulette)
    F4Half   thr_next;   // xyz throughput once the in-flight extension ray lands on a surface
    F4Half   col;        // xyz radiance gathered so far for the current sample
    F4Half   pend_l;     // xyz unresolved NEE light sample (already MIS-weighted), valid with YF_PEND_L
    uint4*   meta;       // x pixel, y sample, z index of the vertex the extension ray will reach, w flags
    int*     evt_idx;    // valid with YF_PEND_EVT
    unsigned char* vis_l;   // 1 = NEE shadow ray reached the light (written by the trace kernel)
    // queues
    int*     eq;         // extension queue: slot indices
    float4*  sq_o;       // shadow queue: xyz origin, w = tmax
    float4*  sq_d;       // xyz direction, w = int bits: target (>= 0 slot -> vis_l ; < 0 -> ~target = event*4 + which)
    // MIS events (rare): 3 x float4 each: (Lv.xyz, flags), (BV.xyz, -), (BO.xyz, -); answers in evt_vis[4*e + which]
    float4*  evt;
    unsigned char* evt_vis;
};

// Extra per-slot storage of the bidirectional integrator (bdpt.cu).  Stride YB_MAXV = 32 vertices per slot.
struct BdptPool {
    float4* lp;        // light path: 4 x float4 per vertex: (point, triangle bits) (normal, -) (arriving direction, -) (contribution, -)
    float4* pend_c;    // [0] = emission of the eye vertex being resolved; [j] = unresolved contribution of the connection to light vertex j
    int4*   bmeta;     // x = light-path length, y = mask of connection rays in flight, z = bits(eye_path_weight), w = bits(ks)
    int     bounces;   // BDPT_BOUNCES (bdpt.cl:7)
};

#define YE_HAS_S      1     // a shadow ray toward the light sample is in flight (which = 0)
#define YE_HAS_MV     2     // BRDF-sampled ray of the "light sample visible" branch (which = 1)
#define YE_HAS_MO     4     // BRDF-sampled ray of the "light sample occluded" branch (which = 2)
#define YE_MO_IS_MV   8     // both branches sampled the same direction: which = 1 answers for both

struct RenderArgs {
    DevScene  sc;
    LightSet  lights;
    PathPool  pool;
    float4*   sum;            // W*H fp32 RGBA accumulation buffer (rgb sums, a = sample count)
    long long* fix;           // option "deterministic": W*H x 4 fixed-point (2^-24) sums instead; `sum` is derived from it after the render
    IterCounters* ctr;        // [2]
    Totals*   tot;
    float     cam[20];        // the 80-byte Cam record
    int       width, height;
    int       spp_begin;
    uint32_t  seed;
    int       gi_check;
    int       rr_threshold;
    int       mis;
    int       oren_nayar;
    int       parity;         // iteration parity -> which IterCounters to use
    int       count_work;
    int       tail;           // every sample of the job has been handed out: the pool only drains (k_shade_dense skips dead chunks)
    unsigned char* chunk_live;// [ceil(n_slots / 256)] 1 = the chunk may hold a slot in flight (maintained while tail != 0)
};

} // namespace yune
#endif
