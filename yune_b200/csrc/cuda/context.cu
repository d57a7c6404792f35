// context.cu -- the C ABI of include/yune_cuda.h: device/buffer management (what CLManager did with OpenCL,
// src/CLManager.cpp) and the headless frame loop (what RendererCore::enqueueKernels did, src/RendererCore.cpp:248-469).
//
// There is no CPU execution path in this file or behind it: every entry point either drives the CUDA kernels
// of kernels.cu or fails with an error code.
#include "yune_cuda.h"
#include "kernels.h"
#include "trav_layout.h"
#include "bvh_build.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

using namespace yune;

namespace {
thread_local std::string g_setup_err;

const yune_quad_light kBuiltinLightUdpt = {      // udpt.cl:97-106
    {{-0.1979f, 0.92f, -3.1972f, 1.f}}, {{0.f, -1.f, 0.f, 0.f}}, {{16.f, 16.f, 16.f, 0.f}}, {{0.f, 0.f, 0.f, 0.f}},
    {{0.f, 0.f, 0.f, 0.f}}, {{0.4f, 0.f, 0.f, 0.f}}, {{0.f, 0.f, 0.4f, 0.f}}, 0.f, {0.f, 0.f, 0.f}};
const yune_quad_light kBuiltinLightBdpt = {      // bdpt.cl:106-115
    {{-0.1979f, 0.703f, -3.1972f, 1.f}}, {{0.f, 1.f, 0.f, 0.f}}, {{18.3f, 16.2f, 14.5f, 0.f}}, {{0.f, 0.f, 0.f, 0.f}},
    {{0.f, 0.f, 0.f, 0.f}}, {{0.4f, 0.f, 0.f, 0.f}}, {{0.f, 0.f, 0.4f, 0.f}}, 0.f, {0.f, 0.f, 0.f}};

LightDev unpack_light(const yune_quad_light& q)
{
    LightDev L;
    L.pos = v3(q.pos.s[0], q.pos.s[1], q.pos.s[2]);
    L.normal = v3(q.normal.s[0], q.normal.s[1], q.normal.s[2]);
    L.ke = v3(q.ke.s[0], q.ke.s[1], q.ke.s[2]);
    L.edge_l = v3(q.edge_l.s[0], q.edge_l.s[1], q.edge_l.s[2]);
    L.edge_w = v3(q.edge_w.s[0], q.edge_w.s[1], q.edge_w.s[2]);
    L.la = vlength(L.edge_l);        // length(float4) with w = 0 (udpt.cl:259-260)
    L.lb = vlength(L.edge_w);
    return L;
}
}

enum { INTEGRATOR_NONE = 0, INTEGRATOR_UDPT = 1, INTEGRATOR_BDPT = 2 };

struct yune_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    std::vector<cudaEvent_t> ev_pool;          // 3 per sampled iteration: before shade, before trace, after trace
    std::string err;

    // host copies of the reference-layout buffers (re-laid-out lazily when both are present)
    std::vector<yune_triangle> h_tris; std::vector<yune_bvh_node> h_nodes; std::vector<yune_material> h_mats;
    bool layout_dirty = true, have_tris = false, have_nodes = false, have_mats = false, have_cam = false;
    TravLayoutHost lay;
    float4 *d_pairs = nullptr, *d_tris = nullptr, *d_shade = nullptr, *d_mats = nullptr, *d_leaf_boxes = nullptr;
    unsigned char* d_tri_class = nullptr;      // per original triangle: 1 = its material is specular (shade-stage sorting key)
    bool class_dirty = true;
    DevScene sc{};
    int layout_built_on_device = 0;                  // the current layout's own tree came from bvh_build.cu (uploaded tree or device-built one)
    GpuBvh gpu_bvh; bool bvh_on_device = false;      // yune_build_bvh_on_device: the layout arrays below are the builder's (h_nodes is filled on demand)
    yune_cam cam{};

    int integrator = INTEGRATOR_NONE; int mis = 0; bool postproc = false;
    LightSet lights{}; bool user_lights = false;

    int W = 0, H = 0;
    float4 *d_sum = nullptr, *d_hdr = nullptr, *d_ldr = nullptr;
    long long* d_fix = nullptr;                 // option "deterministic": fixed-point accumulation buffer (4 x int64 per pixel); d_sum is derived from it

    PathPool pool{}; int pool_alloc = 0; int pool_integrator = 0;
    BdptPool bdpt{};
    IterCounters* d_ctr = nullptr; Totals* d_tot = nullptr; Totals* h_tot = nullptr;
    unsigned char* d_chunk_live = nullptr;      // one byte per 256 path slots (RenderArgs::chunk_live)

    // hook scratch
    float4 *hk_o = nullptr, *hk_d = nullptr, *hk_hit = nullptr; int *hk_tri = nullptr, *hk_light = nullptr; float *hk_t = nullptr, *hk_od = nullptr, *hk_tmax = nullptr;
    unsigned char* hk_vis = nullptr; int* hk_cnt = nullptr; int hk_cap = 0;

    // ray capture (measurement aid)
    float4 *cap_eo = nullptr, *cap_ed = nullptr, *cap_so = nullptr, *cap_sd = nullptr; int* cap_cnt = nullptr; int cap_alloc = 0;
    int cap_iteration = -1, cap_max = 0; int cap_counts[4] = {0, 0, 0, 0};

    // options
    int opt_pool_slots = 0, opt_smem_nodes = -1, opt_rr_threshold = -1, opt_bdpt_bounces = 20;
    int opt_trace_block = 1024, opt_trace_blocks_per_sm = 0, opt_refill_idle = YUNE_DEF_REFILL_IDLE, opt_phase_min = YUNE_DEF_PHASE_MIN, opt_inner_min = YUNE_DEF_INNER_MIN, opt_inner_chain = YUNE_DEF_INNER_CHAIN;
    int opt_leaf_split = 2, opt_accel = 1, opt_shade_blocks_per_sm = 0;
    int opt_oren_nayar = 0, opt_isect = 0, opt_max_iterations = 1 << 30, opt_count_work = 0, opt_sync_every = 8, opt_time_stages = 0, opt_deterministic = 1;

    // Option "pipeline" (the reference's interactive call pattern, one sample per pixel per call, src/RendererCore.cpp:248-306,
    // 483-486): a call returns as soon as its samples have all been HANDED OUT; the paths still in flight keep their slots and are
    // carried into the next call (or yune_finish), so the pool stays full across calls instead of draining ~130 nearly empty
    // iterations per frame.  `epoch` counts everything that invalidates paths in flight (scene, camera, program, lights, image size,
    // options): a carry from another epoch is discarded.
    int opt_pipeline = 0, opt_device_layout = -1, opt_device_builder = 1, opt_ploc_radius = 32, opt_own_tree_passes = -1;
    bool carry = false; unsigned epoch = 0, carry_epoch = 0; uint32_t carry_seed = 0; int carry_gi = 0;
    unsigned it_global = 0;                     // iterations since the pool was last reset: its parity selects the counter / event buffers

    // Option "sort_rays" = 1: both ray queues are sorted by the Morton cell of the ray origin between the shade and the trace kernel
    // (cub radix sort on the top `sort_bits` bits of 30-bit keys), so that rays which start next to each other walk the same nodes
    // of a tree that does not fit the caches.  MEASURED ON C4 (10.5 M triangles, profiles/r2_ab_c4_sort.log): the trace kernel gains
    // 7 % (3.95 -> 3.67 ms per launch) and keys + sorts cost 0.6 ms: 754 vs 796 Msamples/s.  Off by default.
    int opt_sort_rays = 0, opt_sort_bits = 18;
    unsigned *d_eq_key = nullptr, *d_sq_key = nullptr, *d_key_tmp = nullptr; int *d_eq_sorted = nullptr, *d_sq_iota = nullptr, *d_sq_idx = nullptr;
    void* d_sort_tmp = nullptr; size_t sort_tmp_bytes = 0; int sort_alloc = 0; int* h_cnt = nullptr;

    int occ_dense[2] = {0, 0}, occ_bdpt[2] = {0, 0};                           // shade-kernel occupancy caches
    int tc_variant = -1, tc_block = 0, tc_per_sm = 0; size_t tc_smem = 0;      // trace_config cache

    yune_stats stats{};
};

// ------------------------------------------------------------------------------------------------------------
#define Y_FAIL(ctx, code, ...) do { char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__); (ctx)->err = _b; return (code); } while (0)
#define Y_CUDA(ctx, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { \
        char _b[512]; snprintf(_b, sizeof(_b), "%s in File: %s at Line number: %d (%s)", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); \
        (ctx)->err = _b; return YUNE_ERR_CUDA; } } while (0)

template <class T> static void dfree(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

static void free_sort(yune_ctx* c)
{
    dfree(c->d_eq_key); dfree(c->d_sq_key); dfree(c->d_key_tmp); dfree(c->d_eq_sorted); dfree(c->d_sq_iota); dfree(c->d_sq_idx);
    if (c->d_sort_tmp) { cudaFree(c->d_sort_tmp); c->d_sort_tmp = nullptr; }
    c->sort_alloc = 0; c->sort_tmp_bytes = 0;
}
static int ensure_sort(yune_ctx* c)
{
    const int n = c->pool.n_slots;
    if (c->sort_alloc >= n) return YUNE_OK;
    free_sort(c);
    const size_t N = (size_t)n;
    Y_CUDA(c, cudaMalloc(&c->d_eq_key, N * 4)); Y_CUDA(c, cudaMalloc(&c->d_eq_sorted, N * 4));
    Y_CUDA(c, cudaMalloc(&c->d_sq_key, 3 * N * 4)); Y_CUDA(c, cudaMalloc(&c->d_sq_iota, 3 * N * 4)); Y_CUDA(c, cudaMalloc(&c->d_sq_idx, 3 * N * 4));
    Y_CUDA(c, cudaMalloc(&c->d_key_tmp, 3 * N * 4));
    c->sort_tmp_bytes = sort_pairs_tmp_bytes(3 * n);
    Y_CUDA(c, cudaMalloc(&c->d_sort_tmp, c->sort_tmp_bytes));
    Y_CUDA(c, launch_iota(c->d_sq_iota, 3 * n, c->stream));
    if (!c->h_cnt) Y_CUDA(c, cudaMallocHost(&c->h_cnt, 4 * sizeof(int)));
    c->sort_alloc = n;
    return YUNE_OK;
}

static void free_pool(yune_ctx* c)
{
    PathPool& P = c->pool;
    free_sort(c);
    dfree(P.ray_o.p); P.ray_d.p = nullptr; dfree(P.hit); dfree(P.thr.p); P.thr_next.p = nullptr; dfree(P.col.p); P.pend_l.p = nullptr;
    dfree(c->d_chunk_live);
    dfree(P.meta); dfree(P.evt_idx); dfree(P.vis_l); dfree(P.eq); dfree(P.sq_o); dfree(P.sq_d); dfree(P.evt); dfree(P.evt_vis);
    dfree(c->bdpt.lp); dfree(c->bdpt.pend_c); dfree(c->bdpt.bmeta);
    P.n_slots = 0; c->pool_alloc = 0;
}

// Pool size.  "pool_slots" = 0 (default) sizes the pool for the job: measured on C1 / C2 at 16.8 M ... 1.07 G samples the best
// pool doubles when the job quadruples (2 M, 4 M, 8 M, 16 M slots): a bigger pool amortises the per-iteration costs (launches,
// kernel tails), a smaller one shortens the ramp at both ends of the job.  Hence 512 * sqrt(samples), as a power of two.
static int ensure_pool(yune_ctx* c, unsigned long long n_samples, bool keep)
{
    const bool bd = c->integrator == INTEGRATOR_BDPT;
    if (keep && c->pool.n_slots > 0) return YUNE_OK;      // paths of the previous call are still in flight in these slots
    int n = c->opt_pool_slots;
    if (n <= 0) {
        const double want = 512.0 * std::sqrt((double)n_samples);
        int e = (int)std::lround(std::log2(want > 1.0 ? want : 1.0));
        const int e_max = bd ? 23 : 24;                      // a BDPT slot carries 4 KB of path vertices: 8 M slots = 35 GB of the 180 GB
        if (c->opt_pipeline) e += 2;                         // calls of ONE sample per pixel: the pool is shared by consecutive calls (measured: 2.9 / 2.3 / 2.2 ms per 1024^2 frame at 0.5 / 1 / 2 M slots)
        if (e < 16) e = 16;
        if (e > e_max) e = e_max;
        n = 1 << e;
    }
    PathPool& P = c->pool;
    if (c->pool_integrator == c->integrator && c->pool_alloc >= n) { P.n_slots = n; return YUNE_OK; }   // capacity is kept
    free_pool(c);
    const size_t N = (size_t)n;
    const size_t V = 32;                         // YB_MAXV of bdpt.cu: vertices stored per slot
    const size_t rays_per_slot = bd ? 3 + V : 3; // shadow rays one slot can emit per iteration
    Y_CUDA(c, cudaMalloc(&P.ray_o.p, N * 32)); P.ray_d.p = P.ray_o.p + 1; Y_CUDA(c, cudaMalloc(&P.hit, N * 16));
    Y_CUDA(c, cudaMalloc(&P.thr.p, N * 32)); P.thr_next.p = P.thr.p + 1; Y_CUDA(c, cudaMalloc(&P.col.p, N * 32)); P.pend_l.p = P.col.p + 1;
    Y_CUDA(c, cudaMalloc(&P.meta, N * 16));
    Y_CUDA(c, cudaMalloc(&P.evt_idx, N * 4)); Y_CUDA(c, cudaMalloc(&P.vis_l, bd ? N * (1 + V) : N)); Y_CUDA(c, cudaMalloc(&P.eq, N * 4));
    Y_CUDA(c, cudaMalloc(&P.sq_o, rays_per_slot * N * 16)); Y_CUDA(c, cudaMalloc(&P.sq_d, rays_per_slot * N * 16));
    if (bd) {
        Y_CUDA(c, cudaMalloc(&c->bdpt.lp, N * V * 4 * 16)); Y_CUDA(c, cudaMalloc(&c->bdpt.pend_c, N * V * 16)); Y_CUDA(c, cudaMalloc(&c->bdpt.bmeta, N * 16));
    }
    Y_CUDA(c, cudaMalloc(&P.evt, 2 * 3 * N * 16)); Y_CUDA(c, cudaMalloc(&P.evt_vis, 2 * 4 * N));
    Y_CUDA(c, cudaMalloc(&c->d_chunk_live, N / YUNE_SHADE_BLOCK + 2));
    P.n_slots = n; c->pool_alloc = n; c->pool_integrator = c->integrator;
    return YUNE_OK;
}

static void set_builtin_lights(yune_ctx* c)
{
    if (c->user_lights) return;
    c->lights.n = 1;
    c->lights.l[0] = unpack_light(c->integrator == INTEGRATOR_BDPT ? kBuiltinLightBdpt : kBuiltinLightUdpt);
}

// The layout arrays a device build produced become the context's (freed like the host-built ones); `G` keeps the node array.
static void adopt_device_layout(yune_ctx* c, GpuBvh& G)
{
    dfree(c->d_pairs); dfree(c->d_tris); dfree(c->d_shade); dfree(c->d_leaf_boxes); dfree(c->d_tri_class);
    c->d_pairs = G.pairs; c->d_tris = G.tris; c->d_shade = G.shade; c->d_leaf_boxes = G.leaf_boxes; c->d_tri_class = G.tri_class;
    G.pairs = G.tris = G.shade = G.leaf_boxes = nullptr; G.tri_class = nullptr;
    DevScene& s = c->sc;
    s.pairs = c->d_pairs; s.tris = c->d_tris; s.shade = c->d_shade; s.mats = c->d_mats; s.leaf_boxes = c->d_leaf_boxes; s.tri_class = c->d_tri_class;
    s.accel = 1; s.isect = 0; s.n_inner = G.n_inner; s.n_tris = G.n_tris; s.n_mats = (int)c->h_mats.size(); s.root_ref = G.root_ref;
    for (int k = 0; k < 3; k++) { s.root_lo[k] = G.root_lo[k]; s.root_hi[k] = G.root_hi[k]; }
}

// (re)build the traversal layout and upload it
static int ensure_scene(yune_ctx* c)
{
    if (!c->have_tris || !c->have_nodes || !c->have_mats)
        Y_FAIL(c, YUNE_ERR_STATE, "scene incomplete: vertex, material and BVH buffers must all be set up before rendering");
    if (c->class_dirty || c->layout_dirty) {
        for (const yune_triangle& t : c->h_tris)
            if (t.matID < 0 || t.matID >= (int)c->h_mats.size()) Y_FAIL(c, YUNE_ERR_INVALID, "triangle references material %d of %d", t.matID, (int)c->h_mats.size());
        std::vector<unsigned char> cls(std::max<size_t>(c->h_tris.size(), 1), 0);
        for (size_t i = 0; i < c->h_tris.size(); i++) cls[i] = c->h_mats[c->h_tris[i].matID].is_specular != 0 ? 1 : 0;
        dfree(c->d_tri_class);
        Y_CUDA(c, cudaMalloc(&c->d_tri_class, cls.size()));
        Y_CUDA(c, cudaMemcpyAsync(c->d_tri_class, cls.data(), cls.size(), cudaMemcpyHostToDevice, c->stream));
        Y_CUDA(c, cudaStreamSynchronize(c->stream));
        c->sc.tri_class = c->d_tri_class;
        c->class_dirty = false;
    }
    if (!c->layout_dirty) return YUNE_OK;
    std::string err;
    if (c->bvh_on_device) {
        // an option that changes the layout (accel, leaf_split, isect) was set after the device build: the host re-layout takes
        // over from the device-built tree, which is a valid reference-format array like any other
        c->h_nodes.resize((size_t)c->gpu_bvh.n_nodes);
        Y_CUDA(c, cudaMemcpy(c->h_nodes.data(), c->gpu_bvh.nodes, c->h_nodes.size() * sizeof(yune_bvh_node), cudaMemcpyDeviceToHost));
        c->gpu_bvh.free_all(); c->bvh_on_device = false;
    }
    // The walk's OWN tree (accel 1) is built on the device (bvh_build.cu, PLOC) rather than by the host's binned-SAH builder: at
    // 10.5 M triangles the host re-layout is 5 s of every upload against 0.03 s, and the tree is as good (C4: trace 1.77 vs 1.76 ms
    // per launch) or better (C2: 0.735 vs 0.749 ms).  The uploaded tree still decides every hit: the triangle records carry its
    // leaves / visiting ranks, the leaf-box filter tests its boxes (the soundness argument does not depend on the own tree's
    // topology).  Option "device_layout": 1 = on the device when the uploaded tree allows it, 0 = on the host, -1 (default) = on the
    // device above 2^16 triangles; below, the host builder + reinsertion passes (relayout.cpp: OptTree) give the better tree for a
    // few milliseconds (teapot: 25.2 -> 23.4 box tests per ray).
    // A tree the device builder cannot take (deeper than the traversal stack) falls back to the host path.
    if (c->opt_accel == 1 && c->opt_isect == 0 && !c->h_nodes.empty() && !c->h_tris.empty()
        && (c->opt_device_layout == 1 || (c->opt_device_layout < 0 && c->h_tris.size() > ((size_t)1 << 16)))) {
        std::vector<int> leaf_of_tri, rank_of_tri; std::vector<F4> leaf_boxes; bool usable = false;
        if (!referenceLeavesForDevice(c->h_tris.data(), (int)c->h_tris.size(), c->h_nodes.data(), (int)c->h_nodes.size(), leaf_of_tri, rank_of_tri, leaf_boxes, usable, err))
            Y_FAIL(c, YUNE_ERR_LIMIT, "BVH/triangle buffers rejected: %s", err.c_str());
        if (usable) {
            RefLeaves R; R.leaf_of_tri = leaf_of_tri.data(); R.rank_of_tri = rank_of_tri.data(); R.leaf_boxes = &leaf_boxes[0].x; R.n_leaves = (int)(leaf_boxes.size() / 2);
            GpuBvh G;
            if (buildBvhOnDevice(c->h_tris.data(), (int)c->h_tris.size(), reinterpret_cast<const yune_material*>(c->d_mats), (int)c->h_mats.size(), c->opt_leaf_split > 0 ? c->opt_leaf_split : 2, c->stream, G, err, &R, c->opt_device_builder, c->opt_ploc_radius)) {
                adopt_device_layout(c, G);
                c->lay = TravLayoutHost(); c->lay.accel = 1; c->lay.n_inner = G.n_inner; c->lay.n_tris = G.n_tris; c->lay.max_depth = G.depth;
                c->layout_dirty = false; c->class_dirty = false; c->layout_built_on_device = 1;
                return YUNE_OK;
            }
            G.free_all();
            cudaGetLastError();
            err.clear();
        }
    }
    c->layout_built_on_device = 0;
    if (!buildTravLayout(c->h_tris.data(), (int)c->h_tris.size(), c->h_nodes.data(), (int)c->h_nodes.size(), c->lay, err, c->opt_leaf_split, c->opt_accel, c->opt_isect, c->opt_own_tree_passes))
        Y_FAIL(c, YUNE_ERR_LIMIT, "BVH/triangle buffers rejected: %s", err.c_str());
    dfree(c->d_pairs); dfree(c->d_tris); dfree(c->d_shade); dfree(c->d_leaf_boxes);
    const TravLayoutHost& L = c->lay;
    Y_CUDA(c, cudaMalloc(&c->d_leaf_boxes, std::max<size_t>(L.leaf_boxes.size(), 2) * 16));
    if (!L.leaf_boxes.empty()) Y_CUDA(c, cudaMemcpyAsync(c->d_leaf_boxes, L.leaf_boxes.data(), L.leaf_boxes.size() * 16, cudaMemcpyHostToDevice, c->stream));
    Y_CUDA(c, cudaMalloc(&c->d_pairs, std::max<size_t>(L.pairs.size(), 4) * 16));
    Y_CUDA(c, cudaMalloc(&c->d_tris, std::max<size_t>(L.tris.size(), 3) * 16));
    Y_CUDA(c, cudaMalloc(&c->d_shade, std::max<size_t>(L.shade.size(), 4) * 16));
    const std::vector<F4>& node_records = L.accel == 2 ? L.quads : L.pairs;       // accel 2: the kernel walks the 4-wide records instead
    if (L.accel == 2) { dfree(c->d_pairs); Y_CUDA(c, cudaMalloc(&c->d_pairs, std::max<size_t>(node_records.size(), 7) * 16)); }
    if (!node_records.empty()) Y_CUDA(c, cudaMemcpyAsync(c->d_pairs, node_records.data(), node_records.size() * 16, cudaMemcpyHostToDevice, c->stream));
    if (!L.tris.empty()) Y_CUDA(c, cudaMemcpyAsync(c->d_tris, L.tris.data(), L.tris.size() * 16, cudaMemcpyHostToDevice, c->stream));
    if (!L.shade.empty()) Y_CUDA(c, cudaMemcpyAsync(c->d_shade, L.shade.data(), L.shade.size() * 16, cudaMemcpyHostToDevice, c->stream));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    DevScene& s = c->sc;
    s.pairs = c->d_pairs; s.tris = c->d_tris; s.shade = c->d_shade; s.mats = c->d_mats; s.leaf_boxes = c->d_leaf_boxes; s.accel = L.accel; s.isect = L.isect;
    s.n_inner = L.n_inner; s.n_tris = L.n_tris; s.n_mats = (int)c->h_mats.size(); s.root_ref = L.root_ref;
    if (L.accel == 2) { s.n_inner = L.n_wide; s.root_ref = L.root_wide_ref; }
    for (int k = 0; k < 3; k++) { s.root_lo[k] = L.root_lo[k]; s.root_hi[k] = L.root_hi[k]; }
    c->layout_dirty = false;
    return YUNE_OK;
}

// new geometry / a new uploaded tree: the device-built one (and the layout arrays it owns) goes
static void drop_device_bvh(yune_ctx* c)
{
    if (!c->bvh_on_device) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->gpu_bvh.free_all(); c->bvh_on_device = false;
    c->class_dirty = true; c->have_nodes = false;
}

struct TraceLaunch { int grid, block; size_t smem; };
static int trace_config(yune_ctx* c, TraceLaunch& tl, bool count)
{
    // "smem_nodes" < 0 (default): the whole tree if it fits (the kernel variant without a global node path, +3.5 % on C2), else
    // the first 2340 records -- a bigger staging area would take the L1 capacity that triangles and stacks need (measured).
    const bool wide = c->sc.accel == 2;                       // 4-wide records: 112 B each, half as many
    tl.block = c->opt_trace_block;
    const int ray_bytes = wide ? 32 * tl.block : 0;           // the wide kernel keeps (o, d, u, v) of every lane in shared memory
    const int rec_bytes = wide ? 112 : 56, rec_max = wide ? (232448 - 1024 - ray_bytes) / 112 : 3900, rec_part = wide ? 1170 : 2340;
    int want = c->opt_smem_nodes;
    if (want < 0) want = c->sc.n_inner <= rec_max ? c->sc.n_inner : rec_part;
    int n_smem = want < c->sc.n_inner ? want : c->sc.n_inner;
    if (n_smem > rec_max) n_smem = rec_max;                   // 3900 * 56 B = 1950 * 112 B = 213 KB < 227 KB
    c->sc.n_smem_pairs = n_smem;
    tl.smem = (size_t)n_smem * rec_bytes + ray_bytes;         // 48 B of boxes + 8 B of child refs per staged record
    if (tl.smem < 16) tl.smem = 16;
    // the attribute and the occupancy belong to ONE instantiation, block size and staging size: asked once per combination and
    // kept in the context (not in function statics: contexts live on different devices)
    const bool def_knobs = c->opt_refill_idle == YUNE_DEF_REFILL_IDLE && c->opt_phase_min == YUNE_DEF_PHASE_MIN && c->opt_inner_min == YUNE_DEF_INNER_MIN && c->opt_inner_chain == YUNE_DEF_INNER_CHAIN;
    const int variant = trace_variant_id(c->sc, count, def_knobs);
    if (c->tc_variant != variant || c->tc_block != tl.block || c->tc_smem != tl.smem) {
        int per_sm = 0;
        Y_CUDA(c, trace_prepare(c->sc, count, def_knobs, tl.block, tl.smem, &per_sm));
        if (per_sm < 1) Y_FAIL(c, YUNE_ERR_CUDA, "trace kernel does not fit on an SM with %zu bytes of shared memory", tl.smem);
        c->tc_variant = variant; c->tc_block = tl.block; c->tc_smem = tl.smem; c->tc_per_sm = per_sm;
    }
    int per_sm = c->tc_per_sm;
    if (c->opt_trace_blocks_per_sm > 0 && per_sm > c->opt_trace_blocks_per_sm) per_sm = c->opt_trace_blocks_per_sm;
    tl.grid = c->sm_count * per_sm;
    return YUNE_OK;
}

static RenderArgs make_args(yune_ctx* c)
{
    RenderArgs a{};
    a.sc = c->sc; a.lights = c->lights; a.pool = c->pool; a.sum = c->d_sum; a.ctr = c->d_ctr; a.tot = c->d_tot;
    std::memcpy(a.cam, &c->cam, 80);
    a.width = c->W; a.height = c->H;
    a.rr_threshold = c->opt_rr_threshold >= 0 ? c->opt_rr_threshold : (c->integrator == INTEGRATOR_BDPT ? 4 : 6);
    a.mis = c->mis; a.oren_nayar = c->opt_oren_nayar; a.count_work = c->opt_count_work;
    return a;
}

// ------------------------------------------------------------------------------------------------------------
extern "C" {

const char* yune_last_error(const yune_ctx* ctx) { return ctx ? ctx->err.c_str() : g_setup_err.c_str(); }

int yune_setup(int device, yune_ctx** out)
{
    if (!out) { g_setup_err = "yune_setup: out_ctx is NULL"; return YUNE_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_setup_err = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return YUNE_ERR_NODEVICE;
    }
    if (device < 0 || device >= n) { g_setup_err = "yune_setup: device index out of range"; return YUNE_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_setup_err = cudaGetErrorString(e); return YUNE_ERR_CUDA; }
    if (prop.major != 10) {
        char b[256]; snprintf(b, sizeof(b), "device %d (%s) is sm_%d%d; this build contains sm_100a code only", device, prop.name, prop.major, prop.minor);
        g_setup_err = b; return YUNE_ERR_NODEVICE;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_setup_err = cudaGetErrorString(e); return YUNE_ERR_CUDA; }
    yune_ctx* c = new yune_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess
           && cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess && cudaEventCreate(&c->ev2) == cudaSuccess
           && cudaMalloc(&c->d_ctr, 2 * sizeof(IterCounters)) == cudaSuccess && cudaMalloc(&c->d_tot, sizeof(Totals)) == cudaSuccess
           && cudaMallocHost(&c->h_tot, sizeof(Totals)) == cudaSuccess;
    if (!ok) { g_setup_err = std::string("context allocation failed: ") + cudaGetErrorString(cudaGetLastError()); yune_destroy(c); return YUNE_ERR_CUDA; }
    c->integrator = INTEGRATOR_UDPT; set_builtin_lights(c);
    *out = c;
    return YUNE_OK;
}

void yune_destroy(yune_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_pool(c);
    c->gpu_bvh.free_all();
    dfree(c->d_pairs); dfree(c->d_tris); dfree(c->d_shade); dfree(c->d_mats); dfree(c->d_leaf_boxes); dfree(c->d_tri_class);
    dfree(c->d_sum); dfree(c->d_hdr); dfree(c->d_ldr); dfree(c->d_fix); dfree(c->d_ctr); dfree(c->d_tot);
    dfree(c->cap_eo); dfree(c->cap_ed); dfree(c->cap_so); dfree(c->cap_sd); dfree(c->cap_cnt);
    dfree(c->hk_o); dfree(c->hk_d); dfree(c->hk_hit); dfree(c->hk_tri); dfree(c->hk_light); dfree(c->hk_t); dfree(c->hk_od); dfree(c->hk_tmax); dfree(c->hk_vis); dfree(c->hk_cnt);
    if (c->h_tot) cudaFreeHost(c->h_tot);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    for (auto& e : c->ev_pool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static bool name_is(const char* s, const char* base)
{
    if (!s) return false;
    std::string a(s);
    size_t slash = a.find_last_of('/'); if (slash != std::string::npos) a = a.substr(slash + 1);
    return a == base || a == std::string(base) + ".cl";
}

int yune_create_render_program(yune_ctx* c, const char* kernel, const char* opts)
{
    if (!c) return YUNE_ERR_INVALID;
    c->epoch++;
    int integ = name_is(kernel, "udpt") ? INTEGRATOR_UDPT : name_is(kernel, "bdpt") ? INTEGRATOR_BDPT : INTEGRATOR_NONE;
    if (integ == INTEGRATOR_NONE) Y_FAIL(c, YUNE_ERR_INVALID, "unknown render kernel '%s' (built-in: udpt.cl, bdpt.cl)", kernel ? kernel : "(null)");
    int mis = 0;
    if (opts && *opts) {
        if (std::strcmp(opts, "-DMIS") == 0 || std::strcmp(opts, "-D MIS") == 0) mis = 1;
        else Y_FAIL(c, YUNE_ERR_INVALID, "unsupported compiler-opts '%s' (supported: -DMIS)", opts);
    }
    c->integrator = integ; c->mis = mis;
    set_builtin_lights(c);
    return YUNE_OK;
}

int yune_create_postproc_program(yune_ctx* c, const char* kernel, const char* opts)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!name_is(kernel, "tonemap")) Y_FAIL(c, YUNE_ERR_INVALID, "unknown post-processing kernel '%s' (built-in: tonemap.cl)", kernel ? kernel : "(null)");
    if (opts && *opts) Y_FAIL(c, YUNE_ERR_INVALID, "unsupported compiler-opts '%s'", opts);
    c->postproc = true;
    return YUNE_OK;
}

int yune_setup_vertex_buffer(yune_ctx* c, const yune_triangle* tris, int n)
{
    if (!c || n < 0 || (n > 0 && !tris)) { if (c) c->err = "yune_setup_vertex_buffer: bad arguments"; return YUNE_ERR_INVALID; }
    c->epoch++;
    c->h_tris.assign(tris, tris + n);
    drop_device_bvh(c);
    c->have_tris = true; c->layout_dirty = true;
    return YUNE_OK;
}

int yune_setup_mat_buffer(yune_ctx* c, const yune_material* mats, int n)
{
    if (!c || n <= 0 || !mats) { if (c) c->err = "yune_setup_mat_buffer: bad arguments"; return YUNE_ERR_INVALID; }
    c->epoch++;
    Y_CUDA(c, cudaSetDevice(c->device));
    c->h_mats.assign(mats, mats + n);
    dfree(c->d_mats);
    Y_CUDA(c, cudaMalloc(&c->d_mats, (size_t)n * 80));
    Y_CUDA(c, cudaMemcpyAsync(c->d_mats, mats, (size_t)n * 80, cudaMemcpyHostToDevice, c->stream));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    c->sc.mats = c->d_mats; c->sc.n_mats = n;
    c->have_mats = true; c->class_dirty = true;
    return YUNE_OK;
}

int yune_setup_bvh_buffer(yune_ctx* c, const yune_bvh_node* nodes, int n)
{
    if (!c || n < 0 || (n > 0 && !nodes)) { if (c) c->err = "yune_setup_bvh_buffer: bad arguments"; return YUNE_ERR_INVALID; }
    c->epoch++;
    // n == 0: the reference's brute-force mode (kernel arg 6 bvh_size == 0, udpt.cl:280-284) -- same hits, see relayout.cpp
    if (n > 0) c->h_nodes.assign(nodes, nodes + n); else c->h_nodes.clear();
    drop_device_bvh(c);
    c->have_nodes = true; c->layout_dirty = true;
    return YUNE_OK;
}

int yune_setup_camera_buffer(yune_ctx* c, const yune_cam* cam)
{
    if (!c || !cam) { if (c) c->err = "yune_setup_camera_buffer: bad arguments"; return YUNE_ERR_INVALID; }
    c->epoch++;
    c->cam = *cam; c->have_cam = true;
    return YUNE_OK;
}

int yune_setup_image_buffers(yune_ctx* c, int W, int H)
{
    if (!c || W <= 0 || H <= 0 || (long long)W * H > (1ll << 30)) { if (c) c->err = "yune_setup_image_buffers: bad size"; return YUNE_ERR_INVALID; }
    c->epoch++;
    Y_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)W * H;
    if (c->d_sum && c->W == W && c->H == H) {          // same size: keep the allocation (and any pointer handed out), just clear
        Y_CUDA(c, cudaMemsetAsync(c->d_sum, 0, n * 16, c->stream));
        if (c->d_fix) Y_CUDA(c, cudaMemsetAsync(c->d_fix, 0, n * 32, c->stream));
        Y_CUDA(c, cudaStreamSynchronize(c->stream));
        return YUNE_OK;
    }
    dfree(c->d_sum); dfree(c->d_hdr); dfree(c->d_ldr); dfree(c->d_fix);
    Y_CUDA(c, cudaMalloc(&c->d_sum, n * 16)); Y_CUDA(c, cudaMalloc(&c->d_hdr, n * 16)); Y_CUDA(c, cudaMalloc(&c->d_ldr, n * 16));
    Y_CUDA(c, cudaMemsetAsync(c->d_sum, 0, n * 16, c->stream));
    Y_CUDA(c, cudaMemsetAsync(c->d_hdr, 0, n * 16, c->stream));
    Y_CUDA(c, cudaMemsetAsync(c->d_ldr, 0, n * 16, c->stream));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    c->W = W; c->H = H;
    return YUNE_OK;
}

int yune_set_light_sources(yune_ctx* c, const yune_quad_light* lights, int n)
{
    if (!c || n < 0 || (n > 0 && !lights)) { if (c) c->err = "yune_set_light_sources: bad arguments"; return YUNE_ERR_INVALID; }
    c->epoch++;
    if (n > YUNE_MAX_LIGHTS) Y_FAIL(c, YUNE_ERR_LIMIT, "at most %d quad lights are supported", YUNE_MAX_LIGHTS);
    if (n == 0) { c->user_lights = false; set_builtin_lights(c); return YUNE_OK; }
    c->lights.n = n;
    for (int i = 0; i < n; i++) c->lights.l[i] = unpack_light(lights[i]);
    c->user_lights = true;
    return YUNE_OK;
}

static int* option_slot(yune_ctx* c, const char* key)
{
    if (!key) return nullptr;
    struct { const char* k; int* p; } tab[] = {
        {"pool_slots", &c->opt_pool_slots}, {"pool_slots_in_use", &c->pool.n_slots}, {"layout_built_on_device", &c->layout_built_on_device}, {"smem_nodes", &c->opt_smem_nodes}, {"rr_threshold", &c->opt_rr_threshold},
        {"bdpt_bounces", &c->opt_bdpt_bounces}, {"oren_nayar", &c->opt_oren_nayar}, {"isect", &c->opt_isect},
        {"max_iterations", &c->opt_max_iterations}, {"count_work", &c->opt_count_work}, {"sync_every", &c->opt_sync_every},
        {"time_stages", &c->opt_time_stages}, {"trace_block", &c->opt_trace_block}, {"trace_blocks_per_sm", &c->opt_trace_blocks_per_sm},
        {"refill_idle", &c->opt_refill_idle}, {"phase_min", &c->opt_phase_min}, {"inner_min", &c->opt_inner_min}, {"inner_chain", &c->opt_inner_chain}, {"leaf_split", &c->opt_leaf_split}, {"shade_blocks_per_sm", &c->opt_shade_blocks_per_sm}, {"accel", &c->opt_accel}, {"deterministic", &c->opt_deterministic}, {"pipeline", &c->opt_pipeline}, {"device_layout", &c->opt_device_layout}, {"device_builder", &c->opt_device_builder}, {"ploc_radius", &c->opt_ploc_radius}, {"own_tree_passes", &c->opt_own_tree_passes}, {"sort_rays", &c->opt_sort_rays}, {"sort_bits", &c->opt_sort_bits},
    };
    for (auto& t : tab) if (std::strcmp(t.k, key) == 0) return t.p;
    return nullptr;
}
int yune_set_option(yune_ctx* c, const char* key, double value)
{
    if (!c) return YUNE_ERR_INVALID;
    int* p = option_slot(c, key);
    if (!p) Y_FAIL(c, YUNE_ERR_INVALID, "unknown option '%s'", key ? key : "(null)");
    const int v = (int)value;
    if (p == &c->opt_pool_slots && v != 0 && (v < 1024 || v > (1 << 26))) Y_FAIL(c, YUNE_ERR_INVALID, "pool_slots must be 0 (sized per job) or in [1024, 2^26]");
    if (p == &c->pool.n_slots || p == &c->layout_built_on_device) Y_FAIL(c, YUNE_ERR_INVALID, "%s is read-only", key);
    if (p == &c->opt_trace_block && (v < 32 || v > YUNE_TRACE_MAX_BLOCK || (v & 31))) Y_FAIL(c, YUNE_ERR_INVALID, "trace_block must be a multiple of 32 in [32, 1024]");
    if ((p == &c->opt_refill_idle || p == &c->opt_phase_min) && (v < 1 || v > 32) || (p == &c->opt_inner_min && (v < 1 || v > 33)) || (p == &c->opt_inner_chain && (v < 0 || v > 64))) Y_FAIL(c, YUNE_ERR_INVALID, "refill_idle / phase_min must be in [1, 32]");
    if (p == &c->opt_bdpt_bounces && (v < 2 || v > 32)) Y_FAIL(c, YUNE_ERR_INVALID, "bdpt_bounces must be in [2, 32]");
    if (p == &c->opt_sync_every && v < 1) Y_FAIL(c, YUNE_ERR_INVALID, "sync_every must be >= 1");
    if (p == &c->opt_isect) { if (v != 0 && v != 1) Y_FAIL(c, YUNE_ERR_INVALID, "isect must be 0 (the reference's Moller-Trumbore, bit-exact hit records) or 1 (watertight, perf mode; needs accel 1)"); if (v != *p) c->layout_dirty = true; }
    if (p == &c->opt_accel) { if (v != 0 && v != 1 && v != 2) Y_FAIL(c, YUNE_ERR_INVALID, "accel must be 0 (walk the reference tree), 1 (own tree + exact leaf-box filter) or 2 (own tree, 4-wide records)"); if (v != *p) c->layout_dirty = true; }
    if (p == &c->opt_device_layout && v != *p) c->layout_dirty = true;
    if (p == &c->opt_sort_bits && (v < 1 || v > 30)) Y_FAIL(c, YUNE_ERR_INVALID, "sort_bits must be in [1, 30]");
    if (p == &c->opt_own_tree_passes) { if (v < -1 || v > 16) Y_FAIL(c, YUNE_ERR_INVALID, "own_tree_passes must be in [-1, 16]"); if (v != *p) c->layout_dirty = true; }
    if (p == &c->opt_ploc_radius) { if (v < 1 || v > 64) Y_FAIL(c, YUNE_ERR_INVALID, "ploc_radius must be in [1, 64]"); if (v != *p) c->layout_dirty = true; }
    if (p == &c->opt_device_builder) { if (v != 0 && v != 1) Y_FAIL(c, YUNE_ERR_INVALID, "device_builder must be 0 (linear BVH) or 1 (PLOC)"); if (v != *p) c->layout_dirty = true; }
    if (p == &c->opt_leaf_split) { if (v < 0 || v > 10) Y_FAIL(c, YUNE_ERR_INVALID, "leaf_split must be in [0, 10]"); if (v != *p) c->layout_dirty = true; }
    if (p == &c->opt_deterministic && (v != 0) != (*p != 0) && c->d_sum) {
        // switching modes carries the image over: the fixed-point buffer is (re)built from the float sums, or dropped
        Y_CUDA(c, cudaSetDevice(c->device));
        const size_t n = (size_t)c->W * c->H;
        if (v != 0) {
            if (!c->d_fix) Y_CUDA(c, cudaMalloc(&c->d_fix, n * 32));
            Y_CUDA(c, launch_sum_to_fix(c->d_sum, c->d_fix, n, c->stream));
            Y_CUDA(c, cudaStreamSynchronize(c->stream));
        }
    }
    // anything but the pure measurement / call-pattern knobs invalidates paths that a pipelined call left in flight
    if (v != *p && p != &c->opt_pipeline && p != &c->opt_time_stages && p != &c->opt_max_iterations && p != &c->opt_sync_every) c->epoch++;
    *p = v;
    return YUNE_OK;
}
int yune_get_option(yune_ctx* c, const char* key, double* value)
{
    if (!c || !value) return YUNE_ERR_INVALID;
    int* p = option_slot(c, key);
    if (!p) Y_FAIL(c, YUNE_ERR_INVALID, "unknown option '%s'", key ? key : "(null)");
    *value = *p;
    return YUNE_OK;
}

// ---- BVH construction on the device (bvh_build.cu) ----
int yune_build_bvh_on_device(yune_ctx* c, int leaf_max)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!c->have_tris || c->h_tris.empty()) Y_FAIL(c, YUNE_ERR_STATE, "yune_build_bvh_on_device: set up the vertex buffer first");
    if (!c->have_mats) Y_FAIL(c, YUNE_ERR_STATE, "yune_build_bvh_on_device: set up the material buffer first");
    if (c->opt_accel != 1 || c->opt_isect != 0) Y_FAIL(c, YUNE_ERR_STATE, "yune_build_bvh_on_device emits the layout of accel 1 / isect 0 (the defaults); set those options back first");
    for (const yune_triangle& t : c->h_tris)
        if (t.matID < 0 || t.matID >= (int)c->h_mats.size()) Y_FAIL(c, YUNE_ERR_INVALID, "triangle references material %d of %d", t.matID, (int)c->h_mats.size());
    Y_CUDA(c, cudaSetDevice(c->device));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    c->epoch++;
    drop_device_bvh(c);
    std::string err;
    if (!buildBvhOnDevice(c->h_tris.data(), (int)c->h_tris.size(), reinterpret_cast<const yune_material*>(c->d_mats), (int)c->h_mats.size(), leaf_max, c->stream, c->gpu_bvh, err, nullptr, c->opt_device_builder, c->opt_ploc_radius)) {
        c->gpu_bvh.free_all();
        Y_FAIL(c, YUNE_ERR_LIMIT, "device BVH build failed: %s", err.c_str());
    }
    GpuBvh& G = c->gpu_bvh;
    c->bvh_on_device = true; c->h_nodes.clear();
    adopt_device_layout(c, G);
    c->lay = TravLayoutHost(); c->lay.accel = 1; c->lay.n_inner = G.n_inner; c->lay.n_tris = G.n_tris; c->lay.max_depth = G.depth;
    c->have_nodes = true; c->layout_dirty = false; c->class_dirty = false; c->layout_built_on_device = 1;
    return YUNE_OK;
}

int yune_bvh_info(yune_ctx* c, int* n_nodes, int* n_inner, int* depth, float* build_ms)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!c->have_nodes) Y_FAIL(c, YUNE_ERR_STATE, "no BVH has been set up or built");
    if (n_nodes) *n_nodes = c->bvh_on_device ? c->gpu_bvh.n_nodes : (int)c->h_nodes.size();
    if (n_inner) *n_inner = c->bvh_on_device ? c->gpu_bvh.n_inner : c->lay.n_inner;
    if (depth) *depth = c->bvh_on_device ? c->gpu_bvh.depth : c->lay.max_depth;
    if (build_ms) *build_ms = c->bvh_on_device ? c->gpu_bvh.build_ms : 0.0f;
    return YUNE_OK;
}

int yune_read_bvh_buffer(yune_ctx* c, yune_bvh_node* nodes, int capacity)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!c->have_nodes) Y_FAIL(c, YUNE_ERR_STATE, "no BVH has been set up or built");
    const int n = c->bvh_on_device ? c->gpu_bvh.n_nodes : (int)c->h_nodes.size();
    if (!nodes || capacity < n) Y_FAIL(c, YUNE_ERR_INVALID, "yune_read_bvh_buffer: the array holds %d nodes, capacity %d", n, capacity);
    if (c->bvh_on_device) {
        Y_CUDA(c, cudaSetDevice(c->device));
        Y_CUDA(c, cudaMemcpy(nodes, c->gpu_bvh.nodes, (size_t)n * sizeof(yune_bvh_node), cudaMemcpyDeviceToHost));
    } else if (n > 0) std::memcpy(nodes, c->h_nodes.data(), (size_t)n * sizeof(yune_bvh_node));
    return YUNE_OK;
}

// One call of the frame loop.  `wait_all`: return only when no path is in flight (the default contract); otherwise (option
// "pipeline") return as soon as every sample of this call has been handed out -- the rest is carried (see yune_ctx::carry).
static int render_impl(yune_ctx* c, int spp_begin, int spp_count, int gi_check, uint32_t seed, int reset, bool wait_all)
{
    if (spp_begin < 0 || spp_count < 0) Y_FAIL(c, YUNE_ERR_INVALID, "yune_render: negative sample range");
    if ((long long)spp_begin + (long long)spp_count > 2147483647ll) Y_FAIL(c, YUNE_ERR_INVALID, "yune_render: spp_begin + spp_count exceeds INT_MAX");
    if (c->integrator != INTEGRATOR_UDPT && c->integrator != INTEGRATOR_BDPT) Y_FAIL(c, YUNE_ERR_STATE, "yune_render: no render program selected");
    if (!c->d_sum) Y_FAIL(c, YUNE_ERR_STATE, "yune_render: image buffers not set up");
    if (!c->have_cam) Y_FAIL(c, YUNE_ERR_STATE, "yune_render: camera buffer not set up");
    Y_CUDA(c, cudaSetDevice(c->device));
    int rc;
    // paths a pipelined call left in flight continue in this call, unless the image is reset or anything they depend on changed
    bool cont = c->carry && !reset && c->carry_epoch == c->epoch && c->pool_integrator == c->integrator;
    if (cont && (seed != c->carry_seed || (gi_check != 0) != (c->carry_gi != 0))) {
        // they were started under another seed / GI switch: finish them under those before this call's samples start
        if ((rc = render_impl(c, 0, 0, c->carry_gi, c->carry_seed, 0, true)) != YUNE_OK) return rc;
        cont = false;
    }
    c->carry = false;
    if ((rc = ensure_scene(c)) != YUNE_OK) return rc;
    if ((rc = ensure_pool(c, (unsigned long long)c->W * c->H * (unsigned long long)spp_count, cont)) != YUNE_OK) return rc;
    TraceLaunch tl;
    if ((rc = trace_config(c, tl, c->opt_count_work != 0)) != YUNE_OK) return rc;

    const size_t n_pix = (size_t)c->W * c->H;
    const bool det = c->opt_deterministic != 0;
    if (det && !c->d_fix) {
        Y_CUDA(c, cudaMalloc(&c->d_fix, n_pix * 32));
        Y_CUDA(c, launch_sum_to_fix(c->d_sum, c->d_fix, n_pix, c->stream));
    }
    if (reset) {
        Y_CUDA(c, cudaMemsetAsync(c->d_sum, 0, n_pix * 16, c->stream));
        if (det) Y_CUDA(c, cudaMemsetAsync(c->d_fix, 0, n_pix * 32, c->stream));
    }
    std::memset(c->h_tot, 0, sizeof(Totals));
    c->h_tot->n_samples = (unsigned long long)n_pix * (unsigned long long)spp_count;
    c->h_tot->live_last = 1;
    Y_CUDA(c, cudaMemcpyAsync(c->d_tot, c->h_tot, sizeof(Totals), cudaMemcpyHostToDevice, c->stream));
    if (!cont) {
        Y_CUDA(c, cudaMemsetAsync(c->d_ctr, 0, 2 * sizeof(IterCounters), c->stream));
        Y_CUDA(c, launch_pool_reset(c->pool, c->stream));
        c->it_global = 0;
    } else if (spp_count > 0) Y_CUDA(c, launch_pool_revive(c->pool, c->stream));      // slots the previous call's drain retired
    Y_CUDA(c, cudaMemsetAsync(c->d_chunk_live, 1, (size_t)c->pool.n_slots / YUNE_SHADE_BLOCK + 2, c->stream));

    // sorted ray queues (see yune_ctx::opt_sort_rays); the unidirectional integrator only
    const bool sort_rays = c->integrator == INTEGRATOR_UDPT && c->opt_sort_rays == 1;
    if (sort_rays) { if ((rc = ensure_sort(c)) != YUNE_OK) return rc; }
    RenderArgs a = make_args(c);
    a.tail = 0; a.chunk_live = c->d_chunk_live; a.fix = det ? c->d_fix : nullptr;
    a.spp_begin = spp_begin; a.seed = seed; a.gi_check = gi_check;
    c->bdpt.bounces = c->opt_bdpt_bounces;
    TraceArgs t{};
    t.sc = c->sc; t.eq = c->pool.eq; t.ray_o = c->pool.ray_o.p; t.ray_d = c->pool.ray_d.p; t.ray_stride = 2; t.hit = c->pool.hit;
    t.sq_o = c->pool.sq_o; t.sq_d = c->pool.sq_d; t.vis_a = c->pool.vis_l; t.vis_b = c->pool.evt_vis; t.tot = c->d_tot;
    t.refill_idle = c->opt_refill_idle; t.phase_min = c->opt_phase_min; t.inner_min = c->opt_inner_min; t.inner_chain = c->opt_inner_chain;

    yune_stats st{};
    // Stage timing: every `time_stages`-th iteration is bracketed by CUDA events on the launching stream (no host
    // synchronisation inside the loop); at most 128 iterations are sampled per call.
    const int kMaxTimed = 128;
    int n_timed = 0;
    if (c->opt_time_stages > 0 && c->ev_pool.empty()) {
        c->ev_pool.resize(4 * kMaxTimed);
        for (auto& e : c->ev_pool) Y_CUDA(c, cudaEventCreate(&e));
    }
    Y_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    int it = 0;
    bool done = spp_count == 0 && !cont;
    const int sync_every = wait_all ? c->opt_sync_every : 1;      // a pipelined call is a few iterations long: look after each one
    // Steady state = the windows between two host syncs in which the pool was full throughout: past the first iterations (the
    // path mix has settled) and with samples still left to hand out after the window (every finished slot was regenerated).
    // Their rays and timed launches are reported separately so that a roofline can divide like by like.
    std::vector<int> timed_it; timed_it.reserve(kMaxTimed);
    std::vector<std::pair<int, int>> steady_windows;
    unsigned long long prev_ext = 0, prev_sh = 0;
    while (!done) {
        const int window_begin = it;
        for (int b = 0; b < sync_every && it < c->opt_max_iterations; b++, it++, c->it_global++) {
            const int p = (int)(c->it_global & 1u);
            a.parity = p;
            t.n_extend = &c->d_ctr[p].n_extend; t.fetch_extend = &c->d_ctr[p].fetch_extend;
            t.n_shadow = &c->d_ctr[p].n_shadow; t.fetch_shadow = &c->d_ctr[p].fetch_shadow;
            const bool timed = c->opt_time_stages > 0 && (it % c->opt_time_stages) == 0 && n_timed < kMaxTimed;
            if (timed) Y_CUDA(c, cudaEventRecord(c->ev_pool[4 * n_timed], c->stream));
            if (c->integrator == INTEGRATOR_BDPT) Y_CUDA(c, launch_shade_bdpt(a, c->bdpt, c->sm_count, c->occ_bdpt, c->stream));
            else Y_CUDA(c, launch_shade_dense(a, c->sm_count, c->opt_shade_blocks_per_sm, c->occ_dense, c->stream));
            if (timed) Y_CUDA(c, cudaEventRecord(c->ev_pool[4 * n_timed + 1], c->stream));
            if (it == c->cap_iteration && c->cap_max > 0)
                Y_CUDA(c, launch_capture(c->pool, c->d_ctr + p, c->cap_max, c->cap_eo, c->cap_ed, c->cap_so, c->cap_sd, c->cap_cnt, c->stream));
            t.eq = c->pool.eq; t.sq_idx = nullptr;
            if (sort_rays && !a.tail) {
                // queue lengths to the host (an iteration of a scene this size is milliseconds long), then two radix sorts by origin cell
                Y_CUDA(c, cudaMemcpyAsync(&c->h_cnt[0], &c->d_ctr[p].n_extend, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                Y_CUDA(c, cudaMemcpyAsync(&c->h_cnt[1], &c->d_ctr[p].n_shadow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                Y_CUDA(c, cudaStreamSynchronize(c->stream));
                const int ne = c->h_cnt[0], ns = c->h_cnt[1];
                if (ne <= c->pool.n_slots && (long long)ns <= 3ll * c->pool.n_slots)
                    Y_CUDA(c, launch_ray_keys(c->sc, c->pool, ne, ns, c->d_eq_key, c->d_sq_key, c->stream));
                if (ne > 0 && ne <= c->pool.n_slots) {
                    Y_CUDA(c, sort_pairs_by_key(c->d_eq_key, c->d_key_tmp, c->pool.eq, c->d_eq_sorted, ne, c->opt_sort_bits, c->d_sort_tmp, c->sort_tmp_bytes, c->stream));
                    t.eq = c->d_eq_sorted;
                }
                if (ns > 0 && (long long)ns <= 3ll * c->pool.n_slots) {
                    Y_CUDA(c, sort_pairs_by_key(c->d_sq_key, c->d_key_tmp, c->d_sq_iota, c->d_sq_idx, ns, c->opt_sort_bits, c->d_sort_tmp, c->sort_tmp_bytes, c->stream));
                    t.sq_idx = c->d_sq_idx;
                }
                st.kernel_launches += 3;
            }
            if (timed) Y_CUDA(c, cudaEventRecord(c->ev_pool[4 * n_timed + 2], c->stream));
            Y_CUDA(c, launch_trace(t, tl.grid, tl.block, tl.smem, c->opt_count_work != 0, c->stream));
            if (timed) { Y_CUDA(c, cudaEventRecord(c->ev_pool[4 * n_timed + 3], c->stream)); n_timed++; timed_it.push_back(it); }
            Y_CUDA(c, launch_iter_end(c->d_ctr, c->d_tot, p, c->stream));
            st.kernel_launches += 3; st.trace_launches += 1;
        }
        Y_CUDA(c, cudaMemcpyAsync(c->h_tot, c->d_tot, sizeof(Totals), cudaMemcpyDeviceToHost, c->stream));
        Y_CUDA(c, cudaStreamSynchronize(c->stream));
        a.tail = c->h_tot->next_sample >= c->h_tot->n_samples ? 1 : 0;      // the pool only drains from here on
        if (!a.tail && window_begin >= 8) {
            steady_windows.push_back({window_begin, it});
            st.steady_iterations += (uint32_t)(it - window_begin);
            st.steady_extend_rays += c->h_tot->extend_rays - prev_ext; st.steady_shadow_rays += c->h_tot->shadow_rays - prev_sh;
        }
        prev_ext = c->h_tot->extend_rays; prev_sh = c->h_tot->shadow_rays;
        if (c->h_tot->live_last == 0) done = true;
        else if (!wait_all && a.tail) { done = true; c->carry = true; c->carry_epoch = c->epoch; c->carry_seed = seed; c->carry_gi = gi_check; }
        else if (it >= c->opt_max_iterations) Y_FAIL(c, YUNE_ERR_LIMIT, "yune_render: max_iterations (%d) reached with %d paths alive", c->opt_max_iterations, c->h_tot->live_last);
    }
    float shade_ms = 0.f, trace_ms = 0.f, sort_ms = 0.f;
    for (int i = 0; i < n_timed; i++) {
        float m1 = 0, m2 = 0, m3 = 0;
        cudaEventElapsedTime(&m1, c->ev_pool[4 * i], c->ev_pool[4 * i + 1]);
        cudaEventElapsedTime(&m3, c->ev_pool[4 * i + 1], c->ev_pool[4 * i + 2]);      // option "sort_rays": count read-back + the two radix sorts
        cudaEventElapsedTime(&m2, c->ev_pool[4 * i + 2], c->ev_pool[4 * i + 3]);
        shade_ms += m1; trace_ms += m2; sort_ms += m3;
        for (const auto& w : steady_windows)
            if (timed_it[i] >= w.first && timed_it[i] < w.second) { st.steady_shade_ms += m1; st.steady_trace_ms += m2; st.steady_timed_iterations++; break; }
    }
    st.timed_iterations = (uint32_t)n_timed;
    if (det) Y_CUDA(c, launch_fix_to_sum(c->d_fix, c->d_sum, n_pix, c->stream));      // inside the timed region: part of the mode's cost
    Y_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    Y_CUDA(c, cudaEventSynchronize(c->ev1));
    float ms = 0; Y_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    st.render_ms = ms; st.shade_ms = shade_ms; st.trace_ms = trace_ms; st.sort_ms = sort_ms;
    st.samples = c->h_tot->n_samples; st.extend_rays = c->h_tot->extend_rays; st.shadow_rays = c->h_tot->shadow_rays;
    st.diffuse_visits = c->h_tot->visits_d; st.specular_visits = c->h_tot->visits_s; st.regenerations = c->h_tot->visits_r;
    st.slot_visits = (uint64_t)c->pool.n_slots * (uint64_t)it;
    st.box_tests = c->h_tot->box_tests; st.tri_tests = c->h_tot->tri_tests; st.iterations = (uint32_t)it;
    st.tonemap_ms = c->stats.tonemap_ms;
    st.carried_paths = c->carry ? (uint32_t)c->h_tot->live_last : 0u;
    c->stats = st;
    return YUNE_OK;
}

int yune_render(yune_ctx* c, int spp_begin, int spp_count, int gi_check, uint32_t seed, int reset)
{
    if (!c) return YUNE_ERR_INVALID;
    return render_impl(c, spp_begin, spp_count, gi_check, seed, reset, c->opt_pipeline == 0);
}

int yune_finish(yune_ctx* c)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!c->carry) return YUNE_OK;
    if (c->carry_epoch != c->epoch) { c->carry = false; return YUNE_OK; }      // what they depended on changed: discarded
    const yune_stats before = c->stats;
    const int rc = render_impl(c, 0, 0, c->carry_gi, c->carry_seed, 0, true);
    if (rc == YUNE_OK) { const float ms = c->stats.render_ms; c->stats = before; c->stats.finish_ms = ms; c->stats.carried_paths = 0; }
    return rc;
}

int yune_tonemap(yune_ctx* c)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!c->d_sum) Y_FAIL(c, YUNE_ERR_STATE, "yune_tonemap: image buffers not set up");
    Y_CUDA(c, cudaSetDevice(c->device));
    Y_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    Y_CUDA(c, launch_tonemap(c->d_sum, c->d_hdr, c->d_ldr, c->W * c->H, c->stream));
    Y_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    Y_CUDA(c, cudaEventSynchronize(c->ev1));
    float ms = 0; Y_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.tonemap_ms = ms;
    return YUNE_OK;
}

static int read_image(yune_ctx* c, const float4* src, float* dst)
{
    if (!dst) Y_FAIL(c, YUNE_ERR_INVALID, "read: destination is NULL");
    if (!src) Y_FAIL(c, YUNE_ERR_STATE, "read: image buffers not set up");
    Y_CUDA(c, cudaSetDevice(c->device));
    Y_CUDA(c, cudaMemcpyAsync(dst, src, (size_t)c->W * c->H * 16, cudaMemcpyDeviceToHost, c->stream));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    return YUNE_OK;
}
int yune_read_hdr(yune_ctx* c, float* rgba)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!c->d_sum) Y_FAIL(c, YUNE_ERR_STATE, "read: image buffers not set up");
    Y_CUDA(c, cudaSetDevice(c->device));
    Y_CUDA(c, launch_tonemap(c->d_sum, c->d_hdr, nullptr, c->W * c->H, c->stream));
    return read_image(c, c->d_hdr, rgba);
}
int yune_read_sum(yune_ctx* c, float* rgba) { return c ? read_image(c, c->d_sum, rgba) : YUNE_ERR_INVALID; }
int yune_read_ldr(yune_ctx* c, float* rgba) { return c ? read_image(c, c->d_ldr, rgba) : YUNE_ERR_INVALID; }
int yune_write_sum(yune_ctx* c, const float* rgba)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!rgba) Y_FAIL(c, YUNE_ERR_INVALID, "yune_write_sum: source is NULL");
    if (!c->d_sum) Y_FAIL(c, YUNE_ERR_STATE, "yune_write_sum: image buffers not set up");
    Y_CUDA(c, cudaSetDevice(c->device));
    Y_CUDA(c, cudaMemcpyAsync(c->d_sum, rgba, (size_t)c->W * c->H * 16, cudaMemcpyHostToDevice, c->stream));
    if (c->opt_deterministic && c->d_fix) Y_CUDA(c, launch_sum_to_fix(c->d_sum, c->d_fix, (size_t)c->W * c->H, c->stream));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    return YUNE_OK;
}
int yune_read_sum_fixed(yune_ctx* c, int64_t* rgba)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!rgba) Y_FAIL(c, YUNE_ERR_INVALID, "yune_read_sum_fixed: destination is NULL");
    if (!c->opt_deterministic || !c->d_fix) Y_FAIL(c, YUNE_ERR_STATE, "yune_read_sum_fixed: option \"deterministic\" is off (no fixed-point buffer)");
    Y_CUDA(c, cudaSetDevice(c->device));
    Y_CUDA(c, cudaMemcpyAsync(rgba, c->d_fix, (size_t)c->W * c->H * 32, cudaMemcpyDeviceToHost, c->stream));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    return YUNE_OK;
}
int yune_sum_fixed_device_ptr(yune_ctx* c, void** dptr, size_t* n_bytes)
{
    if (!c || !dptr) return YUNE_ERR_INVALID;
    if (!c->opt_deterministic || !c->d_fix) Y_FAIL(c, YUNE_ERR_STATE, "option \"deterministic\" is off (no fixed-point buffer)");
    *dptr = c->d_fix; if (n_bytes) *n_bytes = (size_t)c->W * c->H * 32;
    return YUNE_OK;
}
int yune_sum_refresh(yune_ctx* c)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!c->opt_deterministic || !c->d_fix) return YUNE_OK;           // nothing to derive: the float buffer is the accumulator
    Y_CUDA(c, cudaSetDevice(c->device));
    Y_CUDA(c, launch_fix_to_sum(c->d_fix, c->d_sum, (size_t)c->W * c->H, c->stream));
    return YUNE_OK;
}
int yune_sum_device_ptr(yune_ctx* c, void** dptr, size_t* n_bytes)
{
    if (!c || !dptr) return YUNE_ERR_INVALID;
    if (!c->d_sum) Y_FAIL(c, YUNE_ERR_STATE, "image buffers not set up");
    *dptr = c->d_sum; if (n_bytes) *n_bytes = (size_t)c->W * c->H * 16;
    return YUNE_OK;
}
int yune_stream(yune_ctx* c, void** s) { if (!c || !s) return YUNE_ERR_INVALID; *s = (void*)c->stream; return YUNE_OK; }
int yune_synchronize(yune_ctx* c) { if (!c) return YUNE_ERR_INVALID; Y_CUDA(c, cudaSetDevice(c->device)); Y_CUDA(c, cudaStreamSynchronize(c->stream)); return YUNE_OK; }
int yune_get_stats(yune_ctx* c, yune_stats* out) { if (!c || !out) return YUNE_ERR_INVALID; *out = c->stats; return YUNE_OK; }

// ---- measurement aid: capture the rays of one wavefront iteration ----
int yune_debug_capture_rays(yune_ctx* c, int iteration, int max_rays)
{
    if (!c) return YUNE_ERR_INVALID;
    if (max_rays < 0 || max_rays > (1 << 24)) Y_FAIL(c, YUNE_ERR_INVALID, "yune_debug_capture_rays: max_rays out of range");
    Y_CUDA(c, cudaSetDevice(c->device));
    if (max_rays > c->cap_alloc) {
        dfree(c->cap_eo); dfree(c->cap_ed); dfree(c->cap_so); dfree(c->cap_sd); dfree(c->cap_cnt); c->cap_alloc = 0;
        const size_t N = (size_t)max_rays;
        Y_CUDA(c, cudaMalloc(&c->cap_eo, N * 16)); Y_CUDA(c, cudaMalloc(&c->cap_ed, N * 16));
        Y_CUDA(c, cudaMalloc(&c->cap_so, N * 16)); Y_CUDA(c, cudaMalloc(&c->cap_sd, N * 16)); Y_CUDA(c, cudaMalloc(&c->cap_cnt, 16));
        c->cap_alloc = max_rays;
    }
    if (c->cap_cnt) Y_CUDA(c, cudaMemset(c->cap_cnt, 0, 16));
    c->cap_iteration = max_rays > 0 ? iteration : -1; c->cap_max = max_rays;
    return YUNE_OK;
}
int yune_debug_read_captured(yune_ctx* c, int which, float* od6, float* tmax, int* n_captured, int* n_in_queue)
{
    if (!c || !n_captured) return YUNE_ERR_INVALID;
    if (!c->cap_cnt) Y_FAIL(c, YUNE_ERR_STATE, "no capture was armed");
    Y_CUDA(c, cudaSetDevice(c->device));
    int cnt[4];
    Y_CUDA(c, cudaMemcpy(cnt, c->cap_cnt, 16, cudaMemcpyDeviceToHost));
    const int n = which ? cnt[1] : cnt[0];
    *n_captured = n; if (n_in_queue) *n_in_queue = which ? cnt[3] : cnt[2];
    if (n == 0 || !od6) return YUNE_OK;
    std::vector<float> o((size_t)n * 4), d((size_t)n * 4);
    Y_CUDA(c, cudaMemcpy(o.data(), which ? c->cap_so : c->cap_eo, (size_t)n * 16, cudaMemcpyDeviceToHost));
    Y_CUDA(c, cudaMemcpy(d.data(), which ? c->cap_sd : c->cap_ed, (size_t)n * 16, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        od6[6 * (size_t)i + 0] = o[4 * (size_t)i]; od6[6 * (size_t)i + 1] = o[4 * (size_t)i + 1]; od6[6 * (size_t)i + 2] = o[4 * (size_t)i + 2];
        od6[6 * (size_t)i + 3] = d[4 * (size_t)i]; od6[6 * (size_t)i + 4] = d[4 * (size_t)i + 1]; od6[6 * (size_t)i + 5] = d[4 * (size_t)i + 2];
        if (tmax) tmax[i] = o[4 * (size_t)i + 3];
    }
    return YUNE_OK;
}

// ---- parity hooks ----
static int ensure_hook(yune_ctx* c, int n)
{
    if (c->hk_cap >= n) return YUNE_OK;
    dfree(c->hk_o); dfree(c->hk_d); dfree(c->hk_hit); dfree(c->hk_tri); dfree(c->hk_light); dfree(c->hk_t); dfree(c->hk_od); dfree(c->hk_tmax); dfree(c->hk_vis); dfree(c->hk_cnt);
    c->hk_cap = 0;
    const size_t N = (size_t)n;
    Y_CUDA(c, cudaMalloc(&c->hk_o, N * 16)); Y_CUDA(c, cudaMalloc(&c->hk_d, N * 16)); Y_CUDA(c, cudaMalloc(&c->hk_hit, N * 16));
    Y_CUDA(c, cudaMalloc(&c->hk_tri, N * 4)); Y_CUDA(c, cudaMalloc(&c->hk_light, N * 4)); Y_CUDA(c, cudaMalloc(&c->hk_t, N * 4));
    Y_CUDA(c, cudaMalloc(&c->hk_od, N * 24)); Y_CUDA(c, cudaMalloc(&c->hk_tmax, N * 4)); Y_CUDA(c, cudaMalloc(&c->hk_vis, N));
    Y_CUDA(c, cudaMalloc(&c->hk_cnt, 4 * sizeof(int)));
    c->hk_cap = n;
    return YUNE_OK;
}

// run k_trace over hook rays [0, n): closest (any == 0) or any-hit
static int hook_trace(yune_ctx* c, int n, int any)
{
    TraceLaunch tl; int rc;
    if ((rc = trace_config(c, tl, false)) != YUNE_OK) return rc;
    int h_cnt[4] = {any ? 0 : n, 0, any ? n : 0, 0};       // n_extend, fetch_extend, n_shadow, fetch_shadow
    Y_CUDA(c, cudaMemcpyAsync(c->hk_cnt, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, c->stream));
    TraceArgs t{};
    t.sc = c->sc; t.eq = nullptr; t.ray_o = c->hk_o; t.ray_d = c->hk_d; t.ray_stride = 1; t.hit = c->hk_hit;
    t.n_extend = c->hk_cnt + 0; t.fetch_extend = c->hk_cnt + 1; t.n_shadow = c->hk_cnt + 2; t.fetch_shadow = c->hk_cnt + 3;
    t.sq_o = c->hk_o; t.sq_d = c->hk_d; t.vis_a = c->hk_vis; t.vis_b = c->hk_vis; t.tot = nullptr;
    t.refill_idle = c->opt_refill_idle; t.phase_min = c->opt_phase_min; t.inner_min = c->opt_inner_min; t.inner_chain = c->opt_inner_chain;
    Y_CUDA(c, launch_trace(t, tl.grid, tl.block, tl.smem, false, c->stream));
    return YUNE_OK;
}

int yune_trace_primary(yune_ctx* c, int jitter_mode, uint32_t rand, int32_t* tri_id, int32_t* light_id, float* t_hit)
{
    if (!c) return YUNE_ERR_INVALID;
    if (!tri_id || !light_id) Y_FAIL(c, YUNE_ERR_INVALID, "yune_trace_primary: output pointers are NULL");
    if (c->W <= 0) Y_FAIL(c, YUNE_ERR_STATE, "yune_trace_primary: image buffers not set up");
    if (!c->have_cam) Y_FAIL(c, YUNE_ERR_STATE, "yune_trace_primary: camera buffer not set up");
    Y_CUDA(c, cudaSetDevice(c->device));
    int rc; const int n = c->W * c->H;
    if ((rc = ensure_scene(c)) != YUNE_OK) return rc;
    if ((rc = ensure_hook(c, n)) != YUNE_OK) return rc;
    RenderArgs a = make_args(c);
    Y_CUDA(c, launch_hook_primary(a, jitter_mode, rand, c->hk_o, c->hk_d, c->stream));
    // light ids were stored in ray_d.w by the raygen; copy them out through the finish kernel's light_id input
    Y_CUDA(c, cudaMemsetAsync(c->hk_light, 0xff, (size_t)n * 4, c->stream));
    if ((rc = hook_trace(c, n, 0)) != YUNE_OK) return rc;
    Y_CUDA(c, launch_hook_finish(n, 0, c->hk_o, c->hk_hit, c->hk_vis, c->hk_tri, c->hk_light, c->hk_t, c->stream));
    std::vector<float> raydw((size_t)n * 4);
    Y_CUDA(c, cudaMemcpyAsync(tri_id, c->hk_tri, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (t_hit) Y_CUDA(c, cudaMemcpyAsync(t_hit, c->hk_t, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    Y_CUDA(c, cudaMemcpyAsync(raydw.data(), c->hk_d, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) {
        int lid; std::memcpy(&lid, &raydw[(size_t)i * 4 + 3], 4);
        light_id[i] = tri_id[i] >= 0 ? -1 : lid;
    }
    return YUNE_OK;
}

int yune_trace_rays(yune_ctx* c, int n, const float* od6, const float* tmax, int any_hit, int32_t* tri_id, int32_t* light_id, float* t_hit)
{
    if (!c) return YUNE_ERR_INVALID;
    if (n < 0 || !od6 || !tri_id || !light_id) Y_FAIL(c, YUNE_ERR_INVALID, "yune_trace_rays: bad arguments");
    if (n == 0) return YUNE_OK;
    Y_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_scene(c)) != YUNE_OK) return rc;
    if ((rc = ensure_hook(c, n)) != YUNE_OK) return rc;
    Y_CUDA(c, cudaMemcpyAsync(c->hk_od, od6, (size_t)n * 24, cudaMemcpyHostToDevice, c->stream));
    if (tmax) Y_CUDA(c, cudaMemcpyAsync(c->hk_tmax, tmax, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    Y_CUDA(c, launch_hook_prepare(c->lights, n, c->hk_od, tmax ? c->hk_tmax : nullptr, any_hit, c->hk_o, c->hk_d, c->hk_light, c->hk_vis, c->stream));
    if ((rc = hook_trace(c, n, any_hit)) != YUNE_OK) return rc;
    Y_CUDA(c, launch_hook_finish(n, any_hit, c->hk_o, c->hk_hit, c->hk_vis, c->hk_tri, c->hk_light, c->hk_t, c->stream));
    Y_CUDA(c, cudaMemcpyAsync(tri_id, c->hk_tri, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    Y_CUDA(c, cudaMemcpyAsync(light_id, c->hk_light, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (t_hit) Y_CUDA(c, cudaMemcpyAsync(t_hit, c->hk_t, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    Y_CUDA(c, cudaStreamSynchronize(c->stream));
    return YUNE_OK;
}

} // extern "C"
