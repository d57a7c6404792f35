// kernels.h -- host-callable launch wrappers of kernels.cu.
#ifndef YUNE_KERNELS_H
#define YUNE_KERNELS_H

#include "wavefront_types.h"

/* defaults of the trace kernel's scheduling knobs (options "refill_idle", "phase_min", "inner_min", "inner_chain"); with all four
   at these values the kernel variants that have them compiled in are launched */
#define YUNE_DEF_REFILL_IDLE 12
#define YUNE_DEF_PHASE_MIN   24
#define YUNE_DEF_INNER_MIN   16
#define YUNE_DEF_INNER_CHAIN 8
#define YUNE_TRACE_MAX_BLOCK  1024     /* k_trace is compiled for <= 64 registers so any block size up to this fits */
#ifndef YUNE_SHADE_BLOCK
#define YUNE_SHADE_BLOCK      256
#endif
#ifndef YUNE_CLASSIFY_N
#define YUNE_CLASSIFY_N       2        /* slots a thread of k_shade_dense classifies per chunk (see the kernel) */
#endif
#ifndef YUNE_SHADE_MIN_BLOCKS
#define YUNE_SHADE_MIN_BLOCKS 3
#endif

namespace yune {

// Work description of one k_trace launch.  Counts are read from device memory so the launch shape never
// depends on them (persistent grid; lets a whole batch of iterations be replayed as one CUDA graph).
struct TraceArgs {
    DevScene sc;
    // extension rays (closest hit): ray = ray_o/ray_d[eq ? eq[q] : q], answer -> hit[same index]
    const int* eq; const float4* ray_o; const float4* ray_d; int ray_stride; float4* hit;   // ray k at ray_o/ray_d[k * ray_stride]
    const int* n_extend; int* fetch_extend;
    // shadow rays (any hit): answer -> vis_a[target] (target >= 0) or vis_b[~target]; 1 = unoccluded
    const float4* sq_o; const float4* sq_d; unsigned char* vis_a; unsigned char* vis_b;
    const int* sq_idx;    // option "sort_rays": shadow ray q is sq_o / sq_d[sq_idx[q]] (the queue sorted by origin cell), else NULL
    const int* n_shadow; int* fetch_shadow;
    Totals* tot;
    int refill_idle;      // refill a warp once this many of its lanes are idle
    int phase_min;        // run a TRI step as soon as this many lanes hold postponed triangles
    int inner_min;        // chain further INNER steps without a new vote while this many lanes can still take one ...
    int inner_chain;      // ... at most this many
};

cudaError_t launch_trace(const TraceArgs& a, int grid, int block, size_t smem_bytes, bool count, cudaStream_t st);
cudaError_t trace_prepare(const DevScene& sc, bool count, bool default_knobs, int block, size_t smem_bytes, int* blocks_per_sm);
int         trace_variant_id(const DevScene& sc, bool count, bool default_knobs);
cudaError_t launch_iter_end(IterCounters* ctr, Totals* tot, int parity, cudaStream_t st);
cudaError_t launch_shade_dense(const RenderArgs& a, int sm_count, int blocks_per_sm, int* occ_cache, cudaStream_t st);   // persistent, in-block sorted (default)
cudaError_t launch_shade_bdpt(const RenderArgs& a, const BdptPool& b, int sm_count, int* occ_cache, cudaStream_t st);
cudaError_t launch_capture(const PathPool& p, const IterCounters* c, int max_rays, float4* ext_o, float4* ext_d, float4* sh_o, float4* sh_d, int* counts, cudaStream_t st);
cudaError_t launch_pool_reset(const PathPool& p, cudaStream_t st);
// option "sort_rays" (bvh_build.cu, cub): values[0 .. n) reordered by the top `bits` bits of their 30-bit keys; tmp from sort_pairs_tmp_bytes
size_t      sort_pairs_tmp_bytes(int n_max);
cudaError_t sort_pairs_by_key(const unsigned* keys_in, unsigned* keys_out, const int* vals_in, int* vals_out, int n, int bits, void* tmp, size_t tmp_bytes, cudaStream_t st);
cudaError_t launch_iota(int* p, int n, cudaStream_t st);
cudaError_t launch_ray_keys(const DevScene& sc, const PathPool& p, int n_ext, int n_sh, unsigned* eq_key, unsigned* sq_key, cudaStream_t st);
cudaError_t launch_pool_revive(const PathPool& p, cudaStream_t st);
cudaError_t launch_fill_f4(float4* p, size_t n, float4 v, cudaStream_t st);
cudaError_t launch_fix_to_sum(const long long* fix, float4* sum, size_t n, cudaStream_t st);
cudaError_t launch_sum_to_fix(const float4* sum, long long* fix, size_t n, cudaStream_t st);
cudaError_t launch_tonemap(const float4* sum, float4* hdr, float4* ldr, int n, cudaStream_t st);
cudaError_t launch_hook_primary(const RenderArgs& a, int jitter_mode, uint32_t rand, float4* ray_o, float4* ray_d, cudaStream_t st);
cudaError_t launch_hook_prepare(const LightSet& L, int n, const float* od6, const float* tmax, int any_hit,
                                float4* ray_o, float4* ray_d, int* light_id, unsigned char* vis, cudaStream_t st);
cudaError_t launch_hook_finish(int n, int any_hit, const float4* ray_o, const float4* hit, const unsigned char* vis,
                               int* tri_id, int* light_id, float* t_hit, cudaStream_t st);

} // namespace yune
#endif
