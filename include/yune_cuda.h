/* yune_cuda.h -- the drop-in boundary: a C ABI that replaces the reference's CLManager + kernel launch.
 *
 * The reference has no FFI; its seam is (1) the public methods of yune::CLManager as called by
 * RendererCore/RendererGUI (include/CLManager.h:50-88) and (2) the OpenCL kernel-argument contract
 * (template/kernel.cl:66-68, set at src/RendererCore.cpp:202-236, 284, 516-577).  Every entry point below
 * names the reference interface it stands in for.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions (the C rendering of the reference's "bool + GUI message" protocol, src/CLManager.cpp:261-265):
 *   - every call returns 0 on success or a negative YUNE_ERR_* code; the text is at yune_last_error();
 *   - no exception crosses the boundary; a context is NOT thread-safe (the reference is single-threaded,
 *     include/CLManager.h:40-42);
 *   - host buffers passed in are copied before the call returns (CL_MEM_COPY_HOST_PTR semantics,
 *     src/CLManager.cpp:428, 451, 475); the caller keeps ownership;
 *   - images are RGBA float32, row 0 = bottom row (GL renderbuffer convention, udpt.cl:220);
 *   - there is no CPU fallback: with no usable sm_100 device yune_setup fails.
 */
#ifndef YUNE_CUDA_H
#define YUNE_CUDA_H

#include <stddef.h>
#include <stdint.h>
#include "yune_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define YUNE_OK            0
#define YUNE_ERR_INVALID  -1   /* bad argument / null pointer / size mismatch            */
#define YUNE_ERR_CUDA     -2   /* a CUDA runtime call failed (text has file:line + name) */
#define YUNE_ERR_STATE    -3   /* call made before the buffers it needs were set up      */
#define YUNE_ERR_LIMIT    -4   /* scene exceeds a documented limit (BVH depth, lights)   */
#define YUNE_ERR_NODEVICE -5   /* no CUDA device / not a Blackwell (sm_100) GPU          */

typedef struct yune_ctx yune_ctx;

/* ---- CLManager::setup() (src/CLManager.cpp:118-156): pick the device, create the context/queue ---- */
int  yune_setup(int device, yune_ctx** out_ctx);
void yune_destroy(yune_ctx* ctx);                      /* ~CLManager (src/CLManager.cpp:83-111) */
/* CLManager::checkError text (src/CLManager.cpp:755-855); ctx may be NULL for a failed yune_setup. */
const char* yune_last_error(const yune_ctx* ctx);

/* ---- CLManager::createRenderProgram(fn, path, reload) (src/CLManager.cpp:158-268) ----
 * There is no run-time compilation.  `kernel` selects a built-in integrator by the reference's file
 * name: "udpt.cl" | "udpt" (kernels/legacy/udpt.cl) or "bdpt.cl" | "bdpt" (kernels/legacy/bdpt.cl).
 * `compiler_opts` is the one-token '#yune-preproc compiler-opts' string (src/CLManager.cpp:182-204):
 * "-DMIS" enables udpt.cl's #ifdef MIS branch (udpt.cl:9, 566-608); NULL or "" = none.
 * Selecting a program also installs that kernel's built-in __constant light_sources[] (udpt.cl:97-106 /
 * bdpt.cl:106-115) unless yune_set_light_sources was called. */
int yune_create_render_program(yune_ctx* ctx, const char* kernel, const char* compiler_opts);
/* CLManager::createPostProcProgram (src/CLManager.cpp:270-379): only "tonemap.cl" | "tonemap" exists. */
int yune_create_postproc_program(yune_ctx* ctx, const char* kernel, const char* compiler_opts);

/* ---- buffer set-up: CLManager::setup*Buffer (src/CLManager.cpp:381-484) = kernel args 2-7 ---- */
int yune_setup_vertex_buffer(yune_ctx* ctx, const yune_triangle* tris, int n_triangles);   /* args 3,4 */
int yune_setup_mat_buffer(yune_ctx* ctx, const yune_material* mats, int n_materials);      /* arg 5    */
int yune_setup_bvh_buffer(yune_ctx* ctx, const yune_bvh_node* nodes, int n_nodes);         /* args 6,7; n = 0 (nodes may be NULL): the reference's brute-force mode, udpt.cl:280-284 -- same hit records (every triangle, index order, no box test), answered by a walk over our own tree */
/* BVH construction ON THE DEVICE, behind the same BVHNodeGPU contract (the reference builds on the host,
 * src/BVH.cpp:56-173, and at 10 M triangles that is seconds before the first sample).  Needs the vertex and material buffers;
 * replaces any uploaded BVH.  Builds a BVH over the Morton-sorted triangles (option "device_builder": PLOC -- bottom-up merging by surface area --
 * or a linear BVH with Karras' parallel hierarchy) with leaves of <= leaf_max (1..10, 0 = 2) triangles and emits, without leaving the GPU, both the reference-format node array (breadth-first, siblings adjacent,
 * nested boxes, the reference's +0.2 rule for flat boxes; yune_read_bvh_buffer hands it out) and the traversal layout of it.
 * Hit records are those of the reference's walk (udpt.cl:288-431) over THAT array, bit for bit; it is not the reference
 * builder's tree (yune_scene_load_bvh reproduces that one byte for byte). */
int yune_build_bvh_on_device(yune_ctx* ctx, int leaf_max);
int yune_bvh_info(yune_ctx* ctx, int* n_nodes, int* n_inner_nodes, int* depth, float* device_build_ms);
int yune_read_bvh_buffer(yune_ctx* ctx, yune_bvh_node* nodes, int capacity);      /* the uploaded or device-built array */
int yune_setup_camera_buffer(yune_ctx* ctx, const yune_cam* cam);                          /* arg 2    */
int yune_setup_image_buffers(yune_ctx* ctx, int width, int height);                        /* args 0,1 */
/* Replaces the kernels' __constant Quad light_sources[LIGHT_SIZE]; n in [1, 8].  n = 0 restores the
 * built-in light of the selected program. */
int yune_set_light_sources(yune_ctx* ctx, const yune_quad_light* lights, int n_lights);

/* Tunables that the reference exposes as '#define's at the top of its kernels or as GUI widgets:
 *   "rr_threshold" (udpt.cl:6 / bdpt.cl:4), "bdpt_bounces" (bdpt.cl:7), "oren_nayar" (0/1: use
 *   udpt-primitives.cl:681-725 for pure-diffuse lobes with sigma^2 = alpha_x),
 *   and engine knobs: "pool_slots" (path slots in flight; 0 = sized per job as 512*sqrt(samples), default;
 *   "pool_slots_in_use" reads back the size of the last render), "smem_nodes" (pair records staged in shared memory; < 0 = all if they fit, else 2340, default),
 *   "accel" (1 = walk our own SAH tree and filter candidates with the exact box test of their reference leaf, default;
 *   0 = walk the reference tree itself; 2 = the own tree collapsed into 4-wide records: half the node steps, bit-exact like the others,
 *   measured SLOWER than 1 on B200 because its staging area leaves the L1 2 KB -- DESIGN.md section 5), "leaf_split" (accel 0: refine reference leaves holding more than N triangles; 0 = off),
 *   "trace_block", "trace_blocks_per_sm" (trace-kernel launch shape), "refill_idle" (refill a warp once this many lanes are
 *   idle), "phase_min" (run a triangle step once this many lanes hold postponed triangles), "inner_min" / "inner_chain"
 *   (chain up to inner_chain further node steps without a new vote while inner_min lanes can take one),
 *   "shade_blocks_per_sm" (persistent shade grid; 0 = what the occupancy query returns),
 *   "isect" (0 = the reference's Moller-Trumbore behind the exact leaf-box filter: hit records bit-identical to udpt.cl:326-431, default;
 *   1 = PERF MODE: watertight edge-function test on the raw vertices, no filter -- needs accel 1; differs from 0 only for rays within rounding distance
 *   of an edge, a vertex or a reference box face), "deterministic" (0/1, below), "pipeline" (0/1, see yune_finish), "device_layout" (who builds the own tree of accel 1 for an UPLOADED BVH: 0 = the host's binned-SAH builder,
 *   1 = the device builder of yune_build_bvh_on_device -- the uploaded tree still decides every hit, bit for bit; falls back to the
 *   host builder for trees it cannot take; -1 (default) = the device above 2^16 triangles, the host below; "layout_built_on_device" reads back which one built the layout in use), "device_builder" (the device builder's algorithm: 1 = PLOC, default; 0 = linear BVH), "sort_rays" (sort both ray queues by the Morton cell of the ray
 *   origin between the shade and the trace kernel: 1 = on, 0 = off, default -- measured on the 10.5 M-triangle scene the trace kernel gains 7 % and the sorts
 *   cost more than that; "sort_bits": how many of the key's 30 bits, default 18), "own_tree_passes" (reinsertion passes over the
 *   host-built own tree, -1 = 2 up to 2^18 triangles, default), "ploc_radius" (PLOC's neighbour search radius in Morton positions, 1..64, default 32), "max_iterations", "sync_every", "time_stages", "count_work".
 *   Unknown key -> YUNE_ERR_INVALID.  None of them changes a result: tests/test_gpu_parity.py pins that. */
int yune_set_option(yune_ctx* ctx, const char* key, double value);
int yune_get_option(yune_ctx* ctx, const char* key, double* value);

/* ---- one call = RendererCore::enqueueKernels frames (src/RendererCore.cpp:248-469) ----
 * Renders samples [spp_begin, spp_begin + spp_count) of every pixel and adds them to the fp32 SUM buffer
 * (rgb = sum of sample radiance, a = number of samples; the reference's running mean is sum/a,
 * udpt.cl:196-209).  reset != 0 clears the sum first (kernel arg 9).  gi_check = kernel arg 8.
 * `seed` keys the counter-based RNG (it replaces the per-frame `rand`, kernel arg 10): sample s of pixel p
 * always sees the same random numbers, whatever the spp split or GPU count. */
int yune_render(yune_ctx* ctx, int spp_begin, int spp_count, int gi_check, uint32_t seed, int reset);
/* Option "pipeline" = 1: the reference's interactive call pattern -- ONE sample per pixel per call
 * (src/RendererCore.cpp:248-306, 483-486) -- without paying a full drain of the wavefront per call.  yune_render then returns
 * as soon as every sample of the call has been handed out; the paths still in flight keep their slots and finish during the
 * NEXT call (the pool stays full across calls), or in yune_finish.  Until then the image lacks those samples (each pixel's
 * count `a` says how many it holds, so the running mean stays a mean of complete samples).  After yune_finish the image is
 * bit-for-bit the image of one big call over the same sample range (tests pin that).  reset != 0, or any change of scene,
 * camera, program, lights, image size or options, discards the paths in flight; a change of seed / gi_check finishes them
 * first under the values they were started with.  0 (default): every call returns a complete image. */
int yune_finish(yune_ctx* ctx);
/* Post-processing launch (src/RendererCore.cpp:340-371): mean = sum/a, Reinhard + gamma of tonemap.cl:14-47. */
int yune_tonemap(yune_ctx* ctx);

/* Read-backs (the reference reads the GL renderbuffer, src/RendererCore.cpp:608-646).
 * hdr: running mean in rgb, sample count in a -- exactly the reference's image contents.
 * sum: the raw accumulation buffer.  ldr: output of yune_tonemap. */
int yune_read_hdr(yune_ctx* ctx, float* rgba);
int yune_read_sum(yune_ctx* ctx, float* rgba);
int yune_read_ldr(yune_ctx* ctx, float* rgba);
/* Load a (e.g. all-reduced) sum buffer back from the host. */
int yune_write_sum(yune_ctx* ctx, const float* rgba);

/* Option "deterministic" = 1 (DEFAULT; measured cost 0.5 % of a C2 render): samples are accumulated in 64-bit FIXED POINT (unit 2^-24, 4 x int64 per pixel: r, g, b, count)
 * instead of with fp32 atomics.  Integer addition is associative, so the accumulated image is bit-for-bit the same from run to
 * run, for any pool size, any split of the sample range into calls, and any number of GPUs (yune_group_reduce then reduces
 * the integer buffers).  The float sum buffer (yune_read_sum, yune_sum_device_ptr, tonemap input) is derived from it at the
 * end of every yune_render.  0 = float atomics on the sum buffer itself (the reference-like running sum), order-dependent in the
 * last bits. */
int yune_read_sum_fixed(yune_ctx* ctx, int64_t* rgba);                          /* W*H*4 int64; needs "deterministic" */
int yune_sum_fixed_device_ptr(yune_ctx* ctx, void** dptr, size_t* n_bytes);     /* for an external (NCCL) int64 sum-reduce */
int yune_sum_refresh(yune_ctx* ctx);       /* re-derive the float sum buffer after the fixed-point one was reduced externally */

/* Device pointer of the fp32 RGBA sum buffer (width*height*4 floats) so a collective library can reduce it
 * in place across GPUs; and the CUDA stream the context launches on. */
int yune_sum_device_ptr(yune_ctx* ctx, void** dptr, size_t* n_bytes);
int yune_stream(yune_ctx* ctx, void** cuda_stream);
int yune_synchronize(yune_ctx* ctx);

/* ---- parity hooks (test surface; they run the same raygen/extend kernels as yune_render) ---- */
/* Primary rays only.  jitter_mode 0: pixel centres; 1: the reference's jitter for `rand`
 * (wang_hash/xor_shift, udpt.cl:175-187).  Outputs per pixel: triangle id (-1 none), light id (-1 none), t. */
int yune_trace_primary(yune_ctx* ctx, int jitter_mode, uint32_t rand, int32_t* tri_id, int32_t* light_id, float* t_hit);
/* Arbitrary rays: od6 = n x (origin xyz, dir xyz), tmax may be NULL (= infinity). any_hit != 0 runs the
 * shadow-ray query (udpt.cl:306-308): tri_id is then 0 (occluded) or -1. */
int yune_trace_rays(yune_ctx* ctx, int n, const float* od6, const float* tmax, int any_hit,
                    int32_t* tri_id, int32_t* light_id, float* t_hit);

/* ---- measurement aid ----
 * Arm a capture of (up to max_rays of each of) the extension and shadow rays of wavefront iteration `iteration` of the
 * next yune_render; max_rays = 0 disarms.  After the render, read them back (which: 0 = extension rays, 1 = shadow rays;
 * tmax = the length each ray starts with).  bench.py gives these rays to the CPU oracle, whose ordered walk counts the
 * box / triangle tests that define the roofline's algorithmic bytes. */
int yune_debug_capture_rays(yune_ctx* ctx, int iteration, int max_rays);
int yune_debug_read_captured(yune_ctx* ctx, int which, float* od6, float* tmax, int* n_captured, int* n_in_queue);

/* ---- metrics (the reference's benchmark window, src/RendererCore.cpp:483-505) ---- */
typedef struct yune_stats {
    double   render_ms;        /* device time of the last yune_render (CUDA events on the context's stream) */
    double   trace_ms;         /* sum over the TIMED iterations of the trace-kernel duration (option "time_stages" = n: every n-th) */
    double   shade_ms;         /* same for the logic/shade kernel                                        */
    double   tonemap_ms;       /* device time of the last yune_tonemap                                   */
    uint64_t samples;          /* samples completed by the last yune_render                               */
    uint64_t extend_rays;      /* closest-hit rays traced                                                 */
    uint64_t shadow_rays;      /* any-hit rays traced (NEE + MIS + BDPT connections)                      */
    uint64_t box_tests;        /* only counted when option "count_work" = 1                               */
    uint64_t tri_tests;
    uint32_t iterations;       /* wavefront iterations                                                    */
    uint32_t kernel_launches;  /* kernels launched by the last yune_render                                */
    uint32_t trace_launches;   /* ... of which trace kernels                                              */
    uint32_t timed_iterations; /* iterations that trace_ms / shade_ms were summed over                    */
    uint64_t diffuse_visits;   /* slot visits shaded as a diffuse/glossy surface hit (NEE + bounce)       */
    uint64_t specular_visits;  /* slot visits shaded as a mirror/glass surface hit                         */
    uint64_t regenerations;    /* slot visits that started a fresh sample                                  */
    uint64_t slot_visits;      /* pool slots x shade launches (every launch classifies every slot)        */
    /* steady state = host-sync windows in which the pool was full throughout (past the first 8 iterations, samples still left
     * to hand out afterwards): rays traced in them and the TIMED launches that fell into them -- like divided by like      */
    uint32_t steady_iterations, steady_timed_iterations;
    uint64_t steady_extend_rays, steady_shadow_rays;
    double   steady_trace_ms, steady_shade_ms;
    uint32_t carried_paths;    /* option "pipeline": paths the last yune_render left in flight (0 after yune_finish)        */
    uint32_t reserved0;
    double   finish_ms;        /* device time of the last yune_finish                                                       */
    double   sort_ms;          /* option "sort_rays": sum over the TIMED iterations of queue-length read-back + the two sorts  */
} yune_stats;
int yune_get_stats(yune_ctx* ctx, yune_stats* out);

/* ---- multi-GPU: the reference is single-device; a drop-in that scales needs this surface ----
 * The path shards by SAMPLE INDEX: rank r of G renders all pixels for its slice of the sample range into its own fp32
 * sum buffer (no traffic while rendering), one ncclReduce(sum) over NVLink merges the buffers on the root, the root
 * tonemaps (yune_tonemap on yune_group_ctx(g, root)).  One process drives every device (ncclCommInitAll; NCCL is bound
 * at run time with dlopen, so the library has no link-time dependency on it).  A group of one needs no NCCL.
 * yune_shard_samples is the split rule: contiguous, balanced, the first (count mod n) ranks take one sample more. */
typedef struct yune_group yune_group;
typedef struct yune_group_stats {
    int      n_devices;
    double   render_ms_max, render_ms_min;   /* slowest / fastest rank of the last yune_group_render (device time) */
    double   reduce_ms;                      /* last yune_group_reduce, CUDA events on the root's stream            */
    uint64_t samples, extend_rays, shadow_rays;   /* summed over ranks                                             */
} yune_group_stats;
void yune_shard_samples(int spp_begin, int spp_count, int rank, int n_ranks, int* begin, int* count);
/* devices = NULL: devices 0 .. n-1; n_devices <= 0: every device of the node. */
int  yune_group_create(int n_devices, const int* devices, yune_group** out_group);
void yune_group_destroy(yune_group* g);
int  yune_group_size(const yune_group* g);
yune_ctx* yune_group_ctx(yune_group* g, int rank);              /* the rank's context, e.g. for read-backs and hooks */
const char* yune_group_last_error(const yune_group* g);          /* g may be NULL for a failed yune_group_create      */
/* the set-up calls above, applied to every rank (the scene is replicated) */
int yune_group_create_render_program(yune_group* g, const char* kernel, const char* compiler_opts);
int yune_group_create_postproc_program(yune_group* g, const char* kernel, const char* compiler_opts);
int yune_group_setup_vertex_buffer(yune_group* g, const yune_triangle* tris, int n_triangles);
int yune_group_setup_mat_buffer(yune_group* g, const yune_material* mats, int n_materials);
int yune_group_setup_bvh_buffer(yune_group* g, const yune_bvh_node* nodes, int n_nodes);
int yune_group_build_bvh_on_device(yune_group* g, int leaf_max);     /* instead of yune_group_setup_bvh_buffer */
int yune_group_setup_camera_buffer(yune_group* g, const yune_cam* cam);
int yune_group_setup_image_buffers(yune_group* g, int width, int height);
int yune_group_set_light_sources(yune_group* g, const yune_quad_light* lights, int n_lights);
int yune_group_set_option(yune_group* g, const char* key, double value);
/* Every rank renders its shard of [spp_begin, spp_begin + spp_count) concurrently (yune_render semantics per rank). */
int yune_group_render(yune_group* g, int spp_begin, int spp_count, int gi_check, uint32_t seed, int reset);
/* SUM-reduce the ranks' accumulation buffers into the root's (in place); the other ranks' buffers are unchanged. */
int yune_group_reduce(yune_group* g, int root);
int yune_group_get_stats(yune_group* g, yune_group_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* YUNE_CUDA_H */
