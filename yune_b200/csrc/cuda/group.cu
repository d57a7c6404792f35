// group.cu -- the multi-GPU part of the C ABI (include/yune_cuda.h: yune_group_*).
//
// The path shards by SAMPLE INDEX: rank r of G renders every pixel for its slice of [spp_begin, spp_begin + spp_count) into its
// own fp32 sum buffer (no traffic while rendering: counter-based random numbers make a sample's value independent of who
// renders it), then ONE ncclReduce(sum) over NVLink merges the buffers on the root, which tonemaps.  There is no other
// exchange step on the path, so there is no compute + collective kernel to fuse.
//
// One process drives all devices (ncclCommInitAll, one communicator / context / stream / host thread per device) -- what a C++
// host such as csrc/app/yune_headless.cpp needs.  bench.py's one-process-per-GPU launch reduces through torch.distributed
// instead and never enters this file.  NCCL is bound at run time (dlopen) so that the library has no link-time dependency on
// it: a process that already carries an NCCL (torch) gets that copy, a plain C++ host gets the system one.
#include "yune_cuda.h"

#include <cuda_runtime.h>
#include <nccl.h>          // types and enums only; no NCCL symbol is linked
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load()
    {
        if (handle) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (handle) break; }
        if (!handle) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return false; }
        auto sym = [&](const char* n) { void* p = dlsym(handle, n); if (!p) err = std::string("NCCL symbol missing: ") + n; return p; };
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        Reduce = (decltype(Reduce))sym("ncclReduce");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        return CommInitAll && CommDestroy && Reduce && GroupStart && GroupEnd && GetErrorString;
    }
};
NcclApi g_nccl;
thread_local std::string g_group_err;

}  // namespace

struct yune_group {
    std::vector<int> devices;
    std::vector<yune_ctx*> ctx;
    std::vector<ncclComm_t> comm;        // empty for a group of one
    std::vector<cudaEvent_t> ev;         // [2]: around the reduce on the root's stream
    std::string err;
    yune_group_stats stats{};
};

#define G_FAIL(g, code, ...) do { char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__); (g)->err = _b; return (code); } while (0)

extern "C" {

void yune_shard_samples(int spp_begin, int spp_count, int rank, int n_ranks, int* begin, int* count)
{
    // contiguous and balanced: the first (spp_count mod n) ranks take one sample more (yune_b200/dist.py: shard_samples)
    const int base = n_ranks > 0 ? spp_count / n_ranks : 0, extra = n_ranks > 0 ? spp_count % n_ranks : 0;
    if (begin) *begin = spp_begin + rank * base + (rank < extra ? rank : extra);
    if (count) *count = base + (rank < extra ? 1 : 0);
}

const char* yune_group_last_error(const yune_group* g) { return g ? g->err.c_str() : g_group_err.c_str(); }

int yune_group_create(int n_devices, const int* devices, yune_group** out)
{
    if (!out) { g_group_err = "yune_group_create: out is NULL"; return YUNE_ERR_INVALID; }
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { g_group_err = "no CUDA device available; this library has no CPU fallback"; return YUNE_ERR_NODEVICE; }
    if (n_devices <= 0) n_devices = have;                         // 0 = every device of the node
    if (n_devices > have) { g_group_err = "yune_group_create: more devices requested than the node has"; return YUNE_ERR_INVALID; }
    yune_group* g = new yune_group();
    for (int i = 0; i < n_devices; i++) {
        const int d = devices ? devices[i] : i;
        for (int seen : g->devices) if (seen == d) { g_group_err = "yune_group_create: a device is listed twice"; delete g; return YUNE_ERR_INVALID; }
        g->devices.push_back(d);
    }
    for (int d : g->devices) {
        yune_ctx* c = nullptr;
        const int rc = yune_setup(d, &c);
        if (rc != YUNE_OK) { g_group_err = yune_last_error(nullptr); yune_group_destroy(g); return rc; }
        g->ctx.push_back(c);
    }
    if (n_devices > 1) {
        if (!g_nccl.load()) { g_group_err = g_nccl.err; yune_group_destroy(g); return YUNE_ERR_STATE; }
        g->comm.resize(n_devices);
        const ncclResult_t r = g_nccl.CommInitAll(g->comm.data(), n_devices, g->devices.data());
        if (r != ncclSuccess) { g_group_err = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r); g->comm.clear(); yune_group_destroy(g); return YUNE_ERR_CUDA; }
    }
    cudaSetDevice(g->devices[0]);
    g->ev.resize(2);
    for (auto& e : g->ev) cudaEventCreate(&e);
    *out = g;
    return YUNE_OK;
}

void yune_group_destroy(yune_group* g)
{
    if (!g) return;
    for (size_t i = 0; i < g->comm.size(); i++) if (g->comm[i]) g_nccl.CommDestroy(g->comm[i]);
    if (!g->devices.empty()) cudaSetDevice(g->devices[0]);
    for (auto& e : g->ev) if (e) cudaEventDestroy(e);
    for (yune_ctx* c : g->ctx) yune_destroy(c);
    delete g;
}

int yune_group_size(const yune_group* g) { return g ? (int)g->ctx.size() : 0; }
yune_ctx* yune_group_ctx(yune_group* g, int rank) { return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[rank] : nullptr; }

// The same call on every context (scene, program, camera, image buffers, options are replicated: <= 1.5 GB even for C4).
#define G_EACH(g, call) do { if (!(g)) return YUNE_ERR_INVALID; for (size_t _r = 0; _r < (g)->ctx.size(); _r++) { yune_ctx* ctx = (g)->ctx[_r]; \
        const int _rc = (call); if (_rc != YUNE_OK) { (g)->err = "rank " + std::to_string(_r) + ": " + yune_last_error(ctx); return _rc; } } return YUNE_OK; } while (0)

int yune_group_create_render_program(yune_group* g, const char* kernel, const char* opts) { G_EACH(g, yune_create_render_program(ctx, kernel, opts)); }
int yune_group_create_postproc_program(yune_group* g, const char* kernel, const char* opts) { G_EACH(g, yune_create_postproc_program(ctx, kernel, opts)); }
int yune_group_setup_vertex_buffer(yune_group* g, const yune_triangle* t, int n) { G_EACH(g, yune_setup_vertex_buffer(ctx, t, n)); }
int yune_group_setup_mat_buffer(yune_group* g, const yune_material* m, int n) { G_EACH(g, yune_setup_mat_buffer(ctx, m, n)); }
int yune_group_setup_bvh_buffer(yune_group* g, const yune_bvh_node* b, int n) { G_EACH(g, yune_setup_bvh_buffer(ctx, b, n)); }
int yune_group_build_bvh_on_device(yune_group* g, int leaf_max) { G_EACH(g, yune_build_bvh_on_device(ctx, leaf_max)); }      // every rank builds the same (deterministic) tree
int yune_group_setup_camera_buffer(yune_group* g, const yune_cam* cam) { G_EACH(g, yune_setup_camera_buffer(ctx, cam)); }
int yune_group_setup_image_buffers(yune_group* g, int w, int h) { G_EACH(g, yune_setup_image_buffers(ctx, w, h)); }
int yune_group_set_light_sources(yune_group* g, const yune_quad_light* l, int n) { G_EACH(g, yune_set_light_sources(ctx, l, n)); }
int yune_group_set_option(yune_group* g, const char* key, double v) { G_EACH(g, yune_set_option(ctx, key, v)); }

int yune_group_render(yune_group* g, int spp_begin, int spp_count, int gi_check, uint32_t seed, int reset)
{
    if (!g) return YUNE_ERR_INVALID;
    if (spp_begin < 0 || spp_count < 0) G_FAIL(g, YUNE_ERR_INVALID, "yune_group_render: negative sample range");
    const int n = (int)g->ctx.size();
    std::vector<int> rc(n, YUNE_OK);
    auto work = [&](int r) {
        int b = 0, c = 0;
        yune_shard_samples(spp_begin, spp_count, r, n, &b, &c);
        rc[r] = yune_render(g->ctx[r], b, c, gi_check, seed, reset);      // yune_render selects its own device
    };
    // yune_render drives its wavefront loop from the host (a sync every few iterations): one host thread per device
    std::vector<std::thread> th;
    for (int r = 1; r < n; r++) th.emplace_back(work, r);
    work(0);
    for (auto& t : th) t.join();
    yune_group_stats st{};
    st.n_devices = n;
    for (int r = 0; r < n; r++) {
        if (rc[r] != YUNE_OK) { g->err = "rank " + std::to_string(r) + ": " + yune_last_error(g->ctx[r]); return rc[r]; }
        yune_stats s; yune_get_stats(g->ctx[r], &s);
        st.samples += s.samples; st.extend_rays += s.extend_rays; st.shadow_rays += s.shadow_rays;
        if (s.render_ms > st.render_ms_max) st.render_ms_max = s.render_ms;
        if (r == 0 || s.render_ms < st.render_ms_min) st.render_ms_min = s.render_ms;
    }
    st.reduce_ms = g->stats.reduce_ms;
    g->stats = st;
    return YUNE_OK;
}

int yune_group_reduce(yune_group* g, int root)
{
    if (!g) return YUNE_ERR_INVALID;
    const int n = (int)g->ctx.size();
    if (root < 0 || root >= n) G_FAIL(g, YUNE_ERR_INVALID, "yune_group_reduce: root %d outside the group of %d", root, n);
    g->stats.reduce_ms = 0.0;
    if (n == 1) return YUNE_OK;
    std::vector<void*> buf(n); std::vector<size_t> bytes(n); std::vector<void*> stream(n);
    double det = 0.0;
    yune_get_option(g->ctx[root], "deterministic", &det);       // fixed-point accumulation: reduce the integer buffers (exact, order-free)
    for (int r = 0; r < n; r++) {
        int rc = det != 0.0 ? yune_sum_fixed_device_ptr(g->ctx[r], &buf[r], &bytes[r]) : yune_sum_device_ptr(g->ctx[r], &buf[r], &bytes[r]);
        if (rc == YUNE_OK) rc = yune_stream(g->ctx[r], &stream[r]);
        if (rc != YUNE_OK) { g->err = "rank " + std::to_string(r) + ": " + yune_last_error(g->ctx[r]); return rc; }
        if (bytes[r] != bytes[0]) G_FAIL(g, YUNE_ERR_STATE, "yune_group_reduce: image sizes differ between ranks");
    }
    cudaSetDevice(g->devices[root]);
    cudaEventRecord(g->ev[0], (cudaStream_t)stream[root]);
    ncclResult_t res = g_nccl.GroupStart();
    for (int r = 0; r < n && res == ncclSuccess; r++)      // in place on the root; the other ranks' buffers stay as they are
        res = det != 0.0 ? g_nccl.Reduce(buf[r], buf[root], bytes[r] / sizeof(long long), ncclInt64, ncclSum, root, g->comm[r], (cudaStream_t)stream[r])
                         : g_nccl.Reduce(buf[r], buf[root], bytes[r] / sizeof(float), ncclFloat32, ncclSum, root, g->comm[r], (cudaStream_t)stream[r]);
    if (res == ncclSuccess) res = g_nccl.GroupEnd(); else g_nccl.GroupEnd();
    if (res != ncclSuccess) G_FAIL(g, YUNE_ERR_CUDA, "ncclReduce: %s", g_nccl.GetErrorString(res));
    if (det != 0.0 && yune_sum_refresh(g->ctx[root]) != YUNE_OK) { g->err = yune_last_error(g->ctx[root]); return YUNE_ERR_CUDA; }
    cudaSetDevice(g->devices[root]);
    cudaEventRecord(g->ev[1], (cudaStream_t)stream[root]);
    for (int r = 0; r < n; r++) {
        const int rc = yune_synchronize(g->ctx[r]);
        if (rc != YUNE_OK) { g->err = "rank " + std::to_string(r) + ": " + yune_last_error(g->ctx[r]); return rc; }
    }
    float ms = 0.0f;
    cudaSetDevice(g->devices[root]);
    if (cudaEventElapsedTime(&ms, g->ev[0], g->ev[1]) == cudaSuccess) g->stats.reduce_ms = ms;
    return YUNE_OK;
}

int yune_group_get_stats(yune_group* g, yune_group_stats* out) { if (!g || !out) return YUNE_ERR_INVALID; *out = g->stats; return YUNE_OK; }

}  // extern "C"
