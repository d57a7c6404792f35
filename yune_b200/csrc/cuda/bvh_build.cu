// bvh_build.cu -- BVH construction ON THE DEVICE behind the BVHNodeGPU contract.
//
// The reference builds its BVH on the host (BVH::createBVH, src/BVH.cpp:56-173: binned SAH, recursive) and uploads an array of
// 80-byte BVHNodeGPU records (include/CL_headers.h:86-92) that its kernel walks through `child_idx` / `child_idx + 1`
// (udpt.cl:288-324).  csrc/host/BVH.cpp reproduces that builder byte for byte and stays the parity source; at 10.5 M triangles
// it costs 4 s on 16 cores plus 7 s of re-layout and upload before the first sample.  This file is the fast path to the same
// CONTRACT: from the uploaded TriangleGPU records it produces, on the GPU,
//   (1) a BVHNodeGPU array in the reference's format (breadth-first, siblings adjacent, leaves of <= `leaf_max` <= 10 triangle
//       indices, boxes that contain their children's boxes exactly, the reference's +0.2 rule for flat triangle boxes,
//       src/TriangleCPU.cpp:53-70) -- yune_read_bvh_buffer hands it out, the reference's kernel or the oracle can walk it;
//   (2) the traversal layout of trav_layout.h for that very tree (pair records breadth-first, triangle records grouped by leaf,
//       leaf boxes, shading records, specular flags), so no host re-layout runs.
// Two builders over the Morton-sorted triangles (63-bit codes of the centroids, cub radix sort): builder 1 (default) = PLOC,
// bottom-up merging of mutual nearest neighbours by union surface area (Meister & Bittner 2018); builder 0 = a linear BVH, Karras'
// parallel hierarchy (HPG 2012) with boxes fitted bottom-up.  Then, for either: subtrees of <= leaf_max triangles collapsed into
// leaves, breadth-first numbering level by level with a prefix sum per level, triangle records in depth-first order, one emission
// kernel.  It is NOT the reference's tree -- hit records are those of a reference-style walk of THIS array (tests: device hits ==
// the oracle's walk of the downloaded array, bit for bit).
#include "kernels.h"
#include "bvh_build.h"
#include "strict_math.h"
#include "trav_layout.h"

#include <cub/cub.cuh>
#include <cstdio>
#include <string>
#include <vector>

namespace yune {

namespace {

#define YB_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { char _b[256]; snprintf(_b, sizeof _b, "%s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); err = _b; return false; } } while (0)

struct Box { float lo[3], hi[3]; };

__device__ __forceinline__ unsigned long long spread21(unsigned long long x)
{   // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8)  & 0x100f00f00f00f00full;
    x = (x | x << 4)  & 0x10c30c30c30c30c3ull;
    x = (x | x << 2)  & 0x1249249249249249ull;
    return x;
}

// per triangle: exact box (reference rule), padded box (relayout.cpp: padded_bounds), centroid bounds of the scene
__global__ void k_tri_boxes(const yune_triangle* tris, int n, Box* exact, Box* padded, float* scene_lo, float* scene_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float clo[3] = {3e38f, 3e38f, 3e38f}, chi[3] = {-3e38f, -3e38f, -3e38f};
    if (i < n) {
        const yune_triangle& T = tris[i];
        Box e, p; float ext = 0.0f, mag = 0.0f;
        for (int k = 0; k < 3; k++) {
            e.lo[k] = fminf(T.v1.s[k], fminf(T.v2.s[k], T.v3.s[k]));
            e.hi[k] = fmaxf(T.v1.s[k], fmaxf(T.v2.s[k], T.v3.s[k]));
            p.lo[k] = e.lo[k]; p.hi[k] = e.hi[k];
            ext = fmaxf(ext, e.hi[k] - e.lo[k]); mag = fmaxf(mag, fmaxf(fabsf(e.lo[k]), fabsf(e.hi[k])));
            const float c = 0.5f * (e.lo[k] + e.hi[k]);
            clo[k] = c; chi[k] = c;
            if (e.hi[k] - e.lo[k] == 0.0f) e.hi[k] += 0.2f;            // src/TriangleCPU.cpp:63-67: a flat box could never pass 't_max > t_min'
        }
        const float pad = 2.0e-3f * ext + 4.0e-6f * mag + 1.0e-30f;    // relayout.cpp: the conservative margin of the own tree
        for (int k = 0; k < 3; k++) { p.lo[k] -= pad; p.hi[k] += pad; }
        exact[i] = e; padded[i] = p;
    }
    // block reduction of the centroid bounds, then one atomic per block and axis (floats ordered through their int bits)
    typedef cub::BlockReduce<float, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    for (int k = 0; k < 3; k++) {
        const float lo = BR(tmp).Reduce(clo[k], cub::Min()); __syncthreads();
        const float hi = BR(tmp).Reduce(chi[k], cub::Max()); __syncthreads();
        if (threadIdx.x == 0) {
            // atomicMin / atomicMax on floats via the ordered-int trick (sign handled by two cases)
            if (lo >= 0.0f) atomicMin(reinterpret_cast<int*>(scene_lo + k), __float_as_int(lo)); else atomicMax(reinterpret_cast<unsigned*>(scene_lo + k), __float_as_uint(lo));
            if (hi >= 0.0f) atomicMax(reinterpret_cast<int*>(scene_hi + k), __float_as_int(hi)); else atomicMin(reinterpret_cast<unsigned*>(scene_hi + k), __float_as_uint(hi));
        }
    }
}

__global__ void k_morton(const yune_triangle* tris, int n, const float* scene_lo, const float* scene_hi, unsigned long long* keys, int* idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const yune_triangle& T = tris[i];
    unsigned long long code = 0;
    for (int k = 0; k < 3; k++) {
        const float lo = fminf(T.v1.s[k], fminf(T.v2.s[k], T.v3.s[k])), hi = fmaxf(T.v1.s[k], fmaxf(T.v2.s[k], T.v3.s[k]));
        const float c = 0.5f * (lo + hi), ext = scene_hi[k] - scene_lo[k];
        float u = ext > 0.0f ? (c - scene_lo[k]) / ext : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        const unsigned long long q = (unsigned long long)fminf(u * 2097152.0f, 2097151.0f);
        code |= spread21(q) << (2 - k);
    }
    keys[i] = code; idx[i] = i;
}

// ---- node numbering shared by both builders: node p < n is the triangle at sorted position p; node n + k is inner node k ----

// Karras 2012.  Sorted keys may repeat: ties are broken by the position (the paper's augmented key).
__device__ __forceinline__ int delta(const unsigned long long* keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(i ^ j) : __clzll((long long)(a ^ b));
}
__global__ void k_hierarchy(const unsigned long long* keys, int n, int2* child, int* parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do { t = (t + 1) / 2; if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = lo == gamma ? gamma : n + gamma, right = hi == gamma + 1 ? gamma + 1 : n + gamma + 1;
    child[n + i] = make_int2(left, right);
    parent[left] = n + i; parent[right] = n + i;
    if (i == 0) parent[n] = -1;
}

__device__ __forceinline__ Box box_union(const Box& a, const Box& b)
{
    Box r;
    for (int k = 0; k < 3; k++) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return r;
}
__global__ void k_leaf_boxes(const int* sorted_tri, const Box* tri_exact, const Box* tri_padded, int n, Box* exact, Box* padded, int* size)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    exact[p] = tri_exact[sorted_tri[p]]; padded[p] = tri_padded[sorted_tri[p]]; size[p] = 1;
}
// LBVH: boxes and subtree sizes bottom-up, one atomic per inner node (the second thread to arrive goes on)
__global__ void k_fit(int n, const int2* child, const int* parent, Box* exact, Box* padded, int* size, int* arrived)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || n == 1) return;
    int node = parent[p];
    __threadfence();
    while (node >= 0) {
        if (atomicAdd(&arrived[node - n], 1) == 0) return;
        __threadfence();
        const int2 c = child[node];
        Box ea, eb, pa, pb;
        const volatile Box* v;
        v = exact + c.x;  for (int k = 0; k < 3; k++) { ea.lo[k] = v->lo[k]; ea.hi[k] = v->hi[k]; }
        v = exact + c.y;  for (int k = 0; k < 3; k++) { eb.lo[k] = v->lo[k]; eb.hi[k] = v->hi[k]; }
        v = padded + c.x; for (int k = 0; k < 3; k++) { pa.lo[k] = v->lo[k]; pa.hi[k] = v->hi[k]; }
        v = padded + c.y; for (int k = 0; k < 3; k++) { pb.lo[k] = v->lo[k]; pb.hi[k] = v->hi[k]; }
        exact[node] = box_union(ea, eb); padded[node] = box_union(pa, pb);
        size[node] = ((volatile int*)size)[c.x] + ((volatile int*)size)[c.y];
        __threadfence();
        node = parent[node];
    }
}

// ---- PLOC (Meister & Bittner, "Parallel locally-ordered clustering for bounding volume hierarchy construction", TVCG 2018) ----
// Bottom-up: the clusters stay in Morton order; every round each cluster looks for the neighbour within `radius` positions whose
// union box has the smallest surface area, mutual nearest neighbours merge into a new inner node, the array is compacted.  The
// merge criterion is the SAH's surface area, so large triangles find each other early instead of inflating a chain of boxes the
// way they do in a Morton-split tree.
#define YB_PLOC_R_MAX 64      /* search radius: run-time (option "ploc_radius", default 16), at most this */
__device__ __forceinline__ float box_area(const Box& b) { const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2]; return dx * dy + dx * dz + dy * dz; }
__global__ void k_ploc_nearest(const int* cl, int m, const Box* exact, int* nearest, const int R)
{
    __shared__ Box sb[256 + 2 * YB_PLOC_R_MAX];
    const int base = blockIdx.x * 256 - R;
    for (int t = threadIdx.x; t < 256 + 2 * R; t += 256) {
        const int i = base + t;
        if (i >= 0 && i < m) sb[t] = exact[cl[i]];
    }
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const Box me = sb[threadIdx.x + R];
    float best = 3.0e38f; int best_j = -1;
    for (int dj = -R; dj <= R; dj++) {
        const int j = i + dj;
        if (dj == 0 || j < 0 || j >= m) continue;
        const float a = box_area(box_union(me, sb[threadIdx.x + R + dj]));
        if (a < best) { best = a; best_j = j; }          // ties: the lower position (scan order)
    }
    nearest[i] = best_j;
}
__global__ void k_ploc_flags(const int* nearest, int m, int* creates, int* keeps)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nearest[i];
    const bool mutual = j >= 0 && nearest[j] == i;
    creates[i] = (mutual && i < j) ? 1 : 0;              // the lower partner carries the new node
    keeps[i] = (mutual && i > j) ? 0 : 1;                // the upper partner leaves the array
}
__global__ void k_ploc_merge(const int* cl, const int* nearest, int m, const int* creates, const int* create_rank, const int* keeps, const int* keep_rank,
                             int next_node, int2* child, Box* exact, Box* padded, int* size, int* cl_next)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || !keeps[i]) return;
    int id = cl[i];
    if (creates[i]) {
        const int a = cl[i], b = cl[nearest[i]];
        id = next_node + create_rank[i];
        child[id] = make_int2(a, b);
        exact[id] = box_union(exact[a], exact[b]); padded[id] = box_union(padded[a], padded[b]);
        size[id] = size[a] + size[b];
    }
    cl_next[keep_rank[i]] = id;
}

// ---- breadth-first numbering of the collapsed tree, one level per launch ----
// `order` is the concatenation of the levels' frontiers (node ids).  A node is INNER when it holds more than leaf_max triangles;
// everything else -- a single triangle or a small subtree -- is a leaf.  `tfirst` = where a node's triangles start in the
// triangle-record array: depth-first order (left subtree first), so that a subtree's triangles are contiguous and neighbours in
// space are neighbours in memory.
__device__ __forceinline__ bool is_inner(int node, int n, const int* size, int leaf_max) { return node >= n && size[node] > leaf_max; }
__global__ void k_level_flags(const int* frontier, int m, int n, const int* size, int leaf_max, int* flags)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) flags[k] = is_inner(frontier[k], n, size, leaf_max) ? 1 : 0;
}
__global__ void k_level_expand(const int* frontier, int m, const int* flags, const int* rank, const int2* child, const int* size, int n, int pair_base, int* pair_of, int* tfirst,
                               int* next_frontier, int level_begin, int next_begin, int* first_child_bfs)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    first_child_bfs[level_begin + k] = flags[k] ? next_begin + 2 * rank[k] : -1;
    if (!flags[k]) return;
    const int node = frontier[k];
    pair_of[node - n] = pair_base + rank[k];
    const int2 c = child[node];
    tfirst[c.x] = tfirst[node]; tfirst[c.y] = tfirst[node] + size[c.x];
    next_frontier[2 * rank[k]] = c.x; next_frontier[2 * rank[k] + 1] = c.y;
}
__global__ void k_leaf_flags(const int* order, int n_nodes, int n, const int* size, int leaf_max, int* flags)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_nodes) flags[b] = is_inner(order[b], n, size, leaf_max) ? 0 : 1;
}

// ---- emission: BVHNodeGPU records, pair records, leaf boxes, triangle records ----
__device__ __forceinline__ float4 f4i(float x, float y, float z, int w) { return make_float4(x, y, z, __int_as_float(w)); }
__global__ void k_emit(const int* order, const int* first_child_bfs, int n_nodes, int n, int leaf_max, const int2* child, const int* size, const int* tfirst, const int* pair_of,
                       const int* leaf_rank, const Box* exact, const Box* padded, const int* sorted_tri, const yune_triangle* tris,
                       yune_bvh_node* nodes, float4* pairs, float4* leaf_boxes, float4* tri_rec, const int* ref_leaf_of_tri, const int* ref_rank_of_tri)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_nodes) return;
    const int node = order[b];
    const Box e = exact[node];
    yune_bvh_node nd;
    for (int k = 0; k < 3; k++) { nd.aabb.p_min.s[k] = e.lo[k]; nd.aabb.p_max.s[k] = e.hi[k]; }
    nd.aabb.p_min.s[3] = 1.0f; nd.aabb.p_max.s[3] = 1.0f;
    for (int j = 0; j < 10; j++) nd.vert_list[j] = 0;
    if (is_inner(node, n, size, leaf_max)) {
        nd.child_idx = first_child_bfs[b]; nd.vert_len = -1;
        const int2 c = child[node];
        int refs[2]; Box pb[2];
        for (int s = 0; s < 2; s++) {
            const int cn = s ? c.y : c.x;
            pb[s] = padded[cn];
            refs[s] = is_inner(cn, n, size, leaf_max) ? pair_of[cn - n] : ~((tfirst[cn] << 4) | size[cn]);
        }
        float4* q = pairs + 4 * (size_t)pair_of[node - n];
        q[0] = make_float4(pb[0].lo[0], pb[0].hi[0], pb[0].lo[1], pb[0].hi[1]);
        q[1] = make_float4(pb[1].lo[0], pb[1].hi[0], pb[1].lo[1], pb[1].hi[1]);
        q[2] = make_float4(pb[0].lo[2], pb[0].hi[2], pb[1].lo[2], pb[1].hi[2]);
        q[3] = make_float4(__int_as_float(refs[0]), __int_as_float(refs[1]), 0.0f, 0.0f);
    } else {
        const int first = tfirst[node], cnt = size[node];
        const int lr = leaf_rank[b];
        nd.child_idx = -1; nd.vert_len = cnt;
        if (!ref_leaf_of_tri) {
            leaf_boxes[2 * (size_t)lr] = make_float4(e.lo[0], e.lo[1], e.lo[2], 0.0f);
            leaf_boxes[2 * (size_t)lr + 1] = make_float4(e.hi[0], e.hi[1], e.hi[2], 0.0f);
        }
        // the leaf's triangles: depth-first over its (<= leaf_max <= 10 triangles) subtree, left to right
        int stack[12]; int sp = 0, j = 0;
        stack[sp++] = node;
        while (sp > 0) {
            const int x = stack[--sp];
            if (x >= n) { const int2 c = child[x]; stack[sp++] = c.y; stack[sp++] = c.x; continue; }
            const int t = sorted_tri[x];
            nd.vert_list[j] = t;
            const yune_triangle& T = tris[t];
            const V3 v1 = v3(T.v1.s[0], T.v1.s[1], T.v1.s[2]);
            const V3 e1 = vsub(v3(T.v2.s[0], T.v2.s[1], T.v2.s[2]), v1);     // udpt.cl:328-329, the reference's own expressions (relayout.cpp)
            const V3 e2 = vsub(v3(T.v3.s[0], T.v3.s[1], T.v3.s[2]), v1);
            float4* r = tri_rec + 3 * (size_t)(first + j);
            r[0] = f4i(v1.x, v1.y, v1.z, t);
            // visiting rank: leaves in breadth-first order, then the slot -- of THIS tree, or of the uploaded reference tree
            r[1] = f4i(e1.x, e1.y, e1.z, ref_leaf_of_tri ? ref_rank_of_tri[t] : lr * 16 + j);
            r[2] = f4i(e2.x, e2.y, e2.z, ref_leaf_of_tri ? ref_leaf_of_tri[t] : lr);
            j++;
        }
    }
    if (nodes) nodes[b] = nd;
}

__global__ void k_iota(int* p, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }

// shading records and specular flags by ORIGINAL triangle index (relayout.cpp: shade_records; context.cu: tri_class)
__global__ void k_shade_records(const yune_triangle* tris, int n, const yune_material* mats, int n_mats, float4* shade, unsigned char* tri_class)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const yune_triangle& T = tris[i];
    const V3 n1 = vnormalize(v3(T.vn1.s[0], T.vn1.s[1], T.vn1.s[2]));       // udpt.cl:377-379
    const V3 n2 = vnormalize(v3(T.vn2.s[0], T.vn2.s[1], T.vn2.s[2]));
    const V3 n3 = vnormalize(v3(T.vn3.s[0], T.vn3.s[1], T.vn3.s[2]));
    float4* s = shade + 4 * (size_t)i;
    s[0] = f4i(n1.x, n1.y, n1.z, T.matID); s[1] = make_float4(n2.x, n2.y, n2.z, 0.0f); s[2] = make_float4(n3.x, n3.y, n3.z, 0.0f); s[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int mid = T.matID;
    tri_class[i] = (mid >= 0 && mid < n_mats && mats[mid].is_specular != 0) ? 1 : 0;
}

template <class T> struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
    T* release() { T* q = p; p = nullptr; return q; }
};

} // namespace

// ---- option "sort_rays": the radix sort of a ray queue by origin cell lives here because this is the translation unit with cub ----
size_t sort_pairs_tmp_bytes(int n_max)
{
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr, (int*)nullptr, n_max, 0, 30);
    return b;
}
cudaError_t sort_pairs_by_key(const unsigned* keys_in, unsigned* keys_out, const int* vals_in, int* vals_out, int n, int bits, void* tmp, size_t tmp_bytes, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    if (bits < 1) bits = 1;
    if (bits > 30) bits = 30;
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, n, 30 - bits, 30, st);
}
cudaError_t launch_iota(int* p, int n, cudaStream_t st)
{
    if (n > 0) k_iota<<<(n + 255) / 256, 256, 0, st>>>(p, n);
    return cudaGetLastError();
}

void GpuBvh::free_all()
{
    cudaFree(pairs); cudaFree(tris); cudaFree(leaf_boxes); cudaFree(shade); cudaFree(tri_class); cudaFree(nodes);
    pairs = tris = leaf_boxes = shade = nullptr; tri_class = nullptr; nodes = nullptr; n_nodes = n_inner = n_leaves = n_tris = 0;
}

bool buildBvhOnDevice(const yune_triangle* h_tris, int n, const yune_material* d_mats, int n_mats, int leaf_max, cudaStream_t st, GpuBvh& out, std::string& err, const RefLeaves* ref, int builder, int radius)
{
    if (radius < 1) radius = 32;
    if (radius > YB_PLOC_R_MAX) radius = YB_PLOC_R_MAX;
    out.free_all();
    if (n < 1) { err = "no triangles"; return false; }
    if (n >= (1 << 27)) { err = "more than 2^27 triangles"; return false; }
    if (leaf_max < 1) leaf_max = 2;
    if (leaf_max > 10) leaf_max = 10;                         // vert_list holds 10 indices (include/CL_headers.h:88)
    const int B = 256;
    auto grid = [&](long long m) { return (unsigned)((m + B - 1) / B); };
    struct Events { cudaEvent_t a = nullptr, b = nullptr; ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } ev;      // (freed on every return path)
    YB_CUDA(cudaEventCreate(&ev.a)); YB_CUDA(cudaEventCreate(&ev.b));
    cudaEvent_t e0 = ev.a, e1 = ev.b;

    DevBuf<yune_triangle> d_tris; YB_CUDA(d_tris.alloc(n));
    YB_CUDA(cudaMemcpyAsync(d_tris.p, h_tris, (size_t)n * sizeof(yune_triangle), cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaEventRecord(e0, st));
    DevBuf<Box> tri_exact, tri_padded, exact, padded;
    YB_CUDA(tri_exact.alloc(n)); YB_CUDA(tri_padded.alloc(n)); YB_CUDA(exact.alloc(2 * (size_t)n)); YB_CUDA(padded.alloc(2 * (size_t)n));
    DevBuf<float> scene; YB_CUDA(scene.alloc(6));
    const float init[6] = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
    YB_CUDA(cudaMemcpyAsync(scene.p, init, sizeof init, cudaMemcpyHostToDevice, st));
    k_tri_boxes<<<grid(n), B, 0, st>>>(d_tris.p, n, tri_exact.p, tri_padded.p, scene.p, scene.p + 3);
    DevBuf<unsigned long long> keys, keys_sorted; DevBuf<int> idx, sorted_tri;
    YB_CUDA(keys.alloc(n)); YB_CUDA(keys_sorted.alloc(n)); YB_CUDA(idx.alloc(n)); YB_CUDA(sorted_tri.alloc(n));
    k_morton<<<grid(n), B, 0, st>>>(d_tris.p, n, scene.p, scene.p + 3, keys.p, idx.p);
    size_t tmp_bytes = 0, scan_bytes = 0;
    YB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, idx.p, sorted_tri.p, n, 0, 63, st));
    YB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, 2 * n, st));
    const size_t tmp_cap = tmp_bytes > scan_bytes ? tmp_bytes : scan_bytes;
    DevBuf<unsigned char> tmp; YB_CUDA(tmp.alloc(tmp_cap));
    size_t tb = tmp_cap;
    YB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys_sorted.p, idx.p, sorted_tri.p, n, 0, 63, st));

    // ---- hierarchy over the sorted triangles: nodes [0, n) = triangles, [n, 2n - 1) = inner nodes ----
    DevBuf<int2> child; DevBuf<int> size, parent, arrived;
    YB_CUDA(child.alloc(2 * (size_t)n)); YB_CUDA(size.alloc(2 * (size_t)n));
    k_leaf_boxes<<<grid(n), B, 0, st>>>(sorted_tri.p, tri_exact.p, tri_padded.p, n, exact.p, padded.p, size.p);
    int root = 0, rounds = 0;
    if (n > 1 && builder == 0) {
        YB_CUDA(parent.alloc(2 * (size_t)n)); YB_CUDA(arrived.alloc(n));
        YB_CUDA(cudaMemsetAsync(arrived.p, 0, (size_t)n * sizeof(int), st));
        k_hierarchy<<<grid(n - 1), B, 0, st>>>(keys_sorted.p, n, child.p, parent.p);
        k_fit<<<grid(n), B, 0, st>>>(n, child.p, parent.p, exact.p, padded.p, size.p, arrived.p);
        root = n;
    } else if (n > 1) {
        DevBuf<int> cl_a, cl_b, nearest, creates, keeps, create_rank, keep_rank;
        YB_CUDA(cl_a.alloc(n)); YB_CUDA(cl_b.alloc(n)); YB_CUDA(nearest.alloc(n)); YB_CUDA(creates.alloc(n)); YB_CUDA(keeps.alloc(n));
        YB_CUDA(create_rank.alloc(n)); YB_CUDA(keep_rank.alloc(n));
        k_iota<<<grid(n), B, 0, st>>>(cl_a.p, n);
        int m = n, next_node = n;
        int* cur = cl_a.p; int* nxt = cl_b.p;
        while (m > 1) {
            k_ploc_nearest<<<grid(m), 256, 0, st>>>(cur, m, exact.p, nearest.p, radius);
            k_ploc_flags<<<grid(m), B, 0, st>>>(nearest.p, m, creates.p, keeps.p);
            tb = tmp_cap; YB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, creates.p, create_rank.p, m, st));
            tb = tmp_cap; YB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, keeps.p, keep_rank.p, m, st));
            int last[4];
            YB_CUDA(cudaMemcpyAsync(&last[0], creates.p + m - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            YB_CUDA(cudaMemcpyAsync(&last[1], create_rank.p + m - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            YB_CUDA(cudaMemcpyAsync(&last[2], keeps.p + m - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            YB_CUDA(cudaMemcpyAsync(&last[3], keep_rank.p + m - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            YB_CUDA(cudaStreamSynchronize(st));
            const int n_created = last[0] + last[1], m_next = last[2] + last[3];
            if (n_created < 1 || m_next >= m) { err = "PLOC made no progress"; return false; }      // cannot happen: the globally best pair is mutual
            k_ploc_merge<<<grid(m), B, 0, st>>>(cur, nearest.p, m, creates.p, create_rank.p, keeps.p, keep_rank.p, next_node, child.p, exact.p, padded.p, size.p, nxt);
            next_node += n_created; m = m_next; rounds++;
            int* t = cur; cur = nxt; nxt = t;
        }
        YB_CUDA(cudaMemcpyAsync(&root, cur, sizeof(int), cudaMemcpyDeviceToHost, st));
        YB_CUDA(cudaStreamSynchronize(st));
        if (next_node != 2 * n - 1) { err = "PLOC node count mismatch"; return false; }
    }

    // ---- breadth-first numbering ----
    DevBuf<int> order, flags, rank, pair_of, first_child_bfs, tfirst;
    YB_CUDA(order.alloc(2 * (size_t)n)); YB_CUDA(flags.alloc(2 * (size_t)n)); YB_CUDA(rank.alloc(2 * (size_t)n)); YB_CUDA(pair_of.alloc(n)); YB_CUDA(first_child_bfs.alloc(2 * (size_t)n));
    YB_CUDA(tfirst.alloc(2 * (size_t)n));
    YB_CUDA(cudaMemcpyAsync(order.p, &root, sizeof(int), cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaMemsetAsync(tfirst.p + root, 0, sizeof(int), st));
    int level_begin = 0, m = 1, n_pairs = 0, depth = 0;
    while (m > 0) {
        k_level_flags<<<grid(m), B, 0, st>>>(order.p + level_begin, m, n, size.p, leaf_max, flags.p);
        tb = tmp_cap; YB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flags.p, rank.p, m, st));
        int last[2];
        YB_CUDA(cudaMemcpyAsync(&last[0], flags.p + m - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        YB_CUDA(cudaMemcpyAsync(&last[1], rank.p + m - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        YB_CUDA(cudaStreamSynchronize(st));
        const int n_in = last[0] + last[1];
        k_level_expand<<<grid(m), B, 0, st>>>(order.p + level_begin, m, flags.p, rank.p, child.p, size.p, n, n_pairs, pair_of.p, tfirst.p,
                                               order.p + level_begin + m, level_begin, level_begin + m, first_child_bfs.p);
        n_pairs += n_in; level_begin += m; m = 2 * n_in; depth++;
        if (depth > YUNE_STACK_SIZE - 2) { err = "the device-built BVH is deeper than the traversal stack (YUNE_STACK_SIZE): build the BVH on the host"; return false; }
    }
    const int n_nodes = level_begin;
    DevBuf<int> leaf_rank; YB_CUDA(leaf_rank.alloc(n_nodes));
    k_leaf_flags<<<grid(n_nodes), B, 0, st>>>(order.p, n_nodes, n, size.p, leaf_max, flags.p);
    tb = tmp_cap; YB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flags.p, leaf_rank.p, n_nodes, st));
    const int n_leaves = n_nodes - n_pairs;

    DevBuf<yune_bvh_node> nodes; DevBuf<float4> pairs, leaf_boxes, tri_rec, shade; DevBuf<unsigned char> tri_class; DevBuf<int> ref_leaf, ref_rank;
    if (!ref) YB_CUDA(nodes.alloc(n_nodes));
    YB_CUDA(pairs.alloc(4 * (size_t)(n_pairs > 0 ? n_pairs : 1))); YB_CUDA(leaf_boxes.alloc(2 * (size_t)(ref ? ref->n_leaves : n_leaves)));
    YB_CUDA(tri_rec.alloc(3 * (size_t)n)); YB_CUDA(shade.alloc(4 * (size_t)n)); YB_CUDA(tri_class.alloc(n));
    if (ref) {
        YB_CUDA(ref_leaf.alloc(n)); YB_CUDA(ref_rank.alloc(n));
        YB_CUDA(cudaMemcpyAsync(ref_leaf.p, ref->leaf_of_tri, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
        YB_CUDA(cudaMemcpyAsync(ref_rank.p, ref->rank_of_tri, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
        YB_CUDA(cudaMemcpyAsync(leaf_boxes.p, ref->leaf_boxes, 2 * (size_t)ref->n_leaves * sizeof(float4), cudaMemcpyHostToDevice, st));
    }
    k_emit<<<grid(n_nodes), B, 0, st>>>(order.p, first_child_bfs.p, n_nodes, n, leaf_max, child.p, size.p, tfirst.p, pair_of.p, leaf_rank.p, exact.p, padded.p, sorted_tri.p, d_tris.p,
                                        nodes.p, pairs.p, leaf_boxes.p, tri_rec.p, ref ? ref_leaf.p : nullptr, ref ? ref_rank.p : nullptr);
    k_shade_records<<<grid(n), B, 0, st>>>(d_tris.p, n, d_mats, n_mats, shade.p, tri_class.p);
    Box root_box;
    YB_CUDA(cudaMemcpyAsync(&root_box, padded.p + root, sizeof(Box), cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaEventRecord(e1, st));
    YB_CUDA(cudaStreamSynchronize(st));
    YB_CUDA(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);

    out.n_tris = n; out.n_nodes = ref ? 0 : n_nodes; out.n_inner = n_pairs; out.n_leaves = ref ? ref->n_leaves : n_leaves; out.depth = depth; out.build_ms = ms; out.leaf_max = leaf_max;
    out.builder = builder; out.rounds = rounds;
    for (int k = 0; k < 3; k++) { out.root_lo[k] = root_box.lo[k]; out.root_hi[k] = root_box.hi[k]; }
    // the root: pair record 0, or -- a scene of <= leaf_max triangles -- one leaf
    out.root_ref = n_pairs > 0 ? 0 : ~((0 << 4) | n);      // (tfirst of the root is 0)
    out.nodes = nodes.release(); out.pairs = pairs.release(); out.leaf_boxes = leaf_boxes.release(); out.tris = tri_rec.release();
    out.shade = shade.release(); out.tri_class = tri_class.release();
    return true;
}

} // namespace yune
