// oracle/yune_oracle.cpp -- CPU RESTATEMENT of the reference's path-tracing kernels.  TEST INFRASTRUCTURE ONLY:
// nothing under yune_b200/ includes, links or calls this file; only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may use it, and only as the checker or the CPU baseline.
//
// What it restates (file:line into /root/reference):
//   kernels/legacy/udpt.cl:158-1152   pathtracer / createRay / traceRay / traverseBVH / rayTriangleIntersection /
//                                      rayAabbIntersection / shading / evaluateDirectLighting (with and without MIS) /
//                                      evaluateBRDF / sampleLights / samplers / Fresnel / helpers
//   kernels/legacy/bdpt.cl:432-640    createLightPath / createEyePath / shading, plus its variant helpers :718-737,
//                                      :807-878, :1048-1105
//   kernels/legacy/udpt-primitives.cl:696-725, 1147-1210   OrenNayarBRDF / toShadingSpace (opt-in extension)
//   kernels/post-proc/tonemap.cl:14-47
// Arithmetic conventions are those of oracle/clc_shim.inc (IEEE float32, no contraction, float4 lanes including w --
// the w lane matters: any(isnan(color)) at udpt.cl:193 looks at it).
//
// PINNING.  The reference ships no tests or golden vectors, so this restatement is pinned
// against the reference ITSELF: in rng_mode 0 (the reference's wang_hash/xor_shift stream) every frame it renders must
// be bit-identical to oracle/_ref/libyune_ref_kernels.so, which is the reference's own kernel text compiled as C++
// (tests/test_oracle_pinning.py; fixtures from the same library are committed under tests/golden/).
//
// Extensions beyond the reference, all off by default and documented in DESIGN.md:
//   rng_mode 1   counter-based draws addressed by (seed; pixel, sample, vertex, purpose) -- the stream the CUDA product
//                uses, so product and oracle can be compared sample for sample;
//   n_lights > 1 light list as data (the reference hard-codes LIGHT_SIZE 1); createLightPath picks the emitter with
//                probability proportional to |ke| * area;
//   oren_nayar   pure-diffuse lobes evaluate OrenNayarBRDF with sigma^2 = alpha_x, as udpt-primitives.cl:681-686 does;
//   heap_size    the BFS queue capacity (reference: 1500, udpt.cl:8); 0 = unbounded.
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>
#include "yune_types.h"

namespace {

typedef unsigned int uint;
const float PI = 3.14159265359f, INV_PI = 0.31830988618f, EPSILON = 0.0001f;

struct f4 {
    float x, y, z, w;
    f4() {}
    f4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    explicit f4(const yune_float4& v) : x(v.s[0]), y(v.s[1]), z(v.s[2]), w(v.s[3]) {}
};
inline f4 operator+(f4 a, f4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline f4 operator-(f4 a, f4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline f4 operator*(f4 a, f4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline f4 operator*(f4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline f4 operator*(float s, f4 a) { return f4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline f4 operator/(f4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline f4 operator-(f4 a) { return f4(-a.x, -a.y, -a.z, -a.w); }
inline float dot(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline f4 cross(f4 a, f4 b) { return f4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f); }
inline float length(f4 a) { return sqrtf(dot(a, a)); }
inline float length3(f4 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline f4 normalize(f4 a) { float l = sqrtf(dot(a, a)); return f4(a.x / l, a.y / l, a.z / l, a.w / l); }
inline float clmin(float x, float y) { return (y < x) ? y : x; }      // OpenCL min: "y if y < x, otherwise x"
inline float clmax(float x, float y) { return (x < y) ? y : x; }      // OpenCL max: "y if x < y, otherwise x"
inline bool anynan(f4 c) { return c.x != c.x || c.y != c.y || c.z != c.z || c.w != c.w; }

struct Ray { f4 origin, dir; float length; bool is_shadow_ray; };
struct HitInfo { int triangle_ID, light_ID; f4 hit_point, normal; };
struct PathInfo { HitInfo hit_info; f4 dir, contrib; float fwd_pdf, rev_pdf; };
struct Light { f4 pos, normal, ke, edge_l, edge_w; };

// ---- Philox4x32-10, restated from the published algorithm (Salmon et al. 2011) ----
inline void philox(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// purposes of a draw at one path vertex; (block, word) address inside the vertex's counter space
enum Purpose { P_LOBE = 0, P_DIR1 = 1, P_DIR2 = 2, P_FRESNEL = 3, P_RR = 4, P_NEE_LOBE = 5, P_MIS1 = 6, P_MIS2 = 7,
               P_LIGHT1 = 8, P_LIGHT2 = 9, P_PICK = 10 /* light i adds 4*i to P_LIGHT1/2 */ };
const uint32_t VERTEX_CAMERA = 0xFFFFFFFFu;

struct Rng {
    int mode;                 // 0 = reference stream, 1 = counter-based
    uint state;               // mode 0: xorshift state (the kernel's `seed`)
    uint32_t seed, pixel, sample, vertex;
    float next(int purpose)
    {
        if (mode == 0) {      // '*seed = xor_shift(*seed); r = *seed / (float) UINT_MAX;' (udpt.cl:1141-1147 and every draw site)
            state ^= state << 13; state ^= state >> 17; state ^= state << 5;
            return state / (float)UINT_MAX;
        }
        uint32_t c[4] = {pixel, sample, vertex, (uint32_t)(purpose >> 2)};
        philox(c, seed, 0x59554e45u);
        return c[purpose & 3] / (float)UINT_MAX;
    }
    // the one draw the reference converts in double precision: 'float r1 = *seed / (double) UINT_MAX;' (bdpt.cl:439)
    float next_via_double(int purpose)
    {
        if (mode == 0) {
            state ^= state << 13; state ^= state >> 17; state ^= state << 5;
            return (float)(state / (double)UINT_MAX);
        }
        uint32_t c[4] = {pixel, sample, vertex, (uint32_t)(purpose >> 2)};
        philox(c, seed, 0x59554e45u);
        return (float)(c[purpose & 3] / (double)UINT_MAX);
    }
};
inline uint wang_hash(uint seed)
{
    seed = (seed ^ 61) ^ (seed >> 16); seed *= 9; seed = seed ^ (seed >> 4); seed *= 0x27d4eb2d; seed = seed ^ (seed >> 15);
    return seed;
}

struct Scene {
    const yune_triangle* tris; int ntri;
    const yune_material* mats;
    const yune_bvh_node* bvh; int nnodes;
    Light lights[8]; int n_lights;
    int mis, rr_threshold, oren_nayar, heap_size, bdpt, bdpt_bounces;
    // work counters (per call, not thread safe by themselves: callers keep thread-local copies)
};
struct Work { unsigned long long box, tri, overflow; int max_queue; };

inline f4 tv(const yune_float4& v) { return f4(v); }

// ---- rayAabbIntersection (udpt.cl:392-431) ----
bool rayAabb(const Ray& ray, const yune_aabb& bb, Work* w)
{
    if (w) w->box++;
    float t_max = INFINITY, t_min = -INFINITY;
    const float ix = 1 / ray.dir.x, iy = 1 / ray.dir.y, iz = 1 / ray.dir.z;
    const float minx = (bb.p_min.s[0] - ray.origin.x) * ix, miny = (bb.p_min.s[1] - ray.origin.y) * iy, minz = (bb.p_min.s[2] - ray.origin.z) * iz;
    const float maxx = (bb.p_max.s[0] - ray.origin.x) * ix, maxy = (bb.p_max.s[1] - ray.origin.y) * iy, maxz = (bb.p_max.s[2] - ray.origin.z) * iz;
    if (!(minx != minx)) { t_min = fmaxf(clmin(minx, maxx), t_min); t_max = clmin(fmaxf(minx, maxx), t_max); }
    if (!(miny != miny)) { t_min = fmaxf(clmin(miny, maxy), t_min); t_max = clmin(fmaxf(miny, maxy), t_max); }
    if (t_max < t_min) return false;
    if (!(minz != minz)) { t_min = fmaxf(clmin(minz, maxz), t_min); t_max = clmin(fmaxf(minz, maxz), t_max); }
    return t_max > fmaxf(t_min, 0.0f);
}

// ---- rayTriangleIntersection (udpt.cl:326-390) ----
bool rayTriangle(const Scene& sc, Ray& ray, HitInfo& hit, int idx, Work* w)
{
    if (w) w->tri++;
    const yune_triangle& T = sc.tris[idx];
    const f4 v1 = tv(T.v1), v1v2 = tv(T.v2) - v1, v1v3 = tv(T.v3) - v1;
    const f4 pvec = cross(ray.dir, v1v3);
    const float det = dot(v1v2, pvec);
    const float inv_det = 1.0f / det;
    const f4 dist = ray.origin - v1;
    const float u = dot(pvec, dist) * inv_det;
    if (u < 0.0 || u > 1.0f) return false;
    const f4 qvec = cross(dist, v1v2);
    const float v = dot(qvec, ray.dir) * inv_det;
    if (v < 0.0 || u + v > 1.0) return false;
    const float t = dot(v1v3, qvec) * inv_det;
    if (t > 0 && t < ray.length) {
        ray.length = t;
        const f4 N1 = normalize(tv(T.vn1)), N2 = normalize(tv(T.vn2)), N3 = normalize(tv(T.vn3));
        const float ww = 1 - u - v;
        hit.hit_point = ray.origin + ray.dir * t;
        hit.normal = normalize(N1 * ww + N2 * u + N3 * v);
        hit.triangle_ID = idx; hit.light_ID = -1;
        return true;
    }
    return false;
}

// ---- traverseBVH (udpt.cl:288-324): breadth-first queue, no ordering, no pruning ----
bool traverseBVH(const Scene& sc, Ray& ray, HitInfo& hit, Work* w)
{
    static thread_local std::vector<int> queue;
    const int cap = sc.heap_size > 0 ? sc.heap_size : INT_MAX;
    queue.clear(); queue.push_back(0);
    bool intersect = false;
    if (!rayAabb(ray, sc.bvh[0].aabb, w)) return intersect;
    for (int i = 0; i < (int)queue.size() && (int)queue.size() < cap; i++) {
        const yune_bvh_node& nd = sc.bvh[queue[i]];
        const float c_idx = nd.child_idx;                 // 'float c_idx' in the reference (:300)
        if (c_idx == -1 && nd.vert_len > 0) {
            for (int j = 0; j < nd.vert_len; j++) {
                intersect |= rayTriangle(sc, ray, hit, nd.vert_list[j], w);
                if (ray.is_shadow_ray && intersect) return true;
            }
            continue;
        }
        for (int j = c_idx; j < c_idx + 2; j++) {
            if ((sc.bvh[j].vert_len > 0 || sc.bvh[j].child_idx > 0) && rayAabb(ray, sc.bvh[j].aabb, w)) queue.push_back(j);
        }
    }
    if (w) { if ((int)queue.size() > w->max_queue) w->max_queue = (int)queue.size(); if ((int)queue.size() >= cap) w->overflow++; }
    return intersect;
}

// ---- traceRay (udpt.cl:240-286) ----
bool traceRay(const Scene& sc, Ray& ray, HitInfo& hit, Work* w = nullptr)
{
    bool flag = false;
    for (int i = 0; i < sc.n_lights; i++) {
        const Light& L = sc.lights[i];
        const float DdotN = dot(ray.dir, L.normal);
        if (fabsf(DdotN) > 0.0001) {
            const float t = dot(L.normal, L.pos - ray.origin) / DdotN;
            if (t > 0.0 && t < ray.length) {
                f4 temp = ray.origin + (ray.dir * t);
                temp = temp - L.pos;
                float proj1 = dot(temp, L.edge_l), proj2 = dot(temp, L.edge_w);
                const float la = length(L.edge_l), lb = length(L.edge_w);
                proj1 /= la; proj2 /= lb;
                if ((proj1 >= 0.0 && proj2 >= 0.0) && (proj1 <= la && proj2 <= lb)) {
                    ray.length = t;
                    hit.hit_point = ray.origin + (ray.dir * t);
                    hit.light_ID = i; hit.triangle_ID = -1;
                    flag = true;
                }
            }
        }
    }
    if (sc.nnodes > 0) flag |= traverseBVH(sc, ray, hit, w);
    else for (int i = 0; i < sc.ntri; i++) flag |= rayTriangle(sc, ray, hit, i, w);
    return flag;
}

inline HitInfo noHit() { HitInfo h; h.triangle_ID = -1; h.light_ID = -1; h.hit_point = f4(0, 0, 0, 1); h.normal = f4(0, 0, 0, 0); return h; }
inline const yune_material& matOf(const Scene& sc, const HitInfo& h) { return sc.mats[sc.tris[h.triangle_ID].matID]; }

// ---- createRay (udpt.cl:213-238) ----
void createRay(float pixel_x, float pixel_y, int img_width, int img_height, Ray& eye_ray, const yune_cam& cam)
{
    f4 dir;
    float aspect_ratio = (img_width * 1.0) / img_height;
    dir.x = aspect_ratio * ((2.0 * pixel_x / img_width) - 1);
    dir.y = (2.0 * pixel_y / img_height) - 1;
    dir.z = -cam.view_plane_dist;
    dir.w = 0;
    eye_ray.dir = normalize(f4(dot(tv(cam.r1), dir), dot(tv(cam.r2), dir), dot(tv(cam.r3), dir), dot(tv(cam.r4), dir)));
    eye_ray.origin = f4(cam.r1.s[3], cam.r2.s[3], cam.r3.s[3], cam.r4.s[3]);
    eye_ray.is_shadow_ray = false;
    eye_ray.length = INFINITY;
}

// ---- reflect (udpt.cl:949-961) and the un-flipped bdpt form (bdpt.cl:723, 814) ----
f4 reflectFlip(f4 w_i, f4 normal)
{
    if (dot(w_i, normal) < 0) normal = normal * -1.0f;
    return normalize(2 * (dot(w_i, normal)) * normal - w_i);
}
f4 reflectDir(const Scene& sc, f4 w_i, f4 normal) { return sc.bdpt ? (2 * dot(w_i, normal)) * normal - w_i : reflectFlip(w_i, normal); }

// ---- Oren-Nayar (udpt-primitives.cl:696-725, 1147-1210) ----
void basis(f4 Nz, f4& Nx, f4& Ny)
{
    if (fabsf(Nz.y) > fabsf(Nz.z)) Nx = f4(Nz.y, -Nz.x, 0, 0.f); else Nx = f4(Nz.z, 0, -Nz.x, 0.f);
    Nx = normalize(Nx);
    Ny = normalize(cross(Nz, Nx));
}
f4 toShadingSpace(f4 w, const HitInfo& h)
{
    f4 Nx, Ny, Nz = h.normal; basis(Nz, Nx, Ny);
    const f4 r1(Nx.x, Nx.y, Nx.z, dot(h.hit_point, Nx)), r2(Ny.x, Ny.y, Ny.z, dot(h.hit_point, Ny)), r3(Nz.x, Nz.y, Nz.z, dot(h.hit_point, Nz)), r4(0.f, 0.f, 0.f, 1.0f);
    return f4(dot(r1, w), dot(r2, w), dot(r3, w), dot(r4, w));
}
inline float sinTheta(f4 w) { return sqrtf(1 - w.z * w.z); }
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float cosPhi(f4 w) { float s = sinTheta(w); if (s <= EPSILON && s >= -EPSILON) return 0; return clampf(w.x / s, -1.0f, 1.0f); }
inline float sinPhi(f4 w) { float s = sinTheta(w); if (s <= EPSILON && s >= -EPSILON) return 0; return clampf(w.y / s, -1.0f, 1.0f); }
f4 orenNayar(const Scene& sc, f4 w_i, f4 w_o, const HitInfo& h)
{
    w_i = toShadingSpace(w_i, h); w_o = toShadingSpace(w_o, h);
    float costerm = 0.0f;
    if (sinTheta(w_i) >= EPSILON && sinTheta(w_o) >= EPSILON) costerm = fmaxf(0.0f, ((cosPhi(w_i) * cosPhi(w_o)) + (sinPhi(w_i) * sinPhi(w_o))));
    float sin_alpha, tan_beta;
    if (fabsf(w_i.z) < fabsf(w_o.z)) { sin_alpha = sinTheta(w_i); tan_beta = sinTheta(w_o) / fabsf(w_o.z); }
    else { sin_alpha = sinTheta(w_o); tan_beta = sinTheta(w_i) / fabsf(w_i.z); }
    const float sigma_sq = matOf(sc, h).alpha_x;
    const float A = 1 - (sigma_sq / (2 * (sigma_sq + 0.33)));
    const float B = 0.45 * sigma_sq / (sigma_sq + 0.09);
    return tv(matOf(sc, h).kd) * INV_PI * (A + B * (costerm * sin_alpha * tan_beta));
}

// ---- evaluateBRDF (udpt.cl:611-630, bdpt.cl:718-737) ----
f4 evaluateBRDF(const Scene& sc, f4 w_i, f4 w_o, const HitInfo& h, bool sample_glossy, float rr_prob)
{
    f4 refl_vec = reflectDir(sc, w_i, h.normal);
    refl_vec = normalize(refl_vec);
    const yune_material& m = matOf(sc, h);
    if (!sample_glossy) {
        if (sc.oren_nayar && rr_prob == 1.0f) return orenNayar(sc, w_i, w_o, h);
        return tv(m.kd) * INV_PI / rr_prob;
    }
    const float cos_alpha = powf(fmaxf(dot(w_o, refl_vec), 0.0f), m.px + m.py);
    const int phong_exp = m.px + m.py;
    return tv(m.ks) * cos_alpha * (phong_exp + 2) * INV_PI * 0.5f / rr_prob;
}

// ---- sampleGlossyPdf (udpt.cl:1062-1119, bdpt.cl:1048-1105).  The draw happens first, always. ----
bool sampleGlossyPdf(const Scene& sc, const HitInfo& h, Rng& rng, int purpose, float& prob)
{
    const float r = rng.next(purpose);
    const f4 ks = tv(matOf(sc, h).ks), kd = tv(matOf(sc, h).kd);
    if (length3(ks) == 0.0f) { prob = 1.0f; return false; }
    else if (length3(kd) == 0.0f || kd.x + kd.y + kd.z == 0.0f) { prob = 1.0f; return true; }
    const f4 sum = ks + kd;
    const float max_val = clmax(sum.x, clmax(sum.y, sum.z));
    float pd, ps;
    if (max_val == sum.x) { pd = kd.x; ps = ks.x; } else if (max_val == sum.y) { pd = kd.y; ps = ks.y; } else { pd = kd.z; ps = ks.z; }
    if (sc.bdpt) {
        if (max_val < 1.0f) { pd += (1 - max_val) / 2.0f; ps += (1 - max_val) / 2.0f; }
        if (r < pd) { prob = pd; return false; }
        prob = ps; return true;
    }
    if (r < pd) { prob = pd; return false; }
    else if (r < pd + ps && r >= pd) { prob = ps; return true; }
    prob = 0; return false;
}

// ---- hemisphere samplers (udpt.cl:700-771, 843-910) ----
void finishSample(Ray& ray, f4 Nx, f4 Ny, f4 Nz, const HitInfo& h, float x, float y, float z)
{
    const f4 r1(Nx.x, Ny.x, Nz.x, h.hit_point.x), r2(Nx.y, Ny.y, Nz.y, h.hit_point.y), r3(Nx.z, Ny.z, Nz.z, h.hit_point.z);
    const f4 ray_dir(x, y, z, 0);
    ray.dir = normalize(f4(dot(r1, ray_dir), dot(r2, ray_dir), dot(r3, ray_dir), 0));
    ray.origin = h.hit_point + ray.dir * EPSILON;
    ray.is_shadow_ray = false; ray.length = INFINITY;
}
void cosineWeightedHemisphere(Ray& ray, float& pdf, const HitInfo& h, Rng& rng, int p1, int p2)
{
    f4 Nx, Ny, Nz = h.normal; basis(Nz, Nx, Ny);
    const float r1 = rng.next(p1), r2 = rng.next(p2);
    const float phi = 2 * PI * r2, sinT = sqrtf(r1);
    const float x = sinT * cosf(phi), y = sinT * sinf(phi), z = sqrtf(1 - r1);
    finishSample(ray, Nx, Ny, Nz, h, x, y, z);
    pdf = z * INV_PI;
}
void phongSampleHemisphere(const Scene& sc, Ray& ray, float& pdf, f4 w_i, const HitInfo& h, Rng& rng, int p1, int p2)
{
    f4 Nx, Ny, Nz = reflectDir(sc, w_i, h.normal);
    Nz = normalize(Nz);
    basis(Nz, Nx, Ny);
    const float r1 = rng.next(p1), r2 = rng.next(p2);
    const yune_material& m = matOf(sc, h);
    const int phong_exponent = m.px + m.py;
    const float phi = 2 * PI * r2;
    const float costheta = powf(r1, 1.0f / (phong_exponent + 1));
    float sintheta = 1 - powf(r1, 2.0f / (phong_exponent + 1));
    sintheta = sqrtf(sintheta);
    finishSample(ray, Nx, Ny, Nz, h, sintheta * cosf(phi), sintheta * sinf(phi), costheta);
    if (dot(ray.dir, h.normal) < 0) pdf = 0;
    else pdf = (phong_exponent + 1) * 0.5 * INV_PI * powf(costheta, (float)phong_exponent);
}

// ---- Fresnel (udpt.cl:912-1031) ----
float evalFresnelReflectance(const Scene& sc, f4 w_i, const HitInfo& h, float& ior_factor)
{
    f4 normal = h.normal; float n1, n2;
    const yune_material& m = matOf(sc, h);
    if (dot(w_i, normal) < 0) { n1 = m.n; n2 = 1; normal = normal * -1.0f; } else { n1 = 1; n2 = m.n; }
    const float cosThetaI = dot(w_i, normal);
    const float sinThetaI = sqrtf(1 - cosThetaI * cosThetaI);
    const float sinThetaT = n1 * sinThetaI / n2;
    const float cosThetaT = sqrtf(1 - sinThetaT * sinThetaT);
    if (sinThetaT >= 1.0f && n1 > n2) return 1.0f;
    float r0 = m.ks.s[0] + m.ks.s[1] + m.ks.s[2];
    r0 /= 3.0f;
    ior_factor = (n2 * n2) / (n1 * n1);
    if (n1 > n2) return (r0 + (1 - r0) * (1 - powf(cosThetaI, 5.0f)));
    return (r0 + (1 - r0) * (1 - powf(cosThetaT, 5.0f)));
}
f4 refractDir(const Scene& sc, f4 w_i, const HitInfo& h)
{
    f4 normal = h.normal; float n1, n2;
    const yune_material& m = matOf(sc, h);
    if (dot(w_i, normal) < 0) { n1 = m.n; n2 = 1; normal = normal * -1.0f; } else { n1 = 1; n2 = m.n; }
    const f4 wt_perp = n1 / n2 * (dot(w_i, normal) * normal - w_i);
    const f4 wt_parallel = sqrtf(1 - length(wt_perp) * length(wt_perp)) * -normal;
    return normalize(wt_perp + wt_parallel);
}
void sampleFresnelIncidence(const Scene& sc, Ray& ray, const HitInfo& h, f4 w_i, float& ior_factor, Rng& rng)
{
    const yune_material& m = matOf(sc, h);
    if (m.is_transmissive) {
        const float pdf = evalFresnelReflectance(sc, w_i, h, ior_factor);
        if (pdf == 1.0) { ray.dir = reflectFlip(w_i, h.normal); ior_factor = 1.0f; }
        else {
            const float r = rng.next(P_FRESNEL);
            if (r < pdf) { ray.dir = reflectFlip(w_i, h.normal); ior_factor = 1.0f; }
            else ray.dir = refractDir(sc, w_i, h);
        }
    } else { ior_factor = 1.0f; ray.dir = reflectFlip(w_i, h.normal); }
    ray.length = INFINITY; ray.is_shadow_ray = false;
    ray.origin = h.hit_point + ray.dir * EPSILON;
}

// ---- pdf helpers (udpt.cl:1033-1050, 1121-1124, 1149-1152) ----
float calcPhongPDF(const Scene& sc, f4 w_i, f4 w_o, const HitInfo& h)
{
    f4 refl_dir = 2 * (dot(w_o, h.normal)) * h.normal - w_o;
    refl_dir = normalize(refl_dir);
    const float costheta = fmaxf(0.0f, cosf(dot(refl_dir, w_i)));
    const float phong_exponent = matOf(sc, h).px + matOf(sc, h).py;
    return (phong_exponent + 1) * 0.5 * INV_PI * powf(costheta, phong_exponent);
}
inline float calcCosPDF(f4 w_i, f4 normal) { return fmaxf(dot(w_i, normal), 0.0f) * INV_PI; }
inline float getYluminance(f4 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }
inline float powerHeuristic(float weight, float pdf1, float pdf2) { return (weight * weight) / (pdf1 * pdf1 + pdf2 * pdf2); }

// ---- sampleLights (udpt.cl:632-698) ----
int sampleLights(const Scene& sc, const HitInfo& h, float& light_pdf, f4& w_i, Rng& rng)
{
    float sum = 0, weights[8]; f4 w_is[8];
    for (int i = 0; i < sc.n_lights; i++) {
        const Light& L = sc.lights[i];
        const float r1 = rng.next(P_LIGHT1 + 4 * i), r2 = rng.next(P_LIGHT2 + 4 * i);
        f4 temp_wi = (L.pos + r1 * L.edge_l + r2 * L.edge_w) - h.hit_point;
        const float distance = dot(temp_wi, temp_wi);
        w_is[i] = temp_wi;
        temp_wi = normalize(temp_wi);
        const float cosine_falloff = clmax(dot(temp_wi, h.normal), 0.0f) * clmax(dot(-temp_wi, L.normal), 0.0f);
        if (cosine_falloff <= 0.0) { weights[i] = 0; continue; }
        const float area = length(L.edge_l) * length(L.edge_w);
        if (sc.n_lights == 1) {
            w_i = w_is[i];
            light_pdf = 1 / area;
            light_pdf *= distance / fmaxf(dot(-temp_wi, L.normal), 0.0f);
            return i;
        }
        weights[i] = length(L.ke) * cosine_falloff * area / distance;
        sum += weights[i];
    }
    if (sum == 0) return -1;
    const float r1 = rng.next(P_PICK);
    float cumulative_weight = 0;
    for (int i = 0; i < sc.n_lights; i++) {
        const float weight = weights[i] / sum;
        if (r1 >= cumulative_weight && r1 < (cumulative_weight + weight)) {
            const Light& L = sc.lights[i];
            const float area = length(L.edge_l) * length(L.edge_w);
            w_i = w_is[i];
            light_pdf = (weight / area);
            light_pdf *= (dot(w_i, w_i) / (fmaxf(dot(-normalize(w_i), L.normal), 0.0f)));
            return i;
        }
        cumulative_weight += weight;
    }
    return -1;      // the reference falls off the end here (undefined behaviour); defined as "no light"
}

// ---- evaluateDirectLighting (udpt.cl:535-609; same text in bdpt.cl:642-716) ----
f4 evaluateDirectLighting(const Scene& sc, f4 w_o, const HitInfo& hit, Rng& rng)
{
    const f4 emission = tv(matOf(sc, hit).ke);
    f4 light_sample(0.f, 0.f, 0.f, 0.f), w_i;
    float light_pdf, brdf_prob = 0.0f;
    bool sample_glossy = false;
    const int j = sampleLights(sc, hit, light_pdf, w_i, rng);
    if (j == -1 || light_pdf <= 0.0f) return emission;
    float len = length(w_i);
    len -= EPSILON * 1.5f;
    w_i = normalize(w_i);
    Ray shadow_ray; shadow_ray.origin = hit.hit_point + w_i * EPSILON; shadow_ray.dir = w_i; shadow_ray.is_shadow_ray = true;
    HitInfo shadow_hitinfo = noHit();
    shadow_ray.length = len;
    if (!traceRay(sc, shadow_ray, shadow_hitinfo)) {
        sample_glossy = sampleGlossyPdf(sc, hit, rng, P_NEE_LOBE, brdf_prob);
        if (brdf_prob == 0.0f) return emission;
        light_sample = evaluateBRDF(sc, w_i, w_o, hit, sample_glossy, brdf_prob) * sc.lights[j].ke * fmaxf(dot(w_i, hit.normal), 0.0f);
        light_sample = light_sample * (1 / light_pdf);
    }
    if (!sc.mis) return light_sample + emission;

    f4 brdf_sample(0.f, 0.f, 0.f, 0.f);
    float brdf_pdf, mis_weight;
    if (sample_glossy) brdf_pdf = calcPhongPDF(sc, w_i, w_o, hit); else brdf_pdf = calcCosPDF(w_i, hit.normal);
    mis_weight = powerHeuristic(light_pdf, light_pdf, brdf_pdf);
    light_sample = light_sample * mis_weight;
    Ray brdf_sample_ray;
    if (sample_glossy) phongSampleHemisphere(sc, brdf_sample_ray, brdf_pdf, w_o, hit, rng, P_MIS1, P_MIS2);
    else cosineWeightedHemisphere(brdf_sample_ray, brdf_pdf, hit, rng, P_MIS1, P_MIS2);
    if (brdf_pdf <= 0.0f) return light_sample + emission;
    w_i = brdf_sample_ray.dir;
    HitInfo new_hitinfo = noHit();
    if (!traceRay(sc, brdf_sample_ray, new_hitinfo) || new_hitinfo.light_ID != j) return light_sample + emission;
    mis_weight = powerHeuristic(brdf_pdf, light_pdf, brdf_pdf);
    brdf_sample = evaluateBRDF(sc, w_i, w_o, hit, sample_glossy, brdf_prob) * sc.lights[j].ke * fmaxf(dot(w_i, hit.normal), 0.0f);
    brdf_sample = brdf_sample * (mis_weight / brdf_pdf);
    return (light_sample + brdf_sample + emission);
}

// ---- shading, unidirectional (udpt.cl:433-533) ----
f4 shadingUdpt(const Scene& sc, Ray ray, int GI_CHECK, Rng& rng)
{
    HitInfo hit_info = noHit();
    rng.vertex = 0;
    if (!traceRay(sc, ray, hit_info)) return f4(0.4f, 0.4f, 0.4f, 1.0f);
    if (hit_info.light_ID >= 0) {
        if (dot(ray.dir, sc.lights[hit_info.light_ID].normal) < 0) return f4(1, 1, 1, 1);
        return f4(0.1, .1, .1, 1);
    }
    f4 throughput(1.f, 1.f, 1.f, 1.f), direct_color(0.f, 0.f, 0.f, 0.f), indirect_color(0.f, 0.f, 0.f, 0.f);
    int matID = sc.tris[hit_info.triangle_ID].matID;
    if (!sc.mats[matID].is_specular) direct_color = evaluateDirectLighting(sc, -ray.dir, hit_info, rng);
    if (GI_CHECK) {
        for (int i = 0; i < 100000; i++) {
            HitInfo new_hitinfo = noHit();
            Ray new_ray;
            float pdf = 1, brdf_prob = 0.0f, ior_factor = 1.0f;
            bool is_glossy = false;
            rng.vertex = i;                      // draws made AT vertex i (the current hit)
            if (sc.mats[matID].is_specular) sampleFresnelIncidence(sc, new_ray, hit_info, -ray.dir, ior_factor, rng);
            else {
                is_glossy = sampleGlossyPdf(sc, hit_info, rng, P_LOBE, brdf_prob);
                if (brdf_prob == 0.0f) break;
                if (is_glossy) phongSampleHemisphere(sc, new_ray, pdf, -ray.dir, hit_info, rng, P_DIR1, P_DIR2);
                else cosineWeightedHemisphere(new_ray, pdf, hit_info, rng, P_DIR1, P_DIR2);
            }
            if (pdf <= 0.0f || !traceRay(sc, new_ray, new_hitinfo) || new_hitinfo.light_ID >= 0) {
                if (sc.mats[matID].is_specular && new_hitinfo.light_ID >= 0) indirect_color = indirect_color + throughput * sc.lights[new_hitinfo.light_ID].ke;
                break;
            }
            indirect_color = indirect_color + throughput * tv(matOf(sc, new_hitinfo).ke);
            if (sc.mats[matID].is_specular) throughput = throughput * ior_factor;
            else throughput = throughput * evaluateBRDF(sc, new_ray.dir, -ray.dir, hit_info, is_glossy, brdf_prob) * fmaxf(dot(new_ray.dir, hit_info.normal), 0.0f) / pdf;
            rng.vertex = i + 1;                  // the roulette and the NEE belong to the vertex just reached
            if (i > sc.rr_threshold) {
                const float p = clmin(getYluminance(throughput), 0.95f);
                const float r = rng.next(P_RR);
                if (r >= p) break;
                throughput = throughput * (1 / p);
            }
            matID = sc.tris[new_hitinfo.triangle_ID].matID;
            if (!sc.mats[matID].is_specular) indirect_color = indirect_color + throughput * evaluateDirectLighting(sc, -new_ray.dir, new_hitinfo, rng);
            hit_info = new_hitinfo;
            ray = new_ray;
        }
    }
    return direct_color + indirect_color;
}

// ---- bidirectional (bdpt.cl:432-640) ----
const int MAX_BDPT = 32;
void createLightPath(const Scene& sc, PathInfo* light_path, int& path_length, Rng& rng)
{
    // the reference always emits from light_sources[0]; with several lights the emitter is drawn with probability
    // proportional to |ke| * area (extension) and the vertex weight is divided by that probability
    int li = 0; float pick_prob = 1.0f;
    if (sc.n_lights > 1) {
        float w[8], sum = 0;
        for (int i = 0; i < sc.n_lights; i++) { w[i] = length(sc.lights[i].ke) * (length(sc.lights[i].edge_l) * length(sc.lights[i].edge_w)); sum += w[i]; }
        rng.vertex = 0x40000000u;
        const float r = rng.next(P_PICK);
        float cum = 0; li = sc.n_lights - 1;
        for (int i = 0; i < sc.n_lights; i++) { const float p = w[i] / sum; if (r >= cum && r < cum + p) { li = i; break; } cum += p; }
        pick_prob = w[li] / sum;
    }
    const Light& L = sc.lights[li];
    Ray light_ray;
    HitInfo hit = noHit(); hit.normal = L.normal;
    float pdf = 1;
    rng.vertex = 0x40000000u;                       // light-path vertex 0
    const float r1 = rng.next_via_double(P_LIGHT1);  // bdpt.cl:439
    const float r2 = rng.next(P_LIGHT2);
    const f4 A = L.edge_l * r2, B = L.edge_w * r1;
    hit.hit_point = (A + B) + L.pos;
    hit.normal = L.normal;
    light_path[0].hit_info = hit; light_path[0].hit_info.light_ID = li;
    const float area = length(L.edge_l) * length(L.edge_w);
    light_path[0].fwd_pdf = 1.0f / area; light_path[0].rev_pdf = 1.0;
    light_path[0].contrib = L.ke / light_path[0].fwd_pdf;
    if (sc.n_lights > 1) light_path[0].contrib = light_path[0].contrib / pick_prob;
    cosineWeightedHemisphere(light_ray, pdf, hit, rng, P_DIR1, P_DIR2);
    if (pdf <= 0.0f || !traceRay(sc, light_ray, hit) || hit.light_ID >= 0) return;
    light_path[1].hit_info = hit; light_path[1].dir = light_ray.dir;
    light_path[1].fwd_pdf = pdf; light_path[1].rev_pdf = 1.0f;
    const float c1 = clmax(0.0f, dot(light_path[0].hit_info.normal, light_path[1].dir)) / pdf;
    light_path[1].contrib = f4(c1, c1, c1, c1);
    light_path[1].contrib = light_path[1].contrib * light_path[0].contrib;
    float brdf_prob = 0.0f;
    path_length++;
    for (int i = 2; i < sc.bdpt_bounces; i++) {
        rng.vertex = 0x40000000u + (i - 1);         // draws made at light vertex i-1
        const bool is_glossy_bounce = sampleGlossyPdf(sc, hit, rng, P_LOBE, brdf_prob);
        if (brdf_prob == 0.0f) break;
        if (is_glossy_bounce) phongSampleHemisphere(sc, light_ray, pdf, -light_path[i - 1].dir, hit, rng, P_DIR1, P_DIR2);
        else cosineWeightedHemisphere(light_ray, pdf, hit, rng, P_DIR1, P_DIR2);
        if (!traceRay(sc, light_ray, hit) || hit.light_ID >= 0 || pdf <= 0.0f) break;
        path_length++;
        light_path[i].hit_info = hit; light_path[i].dir = light_ray.dir;
        light_path[i].fwd_pdf = pdf; light_path[i].rev_pdf = 1.0;
        light_path[i].contrib = evaluateBRDF(sc, -light_path[i - 1].dir, light_path[i].dir, light_path[i - 1].hit_info, is_glossy_bounce, brdf_prob)
                                * fmaxf(0.0f, dot(light_path[i].dir, light_path[i - 1].hit_info.normal));
        light_path[i].contrib = light_path[i].contrib / pdf;
        light_path[i].contrib = light_path[i].contrib * light_path[i - 1].contrib;
        if (i > sc.rr_threshold) {
            rng.vertex = 0x40000000u + i;
            const float r = rng.next(P_RR);
            const float p = clmin(getYluminance(light_path[i].contrib), 0.95f);
            if (r >= p) break;
            light_path[i].contrib = light_path[i].contrib * (1.0f / p);
        }
    }
}
void createEyePath(const Scene& sc, PathInfo* eye_path, int& path_length, Ray eye_ray, Rng& rng)
{
    HitInfo hit = noHit();
    eye_path[0].hit_info = hit;
    if (!traceRay(sc, eye_ray, eye_path[0].hit_info)) return;
    // The reference continues here even when the camera ray hit the LIGHT (triangle_ID = -1) and reads scene_data[-1]
    // (bdpt.cl:525-528 -> :1055): an out-of-bounds read whose result never reaches the image, because shading() returns the
    // constant light colour for such a pixel (bdpt.cl:575-581).  The restatement stops instead of reading out of bounds.
    if (eye_path[0].hit_info.light_ID >= 0) return;
    eye_path[0].dir = eye_ray.dir; eye_path[0].fwd_pdf = 1.0f; eye_path[0].contrib = f4(1.0f, 1.0f, 1.0f, 1.0f);
    float pdf = 1, brdf_prob = 0.0f;
    for (int i = 1; i < sc.bdpt_bounces; i++) {
        rng.vertex = i - 1;                          // draws made at eye vertex i-1
        const bool is_glossy_bounce = sampleGlossyPdf(sc, eye_path[i - 1].hit_info, rng, P_LOBE, brdf_prob);
        if (brdf_prob == 0.0f) break;
        if (is_glossy_bounce) phongSampleHemisphere(sc, eye_ray, pdf, -eye_path[i - 1].dir, eye_path[i - 1].hit_info, rng, P_DIR1, P_DIR2);
        else cosineWeightedHemisphere(eye_ray, pdf, eye_path[i - 1].hit_info, rng, P_DIR1, P_DIR2);
        if (pdf <= 0.0f || !traceRay(sc, eye_ray, hit) || hit.light_ID >= 0) break;
        path_length++;
        eye_path[i].hit_info = hit; eye_path[i].dir = eye_ray.dir; eye_path[i].fwd_pdf = pdf;
        eye_path[i].contrib = evaluateBRDF(sc, eye_path[i].dir, -eye_path[i - 1].dir, eye_path[i - 1].hit_info, is_glossy_bounce, brdf_prob)
                              * clmax(0.0f, dot(eye_path[i].dir, eye_path[i - 1].hit_info.normal));
        eye_path[i].contrib = eye_path[i].contrib / pdf;
        eye_path[i].contrib = eye_path[i].contrib * eye_path[i - 1].contrib;
        if (i > sc.rr_threshold) {
            rng.vertex = i;
            const float r = rng.next(P_RR);
            const float p = clmin(getYluminance(eye_path[i].contrib), 0.95f);
            if (r >= p) break;
            eye_path[i].contrib = eye_path[i].contrib * (1.0f / p);
        }
    }
}
f4 shadingBdpt(const Scene& sc, Ray ray, Rng& rng)
{
    PathInfo light_path[MAX_BDPT], eye_path[MAX_BDPT];
    int lp_len = 1, ep_len = 1;
    createLightPath(sc, light_path, lp_len, rng);
    createEyePath(sc, eye_path, ep_len, ray, rng);
    if (eye_path[0].hit_info.light_ID >= 0) {
        if (dot(ray.dir, sc.lights[eye_path[0].hit_info.light_ID].normal) < 0) return f4(1, 1, 1, 1);
        return f4(0.1, 0.1, 0.1, 1);
    } else if (eye_path[0].hit_info.triangle_ID < 0) return f4(0.4f, 0.4f, 0.4f, 1.0f);
    f4 throughput(1.f, 1.f, 1.f, 1.f), color(0.f, 0.f, 0.f, 1.f), subpaths_color;
    float eye_path_weight = 1.0, ks = 0.0;
    for (int i = 0; i < ep_len; i++) {
        if (eye_path_weight == 0) break;
        subpaths_color = f4(0.f, 0.f, 0.f, 1.f);
        const f4 spec_color = tv(matOf(sc, eye_path[i].hit_info).ks), emission = tv(matOf(sc, eye_path[i].hit_info).ke);
        ks = clmax(clmax(spec_color.x, clmax(spec_color.y, spec_color.z)), 0.1f);
        throughput = eye_path[i].contrib;
        rng.vertex = 0x20000000u + i;                // NEE at eye vertex i
        color = color + throughput * (emission + evaluateDirectLighting(sc, -eye_path[i].dir, eye_path[i].hit_info, rng)) * eye_path_weight;
        for (int j = lp_len - 1; j > 0; j--) {
            Ray determ_ray; HitInfo determ_hit = noHit();
            f4 throughput_lp(1.0f, 1.0f, 1.0f, 1.0f);
            determ_ray.dir = normalize(light_path[j].hit_info.hit_point - eye_path[i].hit_info.hit_point);
            determ_ray.origin = eye_path[i].hit_info.hit_point + determ_ray.dir * EPSILON;
            float dist = length(light_path[j].hit_info.hit_point - eye_path[i].hit_info.hit_point);
            dist *= dist;
            determ_ray.length = length(light_path[j].hit_info.hit_point - determ_ray.origin);
            determ_ray.is_shadow_ray = true;
            if (dot(determ_ray.dir, eye_path[i].hit_info.normal) <= 0 || dot(-determ_ray.dir, light_path[j].hit_info.normal) <= 0) continue;
            if (!traceRay(sc, determ_ray, determ_hit)) {
                throughput_lp = light_path[j].contrib;
                const f4 w_i = determ_ray.dir, w_o = -eye_path[i].dir;
                const HitInfo hit = eye_path[i].hit_info;
                const float gf = clmax(dot(w_i, hit.normal), 0.0f) * clmax(dot(-w_i, light_path[j].hit_info.normal), 0.0f) / dist;
                float prob = 0.0f;
                rng.vertex = 0x10000000u + (uint32_t)(i * MAX_BDPT + j);      // the two lobe draws of connection (i, j)
                bool g = sampleGlossyPdf(sc, hit, rng, P_LOBE, prob);
                const f4 eye_to_light_brdf = evaluateBRDF(sc, w_i, w_o, hit, g, prob);
                g = sampleGlossyPdf(sc, light_path[j].hit_info, rng, P_NEE_LOBE, prob);
                const f4 light_to_eye_brdf = evaluateBRDF(sc, -light_path[j].dir, -w_i, light_path[j].hit_info, g, prob);
                throughput_lp = throughput_lp * (gf * eye_to_light_brdf * light_to_eye_brdf);
                throughput_lp = throughput_lp * throughput;
                subpaths_color = subpaths_color + throughput_lp;
            }
        }
        color = color + subpaths_color * (eye_path_weight * (1 - ks));
        eye_path_weight = eye_path_weight * ks;
    }
    return color;
}

Light unpackLight(const yune_quad_light& q) { Light L; L.pos = f4(q.pos); L.normal = f4(q.normal); L.ke = f4(q.ke); L.edge_l = f4(q.edge_l); L.edge_w = f4(q.edge_w); return L; }

} // namespace

extern "C" {

struct yor_config {
    int integrator;      // 0 = udpt.cl, 1 = bdpt.cl
    int mis;             // udpt.cl -DMIS
    int rng_mode;        // 0 = reference stream (uses `rand` per frame), 1 = counter-based (uses seed + sample index)
    int n_lights;        // 0 = the kernel's built-in light
    int rr_threshold;    // < 0 = kernel default (6 / 4)
    int oren_nayar;
    int heap_size;       // < 0 = reference default 1500; 0 = unbounded
    int bdpt_bounces;    // <= 0 = 20
    uint32_t seed;
    int threads;         // <= 0 = all
};

static void fillScene(Scene& sc, const yor_config* cfg, const yune_quad_light* lights, const yune_triangle* tris, int ntri,
                      const yune_material* mats, const yune_bvh_node* nodes, int nnodes)
{
    sc.tris = tris; sc.ntri = ntri; sc.mats = mats; sc.bvh = nodes; sc.nnodes = nnodes;
    sc.bdpt = cfg->integrator == 1; sc.mis = cfg->mis; sc.oren_nayar = cfg->oren_nayar;
    sc.rr_threshold = cfg->rr_threshold >= 0 ? cfg->rr_threshold : (sc.bdpt ? 4 : 6);
    sc.heap_size = cfg->heap_size < 0 ? 1500 : cfg->heap_size;
    sc.bdpt_bounces = cfg->bdpt_bounces > 0 ? (cfg->bdpt_bounces < MAX_BDPT ? cfg->bdpt_bounces : MAX_BDPT) : 20;
    if (cfg->n_lights > 0) { sc.n_lights = cfg->n_lights < 8 ? cfg->n_lights : 8; for (int i = 0; i < sc.n_lights; i++) sc.lights[i] = unpackLight(lights[i]); }
    else {
        sc.n_lights = 1;
        Light& L = sc.lights[0];
        if (sc.bdpt) { L.pos = f4(-0.1979f, 0.703f, -3.1972f, 1.f); L.normal = f4(0.f, 1.f, 0.f, 0.f); L.ke = f4(18.3f, 16.2f, 14.5f, 0.f); }   // bdpt.cl:106-115
        else         { L.pos = f4(-0.1979f, 0.92f, -3.1972f, 1.f);  L.normal = f4(0.f, -1.f, 0.f, 0.f); L.ke = f4(16.f, 16.f, 16.f, 0.f); }     // udpt.cl:97-106
        L.edge_l = f4(0.4f, 0.f, 0.f, 0.f); L.edge_w = f4(0.f, 0.f, 0.4f, 0.f);
    }
}

// One frame = one sample per pixel, the kernel entry of udpt.cl:158-211 / bdpt.cl:158-210 (block loop folded: the
// block -> pixel mapping of :164-170 only decides WHEN a pixel is computed, never its value).
// rng_mode 0: `frame_arg` is the per-frame `rand` (kernel arg 10).  rng_mode 1: `frame_arg` is the sample index.
// out/in: RGBA32F running mean + count in alpha, row 0 = bottom row; reset as kernel arg 9.
void yor_render_frame(const yor_config* cfg, const yune_quad_light* lights, float* out_rgba, const float* in_rgba, const yune_cam* cam,
                      const yune_triangle* tris, int ntri, const yune_material* mats, const yune_bvh_node* nodes, int nnodes,
                      int gi_check, int reset, uint32_t frame_arg, int W, int H)
{
    Scene sc; fillScene(sc, cfg, lights, tris, ntri, mats, nodes, nnodes);
    #pragma omp parallel for schedule(dynamic, 1) num_threads(cfg->threads > 0 ? cfg->threads : 1024) if (cfg->threads != 1)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            Rng rng; rng.mode = cfg->rng_mode; rng.seed = cfg->seed; rng.pixel = (uint32_t)(y * W + x); rng.sample = frame_arg; rng.vertex = VERTEX_CAMERA;
            float r1, r2;
            if (cfg->rng_mode == 0) {
                uint seed = (y + 1) * W + (x + 1);
                seed = frame_arg * seed;
                seed = wang_hash(seed);
                if (seed == 0) seed = wang_hash(seed);
                rng.state = seed;
            }
            r1 = rng.next(0); r2 = rng.next(1);
            Ray eye_ray;
            createRay(x + r1, y + r2, W, H, eye_ray, *cam);
            f4 color = sc.bdpt ? shadingBdpt(sc, eye_ray, rng) : shadingUdpt(sc, eye_ray, gi_check, rng);
            if (anynan(color)) color = f4(0.988f, 0.0588f, 0.7529f, 1.0f);
            float* o = out_rgba + 4 * ((size_t)y * W + x);
            if (reset == 1) { o[0] = color.x; o[1] = color.y; o[2] = color.z; o[3] = 1; }
            else {
                const float* p = in_rgba + 4 * ((size_t)y * W + x);
                const int num_passes = p[3];
                color = color + (f4(p[0], p[1], p[2], p[3]) * (float)num_passes);
                color = color / (float)(num_passes + 1);
                o[0] = color.x; o[1] = color.y; o[2] = color.z; o[3] = num_passes + 1;
            }
        }
}

// Same, but returns the per-sample radiance (no averaging): out_rgb[pixel] = the value `color` after the NaN->PINK rule.
void yor_render_samples(const yor_config* cfg, const yune_quad_light* lights, float* out_rgba, const yune_cam* cam,
                        const yune_triangle* tris, int ntri, const yune_material* mats, const yune_bvh_node* nodes, int nnodes,
                        int gi_check, uint32_t frame_arg, int W, int H)
{
    yor_render_frame(cfg, lights, out_rgba, out_rgba, cam, tris, ntri, mats, nodes, nnodes, gi_check, 1, frame_arg, W, H);
}

// Primary rays (createRay + first traceRay), with traversal work counters of the REFERENCE walk.
void yor_primary(const yor_config* cfg, const yune_quad_light* lights, const yune_cam* cam, const yune_triangle* tris, int ntri,
                 const yune_bvh_node* nodes, int nnodes, uint32_t rand_seed, int jitter_mode, int W, int H,
                 int* tri_id, int* light_id, float* t_hit, float* ray_od6, unsigned long long* work4)
{
    Scene sc; fillScene(sc, cfg, lights, tris, ntri, nullptr, nodes, nnodes);
    Work w = {0, 0, 0, 0};
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float r1 = 0.5f, r2 = 0.5f;
            if (jitter_mode == 1) {
                Rng rng; rng.mode = 0;
                uint seed = (y + 1) * W + (x + 1); seed = rand_seed * seed; seed = wang_hash(seed); if (seed == 0) seed = wang_hash(seed);
                rng.state = seed; r1 = rng.next(0); r2 = rng.next(1);
            }
            Ray ray; createRay(x + r1, y + r2, W, H, ray, *cam);
            HitInfo hit = noHit();
            const bool any = traceRay(sc, ray, hit, &w);
            const size_t i = (size_t)y * W + x;
            tri_id[i] = any ? hit.triangle_ID : -1; light_id[i] = any ? hit.light_ID : -1;
            if (t_hit) t_hit[i] = ray.length;
            if (ray_od6) { float* r = ray_od6 + 6 * i; r[0] = ray.origin.x; r[1] = ray.origin.y; r[2] = ray.origin.z; r[3] = ray.dir.x; r[4] = ray.dir.y; r[5] = ray.dir.z; }
        }
    if (work4) { work4[0] = w.box; work4[1] = w.tri; work4[2] = w.overflow; work4[3] = (unsigned long long)w.max_queue; }
}

// Arbitrary rays through traceRay (closest hit or any-hit), reference walk.
void yor_trace(const yor_config* cfg, const yune_quad_light* lights, int n, const float* od6, const float* tmax, int shadow,
               const yune_triangle* tris, int ntri, const yune_bvh_node* nodes, int nnodes, int* tri_id, int* light_id, float* t_hit)
{
    Scene sc; fillScene(sc, cfg, lights, tris, ntri, nullptr, nodes, nnodes);
    #pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        const float* r = od6 + 6 * (size_t)i;
        Ray ray; ray.origin = f4(r[0], r[1], r[2], 1.0f); ray.dir = f4(r[3], r[4], r[5], 0.0f);
        ray.length = tmax ? tmax[i] : INFINITY; ray.is_shadow_ray = shadow != 0;
        HitInfo hit = noHit();
        const bool any = traceRay(sc, ray, hit);
        tri_id[i] = any ? hit.triangle_ID : -1; light_id[i] = any ? hit.light_ID : -1;
        if (t_hit) t_hit[i] = ray.length;
    }
}

// ---- the roofline's work model: ray-box and ray-triangle tests of a FRONT-TO-BACK ORDERED STACK WALK WITH
// t-PRUNING over the reference BVH, counted here so that neither the product nor the bench can choose the numbers. ----
namespace {
struct Entry { int node; float t; };
float boxEntry(const Ray& ray, const yune_aabb& bb, bool& hit)
{
    float t_max = INFINITY, t_min = -INFINITY;
    for (int k = 0; k < 3; k++) {
        const float o = (&ray.origin.x)[k], inv = 1 / (&ray.dir.x)[k];
        const float a = (bb.p_min.s[k] - o) * inv, b = (bb.p_max.s[k] - o) * inv;
        if (!(a != a)) { t_min = fmaxf(clmin(a, b), t_min); t_max = clmin(fmaxf(a, b), t_max); }
    }
    hit = t_max > fmaxf(t_min, 0.0f);
    return fmaxf(t_min, 0.0f);
}
}
void yor_count_work(int n, const float* od6, const float* tmax, int shadow, const yune_triangle* tris, int ntri,
                    const yune_bvh_node* nodes, int nnodes, unsigned long long* work2)
{
    unsigned long long nb = 0, nt = 0;
    Scene sc; std::memset(&sc, 0, sizeof(sc)); sc.tris = tris; sc.ntri = ntri; sc.bvh = nodes; sc.nnodes = nnodes; sc.n_lights = 0;
    #pragma omp parallel for schedule(dynamic, 256) reduction(+ : nb, nt)
    for (int i = 0; i < n; i++) {
        const float* r = od6 + 6 * (size_t)i;
        Ray ray; ray.origin = f4(r[0], r[1], r[2], 1.0f); ray.dir = f4(r[3], r[4], r[5], 0.0f);
        ray.length = tmax ? tmax[i] : INFINITY; ray.is_shadow_ray = shadow != 0;
        HitInfo hit = noHit();
        bool h; nb++;
        boxEntry(ray, nodes[0].aabb, h);
        if (!h) continue;
        Entry stack[128]; int sp = 0; stack[sp++] = {0, 0.0f};
        bool done = false;
        while (sp > 0 && !done) {
            const Entry e = stack[--sp];
            if (e.t > ray.length) continue;
            const yune_bvh_node& nd = nodes[e.node];
            if (nd.child_idx == -1 && nd.vert_len > 0) {
                for (int j = 0; j < nd.vert_len; j++) { nt++; if (rayTriangle(sc, ray, hit, nd.vert_list[j], nullptr) && shadow) { done = true; break; } }
                continue;
            }
            if (nd.child_idx <= 0) continue;
            Entry c[2]; int nc = 0;
            for (int j = nd.child_idx; j < nd.child_idx + 2; j++) {
                if (!(nodes[j].vert_len > 0 || nodes[j].child_idx > 0)) continue;
                nb++;
                bool hh; const float t = boxEntry(ray, nodes[j].aabb, hh);
                if (hh && !(t > ray.length)) c[nc++] = {j, t};
            }
            if (nc == 2 && c[1].t < c[0].t) { Entry tmp = c[0]; c[0] = c[1]; c[1] = tmp; }
            for (int k = nc - 1; k >= 0; k--) if (sp < 128) stack[sp++] = c[k];
        }
    }
    work2[0] = nb; work2[1] = nt;
}

// tonemap.cl:14-47
void yor_tonemap(const float* in_rgba, float* out_rgba, int n)
{
    for (int i = 0; i < n; i++) {
        const f4 col(in_rgba[4 * i], in_rgba[4 * i + 1], in_rgba[4 * i + 2], in_rgba[4 * i + 3]);
        const float lum_white = 1.0f;
        const float lum_world = 0.212671f * col.x + 0.715160f * col.y + 0.072169f * col.z + 0.001f;
        const float lum_display = lum_world * (1 + lum_world / (lum_white * lum_white)) / (1 + lum_world);
        const f4 q = col / lum_world;
        f4 l = lum_display * f4(powf(q.x, 1.0f), powf(q.y, 1.0f), powf(q.z, 1.0f), powf(q.w, 1.0f));
        const float g = 1 / 2.2f;
        out_rgba[4 * i] = powf(l.x, g); out_rgba[4 * i + 1] = powf(l.y, g); out_rgba[4 * i + 2] = powf(l.z, g); out_rgba[4 * i + 3] = powf(l.w, g);
    }
}

unsigned yor_wang_hash(unsigned s) { return wang_hash(s); }
unsigned yor_xor_shift(unsigned s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
void yor_philox(unsigned seed, unsigned pixel, unsigned sample, unsigned vertex, unsigned block, unsigned* out4)
{
    uint32_t c[4] = {pixel, sample, vertex, block};
    philox(c, seed, 0x59554e45u);
    out4[0] = c[0]; out4[1] = c[1]; out4[2] = c[2]; out4[3] = c[3];
}

}
