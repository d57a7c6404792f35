// kernel_common.cuh -- device helpers shared by the shade kernels (kernels.cu, bdpt.cu).
#ifndef YUNE_KERNEL_COMMON_CUH
#define YUNE_KERNEL_COMMON_CUH

#include "kernels.h"
#include "shade_core.h"
#include "rng.h"

namespace yune {

__device__ __forceinline__ V3 xyz(const float4& f) { return v3(f.x, f.y, f.z); }
__device__ __forceinline__ float4 f4(V3 a, float w) { return make_float4(a.x, a.y, a.z, w); }

// warp-aggregated slot allocation in a global counter: returns this lane's index, or -1 for lanes with want == false
__device__ __forceinline__ int warp_alloc(int* counter, bool want)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return -1;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    return want ? base + __popc(m & ((1u << lane) - 1)) : -1;
}

// Block-wide queue allocation: every thread of the block calls it with up to K request counts (0..3 each); ONE global
// atomic per counter per block.  s_cnt is shared memory of K * (warps + 1) ints.  Returns, per counter, the first index this
// thread owns (contiguous run of `cnt[k]` entries).  Two __syncthreads.  (One atomic per WARP on the same line was the
// bottleneck of the shade stage: ~260 K same-line L2 atomics per iteration.)
template <int K, int NWARPS>
__device__ __forceinline__ void block_alloc(int* const (&counters)[K], const int (&cnt)[K], int (&first)[K], int* s_cnt)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int excl[K];
    #pragma unroll
    for (int k = 0; k < K; k++) {
        int incl = cnt[k];
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        excl[k] = incl - cnt[k];
        if (lane == 31) s_cnt[k * (NWARPS + 1) + warp] = incl;
    }
    __syncthreads();
    if (warp < K) {
        const int k = warp;
        const int v = lane < NWARPS ? s_cnt[k * (NWARPS + 1) + lane] : 0;
        int incl = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 0 && total > 0) base = atomicAdd(counters[k], total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < NWARPS) s_cnt[k * (NWARPS + 1) + lane] = base + incl - v;
    }
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < K; k++) first[k] = s_cnt[k * (NWARPS + 1) + warp] + excl[k];
}

#define YUNE_NW (YUNE_SHADE_BLOCK / 32)

// append to a shared-memory list (persistent shade kernels): one shared atomic per warp
__device__ __forceinline__ void list_push(int* list, int* count, bool want, int value)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

__device__ __forceinline__ void list_push_u16(unsigned short* list, int* count, bool want, unsigned short value)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

#define YUNE_FIX_SCALE 16777216.0f            /* 2^24 */
__device__ __forceinline__ long long to_fixed(float x)
{
    return __float2ll_rn(fminf(fmaxf(x, -2.7487791e11f), 2.7487791e11f) * YUNE_FIX_SCALE);      // clamp to +-2^38, round to nearest even
}

// a finished sample goes to the fp32 accumulation buffer (udpt.cl:193-210; sum instead of running mean)
__device__ __forceinline__ void finish_sample(const RenderArgs& A, unsigned pixel, V3 col)
{
    if (col.x != col.x || col.y != col.y || col.z != col.z) col = v3(0.988f, 0.0588f, 0.7529f);     // PINK (udpt.cl:193-194)
    if (A.fix) {
        // Deterministic accumulation: integer addition is associative, so the sum does not depend on the order in which samples
        // finish, on the pool size, on how a sample range is split into calls or over GPUs.  One unit = 2^-24 (a sample is
        // quantised to ~6e-8 absolute, below the fp32 resolution of any pixel mean >= 1); |x| is clamped to 2^38.
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(A.fix) + 4 * (size_t)pixel;
        atomicAdd(dst + 0, (unsigned long long)to_fixed(col.x)); atomicAdd(dst + 1, (unsigned long long)to_fixed(col.y));
        atomicAdd(dst + 2, (unsigned long long)to_fixed(col.z)); atomicAdd(dst + 3, 1ull);
        return;
    }
    float* dst = reinterpret_cast<float*>(A.sum + pixel);
    atomicAdd(dst + 0, col.x); atomicAdd(dst + 1, col.y); atomicAdd(dst + 2, col.z); atomicAdd(dst + 3, 1.0f);
}

// camera ray (createRay, udpt.cl:213-238); fp64 where the kernel's literals make it fp64
__device__ __forceinline__ void create_ray(const float* cam, int W, int H, float pixel_x, float pixel_y, V3& o, V3& d)
{
    const float aspect_ratio = (float)(((double)W * 1.0) / (double)H);
    V3 dir;
    dir.x = (float)((double)aspect_ratio * ((2.0 * (double)pixel_x / (double)W) - 1.0));
    dir.y = (float)((2.0 * (double)pixel_y / (double)H) - 1.0);
    dir.z = -cam[16];
    V3 w;
    w.x = vdot(v3(cam[0], cam[1], cam[2]), dir);
    w.y = vdot(v3(cam[4], cam[5], cam[6]), dir);
    w.z = vdot(v3(cam[8], cam[9], cam[10]), dir);
    d = vnormalize(w)                  /* exact: primary rays are compared bit for bit */;
    o = v3(cam[3], cam[7], cam[11]);
}

__device__ __forceinline__ MatDev load_material(const float4* mats, int id)
{
    const float4* p = mats + 5 * (size_t)id;
    const float4 ke = __ldg(p), kd = __ldg(p + 1), ks = __ldg(p + 2), a = __ldg(p + 3), b = __ldg(p + 4);
    MatDev m;
    m.ke = xyz(ke); m.kd = xyz(kd); m.ks = xyz(ks);
    m.n = a.x; m.px = a.z; m.py = a.w; m.alpha_x = b.x;
    m.is_specular = __float_as_int(b.z); m.is_transmissive = __float_as_int(b.w);
    return m;
}

// 30-bit Morton code of a point inside the scene's (padded) root box: the sorting key of option "sort_rays" (rays that start
// close to each other walk the same part of a tree that does not fit the caches).
__device__ __forceinline__ unsigned spread10(unsigned x)
{
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8))  & 0x0300f00fu;
    x = (x | (x << 4))  & 0x030c30c3u;
    x = (x | (x << 2))  & 0x09249249u;
    return x;
}
__device__ __forceinline__ unsigned origin_key(const DevScene& sc, V3 o)
{
    const float ux = (o.x - sc.root_lo[0]) / (sc.root_hi[0] - sc.root_lo[0]), uy = (o.y - sc.root_lo[1]) / (sc.root_hi[1] - sc.root_lo[1]),
                uz = (o.z - sc.root_lo[2]) / (sc.root_hi[2] - sc.root_lo[2]);
    const unsigned qx = (unsigned)fminf(fmaxf(ux * 1024.0f, 0.0f), 1023.0f), qy = (unsigned)fminf(fmaxf(uy * 1024.0f, 0.0f), 1023.0f),
                   qz = (unsigned)fminf(fmaxf(uz * 1024.0f, 0.0f), 1023.0f);
    return (spread10(qx) << 2) | (spread10(qy) << 1) | spread10(qz);
}

struct ShadowOut { bool has; V3 o, d; float tmax; };

// What evaluateDirectLighting (udpt.cl:535-609 == bdpt.cl:642-716) leaves to be resolved by rays:
//   visible(S) ? Lv + (visible(MV) ? BV : 0) : (visible(MO) ? BO : 0)
// S = shadow ray to the light sample; MV / MO = BRDF-sampled ray of the branch "S visible" / "S occluded" (MIS only), present
// only when it meets light j analytically.  mo_is_mv: both branches sampled the same direction, MV answers for MO.
struct NeeOut {
    ShadowOut S, MV, MO;
    bool mo_is_mv;
    V3 Lv, BV, BO;
};

// BDPT = the variants bdpt.cl uses inside the same function: un-flipped reflection (bdpt.cl:723, 814), padded lobe
// probabilities (bdpt.cl:1089-1104).
//
// `act` = the lanes of the warp that call this together.  ptxas threads jumps through the early-outs of this function and
// lets the two sides of an earlier branch run the REST of the code as separate groups (ncu: every SASS line executed twice
// per warp at half the lanes); the __syncwarp()s on explicit masks pin the reconvergence points.
template <bool MIS, bool BDPT>
__device__ __forceinline__ void nee_sample(const unsigned act, const LightDev* lights, int n_lights, const MatDev& mat, const LobePrep& lobes,
                                           V3 hp, V3 n, V3 Nx, V3 Ny, V3 w_o,
                                           U4 u_nee, uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t vertex, bool use_on, NeeOut& R)
{
    R.S.has = R.MV.has = R.MO.has = false; R.mo_is_mv = false;
    R.Lv = v3(0, 0, 0); R.BV = v3(0, 0, 0); R.BO = v3(0, 0, 0);
    float u_l[2 * YUNE_MAX_LIGHTS]; float u_pick = 0.0f;
    YUNE_NO_UNROLL
    for (int i = 0; i < n_lights; i++) {
        const U4 ul = draw4(seed, pixel, sample, vertex, YUNE_BLK_LIGHT + i);
        u_l[2 * i] = u01(ul.x); u_l[2 * i + 1] = u01(ul.y);
        if (i == 0) u_pick = u01(ul.z);
    }
    float light_pdf = 0.0f; V3 w_i = v3(0, 0, 0);
    const int j = sample_lights(lights, n_lights, hp, n, u_l, u_pick, light_pdf, w_i);
    const bool lit = !(j == -1 || light_pdf <= 0.0f);
    const unsigned m_lit = __ballot_sync(act, lit);
    if (!lit) return;
    float len = vlength(w_i);
    len = YF_SUB(len, YF_MUL(YUNE_EPS, 1.5f));
    w_i = vnormalize(w_i);
    R.S.o = vadd(hp, vscale(w_i, YUNE_EPS)); R.S.d = w_i; R.S.tmax = len;
    float tl = len;
    R.S.has = !(light_loop(lights, n_lights, R.S.o, R.S.d, tl) >= 0);        // traceRay's light loop (:244-276) can already block it
    // branch "light sample visible": lobe selection happens only then (:559-561)
    float prob = 0.0f;
    const bool glossy = select_lobe_r(lobes, u01(u_nee.y), BDPT, prob);
    const bool v_alive = prob != 0.0f;
    const V3 Lke = lights[j].ke;
    if (v_alive) {
        R.Lv = vscale(vmul(eval_brdf(mat, w_i, w_o, n, glossy, prob, !BDPT, use_on), Lke), fmaxf(vdot(w_i, n), 0.0f));
        R.Lv = vscale(R.Lv, YF_DIV(1.0f, light_pdf));
    }
    if (!MIS) return;
    __syncwarp(m_lit);
    const float r1 = u01(u_nee.z), r2 = u01(u_nee.w);
    float pdfV = 0.0f, pdfO = 0.0f;
    V3 dv = v3(0, 0, 1), dq;
    if (v_alive) {
        const float brdf_pdf = glossy ? phong_pdf(mat, w_i, w_o, n) : cos_pdf(w_i, n);
        R.Lv = vscale(R.Lv, power_heuristic(light_pdf, light_pdf, brdf_pdf));                       // :570-577
        dv = glossy ? sample_phong(w_o, n, mat.px, mat.py, r1, r2, !BDPT, pdfV) : sample_cosine_onb(n, Nx, Ny, r1, r2, pdfV);
        if (pdfV > 0.0f) {                                                                          // :587-588
            R.MV.o = vadd(hp, vscale(dv, YUNE_EPS)); R.MV.d = dv; R.MV.tmax = INFINITY;
            if (light_loop(lights, n_lights, R.MV.o, R.MV.d, R.MV.tmax) == j) {                     // closest light must be j (:594)
                R.MV.has = true;
                R.BV = vscale(vmul(eval_brdf(mat, dv, w_o, n, glossy, prob, !BDPT, use_on), Lke), fmaxf(vdot(dv, n), 0.0f));
                R.BV = vscale(R.BV, YF_DIV(power_heuristic(pdfV, light_pdf, pdfV), pdfV));          // :597-600
            }
        }
    }
    // Branch "light sample occluded": sample_glossy = false and brdf_prob = 0 keep their initial values (:541-542), so the
    // BRDF sample is cosine-distributed and evaluateBRDF divides by zero (:621).  In the reference that makes brdf_sample =
    // (inf, inf, inf, inf) * light ke (.., .., .., 0): the W LANE is inf*0 = NaN, the kernel's any(isnan(color)) fires (:193)
    // and the WHOLE SAMPLE becomes PINK.  We carry that outcome as a NaN contribution (finalisation turns a NaN sample into
    // PINK), not as the infinities of the xyz lanes.
    __syncwarp(m_lit);
    const bool same_dir = v_alive && !glossy;      // both branches then draw the same cosine direction
    if (same_dir) { dq = dv; pdfO = pdfV; } else dq = sample_cosine_onb(n, Nx, Ny, r1, r2, pdfO);
    if (pdfO > 0.0f) {
        bool reaches = false;
        if (same_dir) { reaches = R.MV.has; R.mo_is_mv = R.MV.has; }
        else {
            R.MO.o = vadd(hp, vscale(dq, YUNE_EPS)); R.MO.d = dq; R.MO.tmax = INFINITY;
            reaches = R.MO.has = (light_loop(lights, n_lights, R.MO.o, R.MO.d, R.MO.tmax) == j);
        }
        if (reaches) { const float qnan = __int_as_float(0x7fc00000); R.BO = v3(qnan, qnan, qnan); }
    }
}

} // namespace yune
#endif
