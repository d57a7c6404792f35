// bdpt.cu -- the bidirectional integrator of kernels/legacy/bdpt.cl as a wavefront "logic + material" kernel.
//
// Reference (bdpt.cl): createLightPath :432-511, createEyePath :513-560, shading :562-640 -- one work-item builds a light
// path (<= BDPT_BOUNCES vertices), an eye path, then connects every eye vertex to every light vertex with a shadow ray
// ("naive BDPT": deterministic connections, weights eye_path_weight * (1 - ks), no MIS between strategies).
//
// Wavefront form.  One thread per path slot, same k_trace as the unidirectional integrator.  A slot walks through
//     LIGHT phase: one light-path vertex per iteration (stored in lp[]: point, normal, direction, contribution)
//     EYE   phase: one eye vertex per iteration; on arrival it emits the NEE rays of evaluateDirectLighting (:593) AND one
//                  connection shadow ray per stored light vertex (:594-635); their answers are folded into the sample at
//                  the NEXT visit, in the reference's order (NEE term, then j = lp_len-1 .. 1), so the float sums match.
// Random numbers are addressed per purpose (rng.h), with the vertex codes of oracle/yune_oracle.cpp:
//     eye vertex i: i | NEE at eye vertex i: 0x20000000+i | light vertex k: 0x40000000+k | connection (i,j): 0x10000000+32*i+j.
// Extension: with several lights the emitter of the light path is drawn proportionally to |ke| * area.
#define YUNE_LEAF_NOINLINE 1       // see strict_math.h (YUNE_HD_LEAF)
#include "kernel_common.cuh"

namespace yune {

#define YB_PHASE_LIGHT 32u      // meta.w flag: the in-flight extension ray belongs to the light path
#define YB_PENDING     64u      // meta.w flag: an eye vertex was shaded at the last visit; its NEE / connection answers are due
#define YB_MAXV 32              // storage stride of per-slot vertex arrays (BDPT_BOUNCES <= 32)

__device__ __forceinline__ float u01_via_double(uint32_t w) { return (float)((double)w / 4294967295.0); }   // bdpt.cl:439

// ------------------------------------------------------------------------------------------------------------
// Persistent, in-block sorted form (same scheme as k_shade_dense in kernels.cu).  A block walks chunks of YUNE_SHADE_BLOCK
// slots; phase A settles the answers of the previous eye vertex and sorts the slot into one of four shared-memory lists;
// a list is processed when it holds a full block of entries, so every lane of every warp runs the same code:
//   L  light-path vertex arrived       V  eye-path vertex arrived (NEE + connections + next direction)
//   R  fresh sample: emitter + first light-path direction      E  light path complete: camera ray
// (The one-thread-per-slot predecessor ran 6.5 of 32 lanes per instruction through 244 KB of code: 6.6 ms per launch.)
// ------------------------------------------------------------------------------------------------------------
enum { BL_LIGHT = 0, BL_EYE = 1, BL_REGEN = 2, BL_CAMERA = 3 };

#define YB_MAX_TASKS (YUNE_SHADE_BLOCK * (YB_MAXV - 1))
struct BdptShared {
    int list_l[2 * YUNE_SHADE_BLOCK], list_v[2 * YUNE_SHADE_BLOCK], list_r[4 * YUNE_SHADE_BLOCK], list_e[4 * YUNE_SHADE_BLOCK];
    int n[4];
    int cnt[3 * (YUNE_NW + 1)];
    unsigned long long sample_base;
    int visits[4];
    // eye round: the (eye vertex, light vertex) connections of the round are flattened into one task list so that every
    // lane evaluates a connection, whatever the light-path lengths of the slots that met in the round
    float4 eye[YUNE_SHADE_BLOCK][4];            // (hit point, matID) (normal, vertex) (w_o, pixel) (throughput, sample)
    int eye_slot[YUNE_SHADE_BLOCK];
    unsigned eye_mask[YUNE_SHADE_BLOCK];        // connection rays launched by the entry (bit j = light vertex j)
    unsigned short task[YB_MAX_TASKS];          // (entry << 5) | j
    int n_tasks, n_conn, conn_base;
    __device__ __forceinline__ int* list(int k) { return k == 0 ? list_l : k == 1 ? list_v : k == 2 ? list_r : list_e; }
};

// ---- phase A: fold in the answers of the previous eye vertex (bdpt.cl:593-636), retire finished samples, classify ----
// returns the list the slot goes to, or -1
__device__ __forceinline__ int bdpt_classify(const RenderArgs& A, const BdptPool& B, const int s)
{
    const PathPool& P = A.pool;
    if (s >= P.n_slots) return -1;
    // Everything whose address depends on s alone is requested at once (k_shade_dense's classify phase, DESIGN.md section 5: the
    // flags -> state -> answers chain was three dependent DRAM round trips for all eight warps of the block).
    const float4* pend_c = B.pend_c + (size_t)s * YB_MAXV;
    const uint4 meta = P.meta[s];
    const float hit_w = P.hit[s].w;
    const float4 col4 = P.col[s], thr4 = P.thr[s], pend_l4 = P.pend_l[s], pend_c0 = pend_c[0];
    int4 bm = B.bmeta[s];
    const unsigned char vis_l0 = P.vis_l[s];
    const unsigned state = meta.w & YS_STATE_MASK;
    if (state == YS_FREE) return BL_REGEN;
    if (state == YS_DONE) return -1;
    const bool light_phase = (meta.w & YB_PHASE_LIGHT) != 0;
    const int tri = state == YS_TRACE ? __float_as_int(hit_w) : -1;
    if (light_phase) return tri < 0 ? BL_CAMERA : BL_LIGHT;           // missed or hit a light: the path stops before this vertex (:458)

    V3 col = xyz(col4);
    bool dirty = false;
    float epw = __int_as_float(bm.z);                                   // eye_path_weight BEFORE the vertex whose answers are pending
    if (meta.w & YB_PENDING) {
        const V3 T = xyz(thr4);
        V3 nee = v3(0, 0, 0);
        if (meta.w & YF_PEND_EVT) {
            const int e = P.evt_idx[s];
            const float4 e0 = P.evt[3 * (size_t)e], e1 = P.evt[3 * (size_t)e + 1], e2 = P.evt[3 * (size_t)e + 2];
            const int ef = __float_as_int(e0.w);
            const bool visS = (ef & YE_HAS_S) && P.evt_vis[4 * (size_t)e + 0];
            const bool visMV = (ef & YE_HAS_MV) && P.evt_vis[4 * (size_t)e + 1];
            const bool visMO = (ef & YE_MO_IS_MV) ? visMV : ((ef & YE_HAS_MO) && P.evt_vis[4 * (size_t)e + 2]);
            if (visS) nee = vadd(xyz(e0), visMV ? xyz(e1) : v3(0, 0, 0));
            else      nee = ((ef & (YE_HAS_MO | YE_MO_IS_MV)) && visMO) ? xyz(e2) : v3(0, 0, 0);
        } else if (meta.w & YF_PEND_L) {
            if (vis_l0) nee = xyz(pend_l4);
        }
        const V3 emission = xyz(pend_c0);
        // color += throughput * (emission + NEE) * eye_path_weight          (:593; NEE's own return value already adds emission once)
        col = vadd(col, vscale(vmul(T, vadd(emission, vadd(nee, emission))), epw));
        V3 sub = v3(0, 0, 0);
        unsigned mask = (unsigned)bm.y;                                 // connection rays in flight, folded j = lp_len-1 .. 1 (:594)
        while (mask) {
            const int j = 31 - __clz(mask);
            mask &= ~(1u << j);
            if (P.vis_l[(size_t)P.n_slots + (size_t)s * YB_MAXV + j]) sub = vadd(sub, xyz(pend_c[j]));
        }
        const float ks = __int_as_float(bm.w);
        col = vadd(col, vscale(sub, YF_MUL(epw, YF_SUB(1.0f, ks))));    // :636
        epw = YF_MUL(epw, ks);                                          // :637
        dirty = true;
    }
    bool finished = state == YS_DRAIN;
    if (!finished) {
        if (tri < 0) {
            if (meta.z == 0) {
                const float4 rd = P.ray_d[s];
                const int lid = __float_as_int(rd.w);
                if (lid >= 0) col = (vdot(xyz(rd), A.lights.l[lid].normal) < 0.0f) ? v3(1.0f, 1.0f, 1.0f) : v3(0.1f, 0.1f, 0.1f);   // :575-581
                else col = v3(0.4f, 0.4f, 0.4f);                                                                                  // :582-583
            }
            finished = true;
        } else if (meta.z != 0 && epw == 0.0f) finished = true;         // :593 'if(eye_path_weight == 0) break'
    }
    if (finished) { finish_sample(A, meta.x, col); return BL_REGEN; }   // bdpt.cl:192-209
    if (dirty) { P.col[s] = f4(col, 0.0f); bm.z = __float_as_int(epw); bm.y = 0; B.bmeta[s] = bm; }
    return BL_EYE;
}

// ---- L: arrival at light vertex k = vtx (bdpt.cl:458-510) ----
__device__ __forceinline__ void bdpt_light_round(const RenderArgs& A, const BdptPool& B, const int s, BdptShared& sh, int& live)
{
    const PathPool& P = A.pool;
    IterCounters* C = A.ctr + A.parity;
    bool has_ext = false, to_camera = false;
    V3 ext_o = v3(0, 0, 0), ext_d = v3(0, 0, 1), Tn = v3(1, 1, 1); float ext_t = INFINITY; int ext_lid = -1;
    uint4 meta = make_uint4(0, 0, 0, 0);
    const unsigned act = __ballot_sync(0xffffffffu, s >= 0);
    if (s >= 0) {
        meta = P.meta[s];
        const float4 hit = P.hit[s], ro = P.ray_o[s], rd = P.ray_d[s];
        const int tri = __float_as_int(hit.w);
        const V3 o = xyz(ro), d = xyz(rd);
        const unsigned vtx = meta.z;
        float4* lp = B.lp + (size_t)s * YB_MAXV * 4;
        const float4 s0 = __ldg(A.sc.shade + 4 * (size_t)tri), s1 = __ldg(A.sc.shade + 4 * (size_t)tri + 1), s2 = __ldg(A.sc.shade + 4 * (size_t)tri + 2);
        const MatDev mat = load_material(A.sc.mats, __float_as_int(s0.w));
        const float bw = YF_SUB(YF_SUB(1.0f, hit.y), hit.z);
        const V3 hp = vadd(o, vscale(d, hit.x));
        const V3 n = vnormalize(vmadd3(xyz(s0), bw, xyz(s1), hit.y, xyz(s2), hit.z));
        V3 contrib = xyz(P.thr_next[s]);                                // formed when the ray was sampled
        bool stop = false;
        if ((int)vtx > A.rr_threshold && vtx >= 2) {                    // :498-507 (the loop starts at i = 2)
            const U4 u = draw4(A.seed, meta.x, meta.y, 0x40000000u + vtx, YUNE_BLK_NEE);
            const float r = u01(u.x);
            const float p = cl_min(luminance(contrib), 0.95f);
            if (r >= p) stop = true; else contrib = vscale(contrib, YF_DIV(1.0f, p));
        }
        __syncwarp(act);
        lp[4 * vtx + 0] = f4(hp, __int_as_float(tri));
        lp[4 * vtx + 1] = f4(n, 0.0f);
        lp[4 * vtx + 2] = f4(d, 0.0f);
        lp[4 * vtx + 3] = f4(contrib, 0.0f);
        bool end_path = stop || (int)vtx + 1 >= B.bounces;
        if (!end_path) {
            // sample the next light-path direction at this vertex (:476-483)
            const U4 u = draw4(A.seed, meta.x, meta.y, 0x40000000u + vtx, YUNE_BLK_BOUNCE);
            float prob = 0.0f, pdf = 1.0f;
            const bool glossy = select_lobe(mat, u01(u.x), true, prob);
            const V3 dir = glossy ? sample_phong(vneg(d), n, mat.px, mat.py, u01(u.y), u01(u.z), false, pdf) : sample_cosine(n, u01(u.y), u01(u.z), pdf);
            if (pdf <= 0.0f) end_path = true;                           // :485
            else {
                // contrib_{k+1} = evaluateBRDF(-dir_k, dir_{k+1}, hit_k) * max(0, dot(dir_{k+1}, n_k)) / pdf * contrib_k   (:495-499)
                V3 c = vscale(eval_brdf(mat, vneg(d), dir, n, glossy, prob, false, A.oren_nayar != 0), fmaxf(0.0f, vdot(dir, n)));
                c = vdivs(c, pdf);
                Tn = vmul(c, contrib);
                has_ext = true; ext_d = dir; ext_o = vadd(hp, vscale(dir, YUNE_EPS)); ext_t = INFINITY;
                ext_lid = light_loop(A.lights.l, A.lights.n, ext_o, ext_d, ext_t);
            }
        }
        __syncwarp(act);
        int4 bm = B.bmeta[s];
        bm.x = (int)vtx + 1;                                            // light-path length
        B.bmeta[s] = bm;
        to_camera = !has_ext;
    }
    list_push(sh.list_e, &sh.n[BL_CAMERA], to_camera, s);
    int* const counters[1] = { &C->n_extend };
    const int cnt[1] = { has_ext ? 1 : 0 };
    int first[1];
    block_alloc<1, YUNE_NW>(counters, cnt, first, sh.cnt);
    if (has_ext) {
        P.eq[first[0]] = s;
        P.ray_o[s] = f4(ext_o, ext_t);
        P.ray_d[s] = f4(ext_d, __int_as_float(ext_lid));
        P.thr_next[s] = f4(Tn, 0.0f);
        meta.z += 1; meta.w = YS_TRACE | YB_PHASE_LIGHT;
        P.meta[s] = meta;
        live++;
    }
}

// ---- V: arrival at eye vertex i = vtx (bdpt.cl:513-560) and its shading (:562-640): NEE, one connection ray per light
//      vertex, next direction.  Connections are found in a first pass (geometry + the analytic light test), allocated with
//      one block-wide scan, and evaluated in a second pass. ----
template <bool MIS>
__device__ __forceinline__ void bdpt_eye_round(const RenderArgs& A, const BdptPool& B, const int s, BdptShared& sh, int& live)
{
    const PathPool& P = A.pool;
    IterCounters* C = A.ctr + A.parity;
    const LightDev* lights = A.lights.l;
    const int n_lights = A.lights.n;
    const bool use_on = A.oren_nayar != 0;
    bool has_ext = false;
    V3 ext_o = v3(0, 0, 0), ext_d = v3(0, 0, 1), T = v3(1, 1, 1), Tn = v3(1, 1, 1); float ext_t = INFINITY; int ext_lid = -1;
    NeeOut N; N.S.has = N.MV.has = N.MO.has = false; N.mo_is_mv = false; N.Lv = N.BV = N.BO = v3(0, 0, 0);
    uint4 meta = make_uint4(0, 0, 0, 0);
    int4 bm = make_int4(1, 0, 0, 0);
    V3 e_hp = v3(0, 0, 0), e_n = v3(0, 0, 1), e_wo = v3(0, 0, 1); MatDev e_mat; unsigned e_vtx = 0; bool e_last = false;
    e_mat.ke = e_mat.kd = e_mat.ks = v3(0, 0, 0); e_mat.n = e_mat.px = e_mat.py = e_mat.alpha_x = 1.0f; e_mat.is_specular = e_mat.is_transmissive = 0;
    float epw = 1.0f; int e_mat_id = 0;
    unsigned conn_mask = 0;
    const unsigned act = __ballot_sync(0xffffffffu, s >= 0);
    float4* lp = B.lp + (size_t)(s >= 0 ? s : 0) * YB_MAXV * 4;
    float4* pend_c = B.pend_c + (size_t)(s >= 0 ? s : 0) * YB_MAXV;

    if (s >= 0) {
        meta = P.meta[s];
        bm = B.bmeta[s];
        const float4 hit = P.hit[s], ro = P.ray_o[s], rd = P.ray_d[s];
        const int tri = __float_as_int(hit.w);
        const V3 o = xyz(ro), d = xyz(rd);
        e_vtx = meta.z;
        const float4 s0 = __ldg(A.sc.shade + 4 * (size_t)tri), s1 = __ldg(A.sc.shade + 4 * (size_t)tri + 1), s2 = __ldg(A.sc.shade + 4 * (size_t)tri + 2);
        e_mat_id = __float_as_int(s0.w);
        e_mat = load_material(A.sc.mats, e_mat_id);
        const float bw = YF_SUB(YF_SUB(1.0f, hit.y), hit.z);
        e_hp = vadd(o, vscale(d, hit.x));
        e_n = vnormalize(vmadd3(xyz(s0), bw, xyz(s1), hit.y, xyz(s2), hit.z));
        e_wo = vneg(d);
        epw = __int_as_float(bm.z);
        if (e_vtx == 0) { T = v3(1, 1, 1); epw = 1.0f; }
        else {
            T = xyz(P.thr_next[s]);                                     // eye_path[i].contrib (:546-549)
            if ((int)e_vtx > A.rr_threshold) {                          // :550-558
                const U4 u = draw4(A.seed, meta.x, meta.y, e_vtx, YUNE_BLK_NEE);
                const float r = u01(u.x);
                const float p = cl_min(luminance(T), 0.95f);
                if (r >= p) e_last = true; else T = vscale(T, YF_DIV(1.0f, p));
            }
        }
        if ((int)e_vtx + 1 >= B.bounces) e_last = true;
        __syncwarp(act);
        const float ks = cl_max(cl_max(e_mat.ks.x, cl_max(e_mat.ks.y, e_mat.ks.z)), 0.1f);           // :599
        pend_c[0] = f4(e_mat.ke, 0.0f);
        bm.w = __float_as_int(ks);
        const U4 u_nee = draw4(A.seed, meta.x, meta.y, 0x20000000u + e_vtx, YUNE_BLK_NEE);
        const LobePrep e_lobes = lobe_prepare(e_mat, true);
        V3 e_nx, e_ny; onb(e_n, e_nx, e_ny);
        nee_sample<MIS, true>(act, lights, n_lights, e_mat, e_lobes, e_hp, e_n, e_nx, e_ny, e_wo, u_nee, A.seed, meta.x, meta.y, 0x20000000u + e_vtx, use_on, N);
        __syncwarp(act);
    }
    // ---- continue the eye path (:525-537) ----
    if (s >= 0 && !e_last) {
        const U4 u = draw4(A.seed, meta.x, meta.y, e_vtx, YUNE_BLK_BOUNCE);
        float prob = 0.0f, pdf = 1.0f;
        const bool glossy = select_lobe(e_mat, u01(u.x), true, prob);
        const V3 dir = glossy ? sample_phong(e_wo, e_n, e_mat.px, e_mat.py, u01(u.y), u01(u.z), false, pdf) : sample_cosine(e_n, u01(u.y), u01(u.z), pdf);
        if (pdf > 0.0f) {
            V3 c = vscale(eval_brdf(e_mat, dir, e_wo, e_n, glossy, prob, false, use_on), cl_max(0.0f, vdot(dir, e_n)));
            c = vdivs(c, pdf);
            Tn = vmul(c, T);
            has_ext = true; ext_d = dir; ext_o = vadd(e_hp, vscale(dir, YUNE_EPS)); ext_t = INFINITY;
            ext_lid = light_loop(lights, n_lights, ext_o, ext_d, ext_t);
        }
    }
    __syncwarp();
    // ---- queue space: one block-wide scan ----
    const bool is_event = N.MV.has || N.MO.has;
    const int n_nee = (N.S.has ? 1 : 0) + (N.MV.has ? 1 : 0) + (N.MO.has ? 1 : 0);
    int* const counters[3] = { &C->n_extend, &C->n_shadow, &C->n_events };
    const int cnt[3] = { has_ext ? 1 : 0, n_nee, is_event ? 1 : 0 };
    int first[3];
    block_alloc<3, YUNE_NW>(counters, cnt, first, sh.cnt);
    unsigned new_flags = (has_ext ? YS_TRACE : YS_DRAIN) | YB_PENDING;  // without a next ray: wait for this vertex's answers, then finish
    int qs = first[1];
    const int ev = A.parity * P.n_slots + first[2];
    if (is_event) {
        const int ef = (N.S.has ? YE_HAS_S : 0) | (N.MV.has ? YE_HAS_MV : 0) | (N.MO.has ? YE_HAS_MO : 0) | (N.mo_is_mv ? YE_MO_IS_MV : 0);
        P.evt[3 * (size_t)ev] = f4(N.Lv, __int_as_float(ef));
        P.evt[3 * (size_t)ev + 1] = f4(N.BV, 0.0f);
        P.evt[3 * (size_t)ev + 2] = f4(N.BO, 0.0f);
        P.evt_idx[s] = ev;
        new_flags |= YF_PEND_EVT;
    } else if (N.S.has) new_flags |= YF_PEND_L;
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
    // ---- connections (:594-635), flattened over the round.  Tasks = (entry, j) for every stored light vertex j >= 1. ----
    const int tid = threadIdx.x;
    const int lpn = s >= 0 ? bm.x : 0;
    sh.eye[tid][0] = f4(e_hp, __int_as_float(e_mat_id));
    sh.eye[tid][1] = f4(e_n, __int_as_float((int)e_vtx));
    sh.eye[tid][2] = f4(e_wo, __int_as_float((int)meta.x));
    sh.eye[tid][3] = f4(T, __int_as_float((int)meta.y));
    sh.eye_slot[tid] = s;
    sh.eye_mask[tid] = 0u;
    if (tid == 0) { sh.n_tasks = 0; sh.n_conn = 0; }
    __syncthreads();
    {
        const int lane = tid & 31;
        const int mine = lpn > 1 ? lpn - 1 : 0;
        int incl = mine;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        int base = 0;
        if (lane == 31 && incl > 0) base = atomicAdd(&sh.n_tasks, incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
        for (int j = 1; j <= mine; j++) sh.task[base + j - 1] = (unsigned short)((tid << 5) | j);
        __syncwarp();
    }
    __syncthreads();
    // pass 1: which connections exist geometrically (:604-613)?  Survivors are compacted to the front of the task list.
    const int n_tasks = sh.n_tasks;
    for (int t0 = 0; t0 < n_tasks; t0 += YUNE_SHADE_BLOCK) {
        const int t = t0 + tid;
        bool want = false; unsigned short tk = 0;
        if (t < n_tasks) {
            tk = sh.task[t];
            const int e = tk >> 5, j = tk & 31;
            const float4* lpe = B.lp + (size_t)sh.eye_slot[e] * YB_MAXV * 4;
            const float4 l0 = lpe[4 * j + 0], l1 = lpe[4 * j + 1];
            const V3 hp = xyz(sh.eye[e][0]), en = xyz(sh.eye[e][1]);
            const V3 lpnt = xyz(l0), ln = xyz(l1);
            const V3 cd = vnormalize(vsub(lpnt, hp));
            want = !(vdot(cd, en) <= 0.0f || vdot(vneg(cd), ln) <= 0.0f);                           // :612-613
            if (want) {
                const V3 co = vadd(hp, vscale(cd, YUNE_EPS));
                float tl = vlength(vsub(lpnt, co));
                if (light_loop(lights, n_lights, co, cd, tl) >= 0) want = false;                     // a light inside the segment occludes (traceRay, :244-276)
            }
            if (want) atomicOr(&sh.eye_mask[e], 1u << j);
        }
        __syncthreads();                                                                             // batch read before its survivors are written
        list_push_u16(sh.task, &sh.n_conn, want, tk);
        __syncthreads();
    }
    const int n_conn = sh.n_conn;
    if (tid == 0) sh.conn_base = n_conn > 0 ? atomicAdd(&C->n_shadow, n_conn) : 0;
    __syncthreads();
    // pass 2: throughput of every connection found (:614-635) and its shadow ray
    for (int t = tid; t < n_conn; t += YUNE_SHADE_BLOCK) {
        const unsigned short tk = sh.task[t];
        const int e = tk >> 5, j = tk & 31;
        const int es = sh.eye_slot[e];
        const float4 r0 = sh.eye[e][0], r1 = sh.eye[e][1], r2 = sh.eye[e][2], r3 = sh.eye[e][3];
        const V3 hp = xyz(r0), en = xyz(r1), wo = xyz(r2), Te = xyz(r3);
        const unsigned pvtx = (unsigned)__float_as_int(r1.w), pix = (unsigned)__float_as_int(r2.w), smp = (unsigned)__float_as_int(r3.w);
        const MatDev emat = load_material(A.sc.mats, __float_as_int(r0.w));
        const float4* lpe = B.lp + (size_t)es * YB_MAXV * 4;
        const float4 l0 = lpe[4 * j + 0], l1 = lpe[4 * j + 1], l2 = lpe[4 * j + 2], l3 = lpe[4 * j + 3];
        const V3 lpnt = xyz(l0), ln = xyz(l1);
        const V3 delta = vsub(lpnt, hp);
        const V3 cd = vnormalize(delta);
        const V3 co = vadd(hp, vscale(cd, YUNE_EPS));
        float dist = vlength(delta);
        dist = YF_MUL(dist, dist);
        const float clen = vlength(vsub(lpnt, co));
        const float gf = YF_DIV(YF_MUL(cl_max(vdot(cd, en), 0.0f), cl_max(vdot(vneg(cd), ln), 0.0f)), dist);
        const U4 uc = draw4(A.seed, pix, smp, 0x10000000u + 32u * pvtx + (unsigned)j, YUNE_BLK_BOUNCE);
        const U4 uc2 = draw4(A.seed, pix, smp, 0x10000000u + 32u * pvtx + (unsigned)j, YUNE_BLK_NEE);
        float prob = 0.0f;
        bool g = select_lobe(emat, u01(uc.x), true, prob);
        const V3 e2l = eval_brdf(emat, cd, wo, en, g, prob, false, use_on);
        const MatDev lmat = load_material(A.sc.mats, __float_as_int(__ldg(A.sc.shade + 4 * (size_t)__float_as_int(l0.w)).w));
        g = select_lobe(lmat, u01(uc2.y), true, prob);
        const V3 l2e = eval_brdf(lmat, vneg(xyz(l2)), vneg(cd), ln, g, prob, false, use_on);
        V3 tl3 = vmul(xyz(l3), vmul(vscale(e2l, gf), l2e));                                           // throughput_lp *= gf * e2l * l2e   (:631)
        tl3 = vmul(tl3, Te);                                                                          // *= throughput                       (:632)
        B.pend_c[(size_t)es * YB_MAXV + j] = f4(tl3, 0.0f);
        const int q = sh.conn_base + t;
        P.sq_o[q] = f4(co, clen);
        P.sq_d[q] = f4(cd, __int_as_float(P.n_slots + es * YB_MAXV + j));
    }
    __syncthreads();
    conn_mask = sh.eye_mask[tid];
    if (s >= 0) {
        if (has_ext) {
            P.eq[first[0]] = s;
            P.ray_o[s] = f4(ext_o, ext_t);
            P.ray_d[s] = f4(ext_d, __int_as_float(ext_lid));
            P.thr_next[s] = f4(Tn, 0.0f);
            meta.z = e_vtx + 1;
        }
        meta.w = new_flags;
        P.meta[s] = meta;
        bm.y = (int)conn_mask; bm.z = __float_as_int(epw);
        B.bmeta[s] = bm;
        P.thr[s] = f4(T, 0.0f);
        live++;
    }
}

// ---- R: next (pixel, sample); the light path comes first (createLightPath, bdpt.cl:432-457) ----
__device__ __forceinline__ void bdpt_regen_round(const RenderArgs& A, const BdptPool& B, const int s, const int take, BdptShared& sh, int& live)
{
    const PathPool& P = A.pool;
    IterCounters* C = A.ctr + A.parity;
    const LightDev* lights = A.lights.l;
    const int n_lights = A.lights.n;
    if (threadIdx.x == 0) sh.sample_base = atomicAdd(&A.tot->next_sample, (unsigned long long)take);
    __syncthreads();
    const unsigned long long g = sh.sample_base + threadIdx.x;
    const bool fresh = s >= 0 && g < A.tot->n_samples;
    bool has_ext = false, to_camera = false;
    V3 ext_o = v3(0, 0, 0), ext_d = v3(0, 0, 1), Tn = v3(1, 1, 1); float ext_t = INFINITY; int ext_lid = -1;
    uint4 meta = make_uint4(0, 0, 0, 0);
    if (fresh) {
        const unsigned long long n_pix = (unsigned long long)A.width * A.height;
        meta.x = (unsigned)(g % n_pix);
        meta.y = (unsigned)(A.spp_begin + (int)(g / n_pix));
        float4* lp = B.lp + (size_t)s * YB_MAXV * 4;
        // emitter: light 0 in the reference; drawn proportionally to |ke| * area when there are several
        int li = 0; float pick_prob = 1.0f;
        const U4 ul = draw4(A.seed, meta.x, meta.y, 0x40000000u, YUNE_BLK_LIGHT);
        if (n_lights > 1) {
            float w[YUNE_MAX_LIGHTS], sum = 0.0f;
            for (int i = 0; i < n_lights; i++) { w[i] = YF_MUL(vlength(lights[i].ke), YF_MUL(lights[i].la, lights[i].lb)); sum = YF_ADD(sum, w[i]); }
            const float r = u01(ul.z);
            float cum = 0.0f; li = n_lights - 1;
            for (int i = 0; i < n_lights; i++) { const float p = YF_DIV(w[i], sum); if (r >= cum && r < YF_ADD(cum, p)) { li = i; break; } cum = YF_ADD(cum, p); }
            pick_prob = YF_DIV(w[li], sum);
        }
        const LightDev& L = lights[li];
        const float r1 = u01_via_double(ul.x), r2 = u01(ul.y);
        const V3 lpnt = vadd(vadd(vscale(L.edge_l, r2), vscale(L.edge_w, r1)), L.pos);               // :444-446
        const float area = YF_MUL(L.la, L.lb);
        const float fwd_pdf = YF_DIV(1.0f, area);
        V3 c0 = vdivs(L.ke, fwd_pdf);                                                                  // :454
        if (n_lights > 1) c0 = vdivs(c0, pick_prob);
        lp[0] = f4(lpnt, __int_as_float(-1)); lp[1] = f4(L.normal, 0.0f); lp[2] = f4(v3(0, 0, 0), 0.0f); lp[3] = f4(c0, 0.0f);
        B.bmeta[s] = make_int4(1, 0, __float_as_int(1.0f), 0);
        P.col[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        P.thr[s] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        const U4 ub = draw4(A.seed, meta.x, meta.y, 0x40000000u, YUNE_BLK_BOUNCE);
        float pdf = 1.0f;
        const V3 dir = sample_cosine(L.normal, u01(ub.y), u01(ub.z), pdf);                             // :456
        if (pdf > 0.0f) {
            const float c1 = YF_DIV(cl_max(0.0f, vdot(L.normal, dir)), pdf);                           // :466
            Tn = vmul(v3(c1, c1, c1), c0);
            has_ext = true; ext_d = dir; ext_o = vadd(lpnt, vscale(dir, YUNE_EPS)); ext_t = INFINITY;
            ext_lid = light_loop(lights, n_lights, ext_o, ext_d, ext_t);
        } else to_camera = true;
    }
    __syncwarp();
    list_push(sh.list_e, &sh.n[BL_CAMERA], to_camera, s);
    int* const counters[1] = { &C->n_extend };
    const int cnt[1] = { has_ext ? 1 : 0 };
    int first[1];
    block_alloc<1, YUNE_NW>(counters, cnt, first, sh.cnt);
    if (has_ext) {
        P.eq[first[0]] = s;
        P.ray_o[s] = f4(ext_o, ext_t);
        P.ray_d[s] = f4(ext_d, __int_as_float(ext_lid));
        P.thr_next[s] = f4(Tn, 0.0f);
        meta.z = 1; meta.w = YS_TRACE | YB_PHASE_LIGHT;
        P.meta[s] = meta;
        live++;
    } else if (to_camera) {
        meta.z = 0; meta.w = YS_FREE;                                   // the camera round of this launch fills in the rest
        P.meta[s] = meta;
    } else if (s >= 0) P.meta[s] = make_uint4(0u, 0u, 0u, YS_DONE);
}

// ---- E: light path complete: camera ray (bdpt.cl:172-189) ----
__device__ __forceinline__ void bdpt_camera_round(const RenderArgs& A, const BdptPool& B, const int s, const int take, BdptShared& sh, int& live)
{
    const PathPool& P = A.pool;
    IterCounters* C = A.ctr + A.parity;
    __shared__ int s_ext_base;
    if (threadIdx.x == 0) s_ext_base = atomicAdd(&C->n_extend, take);   // every entry of the round emits its camera ray
    __syncthreads();
    if (s >= 0) {
        uint4 meta = P.meta[s];
        const int px = meta.x % A.width, py = meta.x / A.width;
        const U4 uj = draw4(A.seed, meta.x, meta.y, YUNE_VERTEX_CAMERA, 0u);
        V3 ro, rd;
        create_ray(A.cam, A.width, A.height, (float)px + u01(uj.x), (float)py + u01(uj.y), ro, rd);
        float rt = INFINITY;
        const int rl = light_loop(A.lights.l, A.lights.n, ro, rd, rt);
        P.eq[s_ext_base + threadIdx.x] = s;
        P.ray_o[s] = f4(ro, rt);
        P.ray_d[s] = f4(rd, __int_as_float(rl));
        P.thr[s] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        P.thr_next[s] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        int4 bm = B.bmeta[s];
        bm.y = 0; bm.z = __float_as_int(1.0f);
        B.bmeta[s] = bm;
        meta.z = 0; meta.w = YS_TRACE;
        P.meta[s] = meta;
        live++;
    }
    __syncthreads();
}

template <bool MIS>
__global__ void __launch_bounds__(YUNE_SHADE_BLOCK, 2) k_shade_bdpt(RenderArgs A, BdptPool B)
{
    __shared__ BdptShared sh;
    const int tid = threadIdx.x;
    if (tid < 4) { sh.n[tid] = 0; sh.visits[tid] = 0; }
    __syncthreads();
    const int n_chunks = (A.pool.n_slots + YUNE_SHADE_BLOCK - 1) / YUNE_SHADE_BLOCK;
    int live = 0;
    for (int chunk = blockIdx.x; ; chunk += gridDim.x) {
        const bool flush = chunk >= n_chunks;       // block-uniform: the pass after the last chunk empties the lists
        if (!flush) {
            const int s = chunk * YUNE_SHADE_BLOCK + tid;
            const int cls = bdpt_classify(A, B, s);
            __syncwarp();
            #pragma unroll
            for (int k = 0; k < 4; k++) list_push(sh.list(k), &sh.n[k], cls == k, s);
        }
        __syncthreads();
        // order: L feeds E, V feeds nothing, R feeds E; E last
        YUNE_NO_UNROLL
        for (int k = 0; k < 4; k++) {
            for (;;) {
                const int n = sh.n[k];
                if (!(n >= YUNE_SHADE_BLOCK || (flush && n > 0))) break;
                const int take = n < YUNE_SHADE_BLOCK ? n : YUNE_SHADE_BLOCK;
                const int s = tid < take ? sh.list(k)[n - take + tid] : -1;
                __syncthreads();
                if (tid == 0) { sh.n[k] = n - take; sh.visits[k] += take; }
                if (k == BL_LIGHT) bdpt_light_round(A, B, s, sh, live);
                else if (k == BL_EYE) bdpt_eye_round<MIS>(A, B, s, sh, live);
                else if (k == BL_REGEN) bdpt_regen_round(A, B, s, take, sh, live);
                else bdpt_camera_round(A, B, s, take, sh, live);
                __syncthreads();
            }
        }
        if (flush) break;
        __syncthreads();                            // every warp has read the list counts before the next chunk's classify bumps them
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
    if ((tid & 31) == 0) sh.cnt[tid >> 5] = live;
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int w = 0; w < YUNE_NW; w++) total += sh.cnt[w];
        if (total > 0) atomicAdd(&A.ctr[A.parity].live, total);
        if (sh.visits[BL_EYE]) atomicAdd(&A.tot->visits_d, (unsigned long long)sh.visits[BL_EYE]);
        if (sh.visits[BL_LIGHT]) atomicAdd(&A.tot->visits_s, (unsigned long long)sh.visits[BL_LIGHT]);
        if (sh.visits[BL_REGEN]) atomicAdd(&A.tot->visits_r, (unsigned long long)sh.visits[BL_REGEN]);
    }
}

cudaError_t launch_shade_bdpt(const RenderArgs& a, const BdptPool& b, int sm_count, int* occ, cudaStream_t st)
{      // occ[2]: resident blocks per SM of the two instantiations, cached by the caller per context (0 = not asked yet)
    const int v = a.mis ? 1 : 0;
    if (occ[v] == 0) {
        cudaError_t e = v ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[v], k_shade_bdpt<true>, YUNE_SHADE_BLOCK, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[v], k_shade_bdpt<false>, YUNE_SHADE_BLOCK, 0);
        if (e != cudaSuccess) return e;
        if (occ[v] < 1) occ[v] = 1;
    }
    int grid = sm_count * occ[v];
    const int n_chunks = (a.pool.n_slots + YUNE_SHADE_BLOCK - 1) / YUNE_SHADE_BLOCK;
    if (grid > n_chunks) grid = n_chunks;
    if (grid < 1) grid = 1;
    if (a.mis) k_shade_bdpt<true><<<grid, YUNE_SHADE_BLOCK, 0, st>>>(a, b);
    else       k_shade_bdpt<false><<<grid, YUNE_SHADE_BLOCK, 0, st>>>(a, b);
    return cudaGetLastError();
}

} // namespace yune
