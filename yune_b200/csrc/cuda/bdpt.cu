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
#include "kernel_common.cuh"

namespace yune {

#define YB_PHASE_LIGHT 32u      // meta.w flag: the in-flight extension ray belongs to the light path
#define YB_PENDING     64u      // meta.w flag: an eye vertex was shaded at the last visit; its NEE / connection answers are due
#define YB_MAXV 32              // storage stride of per-slot vertex arrays (BDPT_BOUNCES <= 32)

__device__ __forceinline__ float u01_via_double(uint32_t w) { return (float)((double)w / 4294967295.0); }   // bdpt.cl:439

template <bool MIS>
__global__ void __launch_bounds__(YUNE_SHADE_BLOCK) k_shade_bdpt(RenderArgs A, BdptPool B)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const PathPool& P = A.pool;
    const bool valid = s < P.n_slots;
    IterCounters* C = A.ctr + A.parity;
    const LightDev* lights = A.lights.l;
    const int n_lights = A.lights.n;
    const int max_v = B.bounces;                      // BDPT_BOUNCES (bdpt.cl:7)
    const bool use_on = A.oren_nayar != 0;

    uint4 meta = valid ? P.meta[s] : make_uint4(0, 0, 0, YS_DONE);
    unsigned state = meta.w & YS_STATE_MASK;
    const bool light_phase = (meta.w & YB_PHASE_LIGHT) != 0;
    int4 bm = valid ? B.bmeta[s] : make_int4(1, 0, 0, 0);     // x = lp_len, y = mask of connection rays in flight, z = bits(eye_path_weight), w = bits(ks)
    float4* lp = B.lp + (size_t)s * YB_MAXV * 4;
    float4* pend_c = B.pend_c + (size_t)s * YB_MAXV;

    V3 col = v3(0, 0, 0), T = v3(1, 1, 1), Tn = v3(1, 1, 1);
    unsigned new_flags = 0;
    bool has_ext = false; V3 ext_o = v3(0, 0, 0), ext_d = v3(0, 0, 1); float ext_t = INFINITY; int ext_lid = -1;
    NeeOut N; N.S.has = N.MV.has = N.MO.has = false; N.mo_is_mv = false; N.Lv = N.BV = N.BO = v3(0, 0, 0);
    unsigned conn_mask = 0;
    bool finished = false, start_eye = false, need_new = false, shade_eye = false, e_last = false;
    V3 e_hp = v3(0, 0, 0), e_n = v3(0, 0, 1), e_wo = v3(0, 0, 1); MatDev e_mat; unsigned e_vtx = 0;
    e_mat.ke = e_mat.kd = e_mat.ks = v3(0, 0, 0); e_mat.n = e_mat.px = e_mat.py = e_mat.alpha_x = 1.0f; e_mat.is_specular = e_mat.is_transmissive = 0;
    float epw = __int_as_float(bm.z);                 // eye_path_weight BEFORE the vertex whose answers are pending

    if (state == YS_TRACE || state == YS_DRAIN) {
        col = xyz(P.col[s]);
        T = xyz(P.thr[s]);
    }
    // ---- 1. fold in the answers of the previous eye vertex (bdpt.cl:593-636) ----
    if ((state == YS_TRACE || state == YS_DRAIN) && (meta.w & YB_PENDING)) {
        {
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
                if (P.vis_l[s]) nee = xyz(P.pend_l[s]);
            }
            const V3 emission = xyz(pend_c[0]);
            // color += throughput * (emission + NEE) * eye_path_weight          (:593; NEE's own return value already adds emission once)
            col = vadd(col, vscale(vmul(T, vadd(emission, vadd(nee, emission))), epw));
            V3 sub = v3(0, 0, 0);
            for (int j = bm.x - 1; j > 0; j--)
                if ((bm.y >> j) & 1) { if (P.vis_l[(size_t)P.n_slots + (size_t)s * YB_MAXV + j]) sub = vadd(sub, xyz(pend_c[j])); }
            const float ks = __int_as_float(bm.w);
            col = vadd(col, vscale(sub, YF_MUL(epw, YF_SUB(1.0f, ks))));                                  // :636
            epw = YF_MUL(epw, ks);                                                                          // :637
        }
    }

    if (state == YS_DRAIN) finished = true;
    else if (state == YS_TRACE) {
        const float4 hit = P.hit[s], ro = P.ray_o[s], rd = P.ray_d[s];
        const int tri = __float_as_int(hit.w);
        const V3 o = xyz(ro), d = xyz(rd);
        const unsigned vtx = meta.z;
        if (light_phase) {
            // ---------------- LIGHT PATH: arrival at light vertex k = vtx (bdpt.cl:458-510) ----------------
            bool end_path = tri < 0;                                            // missed, or hit a light: path stops before this vertex
            int lp_len = bm.x;
            if (!end_path) {
                const float4 s0 = __ldg(A.sc.shade + 4 * (size_t)tri), s1 = __ldg(A.sc.shade + 4 * (size_t)tri + 1), s2 = __ldg(A.sc.shade + 4 * (size_t)tri + 2);
                const MatDev mat = load_material(A.sc.mats, __float_as_int(s0.w));
                const float bw = YF_SUB(YF_SUB(1.0f, hit.y), hit.z);
                const V3 hp = vadd(o, vscale(d, hit.x));
                const V3 n = vnormalize(vmadd3(xyz(s0), bw, xyz(s1), hit.y, xyz(s2), hit.z));
                V3 contrib = xyz(P.thr_next[s]);                                // formed when the ray was sampled
                lp_len = (int)vtx + 1;
                bool stop = false;
                if ((int)vtx > A.rr_threshold && vtx >= 2) {                    // :498-507 (the loop starts at i = 2)
                    const U4 u = draw4(A.seed, meta.x, meta.y, 0x40000000u + vtx, YUNE_BLK_NEE);
                    const float r = u01(u.x);
                    const float p = cl_min(luminance(contrib), 0.95f);
                    if (r >= p) stop = true; else contrib = vscale(contrib, YF_DIV(1.0f, p));
                }
                lp[4 * vtx + 0] = f4(hp, __int_as_float(tri));
                lp[4 * vtx + 1] = f4(n, 0.0f);
                lp[4 * vtx + 2] = f4(d, 0.0f);
                lp[4 * vtx + 3] = f4(contrib, 0.0f);
                if (stop || (int)vtx + 1 >= max_v) end_path = true;
                else {
                    // sample the next light-path direction at this vertex (:476-483)
                    const U4 u = draw4(A.seed, meta.x, meta.y, 0x40000000u + vtx, YUNE_BLK_BOUNCE);
                    float prob = 0.0f, pdf = 1.0f;
                    const bool glossy = select_lobe(mat, u01(u.x), true, prob);
                    const V3 dir = glossy ? sample_phong(vneg(d), n, mat.px, mat.py, u01(u.y), u01(u.z), false, pdf) : sample_cosine(n, u01(u.y), u01(u.z), pdf);
                    if (pdf <= 0.0f) end_path = true;                           // :485
                    else {
                        // contrib_{k+1} = evaluateBRDF(-dir_k, dir_{k+1}, hit_k) * max(0, dot(dir_{k+1}, n_k)) / pdf * contrib_k   (:495-499)
                        V3 c = vscale(eval_brdf(mat, vneg(d), dir, n, glossy, prob, false, use_on), fmaxf(0.0f, vdot(dir, n)));
                        c = vdivs(c, pdf);
                        Tn = vmul(c, contrib);
                        has_ext = true; ext_d = dir; ext_o = vadd(hp, vscale(dir, YUNE_EPS)); ext_t = INFINITY;
                        ext_lid = light_loop(lights, n_lights, ext_o, ext_d, ext_t);
                        new_flags = YS_TRACE | YB_PHASE_LIGHT;
                        meta.z = vtx + 1;
                    }
                }
            }
            bm.x = lp_len;
            if (!has_ext) start_eye = true;
        } else {
            // ---------------- EYE PATH: arrival at eye vertex i = vtx (bdpt.cl:513-560 + shading :562-640) ----------------
            if (tri < 0) {
                const int lid = __float_as_int(rd.w);
                if (vtx == 0) {
                    if (lid >= 0) col = (vdot(d, lights[lid].normal) < 0.0f) ? v3(1.0f, 1.0f, 1.0f) : v3(0.1f, 0.1f, 0.1f);   // :575-581
                    else col = v3(0.4f, 0.4f, 0.4f);                                                                       // :582-583
                }
                finished = true;
            } else {
                const float4 s0 = __ldg(A.sc.shade + 4 * (size_t)tri), s1 = __ldg(A.sc.shade + 4 * (size_t)tri + 1), s2 = __ldg(A.sc.shade + 4 * (size_t)tri + 2);
                const MatDev mat = load_material(A.sc.mats, __float_as_int(s0.w));
                const float bw = YF_SUB(YF_SUB(1.0f, hit.y), hit.z);
                const V3 hp = vadd(o, vscale(d, hit.x));
                const V3 n = vnormalize(vmadd3(xyz(s0), bw, xyz(s1), hit.y, xyz(s2), hit.z));
                const V3 w_o = vneg(d);
                bool last = false;
                if (vtx == 0) { T = v3(1, 1, 1); epw = 1.0f; }
                else {
                    T = xyz(P.thr_next[s]);                                     // eye_path[i].contrib (:546-549)
                    if ((int)vtx > A.rr_threshold) {                            // :550-558
                        const U4 u = draw4(A.seed, meta.x, meta.y, vtx, YUNE_BLK_NEE);
                        const float r = u01(u.x);
                        const float p = cl_min(luminance(T), 0.95f);
                        if (r >= p) last = true; else T = vscale(T, YF_DIV(1.0f, p));
                    }
                }
                if ((int)vtx + 1 >= max_v) last = true;
                if (epw == 0.0f) finished = true;                               // :593 'if(eye_path_weight == 0) break'
                else {
                    shade_eye = true; e_hp = hp; e_n = n; e_wo = w_o; e_mat = mat; e_last = last; e_vtx = vtx;
                }
            }
        }
    }

    // ---- 1b. shading of eye vertex i: NEE + one connection ray per light vertex + next direction.  The connection loop has a
    //          warp-uniform trip count because it allocates queue entries with warp votes. ----
    if (shade_eye) {
        const float ks = cl_max(cl_max(e_mat.ks.x, cl_max(e_mat.ks.y, e_mat.ks.z)), 0.1f);           // :599
        pend_c[0] = f4(e_mat.ke, 0.0f);
        bm.w = __float_as_int(ks);
        const U4 u_nee = draw4(A.seed, meta.x, meta.y, 0x20000000u + e_vtx, YUNE_BLK_NEE);
        const LobePrep e_lobes = lobe_prepare(e_mat, true);
        V3 e_nx, e_ny; onb(e_n, e_nx, e_ny);
        nee_sample<MIS, true>(__activemask(), lights, n_lights, e_mat, e_lobes, e_hp, e_n, e_nx, e_ny, e_wo, u_nee, A.seed, meta.x, meta.y, 0x20000000u + e_vtx, use_on, N);
    }
    {
        const int j_top = __reduce_max_sync(0xffffffffu, shade_eye ? bm.x : 0);
        for (int j = j_top - 1; j > 0; j--) {                                                        // :594-635
            bool want = shade_eye && j < bm.x;
            V3 co = v3(0, 0, 0), cd = v3(0, 0, 1); float clen = 0.0f;
            if (want) {
                const float4 l0 = lp[4 * j + 0], l1 = lp[4 * j + 1], l2 = lp[4 * j + 2], l3 = lp[4 * j + 3];
                const V3 lpnt = xyz(l0), ln = xyz(l1);
                const V3 delta = vsub(lpnt, e_hp);
                cd = vnormalize(delta);
                co = vadd(e_hp, vscale(cd, YUNE_EPS));
                float dist = vlength(delta);
                dist = YF_MUL(dist, dist);
                clen = vlength(vsub(lpnt, co));
                want = !(vdot(cd, e_n) <= 0.0f || vdot(vneg(cd), ln) <= 0.0f);                        // :612-613
                if (want) {
                    float tl = clen;
                    if (light_loop(lights, n_lights, co, cd, tl) >= 0) want = false;                 // a light inside the segment occludes (traceRay, :244-276)
                }
                if (want) {
                    const float gf = YF_DIV(YF_MUL(cl_max(vdot(cd, e_n), 0.0f), cl_max(vdot(vneg(cd), ln), 0.0f)), dist);
                    const U4 uc = draw4(A.seed, meta.x, meta.y, 0x10000000u + 32u * e_vtx + (unsigned)j, YUNE_BLK_BOUNCE);
                    const U4 uc2 = draw4(A.seed, meta.x, meta.y, 0x10000000u + 32u * e_vtx + (unsigned)j, YUNE_BLK_NEE);
                    float prob = 0.0f;
                    bool g = select_lobe(e_mat, u01(uc.x), true, prob);
                    const V3 e2l = eval_brdf(e_mat, cd, e_wo, e_n, g, prob, false, use_on);
                    const MatDev lmat = load_material(A.sc.mats, __float_as_int(__ldg(A.sc.shade + 4 * (size_t)__float_as_int(l0.w)).w));
                    g = select_lobe(lmat, u01(uc2.y), true, prob);
                    const V3 l2e = eval_brdf(lmat, vneg(xyz(l2)), vneg(cd), ln, g, prob, false, use_on);
                    V3 tl3 = vmul(xyz(l3), vmul(vscale(e2l, gf), l2e));                               // throughput_lp *= gf * e2l * l2e   (:631)
                    tl3 = vmul(tl3, T);                                                               // *= throughput                       (:632)
                    pend_c[j] = f4(tl3, 0.0f);
                }
            }
            const int q = warp_alloc(&C->n_shadow, want);
            if (want) {
                conn_mask |= 1u << j;
                P.sq_o[q] = f4(co, clen);
                P.sq_d[q] = f4(cd, __int_as_float(P.n_slots + s * YB_MAXV + j));
            }
        }
    }
    if (shade_eye) {
        // ---- continue the eye path (:525-537) ----
        if (!e_last) {
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
                new_flags = YS_TRACE | YB_PENDING;
                meta.z = e_vtx + 1;
            }
        }
        if (!has_ext) new_flags = YS_DRAIN | YB_PENDING;                       // wait for this vertex's rays, then finish
    }

    // ---- 2. a finished sample goes to the accumulation buffer (bdpt.cl:192-209) ----
    if (finished) {
        if (col.x != col.x || col.y != col.y || col.z != col.z) col = v3(0.988f, 0.0588f, 0.7529f);
        float* dst = reinterpret_cast<float*>(A.sum + meta.x);
        atomicAdd(dst + 0, col.x); atomicAdd(dst + 1, col.y); atomicAdd(dst + 2, col.z); atomicAdd(dst + 3, 1.0f);
        state = YS_FREE;
    }
    need_new = valid && (finished || state == YS_FREE) && !has_ext && new_flags == 0 && !start_eye;

    // ---- 3. regenerate: next (pixel, sample); the light path comes first (createLightPath, :432-457) ----
    {
        const long long g = warp_alloc64(&A.tot->next_sample, need_new);
        if (need_new) {
            if ((unsigned long long)g >= A.tot->n_samples) new_flags = YS_DONE;
            else {
                const unsigned long long n_pix = (unsigned long long)A.width * A.height;
                meta.x = (unsigned)((unsigned long long)g % n_pix);
                meta.y = (unsigned)(A.spp_begin + (int)((unsigned long long)g / n_pix));
                col = v3(0, 0, 0);
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
                const V3 lpnt = vadd(vadd(vscale(L.edge_l, r2), vscale(L.edge_w, r1)), L.pos);       // :444-446
                const float area = YF_MUL(L.la, L.lb);
                const float fwd_pdf = YF_DIV(1.0f, area);
                V3 c0 = vdivs(L.ke, fwd_pdf);                                                          // :454
                if (n_lights > 1) c0 = vdivs(c0, pick_prob);
                lp[0] = f4(lpnt, __int_as_float(-1)); lp[1] = f4(L.normal, 0.0f); lp[2] = f4(v3(0, 0, 0), 0.0f); lp[3] = f4(c0, 0.0f);
                bm.x = 1; bm.y = 0; bm.z = __float_as_int(1.0f); bm.w = 0;
                const U4 ub = draw4(A.seed, meta.x, meta.y, 0x40000000u, YUNE_BLK_BOUNCE);
                float pdf = 1.0f;
                const V3 dir = sample_cosine(L.normal, u01(ub.y), u01(ub.z), pdf);                     // :456
                if (pdf > 0.0f) {
                    const float c1 = YF_DIV(cl_max(0.0f, vdot(L.normal, dir)), pdf);                   // :466
                    Tn = vmul(v3(c1, c1, c1), c0);
                    has_ext = true; ext_d = dir; ext_o = vadd(lpnt, vscale(dir, YUNE_EPS)); ext_t = INFINITY;
                    ext_lid = light_loop(lights, n_lights, ext_o, ext_d, ext_t);
                    new_flags = YS_TRACE | YB_PHASE_LIGHT;
                    meta.z = 1;
                } else start_eye = true;
            }
        }
    }
    // ---- 4. light path complete: camera ray (bdpt.cl:172-189) ----
    if (start_eye) {
        const int px = meta.x % A.width, py = meta.x / A.width;
        const U4 uj = draw4(A.seed, meta.x, meta.y, YUNE_VERTEX_CAMERA, 0u);
        create_ray(A.cam, A.width, A.height, (float)px + u01(uj.x), (float)py + u01(uj.y), ext_o, ext_d);
        ext_t = INFINITY;
        ext_lid = light_loop(lights, n_lights, ext_o, ext_d, ext_t);
        has_ext = true; T = v3(1, 1, 1); Tn = v3(1, 1, 1); epw = 1.0f;
        new_flags = YS_TRACE; meta.z = 0; bm.y = 0;
    }

    // ---- 5. queue pushes and state write-back ----
    const int qe = warp_alloc(&C->n_extend, has_ext);
    if (has_ext) {
        P.eq[qe] = s;
        P.ray_o[s] = f4(ext_o, ext_t);
        P.ray_d[s] = f4(ext_d, __int_as_float(ext_lid));
    }
    const bool is_event = N.MV.has || N.MO.has;
    const int ev = A.parity * P.n_slots + warp_alloc(&C->n_events, is_event);
    if (is_event) {
        const int ef = (N.S.has ? YE_HAS_S : 0) | (N.MV.has ? YE_HAS_MV : 0) | (N.MO.has ? YE_HAS_MO : 0) | (N.mo_is_mv ? YE_MO_IS_MV : 0);
        P.evt[3 * (size_t)ev] = f4(N.Lv, __int_as_float(ef));
        P.evt[3 * (size_t)ev + 1] = f4(N.BV, 0.0f);
        P.evt[3 * (size_t)ev + 2] = f4(N.BO, 0.0f);
        P.evt_idx[s] = ev;
        new_flags |= YF_PEND_EVT;
    } else if (N.S.has) new_flags |= YF_PEND_L;
    const int qs = warp_alloc(&C->n_shadow, N.S.has);
    if (N.S.has) {
        P.sq_o[qs] = f4(N.S.o, N.S.tmax);
        P.sq_d[qs] = f4(N.S.d, __int_as_float(is_event ? ~(4 * ev + 0) : s));
        if (!is_event) P.pend_l[s] = f4(N.Lv, 0.0f);
    }
    if (MIS) {
        const int qv = warp_alloc(&C->n_shadow, N.MV.has);
        if (N.MV.has) { P.sq_o[qv] = f4(N.MV.o, N.MV.tmax); P.sq_d[qv] = f4(N.MV.d, __int_as_float(~(4 * ev + 1))); }
        const int qo = warp_alloc(&C->n_shadow, N.MO.has);
        if (N.MO.has) { P.sq_o[qo] = f4(N.MO.o, N.MO.tmax); P.sq_d[qo] = f4(N.MO.d, __int_as_float(~(4 * ev + 2))); }
    }
    if (valid && (meta.w & YS_STATE_MASK) != YS_DONE) {
        meta.w = new_flags;
        P.meta[s] = meta;
        if ((new_flags & YS_STATE_MASK) != YS_DONE) {
            bm.y = (int)conn_mask; bm.z = __float_as_int(epw);
            B.bmeta[s] = bm;
            P.col[s] = f4(col, 0.0f);
            P.thr[s] = f4(T, 0.0f);
            if (has_ext) P.thr_next[s] = f4(Tn, 0.0f);
        }
    }
    const bool live = valid && (new_flags & YS_STATE_MASK) != YS_DONE && (meta.w & YS_STATE_MASK) != YS_DONE;
    const unsigned lm = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0 && lm) atomicAdd(&C->live, __popc(lm));
}

cudaError_t launch_shade_bdpt(const RenderArgs& a, const BdptPool& b, cudaStream_t st)
{
    const int grid = (a.pool.n_slots + YUNE_SHADE_BLOCK - 1) / YUNE_SHADE_BLOCK;
    if (a.mis) k_shade_bdpt<true><<<grid, YUNE_SHADE_BLOCK, 0, st>>>(a, b);
    else       k_shade_bdpt<false><<<grid, YUNE_SHADE_BLOCK, 0, st>>>(a, b);
    return cudaGetLastError();
}

} // namespace yune
