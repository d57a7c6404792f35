// shade_core.h -- the leaf functions of the estimator (BSDF evaluation, lobe selection, direction sampling,
// light sampling, Fresnel), host/device, in the reference's operation order.
//
// Each function names the reference statement it restates (kernels/legacy/udpt.cl unless noted) and keeps
// the quirks listed in SURVEY.md appendix B, because they shape the MEAN of the estimator:
//   B#9  Fresnel uses 1 - cos^5, not (1 - cos)^5          B#10 calcPhongPDF takes cos() of a dot product
//   B#10 Phong normalisation uses (int)(px+py)+2           B#11 lobe selection can absorb (prob = 0)
//   B#21 normals are never flipped toward the ray except inside reflect()
// Random numbers are PARAMETERS here (r, r1, r2): the callers decide which stream they come from.
// Double-precision sub-expressions appear exactly where the kernels have unsuffixed literals (B#5).
#ifndef YUNE_SHADE_CORE_H
#define YUNE_SHADE_CORE_H

#include "strict_math.h"
#include "lights.h"

namespace yune {

#define YUNE_PI      3.14159265359f
#define YUNE_INV_PI  0.31830988618f
#define YUNE_EPS     0.0001f

struct MatDev {
    V3 ke, kd, ks;
    float n, px, py, alpha_x;
    int is_specular, is_transmissive;
};

YUNE_HD V3 vmadd3(V3 a, float sa, V3 b, float sb, V3 c, float sc)   // a*sa + b*sb + c*sc, left to right
{
    return vadd(vadd(vscale(a, sa), vscale(b, sb)), vscale(c, sc));
}

// udpt.cl:1121-1124
YUNE_HD float luminance(V3 c) { return YF_ADD(YF_ADD(YF_MUL(0.212671f, c.x), YF_MUL(0.715160f, c.y)), YF_MUL(0.072169f, c.z)); }

// udpt.cl:949-961 -- mirror direction about the normal turned to the side of w_i; normalised.
YUNE_HD V3 reflect_flip(V3 w_i, V3 n)
{
    if (vdot(w_i, n) < 0.0f) n = vscale(n, -1.0f);
    V3 r = vsub(vscale(n, YF_MUL(2.0f, vdot(w_i, n))), w_i);
    return vnormalize(r);
}
// bdpt.cl:723, 814 -- same without the flip and without the inner normalize.
YUNE_HD V3 reflect_noflip(V3 w_i, V3 n) { return vsub(vscale(n, YF_MUL(2.0f, vdot(w_i, n))), w_i); }

// Orthonormal basis around Nz (udpt.cl:846-857): columns Nx, Ny, Nz.
YUNE_HD void onb(V3 Nz, V3& Nx, V3& Ny)
{
    if (fabsf(Nz.y) > fabsf(Nz.z)) Nx = v3(Nz.y, -Nz.x, 0.0f);
    else                           Nx = v3(Nz.z, 0.0f, -Nz.x);
    Nx = vnormalize(Nx);
    Ny = vnormalize(vcross(Nz, Nx));
}
// rows of normal_to_world dotted with (x, y, z, 0) (udpt.cl:897-900), then normalize
YUNE_HD V3 onb_to_world(V3 Nx, V3 Ny, V3 Nz, float x, float y, float z)
{
    V3 d = v3(YF_ADD(YF_ADD(YF_MUL(Nx.x, x), YF_MUL(Ny.x, y)), YF_MUL(Nz.x, z)),
              YF_ADD(YF_ADD(YF_MUL(Nx.y, x), YF_MUL(Ny.y, y)), YF_MUL(Nz.y, z)),
              YF_ADD(YF_ADD(YF_MUL(Nx.z, x), YF_MUL(Ny.z, y)), YF_MUL(Nz.z, z)));
    return vnormalize(d);
}

// libm calls behind one definition each: real calls in the BDPT translation unit (code size), inlined elsewhere
struct CosSin { float c, s; };
YUNE_HD_LEAF CosSin cos_sin(float phi) { CosSin r; r.c = cosf(phi); r.s = sinf(phi); return r; }
YUNE_HD_LEAF float pow_leaf(float x, float y) { return powf(x, y); }

// cosineWeightedHemisphere (udpt.cl:843-910): direction about the shading normal, pdf = cos/pi.
// (Nx, Ny) = onb(n): a surface point samples up to three directions about the same normal, the frame is built once.
YUNE_HD_CALL V3 sample_cosine_onb(V3 n, V3 Nx, V3 Ny, float r1, float r2, float& pdf)
{
    const float phi = YF_MUL(YF_MUL(2.0f, YUNE_PI), r2);
    const float sinTheta = YF_SQRT(r1);
    const CosSin cs = cos_sin(phi);
    const float x = YF_MUL(sinTheta, cs.c), y = YF_MUL(sinTheta, cs.s), z = YF_SQRT(YF_SUB(1.0f, r1));
    pdf = YF_MUL(z, YUNE_INV_PI);
    return onb_to_world(Nx, Ny, n, x, y, z);
}
YUNE_HD_CALL V3 sample_cosine(V3 n, float r1, float r2, float& pdf)
{
    V3 Nx, Ny; onb(n, Nx, Ny);
    const float phi = YF_MUL(YF_MUL(2.0f, YUNE_PI), r2);
    const float sinTheta = YF_SQRT(r1);
    const CosSin cs = cos_sin(phi);
    const float x = YF_MUL(sinTheta, cs.c), y = YF_MUL(sinTheta, cs.s), z = YF_SQRT(YF_SUB(1.0f, r1));
    pdf = YF_MUL(z, YUNE_INV_PI);
    return onb_to_world(Nx, Ny, n, x, y, z);
}

// phongSampleHemisphere (udpt.cl:700-771): lobe about the mirror direction of w (w = direction back along the
// arriving ray).  flip_normal = udpt behaviour (reflect()), false = bdpt.cl:814.  pdf = 0 below the surface.
struct DirPdf { V3 d; float pdf; };
YUNE_HD_LEAF DirPdf sample_phong_v(V3 w, V3 n, float px, float py, float r1, float r2, bool flip_normal)
{
    float pdf;
    V3 Nz = flip_normal ? reflect_flip(w, n) : reflect_noflip(w, n);
    Nz = vnormalize(Nz);
    V3 Nx, Ny; onb(Nz, Nx, Ny);
    const int phong_exponent = (int)YF_ADD(px, py);
    const float phi = YF_MUL(YF_MUL(2.0f, YUNE_PI), r2);
    const float costheta = pow_leaf(r1, YF_DIV(1.0f, (float)(phong_exponent + 1)));
    float sintheta = YF_SUB(1.0f, pow_leaf(r1, YF_DIV(2.0f, (float)(phong_exponent + 1))));
    sintheta = YF_SQRT(sintheta);
    const CosSin cs = cos_sin(phi);
    const float x = YF_MUL(sintheta, cs.c), y = YF_MUL(sintheta, cs.s), z = costheta;
    V3 dir = onb_to_world(Nx, Ny, Nz, x, y, z);
    if (vdot(dir, n) < 0.0f) pdf = 0.0f;
    else pdf = (float)((phong_exponent + 1) * 0.5 * (double)YUNE_INV_PI * (double)pow_leaf(costheta, (float)phong_exponent));   // :770, double
    DirPdf r; r.d = dir; r.pdf = pdf; return r;
}
YUNE_HD V3 sample_phong(V3 w, V3 n, float px, float py, float r1, float r2, bool flip_normal, float& pdf)
{
    const DirPdf r = sample_phong_v(w, n, px, py, r1, r2, flip_normal);
    pdf = r.pdf;
    return r.d;
}

// sampleGlossyPdf (udpt.cl:1062-1119; bdpt variant bdpt.cl:1048-1105): choose diffuse or glossy lobe with the
// uniform r; returns true for glossy and the selection probability (0 = absorbed, udpt only).
// The part that does not depend on r is computed once per surface point (lobe_prepare) and used by the NEE and the bounce.
struct LobePrep { int mode; float pd, ps; };       // mode 0: diffuse only, 1: glossy only, 2: both with probabilities pd / ps
YUNE_HD_CALL LobePrep lobe_prepare(const MatDev& m, bool bdpt_variant)
{
    LobePrep L; L.pd = L.ps = 0.0f;
    const V3 ks = m.ks, kd = m.kd;
    if (vlength(ks) == 0.0f) { L.mode = 0; return L; }
    if (vlength(kd) == 0.0f || YF_ADD(YF_ADD(kd.x, kd.y), kd.z) == 0.0f) { L.mode = 1; return L; }
    const V3 sum = vadd(ks, kd);
    const float max_val = cl_max(sum.x, cl_max(sum.y, sum.z));
    float pd, ps;
    if (max_val == sum.x) { pd = kd.x; ps = ks.x; }
    else if (max_val == sum.y) { pd = kd.y; ps = ks.y; }
    else { pd = kd.z; ps = ks.z; }
    if (bdpt_variant && max_val < 1.0f) { const float pad = YF_DIV(YF_SUB(1.0f, max_val), 2.0f); pd = YF_ADD(pd, pad); ps = YF_ADD(ps, pad); }
    L.mode = 2; L.pd = pd; L.ps = ps;
    return L;
}
YUNE_HD bool select_lobe_r(const LobePrep& L, float r, bool bdpt_variant, float& prob)
{
    if (L.mode == 0) { prob = 1.0f; return false; }
    if (L.mode == 1) { prob = 1.0f; return true; }
    const float pd = L.pd, ps = L.ps;
    if (bdpt_variant) {
        if (r < pd) { prob = pd; return false; }
        prob = ps; return true;
    }
    if (r < pd) { prob = pd; return false; }
    if (r < YF_ADD(pd, ps) && r >= pd) { prob = ps; return true; }
    prob = 0.0f; return false;
}
YUNE_HD_CALL bool select_lobe(const MatDev& m, float r, bool bdpt_variant, float& prob)
{
    const LobePrep L = lobe_prepare(m, bdpt_variant);
    return select_lobe_r(L, r, bdpt_variant, prob);
}

// toShadingSpace + trig helpers + OrenNayarBRDF (udpt-primitives.cl:696-725, 1147-1210); sigma^2 = alpha_x.
YUNE_HD float on_sin_theta(V3 w) { return YF_SQRT(YF_SUB(1.0f, YF_MUL(w.z, w.z))); }
YUNE_HD float on_cos_phi(V3 w) { const float s = on_sin_theta(w); if (s <= YUNE_EPS && s >= -YUNE_EPS) return 0.0f; return fminf(fmaxf(YF_DIV(w.x, s), -1.0f), 1.0f); }
YUNE_HD float on_sin_phi(V3 w) { const float s = on_sin_theta(w); if (s <= YUNE_EPS && s >= -YUNE_EPS) return 0.0f; return fminf(fmaxf(YF_DIV(w.y, s), -1.0f), 1.0f); }
YUNE_HD_LEAF V3 oren_nayar(V3 kd, float sigma_sq, V3 w_i, V3 w_o, V3 n)
{
    V3 Nx, Ny; onb(n, Nx, Ny);
    const V3 wi = v3(vdot(Nx, w_i), vdot(Ny, w_i), vdot(n, w_i));
    const V3 wo = v3(vdot(Nx, w_o), vdot(Ny, w_o), vdot(n, w_o));
    float costerm = 0.0f;
    if (on_sin_theta(wi) >= YUNE_EPS && on_sin_theta(wo) >= YUNE_EPS)
        costerm = fmaxf(0.0f, YF_ADD(YF_MUL(on_cos_phi(wi), on_cos_phi(wo)), YF_MUL(on_sin_phi(wi), on_sin_phi(wo))));
    float sin_alpha, tan_beta;
    if (fabsf(wi.z) < fabsf(wo.z)) { sin_alpha = on_sin_theta(wi); tan_beta = YF_DIV(on_sin_theta(wo), fabsf(wo.z)); }
    else                           { sin_alpha = on_sin_theta(wo); tan_beta = YF_DIV(on_sin_theta(wi), fabsf(wi.z)); }
    const float A = (float)(1 - ((double)sigma_sq / (2 * ((double)sigma_sq + 0.33))));
    const float B = (float)(0.45 * (double)sigma_sq / ((double)sigma_sq + 0.09));
    const float f = YF_ADD(A, YF_MUL(B, YF_MUL(YF_MUL(costerm, sin_alpha), tan_beta)));
    return vscale(vscale(kd, YUNE_INV_PI), f);
}

// evaluateBRDF (udpt.cl:611-630; bdpt.cl:718-737 does not flip the normal).  rr_prob = lobe-selection probability.
// rr_prob == 0 divides by zero exactly like the reference does in its MIS branch (see engine notes).
YUNE_HD_CALL V3 eval_brdf(const MatDev& m, V3 w_i, V3 w_o, V3 n, bool glossy, float rr_prob, bool flip_normal, bool use_oren_nayar)
{
    if (!glossy) {
        if (use_oren_nayar && rr_prob == 1.0f) return oren_nayar(m.kd, m.alpha_x, w_i, w_o, n);       // udpt-primitives.cl:681-686
        const V3 c = vscale(m.kd, YUNE_INV_PI);
        return rr_prob == 1.0f ? c : vdivs(c, rr_prob);                                 // x / 1 is x, bit for bit
    }
    V3 refl = flip_normal ? reflect_flip(w_i, n) : reflect_noflip(w_i, n);
    refl = vnormalize(refl);
    const float cos_alpha = pow_leaf(fmaxf(vdot(w_o, refl), 0.0f), YF_ADD(m.px, m.py));
    const int phong_exp = (int)YF_ADD(m.px, m.py);
    V3 c = vscale(m.ks, cos_alpha);
    c = vscale(c, (float)(phong_exp + 2));
    c = vscale(c, YUNE_INV_PI);
    c = vscale(c, 0.5f);
    return rr_prob == 1.0f ? c : vdivs(c, rr_prob);
}

// calcPhongPDF (udpt.cl:1033-1045, including the cos() of the dot product) and calcCosPDF (:1047-1050)
YUNE_HD_CALL float phong_pdf(const MatDev& m, V3 w_i, V3 w_o, V3 n)
{
    V3 refl = vsub(vscale(n, YF_MUL(2.0f, vdot(w_o, n))), w_o);
    refl = vnormalize(refl);
    const float costheta = fmaxf(0.0f, cosf(vdot(refl, w_i)));
    const float e = YF_ADD(m.px, m.py);
    return (float)(((double)YF_ADD(e, 1.0f)) * 0.5 * (double)YUNE_INV_PI * (double)pow_leaf(costheta, e));
}
YUNE_HD float cos_pdf(V3 w_i, V3 n) { return YF_MUL(fmaxf(vdot(w_i, n), 0.0f), YUNE_INV_PI); }
// powerHeuristic with beta = 2 (udpt.cl:1149-1152): w^2 / (a^2 + b^2)
YUNE_HD float power_heuristic(float w, float a, float b) { return YF_DIV(YF_MUL(w, w), YF_ADD(YF_MUL(a, a), YF_MUL(b, b))); }

// evalFresnelReflectance (udpt.cl:992-1031).  Returns R; ior_factor is only written when no total internal reflection.
YUNE_HD_CALL float fresnel_reflectance(const MatDev& m, V3 w_i, V3 n, float& ior_factor)
{
    float n1, n2;
    if (vdot(w_i, n) < 0.0f) { n1 = m.n; n2 = 1.0f; n = vscale(n, -1.0f); }
    else { n1 = 1.0f; n2 = m.n; }
    const float cosThetaI = vdot(w_i, n);
    const float sinThetaI = YF_SQRT(YF_SUB(1.0f, YF_MUL(cosThetaI, cosThetaI)));
    const float sinThetaT = YF_DIV(YF_MUL(n1, sinThetaI), n2);
    const float cosThetaT = YF_SQRT(YF_SUB(1.0f, YF_MUL(sinThetaT, sinThetaT)));
    if (sinThetaT >= 1.0f && n1 > n2) return 1.0f;
    float r0 = YF_ADD(YF_ADD(m.ks.x, m.ks.y), m.ks.z);
    r0 = YF_DIV(r0, 3.0f);
    ior_factor = YF_DIV(YF_MUL(n2, n2), YF_MUL(n1, n1));
    const float c = (n1 > n2) ? cosThetaI : cosThetaT;
    return YF_ADD(r0, YF_MUL(YF_SUB(1.0f, r0), YF_SUB(1.0f, pow_leaf(c, 5.0f))));
}
// refract (udpt.cl:963-990)
YUNE_HD_CALL V3 refract_dir(const MatDev& m, V3 w_i, V3 n)
{
    float n1, n2;
    if (vdot(w_i, n) < 0.0f) { n1 = m.n; n2 = 1.0f; n = vscale(n, -1.0f); }
    else { n1 = 1.0f; n2 = m.n; }
    const V3 wt_perp = vscale(vsub(vscale(n, vdot(w_i, n)), w_i), YF_DIV(n1, n2));
    const float lp = vlength(wt_perp);
    const V3 wt_parallel = vscale(vneg(n), YF_SQRT(YF_SUB(1.0f, YF_MUL(lp, lp))));
    return vnormalize(vadd(wt_perp, wt_parallel));
}
// sampleFresnelIncidence (udpt.cl:912-947): mirror or dielectric.  r is consumed only when 0 <= R < 1 decides.
YUNE_HD_CALL V3 sample_specular(const MatDev& m, V3 w_i, V3 n, float r, float& ior_factor)
{
    ior_factor = 1.0f;
    if (m.is_transmissive) {
        float f = 1.0f;
        const float pdf = fresnel_reflectance(m, w_i, n, f);
        if (pdf == 1.0f) return reflect_flip(w_i, n);
        if (r < pdf) return reflect_flip(w_i, n);
        ior_factor = f;
        return refract_dir(m, w_i, n);
    }
    return reflect_flip(w_i, n);
}

// sampleLights (udpt.cl:632-698).  u[2*i], u[2*i+1] = point on light i; u_pick = light choice (n_lights > 1).
// Returns the chosen light or -1; w_i is NOT normalised (the caller needs its length), pdf is w.r.t. solid angle.
YUNE_HD_CALL int sample_lights(const LightDev* lights, int n_lights, V3 p, V3 n, const float* u, float u_pick, float& light_pdf, V3& w_i)
{
    float sum = 0.0f;
    float weights[YUNE_MAX_LIGHTS];
    V3 w_is[YUNE_MAX_LIGHTS];
    YUNE_NO_UNROLL
    for (int i = 0; i < n_lights; i++) {
        const LightDev& L = lights[i];
        const float r1 = u[2 * i], r2 = u[2 * i + 1];
        V3 temp = vsub(vadd(vadd(L.pos, vscale(L.edge_l, r1)), vscale(L.edge_w, r2)), p);
        const float distance = vdot(temp, temp);
        w_is[i] = temp;
        temp = vnormalize(temp);
        const float cl = cl_max(vdot(vneg(temp), L.normal), 0.0f);
        const float cosine_falloff = YF_MUL(cl_max(vdot(temp, n), 0.0f), cl);
        if (cosine_falloff <= 0.0f) { weights[i] = 0.0f; continue; }
        const float area = YF_MUL(L.la, L.lb);
        if (n_lights == 1) {
            w_i = w_is[i];
            light_pdf = YF_DIV(1.0f, area);
            light_pdf = YF_MUL(light_pdf, YF_DIV(distance, fmaxf(vdot(vneg(temp), L.normal), 0.0f)));
            return i;
        }
        weights[i] = YF_DIV(YF_MUL(YF_MUL(vlength(L.ke), cosine_falloff), area), distance);
        sum = YF_ADD(sum, weights[i]);
    }
    if (sum == 0.0f) return -1;
    float cumulative = 0.0f;
    YUNE_NO_UNROLL
    for (int i = 0; i < n_lights; i++) {
        const float weight = YF_DIV(weights[i], sum);
        if (u_pick >= cumulative && u_pick < YF_ADD(cumulative, weight)) {
            const LightDev& L = lights[i];
            const float area = YF_MUL(L.la, L.lb);
            w_i = w_is[i];
            light_pdf = YF_DIV(weight, area);
            light_pdf = YF_MUL(light_pdf, YF_DIV(vdot(w_i, w_i), fmaxf(vdot(vneg(vnormalize(w_i)), L.normal), 0.0f)));
            return i;
        }
        cumulative = YF_ADD(cumulative, weight);
    }
    return -1;      // the reference falls off the end of the function here (undefined); we define it as "no light"
}

} // namespace yune
#endif
