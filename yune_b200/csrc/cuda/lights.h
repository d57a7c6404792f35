// lights.h -- the analytic quad-light intersection that opens traceRay (udpt.cl:244-276), host/device.
// In the reference it runs inside every traceRay call before the BVH walk; here the kernel that CREATES a ray
// runs it once and hands the BVH walk the shortened t, which gives the same closest hit.
#ifndef YUNE_LIGHTS_H
#define YUNE_LIGHTS_H

#include "strict_math.h"

#define YUNE_MAX_LIGHTS 8

namespace yune {

struct LightDev {           // unpacked yune_quad_light + the two edge lengths (udpt.cl:259-260)
    V3 pos, normal, ke, edge_l, edge_w;
    float la, lb;
};

// Loops the lights in index order with the reference's strict comparisons.  On return `t` is the (possibly
// shortened) ray length and the result is the index of the light that owns it, or -1.
YUNE_HD_CALL int light_loop(const LightDev* lights, int n_lights, V3 o, V3 d, float& t_len)
{
    int id = -1;
    YUNE_NO_UNROLL
    for (int i = 0; i < n_lights; i++) {
        const LightDev& L = lights[i];
        const float DdotN = vdot(d, L.normal);
        // 'fabs(DdotN) > 0.0001' compares in double (unsuffixed literal); 0.0001f is the largest float below 0.0001,
        // so the float comparison decides identically.
        if (fabsf(DdotN) > 0.0001f) {
            const float t = YF_DIV(vdot(L.normal, vsub(L.pos, o)), DdotN);
            if (t > 0.0f && t < t_len) {
                V3 temp = vadd(o, vscale(d, t));
                temp = vsub(temp, L.pos);
                float proj1 = vdot(temp, L.edge_l), proj2 = vdot(temp, L.edge_w);
                proj1 = YF_DIV(proj1, L.la);
                proj2 = YF_DIV(proj2, L.lb);
                if ((proj1 >= 0.0f && proj2 >= 0.0f) && (proj1 <= L.la && proj2 <= L.lb)) { t_len = t; id = i; }
            }
        }
    }
    return id;
}

} // namespace yune
#endif
