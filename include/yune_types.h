/* yune_types.h -- plain-old-data records that cross the drop-in boundary.
 *
 * Byte-for-byte the device layouts of the reference (include/CL_headers.h:57-107 on the host,
 * kernels/legacy/udpt.cl:34-91 on the device); SURVEY.md appendix A lists the offsets.  They are written
 * here without any OpenCL typedefs so that C, C++, CUDA and ctypes/numpy users agree on them.
 */
#ifndef YUNE_TYPES_H
#define YUNE_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yune_float4 { float s[4]; } yune_float4;

/* include/CL_headers.h:57-65 -- rows of the view-to-world matrix + view-plane distance. 80 B. */
typedef struct yune_cam {
    yune_float4 r1, r2, r3, r4;
    float view_plane_dist;
    float pad[3];
} yune_cam;

/* include/CL_headers.h:67-77 -- 112 B. v*.w = 1, vn*.w = 0; pad is never written by the reference. */
typedef struct yune_triangle {
    yune_float4 v1, v2, v3;
    yune_float4 vn1, vn2, vn3;
    int32_t matID;
    float pad[3];
} yune_triangle;

/* include/CL_headers.h:79-84 -- 32 B. */
typedef struct yune_aabb { yune_float4 p_min, p_max; } yune_aabb;

/* include/CL_headers.h:86-92 -- 80 B.
 * inner: child_idx = index of first child (> 0, sibling at +1), vert_len = -1
 * leaf : child_idx = -1, 1 <= vert_len <= 10, vert_list = triangle indices
 * empty: child_idx = -2, vert_len = -1 (never traversed, udpt.cl:316) */
typedef struct yune_bvh_node {
    yune_aabb aabb;
    int32_t vert_list[10];
    int32_t child_idx;
    int32_t vert_len;
} yune_bvh_node;

/* include/CL_headers.h:94-107 -- 80 B. */
typedef struct yune_material {
    yune_float4 ke, kd, ks;
    float n, k, px, py, alpha_x, alpha_y;
    int32_t is_specular, is_transmissive;
} yune_material;

/* kernels/legacy/udpt.cl:34-45 -- the kernels' built-in quad light, 128 B under OpenCL alignment rules.
 * In the reference this is a __constant array inside each .cl file; here it is data. */
typedef struct yune_quad_light {
    yune_float4 pos, normal, ke, kd, ks, edge_l, edge_w;
    float phong_exponent;
    float pad[3];
} yune_quad_light;

#ifdef __cplusplus
}
static_assert(sizeof(yune_cam) == 80, "Cam layout");
static_assert(sizeof(yune_triangle) == 112, "TriangleGPU layout");
static_assert(sizeof(yune_aabb) == 32, "AABB layout");
static_assert(sizeof(yune_bvh_node) == 80, "BVHNodeGPU layout");
static_assert(sizeof(yune_material) == 80, "Material layout");
static_assert(sizeof(yune_quad_light) == 128, "Quad layout");
#endif

#endif /* YUNE_TYPES_H */
