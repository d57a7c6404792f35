/* yune_host.h -- C ABI over the host-side scene preparation (the part of the reference that is KEPT:
 * yune::Scene / yune::BVH / yune::Camera, include/Scene.h:40-64, include/BVH.h:37-68, include/Camera.h:50-103).
 * It exists so non-C++ callers (the ctypes test/bench harness) can produce the exact buffers that
 * yune_cuda.h consumes.  Same error convention as yune_cuda.h.
 */
#ifndef YUNE_HOST_H
#define YUNE_HOST_H

#include "yune_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yune_scene yune_scene;

yune_scene* yune_scene_create(void);
void        yune_scene_destroy(yune_scene* s);
const char* yune_scene_last_error(const yune_scene* s);

/* Scene::loadModel(filepath, filename) (src/Scene.cpp:133-383); bins = BVH bin count (reference default 20;
 * 0 = no BVH, <= 2 = median splits only). */
int yune_scene_load_model(yune_scene* s, const char* filepath, int bvh_bins);
/* Same result as loadModel for geometry that is already in memory (synthetic scenes): per-triangle centroid + padded box
 * (src/TriangleCPU.cpp:41-71), scene box (src/Scene.cpp:352-356), BVH (src/BVH.cpp:56-173).  Buffers are copied. */
int yune_scene_set_geometry(yune_scene* s, const yune_triangle* tris, int n_triangles, const yune_material* mats, int n_materials, int bvh_bins);
/* Scene::loadBVH(bins) (src/Scene.cpp:413-416). */
int yune_scene_load_bvh(yune_scene* s, int bvh_bins);
/* Scene::reloadMatFile() (src/Scene.cpp:62-131). */
int yune_scene_reload_mat_file(yune_scene* s);

int yune_scene_num_triangles(const yune_scene* s);
int yune_scene_num_materials(const yune_scene* s);
int yune_scene_num_bvh_nodes(const yune_scene* s);
/* Pointers into the scene's own vectors (valid until the next load/destroy). */
const yune_triangle* yune_scene_vert_data(const yune_scene* s);
const yune_material* yune_scene_mat_data(const yune_scene* s);
yune_material*       yune_scene_mat_data_mut(yune_scene* s);
const yune_bvh_node* yune_scene_bvh_data(const yune_scene* s);
void yune_scene_root_aabb(const yune_scene* s, yune_aabb* out);

/* Camera (src/Camera.cpp:36-117): default pose, vertical FOV in degrees -> the 80-byte Cam record. */
void yune_camera_default(float y_fov_degrees, yune_cam* out);
/* General pose: side/up/look_at/eye as 4-vectors (w = 0,0,0,1), Camera::setViewMatrix + setBuffer. */
void yune_camera_set(const float side[4], const float up[4], const float look_at[4], const float eye[4],
                     float y_fov_degrees, yune_cam* out);

/* Interactive pose updates (Camera::setOrientation, src/Camera.cpp:119-166): a stateful camera object.
 * dir = key direction (one axis per call, priority z, x, y; moves by move_speed = 0.1), pitch / yaw = mouse deltas
 * (multiplied by rotation_speed = 0.25 rad); pitch is ignored once it would turn the up vector below the horizon.
 * After a change render with reset = 1 (src/RendererCore.cpp:531-574). */
typedef struct yune_camera yune_camera;
yune_camera* yune_camera_create(float y_fov_degrees);
void yune_camera_destroy(yune_camera* cam);
void yune_camera_set_orientation(yune_camera* cam, const float dir[4], float pitch, float yaw);
void yune_camera_reset(yune_camera* cam);
int  yune_camera_is_changed(const yune_camera* cam);
void yune_camera_set_buffer(yune_camera* cam, yune_cam* out);       /* Camera::setBuffer: writes the record, clears is_changed */

/* Image export: the stb_image_write calls of RendererCore::saveImage(save_fn, save_ext) (src/RendererCore.cpp:608-646).  `rgba`
 * is a bottom-up RGBA float image as returned by yune_read_hdr / yune_read_ldr; the file is `path` as given and the format
 * follows `save_ext` (".png" ...; NULL or "" = the extension of `path`):
 * ".hdr" (Radiance RGBE) and ".pfm" keep the float values; ".png", ".jpg" (baseline, quality 100) and ".ppm" store
 * clamp(v, 0, 1) * 255 rounded, as the reference's GL_UNSIGNED_BYTE read-back does.  The file's first row is the top of the
 * picture; alpha is dropped.  Returns 0, -1 (bad argument), -2 (unsupported extension) or -3 (file could not be written). */
int yune_write_image(const char* path, const char* save_ext, const float* rgba, int width, int height);

#ifdef __cplusplus
}
#endif
#endif
