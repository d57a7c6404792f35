// host_capi.cpp -- C ABI of include/yune_host.h over yune::Scene / yune::Camera.
#include "yune_host.h"
#include "Scene.h"
#include "ImageIO.h"

#include <new>
#include <string>
#include <exception>

struct yune_scene { yune::Scene scene; std::string err; };

namespace { thread_local std::string g_create_err; }

extern "C" {

yune_scene* yune_scene_create(void)
{
    try { return new yune_scene(); } catch (const std::exception& e) { g_create_err = e.what(); return nullptr; }
}
void yune_scene_destroy(yune_scene* s) { delete s; }
const char* yune_scene_last_error(const yune_scene* s) { return s ? s->err.c_str() : g_create_err.c_str(); }

int yune_scene_load_model(yune_scene* s, const char* filepath, int bvh_bins)
{
    if (!s || !filepath) return -1;
    try {
        std::string fp(filepath);
        std::string fn = fp.substr(fp.find_last_of("/") + 1);
        s->scene.bvh.bins = 20;                       // loadModel always builds with the default (include/BVH.h:43) ...
        if (bvh_bins == 0) s->scene.bvh.bins = 0;     // ... unless the BVH is disabled
        s->scene.loadModel(fp, fn);
        if (bvh_bins > 0 && bvh_bins != 20) s->scene.loadBVH(bvh_bins);
    } catch (const std::exception& e) { s->err = e.what(); return -1; }
    return 0;
}
int yune_scene_set_geometry(yune_scene* s, const yune_triangle* tris, int n_triangles, const yune_material* mats, int n_materials, int bvh_bins)
{
    if (!s || n_triangles < 0 || (n_triangles > 0 && !tris)) return -1;
    try { s->scene.setGeometry(tris, n_triangles, mats, n_materials, bvh_bins); } catch (const std::exception& e) { s->err = e.what(); return -1; }
    return 0;
}
int yune_scene_load_bvh(yune_scene* s, int bvh_bins)
{
    if (!s) return -1;
    try { s->scene.loadBVH(bvh_bins); } catch (const std::exception& e) { s->err = e.what(); return -1; }
    return 0;
}
int yune_scene_reload_mat_file(yune_scene* s)
{
    if (!s) return -1;
    try { s->scene.reloadMatFile(); } catch (const std::exception& e) { s->err = e.what(); return -1; }
    return 0;
}
int yune_scene_num_triangles(const yune_scene* s) { return s ? (int)s->scene.vert_data.size() : 0; }
int yune_scene_num_materials(const yune_scene* s) { return s ? (int)s->scene.mat_data.size() : 0; }
int yune_scene_num_bvh_nodes(const yune_scene* s) { return s ? (int)s->scene.bvh.gpu_node_list.size() : 0; }
const yune_triangle* yune_scene_vert_data(const yune_scene* s) { return s ? s->scene.vert_data.data() : nullptr; }
const yune_material* yune_scene_mat_data(const yune_scene* s) { return s ? s->scene.mat_data.data() : nullptr; }
yune_material* yune_scene_mat_data_mut(yune_scene* s) { return s ? s->scene.mat_data.data() : nullptr; }
const yune_bvh_node* yune_scene_bvh_data(const yune_scene* s) { return s ? s->scene.bvh.gpu_node_list.data() : nullptr; }
void yune_scene_root_aabb(const yune_scene* s, yune_aabb* out) { if (s && out) *out = s->scene.root; }

void yune_camera_default(float fov, yune_cam* out)
{
    if (!out) return;
    yune::Camera cam(fov);
    cam.setBuffer(out);
}
void yune_camera_set(const float side[4], const float up[4], const float look_at[4], const float eye[4], float fov, yune_cam* out)
{
    if (!side || !up || !look_at || !eye || !out) return;
    yune::Camera cam(fov);
    cam.setViewMatrix(yune::Vec4{side[0], side[1], side[2], side[3]}, yune::Vec4{up[0], up[1], up[2], up[3]},
                      yune::Vec4{look_at[0], look_at[1], look_at[2], look_at[3]}, yune::Vec4{eye[0], eye[1], eye[2], eye[3]});
    cam.setBuffer(out);
}


int yune_write_image(const char* path, const char* save_ext, const float* rgba, int width, int height)
{
    if (!path || !rgba || width <= 0 || height <= 0) return -1;
    try {
        const std::string ext = (save_ext && save_ext[0]) ? yune::imageExtension(save_ext) : yune::imageExtension(path);
        if (!yune::imageIsLdr(ext) && !yune::imageIsHdr(ext)) return -2;
        std::string err;
        return yune::writeImage(path, ext, rgba, width, height, err) ? 0 : -3;
    } catch (const std::exception&) { return -3; }
}

struct yune_camera { yune::Camera cam; explicit yune_camera(float fov) : cam(fov) {} };
yune_camera* yune_camera_create(float fov) { return new (std::nothrow) yune_camera(fov); }
void yune_camera_destroy(yune_camera* c) { delete c; }
void yune_camera_set_orientation(yune_camera* c, const float dir[4], float pitch, float yaw)
{
    if (c && dir) c->cam.setOrientation(yune::Vec4{dir[0], dir[1], dir[2], dir[3]}, pitch, yaw);
}
void yune_camera_reset(yune_camera* c) { if (c) c->cam.resetCamera(); }
int  yune_camera_is_changed(const yune_camera* c) { return (c && c->cam.is_changed) ? 1 : 0; }
void yune_camera_set_buffer(yune_camera* c, yune_cam* out) { if (c && out) c->cam.setBuffer(out); }

}
