// oracle/_ref/libyune_ref_host.so -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" driver around the REFERENCE's own host sources, compiled where they lie
// (/root/reference/src/{Scene,BVH,BVHNodeCPU,TriangleCPU}.cpp) against oracle/shim.  It exists so the
// tests can compare the product's Scene/BVH output byte-for-byte with what the reference produces
// (pin 1).  Nothing here is reference code: it only calls yune::Scene.
//
// src/Camera.cpp needs the real GLM (glm::row/rotate/...), which is not in this image, so the two
// Camera special members that yune::Scene's constructor/destructor need are stubbed below; Camera is
// never used by loadModel().
#include "Scene.h"
#include <cstring>
#include <string>
#include <iostream>
#include <sstream>

namespace yune {
Camera::Camera() : y_FOV(60.0f), rotation_speed(0.25f), move_speed(0.1f) { is_changed = true; }
Camera::~Camera() {}
}

namespace {
struct Quiet {   // the reference prints progress banners to std::cout (src/Scene.cpp:159-371)
    std::streambuf* old; std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
thread_local std::string g_err;
}

extern "C" {

const char* yref_last_error() { return g_err.c_str(); }

// bins < 0: keep the BVH that loadModel built with the reference default (20 bins, include/BVH.h:43).
void* yref_scene_load(const char* filepath, const char* filename, int bins)
{
    yune::Scene* s = new yune::Scene();
    try {
        Quiet q;
        s->loadModel(filepath, filename);
        if (bins >= 0 && bins != 20) s->loadBVH(bins);
    } catch (const std::exception& e) {
        g_err = e.what(); delete s; return nullptr;
    }
    return s;
}

void yref_scene_counts(void* h, int* ntri, int* nmat, int* nnodes)
{
    yune::Scene* s = (yune::Scene*)h;
    *ntri = (int)s->vert_data.size(); *nmat = (int)s->mat_data.size(); *nnodes = (int)s->bvh.gpu_node_list.size();
}

void yref_scene_copy(void* h, void* tris, void* mats, void* nodes, float* root8)
{
    yune::Scene* s = (yune::Scene*)h;
    if (tris)  std::memcpy(tris,  s->vert_data.data(), s->vert_data.size() * sizeof(TriangleGPU));
    if (mats)  std::memcpy(mats,  s->mat_data.data(),  s->mat_data.size()  * sizeof(Material));
    if (nodes) std::memcpy(nodes, s->bvh.gpu_node_list.data(), s->bvh.gpu_node_list.size() * sizeof(BVHNodeGPU));
    if (root8) std::memcpy(root8, &s->root, sizeof(AABB));
}

void yref_scene_free(void* h) { delete (yune::Scene*)h; }

void yref_sizes(int* out4) { out4[0] = sizeof(TriangleGPU); out4[1] = sizeof(BVHNodeGPU); out4[2] = sizeof(Material); out4[3] = sizeof(Cam); }

}
