// Scene.h -- OBJ + (custom lowercase) MTL loader producing the device-layout arrays.
// Same public surface as yune::Scene (include/Scene.h:40-64): loadModel / loadBVH / reloadMatFile and the
// public vert_data / mat_data / bvh / main_camera / root members that RendererCore reads.
#ifndef YUNE_SCENE_H
#define YUNE_SCENE_H

#include "CUDA_headers.h"
#include "Camera.h"
#include "BVH.h"

#include <string>
#include <vector>

namespace yune
{
    class Scene
    {
        public:
            Scene();
            /** Parse `filepath` (an .obj; `filename` is its last path component) and the .mtl it names, fill vert_data /
             *  mat_data / root, and build the BVH when bvh.bins > 0 (src/Scene.cpp:133-383).  Throws std::runtime_error
             *  with the reference's messages on unreadable files.  Unlike the reference it never WRITES a default .mtl next
             *  to an .obj that has no mtllib line (src/Scene.cpp:139-168): the default material is used in memory only. */
            void loadModel(std::string filepath, std::string filename);
            /** loadModel without the parsing: take device-layout triangles and materials that are already in memory. */
            void setGeometry(const TriangleGPU* tris, int n_tris, const Material* mats, int n_mats, int bvh_bins);
            void loadBVH(int bvh_bins);                 /**< Rebuild with another bin count (src/Scene.cpp:413-416). */
            void reloadMatFile();                       /**< Re-read mat_file into mat_data (src/Scene.cpp:62-131). */

            Camera main_camera;
            std::vector<TriangleGPU> vert_data;
            std::vector<Material> mat_data;
            std::string scene_file, mat_file, mat_filename;
            AABB root;
            BVH bvh;
            int num_triangles;
            float scene_size_kb, scene_size_mb;

        private:
            void clearValues();
            std::vector<TriangleCPU> cpu_tri_list;
    };
}
#endif
