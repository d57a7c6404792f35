// Camera.h -- right-handed look-down-minus-Z camera with the reference's buffer contract.
// Restates yune::Camera (include/Camera.h:50-103, src/Camera.cpp:36-117) without GLM: the only thing the
// kernels ever see is Cam = rows of the view-to-world matrix + the view-plane distance.
#ifndef YUNE_CAMERA_H
#define YUNE_CAMERA_H

#include "CUDA_headers.h"

namespace yune
{
    struct Vec4 { float x, y, z, w; };

    class Camera
    {
        public:
            Camera();                                   /**< FOV 60, eye at origin, looking down -Z (src/Camera.cpp:36-40, 93-103). */
            Camera(float y_FOV, float rot_speed = 0.25f, float mov_speed = 0.1f);

            /** Columns of view2world = (side, up, -look_at, eye) (src/Camera.cpp:105-117). Inputs are normalised. */
            void setViewMatrix(const Vec4& side, const Vec4& up, const Vec4& look_at, const Vec4& eye);
            /** Keyboard/mouse step (src/Camera.cpp:119-166): translate along the basis, pitch about `side`, yaw about +Y. */
            void setOrientation(const Vec4& dir, float pitch, float yaw);
            /** Write the kernel-side record: r_i = row i of view2world, view_plane_dist (src/Camera.cpp:66-82). */
            void setBuffer(Cam* cam_data);
            void resetCamera();
            void updateViewPlaneDist();                 /**< view_plane_dist = 1/tan(y_FOV*3.14/360), in double (src/Camera.cpp:60-64). */

            bool is_changed;
            Vec4 side, up, look_at, eye;
            float y_FOV, rotation_speed, move_speed;
            float view2world[4][4];                     /**< [column][row], like glm::mat4. */
        private:
            float view_plane_dist;
    };
}
#endif
