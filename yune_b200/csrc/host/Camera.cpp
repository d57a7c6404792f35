#include "Camera.h"
#include <cmath>

namespace yune
{
    namespace
    {
        Vec4 normalized(const Vec4& v)
        {
            float len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
            return Vec4{v.x / len, v.y / len, v.z / len, v.w / len};
        }
        void setColumn(float m[4][4], int c, const Vec4& v) { m[c][0] = v.x; m[c][1] = v.y; m[c][2] = v.z; m[c][3] = v.w; }
        Vec4 column(const float m[4][4], int c) { return Vec4{m[c][0], m[c][1], m[c][2], m[c][3]}; }
        // out = a * b for column-major 4x4 (same convention as glm: out[c] = sum_k a[k] * b[c][k])
        void mul(const float a[4][4], const float b[4][4], float out[4][4])
        {
            float t[4][4];
            for (int c = 0; c < 4; c++)
                for (int r = 0; r < 4; r++)
                    t[c][r] = a[0][r] * b[c][0] + a[1][r] * b[c][1] + a[2][r] * b[c][2] + a[3][r] * b[c][3];
            for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) out[c][r] = t[c][r];
        }
        // rotation by `angle` radians about the unit axis (ax, ay, az): Rodrigues, column-major
        void rotation(float angle, float ax, float ay, float az, float m[4][4])
        {
            float len = std::sqrt(ax * ax + ay * ay + az * az);
            ax /= len; ay /= len; az /= len;
            const float c = std::cos(angle), s = std::sin(angle), t = 1.0f - c;
            m[0][0] = c + t * ax * ax;      m[0][1] = t * ax * ay + s * az; m[0][2] = t * ax * az - s * ay; m[0][3] = 0;
            m[1][0] = t * ay * ax - s * az; m[1][1] = c + t * ay * ay;      m[1][2] = t * ay * az + s * ax; m[1][3] = 0;
            m[2][0] = t * az * ax + s * ay; m[2][1] = t * az * ay - s * ax; m[2][2] = c + t * az * az;      m[2][3] = 0;
            m[3][0] = 0; m[3][1] = 0; m[3][2] = 0; m[3][3] = 1;
        }
    }

    Camera::Camera() : y_FOV(60.0f), rotation_speed(0.25f), move_speed(0.1f) { resetCamera(); }

    Camera::Camera(float fov, float rot_speed, float mov_speed) : y_FOV(fov), rotation_speed(rot_speed), move_speed(mov_speed)
    {
        setViewMatrix(Vec4{1, 0, 0, 0}, Vec4{0, 1, 0, 0}, Vec4{0, 0, -1, 0}, Vec4{0, 0, 0, 1});
        updateViewPlaneDist();
    }

    void Camera::updateViewPlaneDist()
    {
        view_plane_dist = 1 / tan(y_FOV * 3.14 / 360);      // double arithmetic and the reference's 3.14 (src/Camera.cpp:62)
        is_changed = true;
    }

    void Camera::resetCamera()
    {
        setViewMatrix(Vec4{1, 0, 0, 0}, Vec4{0, 1, 0, 0}, Vec4{0, 0, -1, 0}, Vec4{0, 0, 0, 1});
        y_FOV = 60;
        updateViewPlaneDist();
    }

    void Camera::setViewMatrix(const Vec4& s, const Vec4& u, const Vec4& l, const Vec4& e)
    {
        eye = e; side = normalized(s); up = normalized(u); look_at = normalized(l);
        // the reference stores the UN-normalised arguments in the matrix (src/Camera.cpp:115)
        setColumn(view2world, 0, s);
        setColumn(view2world, 1, u);
        setColumn(view2world, 2, Vec4{-l.x, -l.y, -l.z, -l.w});
        setColumn(view2world, 3, e);
        is_changed = true;
    }

    void Camera::setBuffer(Cam* cam)
    {
        for (int c = 0; c < 4; c++) {
            cam->r1.s[c] = view2world[c][0];
            cam->r2.s[c] = view2world[c][1];
            cam->r3.s[c] = view2world[c][2];
            cam->r4.s[c] = view2world[c][3];
        }
        cam->view_plane_dist = view_plane_dist;
        cam->pad[0] = cam->pad[1] = cam->pad[2] = 0.0f;
        is_changed = false;
    }

    void Camera::setOrientation(const Vec4& dir, float pitch, float yaw)
    {
        // translation first: one axis per call, priority z, x, y (src/Camera.cpp:121-135)
        auto move = [&](const Vec4& axis, float sgn) {
            eye.x += sgn * axis.x * move_speed; eye.y += sgn * axis.y * move_speed;
            eye.z += sgn * axis.z * move_speed; eye.w += sgn * axis.w * move_speed;
        };
        if (dir.z > 0) move(look_at, 1); else if (dir.z < 0) move(look_at, -1);
        else if (dir.x > 0) move(side, 1); else if (dir.x < 0) move(side, -1);
        else if (dir.y > 0) move(up, 1); else if (dir.y < 0) move(up, -1);
        if (pitch == 0 && yaw == 0) { setColumn(view2world, 3, eye); is_changed = true; return; }

        float rotx[4][4], roty[4][4];
        rotation(pitch * rotation_speed, side.x, side.y, side.z, rotx);
        rotation(yaw * rotation_speed, 0, 1, 0, roty);
        // pitch is applied only while the rotated up vector keeps pointing up (:147-155)
        const float up_y = rotx[0][1] * up.x + rotx[1][1] * up.y + rotx[2][1] * up.z + rotx[3][1] * up.w;
        setColumn(view2world, 3, Vec4{0, 0, 0, 1});
        if (up_y >= 0) mul(rotx, view2world, view2world);
        mul(roty, view2world, view2world);
        setColumn(view2world, 3, eye);
        side = normalized(column(view2world, 0));
        up = normalized(column(view2world, 1));
        Vec4 z = normalized(column(view2world, 2));
        look_at = Vec4{-z.x, -z.y, -z.z, -z.w};
        is_changed = true;
    }
}
