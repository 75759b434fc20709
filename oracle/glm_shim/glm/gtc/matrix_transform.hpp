// oracle/glm_shim/glm/gtc/matrix_transform.hpp -- TEST INFRASTRUCTURE ONLY.
// Restates the operation order of GLM's ext/matrix_transform.inl and ext/matrix_clip_space.inl
// (scalar path) for the six builders the reference's hot path uses; see ../glm.hpp for scope.
#pragma once
#include "../glm.hpp"

namespace glm
{
    inline mat4 translate(const mat4& m, const vec3& v)
    {
        mat4 Result(m);
        Result[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
        return Result;
    }

    inline mat4 rotate(const mat4& m, float angle, const vec3& v)
    {
        const float a = angle;
        const float c = std::cos(a);
        const float s = std::sin(a);

        const vec3 axis(normalize(v));
        const vec3 temp((1.0f - c) * axis);

        mat4 Rotate;
        Rotate[0][0] = c + temp[0] * axis[0];
        Rotate[0][1] = temp[0] * axis[1] + s * axis[2];
        Rotate[0][2] = temp[0] * axis[2] - s * axis[1];

        Rotate[1][0] = temp[1] * axis[0] - s * axis[2];
        Rotate[1][1] = c + temp[1] * axis[1];
        Rotate[1][2] = temp[1] * axis[2] + s * axis[0];

        Rotate[2][0] = temp[2] * axis[0] + s * axis[1];
        Rotate[2][1] = temp[2] * axis[1] - s * axis[0];
        Rotate[2][2] = c + temp[2] * axis[2];

        mat4 Result;
        Result[0] = m[0] * Rotate[0][0] + m[1] * Rotate[0][1] + m[2] * Rotate[0][2];
        Result[1] = m[0] * Rotate[1][0] + m[1] * Rotate[1][1] + m[2] * Rotate[1][2];
        Result[2] = m[0] * Rotate[2][0] + m[1] * Rotate[2][1] + m[2] * Rotate[2][2];
        Result[3] = m[3];
        return Result;
    }

    inline mat4 scale(const mat4& m, const vec3& v)
    {
        mat4 Result;
        Result[0] = m[0] * v[0];
        Result[1] = m[1] * v[1];
        Result[2] = m[2] * v[2];
        Result[3] = m[3];
        return Result;
    }

    inline mat4 lookAtLH(const vec3& eye, const vec3& center, const vec3& up)
    {
        const vec3 f(normalize(center - eye));
        const vec3 s(normalize(cross(up, f)));
        const vec3 u(cross(f, s));

        mat4 Result(1.0f);
        Result[0][0] = s.x;
        Result[1][0] = s.y;
        Result[2][0] = s.z;
        Result[0][1] = u.x;
        Result[1][1] = u.y;
        Result[2][1] = u.z;
        Result[0][2] = f.x;
        Result[1][2] = f.y;
        Result[2][2] = f.z;
        Result[3][0] = -dot(s, eye);
        Result[3][1] = -dot(u, eye);
        Result[3][2] = -dot(f, eye);
        return Result;
    }

    inline mat4 perspectiveLH_NO(float fovy, float aspect, float zNear, float zFar)
    {
        const float tanHalfFovy = std::tan(fovy / 2.0f);

        mat4 Result(0.0f);
        Result[0][0] = 1.0f / (aspect * tanHalfFovy);
        Result[1][1] = 1.0f / (tanHalfFovy);
        Result[2][2] = (zFar + zNear) / (zFar - zNear);
        Result[2][3] = 1.0f;
        Result[3][2] = -(2.0f * zFar * zNear) / (zFar - zNear);
        return Result;
    }

    inline mat4 orthoLH_NO(float left, float right, float bottom, float top, float zNear, float zFar)
    {
        mat4 Result(1.0f);
        Result[0][0] = 2.0f / (right - left);
        Result[1][1] = 2.0f / (top - bottom);
        Result[2][2] = 2.0f / (zFar - zNear);
        Result[3][0] = -(right + left) / (right - left);
        Result[3][1] = -(top + bottom) / (top - bottom);
        Result[3][2] = -(zFar + zNear) / (zFar - zNear);
        return Result;
    }
}
