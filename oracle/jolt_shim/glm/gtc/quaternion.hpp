// oracle/jolt_shim/glm/gtc/quaternion.hpp -- TEST INFRASTRUCTURE ONLY.  shs/geometry/jolt_adapter.hpp and shs/scene/scene_elements.hpp
// include GLM's quaternion header; the only use on any include path the checkers compile is one line of scene_elements.hpp
// (a matrix -> Euler-angle decomposition that no checker calls).  The three functions below exist so that line compiles and links;
// they follow GLM's published formulas but NOTHING PINNED DEPENDS ON THEM.
#pragma once
#include <cmath>
#include <glm/glm.hpp>

namespace glm
{
    struct quat { float w = 1, x = 0, y = 0, z = 0; };

    inline quat quat_cast(const mat3& m)
    {
        const float fx = m[0][0] - m[1][1] - m[2][2], fy = m[1][1] - m[0][0] - m[2][2], fz = m[2][2] - m[0][0] - m[1][1], fw = m[0][0] + m[1][1] + m[2][2];
        int big = 0;
        float best = fw;
        if (fx > best) { best = fx; big = 1; }
        if (fy > best) { best = fy; big = 2; }
        if (fz > best) { best = fz; big = 3; }
        const float v = std::sqrt(best + 1.0f) * 0.5f, mult = 0.25f / v;
        quat q;
        switch (big)
        {
        case 0: q.w = v; q.x = (m[1][2] - m[2][1]) * mult; q.y = (m[2][0] - m[0][2]) * mult; q.z = (m[0][1] - m[1][0]) * mult; break;
        case 1: q.w = (m[1][2] - m[2][1]) * mult; q.x = v; q.y = (m[0][1] + m[1][0]) * mult; q.z = (m[2][0] + m[0][2]) * mult; break;
        case 2: q.w = (m[2][0] - m[0][2]) * mult; q.x = (m[0][1] + m[1][0]) * mult; q.y = v; q.z = (m[1][2] + m[2][1]) * mult; break;
        default: q.w = (m[0][1] - m[1][0]) * mult; q.x = (m[2][0] + m[0][2]) * mult; q.y = (m[1][2] + m[2][1]) * mult; q.z = v; break;
        }
        return q;
    }

    inline quat normalize(const quat& q)
    {
        const float len = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
        if (len <= 0.0f) return quat{};
        const float k = 1.0f / len;
        return quat{q.w * k, q.x * k, q.y * k, q.z * k};
    }

    inline vec3 eulerAngles(const quat& q) // (pitch, yaw, roll)
    {
        const float py = 2.0f * (q.y * q.z + q.w * q.x), px = q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z;
        const float pitch = (py == 0.0f && px == 0.0f) ? 2.0f * std::atan2(q.x, q.w) : std::atan2(py, px);
        float s = -2.0f * (q.x * q.z - q.w * q.y);
        s = s < -1.0f ? -1.0f : (s > 1.0f ? 1.0f : s);
        const float yaw = std::asin(s);
        const float ry = 2.0f * (q.x * q.y + q.w * q.z), rx = q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z;
        const float roll = (ry == 0.0f && rx == 0.0f) ? 0.0f : std::atan2(ry, rx);
        return vec3(pitch, yaw, roll);
    }
}
