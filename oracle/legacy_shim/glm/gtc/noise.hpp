// oracle/legacy_shim/glm/gtc/noise.hpp -- TEST INFRASTRUCTURE.  The few GLM names the reference's legacy header
// (hello-shs-renderer/shs_renderer.hpp) uses beyond what oracle/glm_shim/glm/glm.hpp already states.  Only `pow` is on the pinned
// path (blinn_phong_fragment_shader, hello_pipeline_blinn_phong_shading.cpp:84: GLM's scalar pow is std::pow); `perspectiveLH`
// is GLM's default-clip-space name for perspectiveLH_NO (the legacy sources never define GLM_FORCE_DEPTH_ZERO_TO_ONE); simplex /
// fract / sin / smoothstep / mat2 are used by sky and noise helpers the harness never calls and are only declared or stated trivially.
#pragma once
#include <cmath>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
namespace glm
{
    inline float pow(float b, float e) { return std::pow(b, e); }
    inline vec3 pow(const vec3& b, const vec3& e) { return vec3(std::pow(b.x, e.x), std::pow(b.y, e.y), std::pow(b.z, e.z)); }
    inline float sin(float x) { return std::sin(x); }
    // ivec2: oracle/glm_shim/glm/glm.hpp
    inline float fract(float x) { return x - std::floor(x); }
    inline float smoothstep(float e0, float e1, float x) { const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
    float simplex(const vec2& p); // declared only: never called by the harness
    struct mat2
    {
        vec2 c[2];
        mat2(float a, float b, float cc, float d) { c[0] = vec2(a, b); c[1] = vec2(cc, d); }
    };
    inline vec2 operator*(const mat2& m, const vec2& v) { return vec2(m.c[0].x * v.x + m.c[1].x * v.y, m.c[0].y * v.x + m.c[1].y * v.y); }
    inline mat4 perspectiveLH(float fovy, float aspect, float zn, float zf) { return perspectiveLH_NO(fovy, aspect, zn, zf); }
}
