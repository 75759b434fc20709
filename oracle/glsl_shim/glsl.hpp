// oracle/glsl_shim/glsl.hpp -- TEST INFRASTRUCTURE ONLY.
//
// The subset of GLSL 4.50 that the reference's Forward+ fragment shader uses in its local-light code
// (shaders/vulkan/fp_stress_scene.frag:100-165, 421-539, 596-678 and shaders/vulkan/common/light_math.glsl), as C++ types and
// functions, so that the shader's OWN TEXT -- extracted mechanically by oracle/extract_glsl_a9.py, never copied into this
// repository -- compiles with g++ and pins oracle.cpp's restatement of row A9.
//
// GLSL leaves the precision of its built-ins implementation-defined; this shim evaluates every one of them in IEEE-754
// binary32 with the formula the GLSL specification gives (mix = x*(1-a) + y*a, smoothstep = t*t*(3 - 2t) with
// t = clamp((x-e0)/(e1-e0), 0, 1), normalize = v * (1/sqrt(dot)), length = sqrt(dot), dot left to right, pow = powf):
// one conforming evaluation, the same choices GLM's scalar path makes (oracle/glm_shim/glm/glm.hpp).  Swizzles are member
// functions (`.xyz` is rewritten to `.xyz()` by the extractor); single components are plain members with the colour aliases.
#pragma once

#include <cmath>
#include <cstdint>

namespace glsl
{
    typedef uint32_t uint;

    struct vec2
    {
        float x, y;
        vec2() : x(0), y(0) {}
        explicit vec2(float s) : x(s), y(s) {}
        vec2(float a, float b) : x(a), y(b) {}
    };

    struct vec3
    {
        union { float x; float r; };
        union { float y; float g; };
        union { float z; float b; };
        vec3() : x(0), y(0), z(0) {}
        explicit vec3(float s) : x(s), y(s), z(s) {}
        vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
        vec2 xy() const { return vec2(x, y); }
    };

    struct vec4
    {
        union { float x; float r; };
        union { float y; float g; };
        union { float z; float b; };
        union { float w; float a; };
        vec4() : x(0), y(0), z(0), w(0) {}
        vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
        vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
        vec3 xyz() const { return vec3(x, y, z); }
        vec3 rgb() const { return vec3(x, y, z); }
        vec2 xy() const { return vec2(x, y); }
    };

    struct uvec2
    {
        uint x, y;
        uvec2() : x(0), y(0) {}
        uvec2(uint a, uint b) : x(a), y(b) {}
        explicit uvec2(vec2 v) : x((uint)v.x), y((uint)v.y) {} // float -> uint conversion truncates toward zero
    };
    struct uvec4 { uint x, y, z, w; };
    struct uvec3 { uint x, y, z; };
    struct ivec2 // fp_stress_depth_reduce.comp: texel coordinates
    {
        int x, y;
        ivec2() : x(0), y(0) {}
        ivec2(int a, int b) : x(a), y(b) {}
        ivec2(uint a, uint b) : x((int)a), y((int)b) {} // uint -> int conversion keeps the bit pattern
    };

    struct mat4
    {
        vec4 c[4]; // column-major, like std140 / glm
    };

    // ---- operators (component-wise, binary32, in the order written)
    inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
    inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
    inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
    inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
    inline vec3 operator*(vec3 a, float k) { return vec3(a.x * k, a.y * k, a.z * k); }
    inline vec3 operator*(float k, vec3 a) { return vec3(k * a.x, k * a.y, k * a.z); }
    inline vec3 operator/(vec3 a, float k) { return vec3(a.x / k, a.y / k, a.z / k); }
    inline vec3 operator+(vec3 a, float k) { return vec3(a.x + k, a.y + k, a.z + k); }
    inline vec3 operator-(float k, vec3 a) { return vec3(k - a.x, k - a.y, k - a.z); }
    inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
    inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
    inline uvec2 operator/(uvec2 a, uint k) { return uvec2(a.x / k, a.y / k); }
    inline vec4 operator*(const mat4& m, vec4 v) // (m0*x + m1*y) + (m2*z + m3*w), the scalar order of glm's mat4 * vec4
    {
        return vec4((m.c[0].x * v.x + m.c[1].x * v.y) + (m.c[2].x * v.z + m.c[3].x * v.w), (m.c[0].y * v.x + m.c[1].y * v.y) + (m.c[2].y * v.z + m.c[3].y * v.w),
                    (m.c[0].z * v.x + m.c[1].z * v.y) + (m.c[2].z * v.z + m.c[3].z * v.w), (m.c[0].w * v.x + m.c[1].w * v.y) + (m.c[2].w * v.z + m.c[3].w * v.w));
    }

    // ---- built-ins (GLSL 4.50 specification, chapter 8)
    inline float max(float a, float b) { return (a < b) ? b : a; }
    inline float min(float a, float b) { return (b < a) ? b : a; }
    inline uint max(uint a, uint b) { return (a < b) ? b : a; }
    inline uint min(uint a, uint b) { return (b < a) ? b : a; }
    inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
    inline float abs(float x) { return std::fabs(x); }
    inline float pow(float x, float y) { return std::pow(x, y); }
    inline float log(float x) { return std::log(x); }
    inline float floor(float x) { return std::floor(x); }
    inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
    inline vec3 mix(vec3 x, vec3 y, float a) { return vec3(mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)); }
    inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
    inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
    inline float smoothstep(float e0, float e1, float x)
    {
        const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
        return t * t * (3.0f - 2.0f * t);
    }
}
