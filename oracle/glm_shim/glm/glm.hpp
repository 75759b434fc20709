// oracle/glm_shim/glm/glm.hpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Minimal GLM-compatible shim, written from scratch for this repository, so that the
// reference's own raster-path headers (which need nothing but std + GLM) compile in a
// container that has no GLM.  It restates the *scalar* code path of
// GLM 0.9.9 / 1.0 (the reference never defines GLM_FORCE_INTRINSICS / GLM_FORCE_SIMD_*),
// i.e. the published operation ORDER of each function, because with -ffp-contract=off the
// order of IEEE-754 binary32 operations is what decides every output bit:
//
//   mat4*vec4      (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w)
//   mat3*vec3      m[0][r]*v.x + m[1][r]*v.y + m[2][r]*v.z              (left to right)
//   dot(vec2/3)    x*x' + y*y' (+ z*z')                                 (left to right)
//   dot(vec4)      (x*x' + y*y') + (z*z' + w*w')
//   normalize      v * (1/sqrt(dot(v,v)))
//   length         sqrt(dot(v,v))
//   mix(a,b,t)     a*(1-t) + b*t
//   clamp(x,l,h)   min(max(x,l),h);  max(a,b) = (a<b)?b:a;  min(a,b) = (b<a)?b:a
//   reflect(I,N)   I - N*dot(N,I)*2
//   inverse/determinant: cofactor expansion, multiply by 1/det
//
// GLM itself is a vcpkg dependency of the reference with NO pinned version
// (reference README.md:154, shs-renderer-lib/CMakeLists.txt:87); its source is not in this
// container, so these orders are stated from GLM's published scalar implementation and
// could not be diffed here.  Only the ~28 symbols the hot-path headers use are provided.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

namespace glm
{
    typedef int length_t;

    // ------------------------------------------------------------------ vectors
    struct vec2
    {
        union { float x, r, s; };
        union { float y, g, t; };
        vec2() = default;
        explicit vec2(float v) : x(v), y(v) {}
        vec2(float X, float Y) : x(X), y(Y) {}
        template <typename A, typename B> vec2(A X, B Y) : x((float)X), y((float)Y) {}
        explicit vec2(const struct vec3& v);
        explicit vec2(const struct vec4& v);
        float& operator[](length_t i) { return (&x)[i]; }
        const float& operator[](length_t i) const { return (&x)[i]; }
        vec2& operator+=(const vec2& o) { x += o.x; y += o.y; return *this; }
        vec2& operator-=(const vec2& o) { x -= o.x; y -= o.y; return *this; }
        vec2& operator*=(float k) { x *= k; y *= k; return *this; }
        vec2& operator*=(const vec2& o) { x *= o.x; y *= o.y; return *this; }
        vec2& operator/=(float k) { x /= k; y /= k; return *this; }
    };

    struct ivec2 // what sw_render/debug_draw.hpp's wireframe draw and the legacy demos' job-tile bounds store: two ints
    {
        int x, y;
        ivec2() : x(0), y(0) {}
        ivec2(int X, int Y) : x(X), y(Y) {}
        explicit operator vec2() const { return vec2((float)x, (float)y); } // glm::vec2(ivec2): int -> float per component
    };

    struct vec3
    {
        union { float x, r, s; };
        union { float y, g, t; };
        union { float z, b, p; };
        vec3() = default;
        explicit vec3(float v) : x(v), y(v), z(v) {}
        vec3(float X, float Y, float Z) : x(X), y(Y), z(Z) {}
        template <typename A, typename B, typename C> vec3(A X, B Y, C Z) : x((float)X), y((float)Y), z((float)Z) {}
        vec3(const vec2& v, float Z) : x(v.x), y(v.y), z(Z) {}
        explicit vec3(const struct vec4& v);
        float& operator[](length_t i) { return (&x)[i]; }
        const float& operator[](length_t i) const { return (&x)[i]; }
        vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
        vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
        vec3& operator*=(float k) { x *= k; y *= k; z *= k; return *this; }
        vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
        vec3& operator/=(float k) { x /= k; y /= k; z /= k; return *this; }
    };

    struct vec4
    {
        union { float x, r, s; };
        union { float y, g, t; };
        union { float z, b, p; };
        union { float w, a, q; };
        vec4() = default;
        explicit vec4(float v) : x(v), y(v), z(v), w(v) {}
        vec4(float X, float Y, float Z, float W) : x(X), y(Y), z(Z), w(W) {}
        template <typename A, typename B, typename C, typename D>
        vec4(A X, B Y, C Z, D W) : x((float)X), y((float)Y), z((float)Z), w((float)W) {}
        vec4(const vec3& v, float W) : x(v.x), y(v.y), z(v.z), w(W) {}
        vec4(const vec2& v, float Z, float W) : x(v.x), y(v.y), z(Z), w(W) {}
        float& operator[](length_t i) { return (&x)[i]; }
        const float& operator[](length_t i) const { return (&x)[i]; }
        vec4& operator+=(const vec4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
        vec4& operator-=(const vec4& o) { x -= o.x; y -= o.y; z -= o.z; w -= o.w; return *this; }
        vec4& operator*=(float k) { x *= k; y *= k; z *= k; w *= k; return *this; }
        vec4& operator/=(float k) { x /= k; y /= k; z /= k; w /= k; return *this; }
    };

    struct uvec4
    {
        uint32_t x, y, z, w;
        uvec4() = default;
        explicit uvec4(uint32_t v) : x(v), y(v), z(v), w(v) {}
        uvec4(uint32_t X, uint32_t Y, uint32_t Z, uint32_t W) : x(X), y(Y), z(Z), w(W) {}
        uint32_t& operator[](length_t i) { return (&x)[i]; }
        const uint32_t& operator[](length_t i) const { return (&x)[i]; }
    };

    inline vec2::vec2(const vec3& v) : x(v.x), y(v.y) {}
    inline vec2::vec2(const vec4& v) : x(v.x), y(v.y) {}
    inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

    // vec2 operators
    inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
    inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
    inline vec2 operator*(const vec2& a, const vec2& b) { return vec2(a.x * b.x, a.y * b.y); }
    inline vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
    inline vec2 operator*(const vec2& a, float k) { return vec2(a.x * k, a.y * k); }
    inline vec2 operator*(float k, const vec2& a) { return vec2(k * a.x, k * a.y); }
    inline vec2 operator/(const vec2& a, float k) { return vec2(a.x / k, a.y / k); }
    inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }

    // vec3 operators
    inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
    inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
    inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
    inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
    inline vec3 operator*(const vec3& a, float k) { return vec3(a.x * k, a.y * k, a.z * k); }
    inline vec3 operator*(float k, const vec3& a) { return vec3(k * a.x, k * a.y, k * a.z); }
    inline vec3 operator/(const vec3& a, float k) { return vec3(a.x / k, a.y / k, a.z / k); }
    inline vec3 operator+(const vec3& a, float k) { return vec3(a.x + k, a.y + k, a.z + k); }
    inline vec3 operator-(const vec3& a, float k) { return vec3(a.x - k, a.y - k, a.z - k); }
    inline vec3 operator-(float k, const vec3& a) { return vec3(k - a.x, k - a.y, k - a.z); }
    inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }

    // vec4 operators
    inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
    inline vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
    inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
    inline vec4 operator*(const vec4& a, float k) { return vec4(a.x * k, a.y * k, a.z * k, a.w * k); }
    inline vec4 operator*(float k, const vec4& a) { return vec4(k * a.x, k * a.y, k * a.z, k * a.w); }
    inline vec4 operator/(const vec4& a, float k) { return vec4(a.x / k, a.y / k, a.z / k, a.w / k); }
    inline vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

    // ------------------------------------------------------------------ scalar helpers
    inline float min(float a, float b) { return (b < a) ? b : a; }
    inline float max(float a, float b) { return (a < b) ? b : a; }
    inline int min(int a, int b) { return (b < a) ? b : a; }
    inline int max(int a, int b) { return (a < b) ? b : a; }
    inline float abs(float a) { return std::fabs(a); }
    inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
    inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
    inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
    inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

    template <typename T> inline T pi() { return T(3.14159265358979323846264338327950288); }
    template <typename T> inline T half_pi() { return T(1.57079632679489661923132169163975144); }

    // ------------------------------------------------------------------ vector functions
    inline vec2 min(const vec2& a, const vec2& b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
    inline vec2 max(const vec2& a, const vec2& b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
    inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
    inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
    inline vec4 min(const vec4& a, const vec4& b) { return vec4(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w)); }
    inline vec4 max(const vec4& a, const vec4& b) { return vec4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }
    inline vec3 abs(const vec3& a) { return vec3(abs(a.x), abs(a.y), abs(a.z)); }
    inline vec3 clamp(const vec3& v, const vec3& lo, const vec3& hi) { return min(max(v, lo), hi); }
    inline vec3 clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }

    inline float dot(const vec2& a, const vec2& b) { const vec2 t(a * b); return t.x + t.y; }
    inline float dot(const vec3& a, const vec3& b) { const vec3 t(a * b); return t.x + t.y + t.z; }
    inline float dot(const vec4& a, const vec4& b) { const vec4 t(a * b); return (t.x + t.y) + (t.z + t.w); }

    inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
    inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
    inline float length(const vec4& v) { return std::sqrt(dot(v, v)); }

    inline vec2 normalize(const vec2& v) { return v * inversesqrt(dot(v, v)); }
    inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
    inline vec4 normalize(const vec4& v) { return v * inversesqrt(dot(v, v)); }

    inline vec3 cross(const vec3& x, const vec3& y)
    {
        return vec3(
            x.y * y.z - y.y * x.z,
            x.z * y.x - y.z * x.x,
            x.x * y.y - y.x * x.y);
    }

    inline vec2 mix(const vec2& a, const vec2& b, float t) { return a * (1.0f - t) + b * t; }
    inline vec3 mix(const vec3& a, const vec3& b, float t) { return a * (1.0f - t) + b * t; }
    inline vec4 mix(const vec4& a, const vec4& b, float t) { return a * (1.0f - t) + b * t; }

    inline vec3 reflect(const vec3& I, const vec3& N) { return I - N * dot(N, I) * 2.0f; }

    // ------------------------------------------------------------------ matrices (column-major)
    struct mat3
    {
        vec3 c[3];
        mat3() = default;
        explicit mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
        mat3(const vec3& a, const vec3& b, const vec3& cc) { c[0] = a; c[1] = b; c[2] = cc; }
        explicit mat3(const struct mat4& m);
        vec3& operator[](length_t i) { return c[i]; }
        const vec3& operator[](length_t i) const { return c[i]; }
    };

    struct mat4
    {
        vec4 c[4];
        mat4() = default;
        explicit mat4(float d)
        {
            c[0] = vec4(d, 0, 0, 0); c[1] = vec4(0, d, 0, 0); c[2] = vec4(0, 0, d, 0); c[3] = vec4(0, 0, 0, d);
        }
        mat4(const vec4& a, const vec4& b, const vec4& cc, const vec4& d) { c[0] = a; c[1] = b; c[2] = cc; c[3] = d; }
        vec4& operator[](length_t i) { return c[i]; }
        const vec4& operator[](length_t i) const { return c[i]; }
    };

    inline mat3::mat3(const mat4& m) { c[0] = vec3(m[0]); c[1] = vec3(m[1]); c[2] = vec3(m[2]); }

    inline vec3 operator*(const mat3& m, const vec3& v)
    {
        return vec3(
            m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
            m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
            m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
    }

    inline vec4 operator*(const mat4& m, const vec4& v)
    {
        const vec4 Mul0 = m[0] * vec4(v.x);
        const vec4 Mul1 = m[1] * vec4(v.y);
        const vec4 Add0 = Mul0 + Mul1;
        const vec4 Mul2 = m[2] * vec4(v.z);
        const vec4 Mul3 = m[3] * vec4(v.w);
        const vec4 Add1 = Mul2 + Mul3;
        return Add0 + Add1;
    }

    inline mat4 operator*(const mat4& a, const mat4& b)
    {
        mat4 r;
        for (int i = 0; i < 4; ++i)
            r[i] = a[0] * b[i][0] + a[1] * b[i][1] + a[2] * b[i][2] + a[3] * b[i][3];
        return r;
    }

    inline mat4 operator*(const mat4& m, float k) { return mat4(m[0] * k, m[1] * k, m[2] * k, m[3] * k); }

    inline mat3 transpose(const mat3& m)
    {
        mat3 r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = m[j][i];
        return r;
    }

    inline mat4 transpose(const mat4& m)
    {
        mat4 r;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[i][j] = m[j][i];
        return r;
    }

    inline float determinant(const mat3& m)
    {
        return
            + m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
            - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
            + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
    }

    inline mat3 inverse(const mat3& m)
    {
        const float OneOverDeterminant = 1.0f / (
            + m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
            - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
            + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
        mat3 Inverse;
        Inverse[0][0] = + (m[1][1] * m[2][2] - m[2][1] * m[1][2]) * OneOverDeterminant;
        Inverse[1][0] = - (m[1][0] * m[2][2] - m[2][0] * m[1][2]) * OneOverDeterminant;
        Inverse[2][0] = + (m[1][0] * m[2][1] - m[2][0] * m[1][1]) * OneOverDeterminant;
        Inverse[0][1] = - (m[0][1] * m[2][2] - m[2][1] * m[0][2]) * OneOverDeterminant;
        Inverse[1][1] = + (m[0][0] * m[2][2] - m[2][0] * m[0][2]) * OneOverDeterminant;
        Inverse[2][1] = - (m[0][0] * m[2][1] - m[2][0] * m[0][1]) * OneOverDeterminant;
        Inverse[0][2] = + (m[0][1] * m[1][2] - m[1][1] * m[0][2]) * OneOverDeterminant;
        Inverse[1][2] = - (m[0][0] * m[1][2] - m[1][0] * m[0][2]) * OneOverDeterminant;
        Inverse[2][2] = + (m[0][0] * m[1][1] - m[1][0] * m[0][1]) * OneOverDeterminant;
        return Inverse;
    }

    inline float determinant(const mat4& m)
    {
        const float SubFactor00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
        const float SubFactor01 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
        const float SubFactor02 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
        const float SubFactor03 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
        const float SubFactor04 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
        const float SubFactor05 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
        const vec4 DetCof(
            + (m[1][1] * SubFactor00 - m[1][2] * SubFactor01 + m[1][3] * SubFactor02),
            - (m[1][0] * SubFactor00 - m[1][2] * SubFactor03 + m[1][3] * SubFactor04),
            + (m[1][0] * SubFactor01 - m[1][1] * SubFactor03 + m[1][3] * SubFactor05),
            - (m[1][0] * SubFactor02 - m[1][1] * SubFactor04 + m[1][2] * SubFactor05));
        return
            m[0][0] * DetCof[0] + m[0][1] * DetCof[1] +
            m[0][2] * DetCof[2] + m[0][3] * DetCof[3];
    }

    inline mat4 inverse(const mat4& m)
    {
        const float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
        const float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
        const float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
        const float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
        const float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
        const float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
        const float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
        const float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
        const float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
        const float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
        const float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
        const float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
        const float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
        const float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
        const float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
        const float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
        const float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
        const float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];

        const vec4 Fac0(Coef00, Coef00, Coef02, Coef03);
        const vec4 Fac1(Coef04, Coef04, Coef06, Coef07);
        const vec4 Fac2(Coef08, Coef08, Coef10, Coef11);
        const vec4 Fac3(Coef12, Coef12, Coef14, Coef15);
        const vec4 Fac4(Coef16, Coef16, Coef18, Coef19);
        const vec4 Fac5(Coef20, Coef20, Coef22, Coef23);

        const vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
        const vec4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
        const vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
        const vec4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);

        const vec4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
        const vec4 Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
        const vec4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
        const vec4 Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);

        const vec4 SignA(+1.0f, -1.0f, +1.0f, -1.0f);
        const vec4 SignB(-1.0f, +1.0f, -1.0f, +1.0f);
        const mat4 Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);

        const vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
        const vec4 Dot0(m[0] * Row0);
        const float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
        const float OneOverDeterminant = 1.0f / Dot1;
        return Inverse * OneOverDeterminant;
    }
}
