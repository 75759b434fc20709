// oracle/jolt_shim/Jolt/Jolt.h -- TEST INFRASTRUCTURE ONLY.
//
// Declaration shim for JoltPhysics (the reference pins v5.2.0 through vcpkg; the library is NOT in /root/reference and not in this
// image).  It exists so that the reference's OWN light-culling headers -- lighting/jolt_light_culling.hpp, geometry/jolt_culling.hpp,
// geometry/scene_shape.hpp, geometry/jolt_adapter.hpp, which are guarded by SHS_HAS_JOLT -- compile where they lie and can be run
// against the restatement (oracle/ref_lightcull_harness.cpp).  Those headers touch Jolt only to fetch a light's bounds:
//     shape->GetWorldSpaceBounds(transform, Vec3::sReplicate(1))  ->  AABox;   AABox::GetCenter / GetExtent;   Vec3::Length
// plus value types for the LH <-> RH conversions.  Everything after that (tile cells, plane extraction, sphere / p-vertex
// classification, list order) is the reference's own glm code.  Stated here, from Jolt's published headers:
//     AABox::GetCenter() = 0.5f * (mMin + mMax)           (Jolt/Geometry/AABox.h)
//     AABox::GetExtent() = 0.5f * (mMax - mMin)
//     Vec3::Length()     = sqrt(x*x + y*y + z*z)          (Jolt/Math/Vec3.inl; its SSE path may sum in another order -- the harness
//                                                          therefore RETURNS the bounds the reference derived, and the restatement
//                                                          is fed exactly those, so the pin does not depend on this line)
// Shapes: only an abstract Shape with GetWorldSpaceBounds; the harness supplies a bounds-carrying subclass.
#pragma once
#include <cmath>

namespace JPH
{
    class Vec3
    {
    public:
        Vec3() = default;
        Vec3(float x, float y, float z) : x_(x), y_(y), z_(z) {}
        static Vec3 sReplicate(float v) { return Vec3(v, v, v); }
        static Vec3 sZero() { return Vec3(0, 0, 0); }
        float GetX() const { return x_; }
        float GetY() const { return y_; }
        float GetZ() const { return z_; }
        Vec3 operator+(const Vec3& o) const { return Vec3(x_ + o.x_, y_ + o.y_, z_ + o.z_); }
        Vec3 operator-(const Vec3& o) const { return Vec3(x_ - o.x_, y_ - o.y_, z_ - o.z_); }
        Vec3 operator*(float k) const { return Vec3(x_ * k, y_ * k, z_ * k); }
        friend Vec3 operator*(float k, const Vec3& v) { return Vec3(k * v.x_, k * v.y_, k * v.z_); }
        float Dot(const Vec3& o) const { return x_ * o.x_ + y_ * o.y_ + z_ * o.z_; }
        float Length() const { return std::sqrt(Dot(*this)); }
    private:
        float x_ = 0, y_ = 0, z_ = 0;
    };

    class Vec4
    {
    public:
        Vec4() = default;
        Vec4(float x, float y, float z, float w) : x_(x), y_(y), z_(z), w_(w) {}
        float GetX() const { return x_; }
        float GetY() const { return y_; }
        float GetZ() const { return z_; }
        float GetW() const { return w_; }
    private:
        float x_ = 0, y_ = 0, z_ = 0, w_ = 0;
    };

    class Mat44
    {
    public:
        Mat44() = default;
        Mat44(const Vec4& c0, const Vec4& c1, const Vec4& c2, const Vec4& c3) : c_{c0, c1, c2, c3} {}
        static Mat44 sIdentity() { return Mat44(Vec4(1, 0, 0, 0), Vec4(0, 1, 0, 0), Vec4(0, 0, 1, 0), Vec4(0, 0, 0, 1)); }
        Vec4 GetColumn4(unsigned i) const { return c_[i]; }
        Vec3 GetTranslation() const { return Vec3(c_[3].GetX(), c_[3].GetY(), c_[3].GetZ()); }
        class Quat GetQuaternion() const; // declared for jolt_debug_draw.hpp, never called (see Shape::GetTriangles*)
    private:
        Vec4 c_[4];
    };

    class Plane
    {
    public:
        Plane() = default;
        Plane(const Vec3& n, float c) : n_(n), c_(c) {}
        Vec3 GetNormal() const { return n_; }
        float GetConstant() const { return c_; }
    private:
        Vec3 n_;
        float c_ = 0;
    };

    class AABox
    {
    public:
        AABox() = default;
        AABox(const Vec3& mn, const Vec3& mx) : mMin(mn), mMax(mx) {}
        Vec3 GetCenter() const { return 0.5f * (mMin + mMax); }
        Vec3 GetExtent() const { return 0.5f * (mMax - mMin); }
        Vec3 mMin, mMax;
    };

    // geometry/jolt_debug_draw.hpp (included by geometry/culling_software.hpp for its DebugMesh type) walks a shape's triangles through
    // Shape::GetTrianglesStart / GetTrianglesNext.  Declarations only: the occlusion harness hands the reference ready-made DebugMesh
    // data (vertices + indices are the INPUT of rasterize_mesh_depth_transformed), so these are never called.
    struct Float3 { float x, y, z; Float3() = default; Float3(float a, float b, float c) : x(a), y(b), z(c) {} };
    class Quat
    {
    public:
        Quat() = default;
        Quat(float x, float y, float z, float w) : x_(x), y_(y), z_(z), w_(w) {}
        static Quat sIdentity() { return Quat(0, 0, 0, 1); }
        float GetX() const { return x_; }
        float GetY() const { return y_; }
        float GetZ() const { return z_; }
        float GetW() const { return w_; }
    private:
        float x_ = 0, y_ = 0, z_ = 0, w_ = 1;
    };
    class PhysicsMaterial;

    class Shape
    {
    public:
        virtual ~Shape() = default;
        virtual AABox GetWorldSpaceBounds(const Mat44& center_of_mass_transform, const Vec3& scale) const = 0;
        struct GetTrianglesContext { unsigned char opaque[4288]; };
        virtual void GetTrianglesStart(GetTrianglesContext&, const AABox&, const Vec3&, const Quat&, const Vec3&) const {}
        virtual int GetTrianglesNext(GetTrianglesContext&, int, Float3*, const PhysicsMaterial** = nullptr) const { return 0; }
    };

    // RefConst<Shape>: a non-owning stand-in (the harness keeps the shapes alive)
    class ShapeRefC
    {
    public:
        ShapeRefC() = default;
        ShapeRefC(const Shape* s) : s_(s) {}
        const Shape* operator->() const { return s_; }
        const Shape& operator*() const { return *s_; }
        const Shape* GetPtr() const { return s_; }
        explicit operator bool() const { return s_ != nullptr; }
        bool operator!() const { return s_ == nullptr; }
        bool operator==(std::nullptr_t) const { return s_ == nullptr; }
        bool operator!=(std::nullptr_t) const { return s_ != nullptr; }
    private:
        const Shape* s_ = nullptr;
    };

    class Factory
    {
    public:
        static inline Factory* sInstance = nullptr;
    };
    inline void RegisterDefaultAllocator() {}
    inline void RegisterTypes() {}
    inline void UnregisterTypes() {}
}
