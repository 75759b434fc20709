// oracle/jolt_shim/Jolt/shapes_decl.h -- TEST INFRASTRUCTURE ONLY.  Declarations that let shs/geometry/jolt_shapes.hpp (the reference's
// shape factory helpers) COMPILE against the shim of Jolt/Jolt.h.  The primitive shapes report their local bounds for an identity
// transform only (sphere +-r, box +-half extents, capsule / cylinder along y); none of the checkers in oracle/ builds a light or an
// object through them -- the light-list harness feeds bounds directly (ref_lightcull_harness.cpp) -- so nothing pinned depends on
// these bodies.  Hull / mesh / tapered-capsule creation always reports an error, which takes the reference's documented fallbacks.
#pragma once
#include <Jolt/Jolt.h>
#include <vector>

namespace JPH
{
    // Float3 lives in Jolt/Jolt.h (geometry/jolt_debug_draw.hpp needs it as well)
    struct Triangle { Float3 mV[3]; Triangle() = default; Triangle(const Float3& a, const Float3& b, const Float3& c) : mV{a, b, c} {} Triangle(const Vec3& a, const Vec3& b, const Vec3& c) : mV{Float3(a.GetX(), a.GetY(), a.GetZ()), Float3(b.GetX(), b.GetY(), b.GetZ()), Float3(c.GetX(), c.GetY(), c.GetZ())} {} };
    using TriangleList = std::vector<Triangle>;

    class LocalBoundsShape : public Shape
    {
    public:
        explicit LocalBoundsShape(const Vec3& half) : half_(half) {}
        AABox GetWorldSpaceBounds(const Mat44& m, const Vec3&) const override
        {
            const Vec4 t = m.GetColumn4(3); // identity rotation assumed (see the header comment)
            const Vec3 c(t.GetX(), t.GetY(), t.GetZ());
            return AABox(c - half_, c + half_);
        }
    private:
        Vec3 half_;
    };
    class SphereShape : public LocalBoundsShape { public: explicit SphereShape(float r) : LocalBoundsShape(Vec3(r, r, r)) {} };
    class BoxShape : public LocalBoundsShape { public: explicit BoxShape(const Vec3& h) : LocalBoundsShape(h) {} };
    class CapsuleShape : public LocalBoundsShape { public: CapsuleShape(float hh, float r) : LocalBoundsShape(Vec3(r, hh + r, r)) {} };
    class CylinderShape : public LocalBoundsShape { public: CylinderShape(float hh, float r) : LocalBoundsShape(Vec3(r, hh, r)) {} };
    class ConvexHullShape : public Shape {};
    class MeshShape : public Shape {};

    struct ShapeResult
    {
        bool HasError() const { return true; }
        ShapeRefC Get() const { return ShapeRefC(); }
    };
    struct TaperedCapsuleShapeSettings { TaperedCapsuleShapeSettings(float, float, float) {} ShapeResult Create() const { return {}; } };
    struct ConvexHullShapeSettings { ConvexHullShapeSettings(const Vec3*, int) {} ShapeResult Create() const { return {}; } };
    struct MeshShapeSettings { explicit MeshShapeSettings(const TriangleList&) {} ShapeResult Create() const { return {}; } };
}
