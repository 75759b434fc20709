// oracle/ref_lightcull_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with the product).
//
// The reference's OWN tile / depth-range / clustered light-list builders,
// shs/lighting/jolt_light_culling.hpp:135-412 with geometry/jolt_culling.hpp, geometry/frustum_culling.hpp, geometry/scene_shape.hpp
// and geometry/jolt_adapter.hpp, compiled where they lie under /root/reference with SHS_HAS_JOLT=1 against the JoltPhysics
// DECLARATION shim oracle/jolt_shim (JoltPhysics v5.2.0 itself is a vcpkg dependency, absent here).  The headers use Jolt only to
// obtain a light's world bounds (SceneShape::bounding_sphere / world_aabb); a light is therefore a shim Shape that carries its
// world AABox.  Everything the lists depend on after that point is the reference's code.  The bounds the reference derived
// (sphere + AABB in SHS space, after its LH <-> RH conversions) are returned so that the restatement is fed the same numbers.
// Built by oracle/Makefile (`make ref`) into oracle/_ref/libshs_lightcull_ref.so.
#include <cstdint>
#include <cstring>
#include <vector>

#define SHS_HAS_JOLT 1
#include "shs/lighting/jolt_light_culling.hpp"

namespace
{
    struct BoundsShape final : JPH::Shape
    {
        JPH::AABox box;
        JPH::AABox GetWorldSpaceBounds(const JPH::Mat44&, const JPH::Vec3&) const override { return box; }
    };

    struct Lights
    {
        std::vector<BoundsShape> shapes;
        std::vector<shs::SceneShape> scene;
        Lights(const float* aabb6, uint32_t n) : shapes(n), scene(n)
        {
            for (uint32_t i = 0; i < n; ++i)
            {
                shs::AABB b{};
                b.minv = glm::vec3(aabb6[6 * i], aabb6[6 * i + 1], aabb6[6 * i + 2]);
                b.maxv = glm::vec3(aabb6[6 * i + 3], aabb6[6 * i + 4], aabb6[6 * i + 5]);
                shapes[i].box = shs::jolt::to_jph(b); // the reference's own SHS (LH) -> Jolt (RH) conversion
                scene[i].shape = JPH::ShapeRefC(&shapes[i]);
                scene[i].stable_id = i;
            }
        }
    };

    void write_lists(const std::vector<std::vector<uint32_t>>& lists, uint32_t max_per_bin, uint32_t* counts, uint32_t* indices)
    {
        for (size_t b = 0; b < lists.size(); ++b)
        {
            counts[b] = (uint32_t)lists[b].size();
            for (size_t k = 0; k < lists[b].size() && k < max_per_bin; ++k) indices[b * max_per_bin + k] = lists[b][k];
        }
    }
}

extern "C"
{
    // aabb6: n x (min xyz, max xyz) in SHS space.  out_bounds10: n x (sphere centre xyz, radius, aabb min xyz, aabb max xyz) as
    // SceneShape::bounding_sphere / world_aabb report them.
    int32_t shsref_light_bounds(const float* aabb6, uint32_t n, float* out_bounds10)
    {
        if (!aabb6 || !out_bounds10) return 1;
        Lights L(aabb6, n);
        for (uint32_t i = 0; i < n; ++i)
        {
            const shs::Sphere s = L.scene[i].bounding_sphere();
            const shs::AABB b = L.scene[i].world_aabb();
            float* o = out_bounds10 + 10 * i;
            o[0] = s.center.x; o[1] = s.center.y; o[2] = s.center.z; o[3] = s.radius;
            o[4] = b.minv.x; o[5] = b.minv.y; o[6] = b.minv.z; o[7] = b.maxv.x; o[8] = b.maxv.y; o[9] = b.maxv.z;
        }
        return 0;
    }

    // mode: 0 cull_lights_tiled, 1 cull_lights_tiled_depth01_range, 2 cull_lights_tiled_view_depth_range, 3 cull_lights_clustered
    // (the numbering of SHSB_LIGHT_CULL_*).  counts[bins] uncapped, indices[bins * max_per_bin] the first max_per_bin entries.
    int32_t shsref_light_cull(const float* aabb6, uint32_t n, const float view_proj[16], uint32_t w, uint32_t h, uint32_t tile_size, uint32_t max_per_bin,
                              int32_t mode, uint32_t depth_slices, float z_near, float z_far, const float* range_min, const float* range_max, uint32_t n_ranges,
                              uint32_t* counts, uint32_t* indices)
    {
        if (!aabb6 || !view_proj || !counts || !indices || tile_size == 0) return 1;
        Lights L(aabb6, n);
        glm::mat4 vp;
        std::memcpy(&vp, view_proj, 64);
        const std::span<const shs::SceneShape> shapes(L.scene.data(), L.scene.size());
        const std::span<const float> lo(range_min, range_min ? n_ranges : 0), hi(range_max, range_max ? n_ranges : 0);
        switch (mode)
        {
        case 0: write_lists(shs::cull_lights_tiled(shapes, vp, w, h, tile_size).tile_light_lists, max_per_bin, counts, indices); return 0;
        case 1: write_lists(shs::cull_lights_tiled_depth01_range(shapes, vp, w, h, tile_size, lo, hi).tile_light_lists, max_per_bin, counts, indices); return 0;
        case 2: write_lists(shs::cull_lights_tiled_view_depth_range(shapes, vp, w, h, tile_size, lo, hi, z_near, z_far).tile_light_lists, max_per_bin, counts, indices); return 0;
        case 3: write_lists(shs::cull_lights_clustered(shapes, vp, w, h, tile_size, depth_slices, z_near, z_far).cluster_light_lists, max_per_bin, counts, indices); return 0;
        default: return 1;
        }
    }
}
