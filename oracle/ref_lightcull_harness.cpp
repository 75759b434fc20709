// oracle/ref_lightcull_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with the product).
//
// The reference's OWN tile / depth-range / clustered light-list builders,
// shs/lighting/jolt_light_culling.hpp:135-412 with geometry/jolt_culling.hpp, geometry/frustum_culling.hpp, geometry/scene_shape.hpp
// and geometry/jolt_adapter.hpp, compiled where they lie under /root/reference with SHS_HAS_JOLT=1 against the JoltPhysics
// DECLARATION shim oracle/jolt_shim (JoltPhysics v5.2.0 itself is a vcpkg dependency, absent here).  The headers use Jolt only to
// obtain a light's world bounds (SceneShape::bounding_sphere / world_aabb); a light is therefore a shim Shape that carries its
// world AABox.  Everything the lists depend on after that point is the reference's code.  The bounds the reference derived
// (sphere + AABB in SHS space, after its LH <-> RH conversions) are returned so that the restatement is fed the same numbers.
// Also here, compiled the same way: cull_vs_frustum over SceneShapes (geometry/jolt_culling.hpp:279-306)
// and collect_object_lights (lighting/light_runtime.hpp:592-616).
// Built by oracle/Makefile (`make ref`) into oracle/_ref/libshs_lightcull_ref.so.
#include <cstdint>
#include <cstring>
#include <vector>

#define SHS_HAS_JOLT 1
#include "shs/lighting/jolt_light_culling.hpp"
#include "shs/lighting/light_runtime.hpp"
#include "shs/lighting/light_culling_runtime.hpp"

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

extern "C"
{
    // cull_vs_frustum(objects, extract_frustum_planes(view_proj)); objects enter as world AABBs like the lights above.
    int32_t shsref_cull_objects(const float* aabb6, uint32_t n, const float view_proj[16], uint8_t* classes, uint32_t* visible, uint32_t counts5[5])
    {
        if (!view_proj || !classes || !visible || !counts5) return 1;
        Lights L(aabb6, n);
        glm::mat4 vp;
        std::memcpy(&vp, view_proj, 64);
        const shs::Frustum fr = shs::extract_frustum_planes(vp);
        const shs::CullResult r = shs::cull_vs_frustum(std::span<const shs::SceneShape>(L.scene.data(), L.scene.size()), fr);
        for (uint32_t i = 0; i < n; ++i) classes[i] = (uint8_t)r.classes[i];
        for (size_t i = 0; i < r.visible_indices.size(); ++i) visible[i] = (uint32_t)r.visible_indices[i];
        counts5[0] = (uint32_t)r.tested; counts5[1] = (uint32_t)r.outside; counts5[2] = (uint32_t)r.intersecting; counts5[3] = (uint32_t)r.inside;
        counts5[4] = (uint32_t)r.visible_indices.size();
        return 0;
    }

    // collect_object_lights per object.  records160: CullingLightGPU records -> LightInstance::packed; LightProperties::position_ws is
    // set from position_range.xyz (what the reference's packers copy there).  The light scene is the identity mapping
    // (scene element i -> user_index i), so `visible` holds light indices.
    int32_t shsref_collect_object_lights(const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const void* records160, uint32_t n_lights,
                                         int32_t cull_mode, uint32_t* out_counts, uint32_t* out_indices8, float* out_dist2_8)
    {
        if (!out_counts || !out_indices8 || !out_dist2_8 || cull_mode < 0 || cull_mode > 2) return 1;
        static_assert(sizeof(shs::CullingLightGPU) == 160, "CullingLightGPU is 160 bytes");
        static_assert(shs::kLightSelectionCapacity == 8u, "LightSelection holds 8 lights");
        std::vector<shs::LightInstance> lights(n_lights);
        shs::SceneElementSet light_scene;
        for (uint32_t i = 0; i < n_lights; ++i)
        {
            std::memcpy(&lights[i].packed, (const uint8_t*)records160 + (size_t)i * 160, 160);
            lights[i].props.position_ws = glm::vec3(lights[i].packed.position_range);
            shs::SceneElement e{};
            e.user_index = i;
            light_scene.add(e);
        }
        for (uint32_t o = 0; o < n_objects; ++o)
        {
            shs::AABB box{};
            box.minv = glm::vec3(object_aabbs6[6 * o], object_aabbs6[6 * o + 1], object_aabbs6[6 * o + 2]);
            box.maxv = glm::vec3(object_aabbs6[6 * o + 3], object_aabbs6[6 * o + 4], object_aabbs6[6 * o + 5]);
            const shs::LightSelection sel = shs::collect_object_lights(box, std::span<const uint32_t>(visible, n_visible), light_scene, lights, (shs::LightObjectCullMode)cull_mode);
            out_counts[o] = sel.count;
            for (uint32_t k = 0; k < 8; ++k) { out_indices8[8 * o + k] = sel.indices[k]; out_dist2_8[8 * o + k] = sel.dist2[k]; }
        }
        return 0;
    }
}

extern "C" int32_t shsref_tile_depth_range_from_scene(const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const float view[16],
                                                       const float view_proj[16], uint32_t viewport_w, uint32_t viewport_h, uint32_t tile_size, float z_near, float z_far,
                                                       float* out_min, float* out_max)
{
    // build_tile_view_depth_range_from_scene (lighting/light_culling_runtime.hpp:188-264) over a SceneElementSet whose geometries carry the AABBs
    if (!view || !view_proj || !out_min || !out_max) return 1;
    Lights L(object_aabbs6, n_objects);
    shs::SceneElementSet scene;
    for (uint32_t i = 0; i < n_objects; ++i)
    {
        shs::SceneElement e{};
        e.geometry = L.scene[i];
        e.geometry.stable_id = i + 1u;
        scene.add(e);
    }
    glm::mat4 v, vp;
    std::memcpy(&v, view, 64);
    std::memcpy(&vp, view_proj, 64);
    const shs::TileViewDepthRange r = shs::build_tile_view_depth_range_from_scene(std::span<const uint32_t>(visible, n_visible), scene, v, vp, viewport_w, viewport_h, tile_size, z_near, z_far);
    for (size_t t = 0; t < r.min_view_depth.size(); ++t) { out_min[t] = r.min_view_depth[t]; out_max[t] = r.max_view_depth[t]; }
    return 0;
}

extern "C" int32_t shsref_select_object_lights_from_bins(const float* object_aabbs6, uint32_t n_objects, const float view[16], const float view_proj[16], uint32_t viewport_w,
                                                          uint32_t viewport_h, int32_t culling_mode /* LightCullingMode 0..3 */, uint32_t tile_size, uint32_t depth_slices, float z_near,
                                                          float z_far, const float* range_min, const float* range_max, uint32_t n_ranges, const float* light_aabbs6,
                                                          const void* records160, uint32_t n_lights, int32_t cull_mode, uint32_t* out_counts, uint32_t* out_indices8,
                                                          float* out_dist2_8, uint32_t* out_candidates)
{
    // build_light_bin_culling (lighting/light_culling_runtime.hpp:266-371) over all lights, then per object
    // gather_light_scene_candidates_for_aabb (:373-449) -> collect_object_lights, as hello_light_types_culling_sw.cpp:968-996 does
    if (!view || !view_proj || !out_counts || !out_indices8 || !out_dist2_8 || !out_candidates) return 1;
    Lights L(light_aabbs6, n_lights);
    std::vector<shs::LightInstance> lights(n_lights);
    shs::SceneElementSet light_scene;
    std::vector<uint32_t> visible(n_lights);
    for (uint32_t i = 0; i < n_lights; ++i)
    {
        std::memcpy(&lights[i].packed, (const uint8_t*)records160 + (size_t)i * 160, 160);
        lights[i].props.position_ws = glm::vec3(lights[i].packed.position_range);
        shs::SceneElement e{};
        e.geometry = L.scene[i];
        e.geometry.stable_id = i + 1u;
        e.user_index = i;
        light_scene.add(e);
        visible[i] = i;
    }
    glm::mat4 v, vp;
    std::memcpy(&v, view, 64);
    std::memcpy(&vp, view_proj, 64);
    shs::LightBinCullingConfig cfg{};
    cfg.mode = (shs::LightCullingMode)culling_mode;
    cfg.tile_size = tile_size;
    cfg.cluster_depth_slices = depth_slices;
    cfg.z_near = z_near;
    cfg.z_far = z_far;
    const shs::LightBinCullingData data = shs::build_light_bin_culling(std::span<const uint32_t>(visible.data(), visible.size()), light_scene, vp, viewport_w, viewport_h, cfg,
                                                                       std::span<const float>(range_min, range_min ? n_ranges : 0), std::span<const float>(range_max, range_max ? n_ranges : 0));
    std::vector<uint32_t> scratch;
    for (uint32_t o = 0; o < n_objects; ++o)
    {
        shs::AABB box{};
        box.minv = glm::vec3(object_aabbs6[6 * o], object_aabbs6[6 * o + 1], object_aabbs6[6 * o + 2]);
        box.maxv = glm::vec3(object_aabbs6[6 * o + 3], object_aabbs6[6 * o + 4], object_aabbs6[6 * o + 5]);
        const std::span<const uint32_t> cand = shs::gather_light_scene_candidates_for_aabb(data, box, v, vp, scratch);
        out_candidates[o] = (uint32_t)cand.size();
        const shs::LightSelection sel = shs::collect_object_lights(box, cand, light_scene, lights, (shs::LightObjectCullMode)cull_mode);
        out_counts[o] = sel.count;
        for (uint32_t k = 0; k < 8; ++k) { out_indices8[8 * o + k] = sel.indices[k]; out_dist2_8[8 * o + k] = sel.dist2[k]; }
    }
    return 0;
}
