// shs_b200/scene_cull_drop_in.hpp -- reference-side binding of the scene-level culling calls (C++20, header-only; SURVEY.md section 8f
// row 1): same argument and result types as the reference functions they stand in for, forwarding to the C-ABI (include/shsb.h).
//
//     shs::cull_vs_frustum(objects, extract_frustum_planes(vp))           -> shs::b200::cull_vs_frustum(ctx, objects, vp)
//     shs::collect_object_lights(box, visible, light_scene, lights, mode) -> shs::b200::collect_object_lights(ctx, boxes, visible, light_scene, lights, mode)   (all draws at once)
//     shs::build_tile_view_depth_range_from_scene(visible, scene, view, vp, w, h, ts, zn, zf) -> shs::b200::build_tile_view_depth_range_from_scene(ctx, ...)
//
// Needs the reference's headers with SHS_HAS_JOLT (geometry/jolt_culling.hpp, lighting/light_runtime.hpp, lighting/light_culling_runtime.hpp):
// the Jolt shape -> bounds step (SceneShape::bounding_sphere / world_aabb) stays on the host, inside the reference's own types.
// No CPU fallback: a failing call returns false / an empty result and leaves the error in shsb_last_error_string(ctx).
#pragma once

#include <cstring>
#include <span>
#include <vector>

#include "shs/geometry/jolt_culling.hpp"
#include "shs/lighting/light_culling_runtime.hpp"
#include "shs/lighting/light_runtime.hpp"

#include "shsb.h"

namespace shs::b200
{
    // geometry/jolt_culling.hpp:279-306
    template <FastCullable T>
        requires HasWorldAABB<T>
    inline bool cull_vs_frustum(shsb_ctx ctx, std::span<const T> objects, const glm::mat4& view_proj, CullResult& out)
    {
        out = CullResult{};
        std::vector<float> bounds;
        bounds.reserve(objects.size() * 10);
        for (const T& o : objects)
        {
            const Sphere s = o.bounding_sphere();
            const AABB b = o.world_aabb();
            const float v[10] = {s.center.x, s.center.y, s.center.z, s.radius, b.minv.x, b.minv.y, b.minv.z, b.maxv.x, b.maxv.y, b.maxv.z};
            bounds.insert(bounds.end(), v, v + 10);
        }
        const uint32_t n = (uint32_t)objects.size();
        std::vector<uint8_t> classes(n);
        std::vector<uint32_t> visible(n);
        uint32_t counts[5] = {0, 0, 0, 0, 0};
        if (shsb_cull_objects_frustum(ctx, bounds.data(), n, &view_proj[0][0], classes.data(), visible.data(), counts) != SHSB_OK) return false;
        out.classes.resize(n);
        for (uint32_t i = 0; i < n; ++i) out.classes[i] = (CullClass)classes[i];
        out.visible_indices.assign(visible.begin(), visible.begin() + counts[4]);
        out.tested = counts[0]; out.outside = counts[1]; out.intersecting = counts[2]; out.inside = counts[3];
        return true;
    }

    // lighting/light_runtime.hpp:592-616, for every box of `boxes` (one per draw) in one call
    inline bool collect_object_lights(shsb_ctx ctx, std::span<const AABB> boxes, std::span<const uint32_t> visible_light_scene_indices, const SceneElementSet& light_scene,
                                      const std::vector<LightInstance>& lights, LightObjectCullMode cull_mode, std::vector<LightSelection>& out)
    {
        out.assign(boxes.size(), LightSelection{});
        static_assert(sizeof(CullingLightGPU) == 160, "CullingLightGPU is 160 bytes");
        std::vector<CullingLightGPU> records(lights.size());
        for (size_t i = 0; i < lights.size(); ++i)
        {
            records[i] = lights[i].packed;
            records[i].position_range = glm::vec4(lights[i].props.position_ws, lights[i].packed.position_range.w); // the distance uses LightProperties::position_ws
        }
        std::vector<uint32_t> visible_lights; // scene index -> light index, the two range checks of :604-606 applied here
        visible_lights.reserve(visible_light_scene_indices.size());
        for (const uint32_t scene_idx : visible_light_scene_indices)
        {
            if (scene_idx >= light_scene.size()) continue;
            const uint32_t light_idx = light_scene[scene_idx].user_index;
            if (light_idx >= lights.size()) continue;
            visible_lights.push_back(light_idx);
        }
        std::vector<float> aabbs;
        aabbs.reserve(boxes.size() * 6);
        for (const AABB& b : boxes) { const float v[6] = {b.minv.x, b.minv.y, b.minv.z, b.maxv.x, b.maxv.y, b.maxv.z}; aabbs.insert(aabbs.end(), v, v + 6); }
        const uint32_t n = (uint32_t)boxes.size();
        std::vector<uint32_t> counts(n), idx((size_t)n * 8);
        std::vector<float> d2((size_t)n * 8);
        if (shsb_collect_object_lights(ctx, aabbs.data(), n, visible_lights.data(), (uint32_t)visible_lights.size(), records.data(), (uint32_t)records.size(), (int32_t)cull_mode,
                                       counts.data(), idx.data(), d2.data()) != SHSB_OK)
            return false;
        for (uint32_t o = 0; o < n; ++o)
        {
            out[o].count = counts[o];
            for (uint32_t k = 0; k < kLightSelectionCapacity; ++k) { out[o].indices[k] = idx[(size_t)o * 8 + k]; out[o].dist2[k] = d2[(size_t)o * 8 + k]; }
        }
        return true;
    }

    // lighting/light_culling_runtime.hpp:188-264; the ranges also stay on the device for shsb_light_cull_ex
    inline bool build_tile_view_depth_range_from_scene(shsb_ctx ctx, std::span<const uint32_t> visible_scene_indices, const SceneElementSet& scene, const glm::mat4& view,
                                                       const glm::mat4& view_proj, uint32_t viewport_w, uint32_t viewport_h, uint32_t tile_size, float z_near, float z_far,
                                                       TileViewDepthRange& out)
    {
        out = TileViewDepthRange{};
        if (viewport_w == 0u || viewport_h == 0u || tile_size == 0u) return true; // the reference returns an empty range
        std::vector<float> aabbs;
        aabbs.reserve(scene.size() * 6);
        for (size_t i = 0; i < scene.size(); ++i)
        {
            const AABB b = scene[i].geometry.world_aabb();
            const float v[6] = {b.minv.x, b.minv.y, b.minv.z, b.maxv.x, b.maxv.y, b.maxv.z};
            aabbs.insert(aabbs.end(), v, v + 6);
        }
        if (shsb_tile_depth_range_from_scene(ctx, aabbs.data(), (uint32_t)scene.size(), visible_scene_indices.data(), (uint32_t)visible_scene_indices.size(), &view[0][0],
                                             &view_proj[0][0], viewport_w, viewport_h, tile_size, z_near, z_far) != SHSB_OK)
            return false;
        out.tiles_x = (viewport_w + tile_size - 1u) / tile_size;
        out.tiles_y = (viewport_h + tile_size - 1u) / tile_size;
        const size_t tiles = (size_t)out.tiles_x * out.tiles_y;
        out.min_view_depth.resize(tiles);
        out.max_view_depth.resize(tiles);
        return shsb_tile_depth_range_download(ctx, out.min_view_depth.data(), out.max_view_depth.data(), tiles) == SHSB_OK;
    }
}
