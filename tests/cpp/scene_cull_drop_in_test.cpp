// tests/cpp/scene_cull_drop_in_test.cpp -- host/shs_b200/scene_cull_drop_in.hpp against the reference functions it stands in for, over the
// reference's own types (SceneShape, SceneElementSet, LightInstance, CullResult, LightSelection, TileViewDepthRange).  Compiled with
// SHS_HAS_JOLT=1 against the JoltPhysics declaration shim (oracle/jolt_shim); objects / lights are shim shapes carrying their bounds.
// Exit code 0 = every result equal, 1 = mismatch, 77 = no CUDA device (after checking that the binding refused and computed nothing).
#include <cstdio>
#include <cstring>
#include <random>

#define SHS_HAS_JOLT 1
#include "shs_b200/scene_cull_drop_in.hpp"

struct BoundsShape final : JPH::Shape
{
    JPH::AABox box;
    JPH::AABox GetWorldSpaceBounds(const JPH::Mat44&, const JPH::Vec3&) const override { return box; }
};

int main()
{
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(-1.0f, 1.0f);
    const uint32_t n_obj = 600, n_lights = 150, W = 640, H = 360, TS = 16;
    const float zn = 0.1f, zf = 120.0f;
    std::vector<BoundsShape> shapes(n_obj + n_lights);
    std::vector<shs::SceneShape> objects(n_obj);
    shs::SceneElementSet scene, light_scene;
    std::vector<shs::AABB> boxes(n_obj);
    for (uint32_t i = 0; i < n_obj; ++i)
    {
        const glm::vec3 c(U(rng) * 40.0f, U(rng) * 6.0f, U(rng) * 40.0f), h(0.05f + std::abs(U(rng)) * 2.5f, 0.05f + std::abs(U(rng)) * 2.5f, 0.05f + std::abs(U(rng)) * 2.5f);
        boxes[i].minv = c - h; boxes[i].maxv = c + h;
        shapes[i].box = shs::jolt::to_jph(boxes[i]);
        objects[i].shape = JPH::ShapeRefC(&shapes[i]);
        objects[i].stable_id = i + 1u;
        shs::SceneElement e{};
        e.geometry = objects[i];
        scene.add(e);
    }
    std::vector<shs::LightInstance> lights(n_lights);
    std::vector<uint32_t> visible_lights, visible_objects;
    for (uint32_t i = 0; i < n_lights; ++i)
    {
        const glm::vec3 p(U(rng) * 30.0f, 1.0f + U(rng) * 3.0f, U(rng) * 30.0f);
        const float r = 1.0f + std::abs(U(rng)) * 9.0f;
        lights[i].props.position_ws = p;
        lights[i].packed.position_range = glm::vec4(p, r);
        lights[i].packed.cull_sphere = glm::vec4(p, r);
        lights[i].packed.cull_aabb_min = glm::vec4(p - glm::vec3(r), 1.0f);
        lights[i].packed.cull_aabb_max = glm::vec4(p + glm::vec3(r), 1.0f);
        shs::SceneElement e{};
        e.user_index = (n_lights - 1u) - i; // a non-identity scene -> light mapping
        light_scene.add(e);
        if (i % 7 != 3) visible_lights.push_back(i);
    }
    visible_lights.push_back(n_lights + 5u); // out of range: skipped
    for (uint32_t i = 0; i < n_obj; i += (i % 5 == 0 ? 2 : 1)) visible_objects.push_back(i);

    const glm::vec3 eye(2.0f, 5.0f, -35.0f);
    const glm::mat4 view = glm::lookAtLH(eye, glm::vec3(0.0f, 1.0f, 0.0f), glm::vec3(0, 1, 0));
    const glm::mat4 proj = glm::perspectiveLH_NO(glm::radians(60.0f), (float)W / (float)H, zn, zf);
    const glm::mat4 vp = proj * view;

    // ---- reference
    const shs::CullResult ref_cull = shs::cull_vs_frustum(std::span<const shs::SceneShape>(objects.data(), objects.size()), shs::extract_frustum_planes(vp));
    std::vector<shs::LightSelection> ref_sel(n_obj);
    for (uint32_t i = 0; i < n_obj; ++i) ref_sel[i] = shs::collect_object_lights(boxes[i], visible_lights, light_scene, lights, shs::LightObjectCullMode::SphereAabb);
    const shs::TileViewDepthRange ref_range = shs::build_tile_view_depth_range_from_scene(visible_objects, scene, view, vp, W, H, TS, zn, zf);
    std::printf("reference side: %zu of %u objects visible (%llu intersecting), %u light links, %zu tiles\n", ref_cull.visible_indices.size(), n_obj,
                (unsigned long long)ref_cull.intersecting, [&] { uint32_t s = 0; for (auto& x : ref_sel) s += x.count; return s; }(), ref_range.min_view_depth.size());

    // ---- B200
    shsb_ctx ctx = nullptr;
    shs::CullResult gpu_cull;
    std::vector<shs::LightSelection> gpu_sel;
    shs::TileViewDepthRange gpu_range;
    if (shsb_context_create(0, &ctx) != SHSB_OK)
    {
        const bool refused = !shs::b200::cull_vs_frustum(nullptr, std::span<const shs::SceneShape>(objects.data(), objects.size()), vp, gpu_cull) && gpu_cull.visible_indices.empty() &&
                             !shs::b200::collect_object_lights(nullptr, boxes, visible_lights, light_scene, lights, shs::LightObjectCullMode::SphereAabb, gpu_sel) &&
                             !shs::b200::build_tile_view_depth_range_from_scene(nullptr, visible_objects, scene, view, vp, W, H, TS, zn, zf, gpu_range) && gpu_range.min_view_depth.empty();
        std::printf("SKIP: no CUDA device (%s)\n", refused ? "every call refused, nothing ran on the CPU" : "A CALL CLAIMED SUCCESS WITHOUT A DEVICE");
        return refused ? 77 : 1;
    }
    bool ok = shs::b200::cull_vs_frustum(ctx, std::span<const shs::SceneShape>(objects.data(), objects.size()), vp, gpu_cull);
    ok = ok && shs::b200::collect_object_lights(ctx, boxes, visible_lights, light_scene, lights, shs::LightObjectCullMode::SphereAabb, gpu_sel);
    ok = ok && shs::b200::build_tile_view_depth_range_from_scene(ctx, visible_objects, scene, view, vp, W, H, TS, zn, zf, gpu_range);
    if (!ok) { std::printf("FAIL: %s\n", shsb_last_error_string(ctx)); return 1; }
    bool same = gpu_cull.classes == ref_cull.classes && gpu_cull.visible_indices == ref_cull.visible_indices && gpu_cull.tested == ref_cull.tested && gpu_cull.outside == ref_cull.outside &&
                gpu_cull.intersecting == ref_cull.intersecting && gpu_cull.inside == ref_cull.inside;
    for (uint32_t i = 0; i < n_obj && same; ++i)
        same = gpu_sel[i].count == ref_sel[i].count && gpu_sel[i].indices == ref_sel[i].indices && std::memcmp(gpu_sel[i].dist2.data(), ref_sel[i].dist2.data(), 32) == 0;
    same = same && gpu_range.tiles_x == ref_range.tiles_x && gpu_range.tiles_y == ref_range.tiles_y && gpu_range.min_view_depth.size() == ref_range.min_view_depth.size() &&
           std::memcmp(gpu_range.min_view_depth.data(), ref_range.min_view_depth.data(), ref_range.min_view_depth.size() * 4) == 0 &&
           std::memcmp(gpu_range.max_view_depth.data(), ref_range.max_view_depth.data(), ref_range.max_view_depth.size() * 4) == 0;
    std::printf("scene-cull drop-in: %s\n", same ? "OK (classes, visible list, counters, light selections, tile depth ranges equal)" : "MISMATCH");
    shsb_context_destroy(ctx);
    return same ? 0 : 1;
}
