// oracle/ref_occlusion_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN software-occlusion pass (geometry/culling_software.hpp: run_software_occlusion_pass, project_aabb_to_screen_rect,
// is_rect_occluded, rasterize_mesh_depth_transformed, view_depth_of_aabb_center) compiled where it lies under /root/reference with
// SHS_HAS_JOLT=1 against the JoltPhysics declaration shim (oracle/jolt_shim: the header pulls Jolt in only through
// geometry/jolt_debug_draw.hpp for the DebugMesh struct; no Jolt function is called here) and driven with the callbacks
// SceneCullingContext::run_software_occlusion uses (scene/scene_culling.hpp:186-222).  Built by `make -C oracle ref` into
// oracle/_ref/libshs_occlusion_ref.so; tests/test_occlusion_cpu.py holds oracle_scene_cull.cpp's restatement to it bit for bit.
#define SHS_HAS_JOLT 1
#include <cstring>
#include <vector>

#include "shs/geometry/culling_software.hpp"

namespace
{
    struct Obj
    {
        shs::AABB box;
        uint32_t mesh = 0xFFFFFFFFu;
        glm::mat4 model{1.0f};
        bool occluded = false, visible = false;
    };
}

extern "C" int32_t shsref_software_occlusion(const float* object_aabbs6, uint32_t n_objects, const uint32_t* frustum_visible, uint32_t n_visible, const uint32_t* object_mesh,
                                             const float* object_models16, const uint32_t* mesh_table3, uint32_t n_meshes, const float* vertices, uint32_t n_vertices,
                                             const uint32_t* indices, uint32_t n_indices, const float view[16], const float view_proj[16], int32_t occ_w, int32_t occ_h,
                                             float depth_epsilon, int32_t enable_occlusion, uint8_t* out_occluded, uint32_t* out_visible, uint32_t out_counts4[4], float* out_depth)
{
    (void)n_vertices; (void)n_indices;
    std::vector<shs::DebugMesh> library(n_meshes);
    for (uint32_t m = 0; m < n_meshes; ++m)
    {
        const uint32_t first = mesh_table3[3 * m], count = mesh_table3[3 * m + 1], base_v = mesh_table3[3 * m + 2];
        // a DebugMesh of its own: vertices re-based so that its indices start at 0
        uint32_t max_i = 0;
        for (uint32_t i = 0; i < count; ++i) max_i = std::max(max_i, indices[first + i]);
        for (uint32_t v = 0; v <= max_i && count; ++v) library[m].vertices.push_back(glm::vec3(vertices[(size_t)(base_v + v) * 3], vertices[(size_t)(base_v + v) * 3 + 1], vertices[(size_t)(base_v + v) * 3 + 2]));
        for (uint32_t i = 0; i < count; ++i) library[m].indices.push_back(indices[first + i]);
    }
    std::vector<Obj> objects(n_objects);
    for (uint32_t i = 0; i < n_objects; ++i)
    {
        const float* b = object_aabbs6 + (size_t)i * 6;
        objects[i].box.minv = glm::vec3(b[0], b[1], b[2]);
        objects[i].box.maxv = glm::vec3(b[3], b[4], b[5]);
        objects[i].mesh = object_mesh[i];
        std::memcpy(&objects[i].model, object_models16 + (size_t)i * 16, 64);
    }
    glm::mat4 v, vp;
    std::memcpy(&v, view, 64);
    std::memcpy(&vp, view_proj, 64);
    std::vector<float> depth((size_t)occ_w * occ_h, 1.0f);
    std::vector<uint32_t> visible;
    const shs::CullingStats st = shs::culling_sw::run_software_occlusion_pass(
        std::span<Obj>(objects.data(), objects.size()), std::span<const uint32_t>(frustum_visible, n_visible), enable_occlusion != 0,
        std::span<float>(depth.data(), depth.size()), occ_w, occ_h, v, vp,
        [](const Obj& o) -> shs::AABB { return o.box; },
        [](const Obj& o, const glm::mat4& view_mtx) -> float { return shs::culling_sw::view_depth_of_aabb_center(o.box, view_mtx); },
        [](Obj& o, bool occluded) { o.occluded = occluded; },
        [](Obj& o, bool vis) { o.visible = vis; },
        [&](const Obj& o, uint32_t, std::span<float> depth_span) {
            if (o.mesh >= library.size()) return;
            shs::culling_sw::rasterize_mesh_depth_transformed(depth_span, occ_w, occ_h, library[o.mesh], o.model, vp);
        },
        visible, depth_epsilon);
    for (uint32_t i = 0; i < n_objects; ++i) out_occluded[i] = objects[i].occluded ? 1 : 0;
    for (size_t k = 0; k < visible.size(); ++k) out_visible[k] = visible[k];
    out_counts4[0] = st.scene_count; out_counts4[1] = st.frustum_visible_count; out_counts4[2] = st.visible_count; out_counts4[3] = st.occluded_count;
    if (out_depth) std::memcpy(out_depth, depth.data(), depth.size() * sizeof(float));
    return 0;
}
