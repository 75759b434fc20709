// oracle/ref_flat_draw_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN flat-shaded software mesh draws, compiled where they lie under /root/reference with SHS_HAS_JOLT=1 against the
// JoltPhysics declaration shim (oracle/jolt_shim):
//   mode 0  debug_draw::draw_mesh_blinn_phong_transformed               sw_render/debug_draw.hpp:153-203
//   mode 1  draw_mesh_multi_light_transformed                           exp-plumbing/hello_light_types_culling_sw.cpp:366-422, a function of
//           the demo's translation unit: its text (with to_u8 and the two ambient constants) is cut out of the .cpp by
//           oracle/extract_flat_draw.py into oracle/_ref/flat_draw_generated.inc, unmodified, and included below
//   both rasterise through debug_draw::draw_filled_triangle (:60-112) into an RT_ColorLDR + float depth buffer; mode 1 shades through
//   PointLightModel / SpotLightModel / RectAreaLightModel / TubeAreaLightModel::sample of lighting/light_runtime.hpp as they are.
// Built by `make -C oracle ref` into oracle/_ref/libshs_flat_draw_ref.so; tests/test_flat_draw_cpu.py holds oracle_flat_draw.cpp's
// restatement to it byte for byte (canvas) and bit for bit (depth buffer).  Same signature as shso_flat_draw.
#define SHS_HAS_JOLT 1
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <span>
#include <vector>

#include "shs/geometry/volumes.hpp"
#include "shs/gfx/rt_types.hpp"
#include "shs/lighting/light_runtime.hpp"
#include "shs/sw_render/debug_draw.hpp"

using namespace shs;

namespace
{
#include "_ref/flat_draw_generated.inc"

    struct Props128 // ShsbLightProperties (include/shsb.h)
    {
        float color[3], intensity, position[3], range, direction[3], inner, right[3], outer, up[3], tube_half_length, rect_half[2], tube_radius, att_power, att_bias, att_cutoff;
        uint32_t att_model, flags, type, reserved[3];
    };
    static_assert(sizeof(Props128) == 128, "ShsbLightProperties is 128 bytes");
}

extern "C" int32_t shsref_flat_draw(int32_t mode, uint32_t n_draws, const uint32_t* draw_mesh, const float* models16, const float* base3, const uint32_t* sel_counts,
                                    const uint32_t* sel_idx8, const uint32_t* mesh_table3, uint32_t n_meshes, const float* vertices, uint32_t n_vertices, const uint32_t* indices,
                                    uint32_t n_indices, const float view_proj[16], const float camera3[3], const float light_dir3[3], const void* lights128, uint32_t n_lights,
                                    int32_t W, int32_t H, uint8_t* canvas_rgba, float* depth)
{
    (void)n_indices;
    static const PointLightModel point_model;
    static const SpotLightModel spot_model;
    static const RectAreaLightModel rect_model;
    static const TubeAreaLightModel tube_model;

    std::vector<DebugMesh> library(n_meshes);
    for (uint32_t m = 0; m < n_meshes; ++m)
    {
        const uint32_t first = mesh_table3[3 * m], count = mesh_table3[3 * m + 1], base_v = mesh_table3[3 * m + 2];
        uint32_t max_i = 0;
        for (uint32_t i = 0; i < count; ++i) max_i = std::max(max_i, indices[first + i]);
        // a DebugMesh of its own: vertices re-based so that its indices start at 0 (every index must address a vertex: the reference reads it unchecked)
        for (uint32_t v = 0; v <= max_i && count && base_v + v < n_vertices; ++v)
            library[m].vertices.push_back(glm::vec3(vertices[(size_t)(base_v + v) * 3], vertices[(size_t)(base_v + v) * 3 + 1], vertices[(size_t)(base_v + v) * 3 + 2]));
        for (uint32_t i = 0; i < count; ++i) library[m].indices.push_back(indices[first + i]);
    }

    std::vector<LightInstance> lights(n_lights);
    const Props128* lp = static_cast<const Props128*>(lights128);
    for (uint32_t i = 0; i < n_lights; ++i)
    {
        LightProperties& p = lights[i].props;
        p.color = glm::vec3(lp[i].color[0], lp[i].color[1], lp[i].color[2]);
        p.intensity = lp[i].intensity;
        p.position_ws = glm::vec3(lp[i].position[0], lp[i].position[1], lp[i].position[2]);
        p.range = lp[i].range;
        p.direction_ws = glm::vec3(lp[i].direction[0], lp[i].direction[1], lp[i].direction[2]);
        p.inner_angle_rad = lp[i].inner;
        p.outer_angle_rad = lp[i].outer;
        p.right_ws = glm::vec3(lp[i].right[0], lp[i].right[1], lp[i].right[2]);
        p.up_ws = glm::vec3(lp[i].up[0], lp[i].up[1], lp[i].up[2]);
        p.rect_half_extents = glm::vec2(lp[i].rect_half[0], lp[i].rect_half[1]);
        p.tube_half_length = lp[i].tube_half_length;
        p.tube_radius = lp[i].tube_radius;
        p.attenuation_model = (LightAttenuationModel)lp[i].att_model;
        p.attenuation_power = lp[i].att_power;
        p.attenuation_bias = lp[i].att_bias;
        p.attenuation_cutoff = lp[i].att_cutoff;
        p.flags = lp[i].flags;
        switch (lp[i].type)
        {
            case 1u: lights[i].model = &point_model; break;
            case 2u: lights[i].model = &spot_model; break;
            case 3u: lights[i].model = &rect_model; break;
            case 4u: lights[i].model = &tube_model; break;
            default: return 2; // the demo registers a model for every light it creates; a light without one would be a null call
        }
    }

    RT_ColorLDR rt(W, H);
    std::memcpy(&rt.color.at(0, 0), canvas_rgba, (size_t)W * H * 4);
    std::vector<float> depth_buffer(depth, depth + (size_t)W * H);
    glm::mat4 vp;
    std::memcpy(&vp, view_proj, 64);
    const glm::vec3 camera(camera3[0], camera3[1], camera3[2]);
    for (uint32_t d = 0; d < n_draws; ++d)
    {
        if (draw_mesh[d] >= n_meshes) return 1;
        glm::mat4 model;
        std::memcpy(&model, models16 + (size_t)d * 16, 64);
        const glm::vec3 base(base3[3 * d], base3[3 * d + 1], base3[3 * d + 2]);
        if (mode == 0)
        {
            debug_draw::draw_mesh_blinn_phong_transformed(rt, std::span<float>(depth_buffer.data(), depth_buffer.size()), library[draw_mesh[d]], model, vp, W, H, camera,
                                                          glm::vec3(light_dir3[0], light_dir3[1], light_dir3[2]), base);
        }
        else
        {
            LightSelection sel{};
            sel.count = std::min(sel_counts[d], kLightSelectionCapacity);
            for (uint32_t k = 0; k < kLightSelectionCapacity; ++k) sel.indices[k] = sel_idx8[(size_t)d * 8 + k];
            draw_mesh_multi_light_transformed(rt, depth_buffer, library[draw_mesh[d]], model, vp, W, H, camera, base, lights, sel);
        }
    }
    std::memcpy(canvas_rgba, &rt.color.at(0, 0), (size_t)W * H * 4);
    std::memcpy(depth, depth_buffer.data(), depth_buffer.size() * sizeof(float));
    return 0;
}
