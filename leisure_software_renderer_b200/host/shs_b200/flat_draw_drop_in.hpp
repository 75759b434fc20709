// shs_b200/flat_draw_drop_in.hpp -- reference-side binding of the flat-shaded software mesh draws (C++20, header-only; SURVEY.md section 8f
// row 1: the consumer of the per-object light selections).  The reference issues one call per visible object,
//
//     debug_draw::draw_mesh_blinn_phong_transformed(rt, depth, mesh, model, vp, w, h, camera_pos, light_dir_ws, base_color)   sw_render/debug_draw.hpp:153-203
//     draw_mesh_multi_light_transformed(rt, depth, mesh, model, vp, w, h, camera_pos, base_color, lights, selection)          exp-plumbing/hello_light_types_culling_sw.cpp:366-422
//
// and the device takes a frame's draws as ONE batch (include/shsb.h: shsb_flat_draw_blinn_phong / shsb_flat_draw_multi_light), so the binding
// records the per-object arguments with the same parameter lists minus the per-frame ones and submits them in flush():
//
//     shs::b200::FlatDrawBatch batch(ctx);
//     for (object : visible) batch.draw_mesh_multi_light_transformed(mesh, model, base_color, selection);
//     batch.flush_multi_light(rt, depth, vp, w, h, camera_pos, lights);         // rt and depth are updated like the reference's loop leaves them
//
// DebugMesh objects are uploaded once per (address, vertex count, index count) and cached; the RT_ColorLDR and the float depth buffer are
// uploaded before and downloaded after the batch (a caller that keeps its targets on the device uses the C-ABI directly).
// Needs the reference's headers with SHS_HAS_JOLT (geometry/jolt_debug_draw.hpp declares DebugMesh inside that guard).
// No CPU fallback: a failing call returns false, leaves rt / depth untouched and the error in shsb_last_error_string(ctx).
#pragma once

#include <cstring>
#include <map>
#include <span>
#include <tuple>
#include <vector>

#include "shs/gfx/rt_types.hpp"
#include "shs/lighting/light_runtime.hpp"
#include "shs/sw_render/debug_draw.hpp"

#include "shsb.h"

namespace shs::b200
{
    inline ShsbLightProperties to_shsb(const LightInstance& light)
    {
        ShsbLightProperties o{};
        const LightProperties& p = light.props;
        std::memcpy(o.color, &p.color[0], 12); o.intensity = p.intensity;
        std::memcpy(o.position_ws, &p.position_ws[0], 12); o.range = p.range;
        std::memcpy(o.direction_ws, &p.direction_ws[0], 12); o.inner_angle_rad = p.inner_angle_rad;
        std::memcpy(o.right_ws, &p.right_ws[0], 12); o.outer_angle_rad = p.outer_angle_rad;
        std::memcpy(o.up_ws, &p.up_ws[0], 12); o.tube_half_length = p.tube_half_length;
        o.rect_half_extents[0] = p.rect_half_extents.x; o.rect_half_extents[1] = p.rect_half_extents.y;
        o.tube_radius = p.tube_radius; o.attenuation_power = p.attenuation_power; o.attenuation_bias = p.attenuation_bias; o.attenuation_cutoff = p.attenuation_cutoff;
        o.attenuation_model = (uint32_t)p.attenuation_model; o.flags = p.flags;
        o.light_type = light.model ? (uint32_t)light.model->type() : 0u; // a light without a model contributes nothing
        return o;
    }

    class FlatDrawBatch
    {
    public:
        explicit FlatDrawBatch(shsb_ctx ctx) : ctx_(ctx) {}
        FlatDrawBatch(const FlatDrawBatch&) = delete;
        FlatDrawBatch& operator=(const FlatDrawBatch&) = delete;
        ~FlatDrawBatch()
        {
            for (auto& [key, mesh] : meshes_) shsb_mesh_destroy(ctx_, mesh);
            release_targets();
        }

        // per-object halves of the two reference draws; false: the mesh could not be uploaded (the draw is not recorded)
        bool draw_mesh_blinn_phong_transformed(const DebugMesh& mesh_local, const glm::mat4& model, const glm::vec3& base_color) { return record(mesh_local, model, base_color, nullptr); }
        bool draw_mesh_multi_light_transformed(const DebugMesh& mesh_local, const glm::mat4& model, const glm::vec3& base_color, const LightSelection& selection)
        {
            return record(mesh_local, model, base_color, &selection);
        }
        size_t size() const { return draws_.size(); }
        void clear() { draws_.clear(); }

        // per-frame halves: submit what was recorded, in order, and forget it
        bool flush_blinn_phong(RT_ColorLDR& rt, std::span<float> depth_buffer, const glm::mat4& vp, int canvas_w, int canvas_h, const glm::vec3& camera_pos, const glm::vec3& light_dir_ws)
        {
            return flush(rt, depth_buffer, canvas_w, canvas_h, [&](shsb_rt canvas, shsb_rt depth) {
                return shsb_flat_draw_blinn_phong(ctx_, draws_.data(), (uint32_t)draws_.size(), &vp[0][0], &camera_pos[0], &light_dir_ws[0], canvas, depth);
            });
        }
        bool flush_multi_light(RT_ColorLDR& rt, std::span<float> depth_buffer, const glm::mat4& vp, int canvas_w, int canvas_h, const glm::vec3& camera_pos,
                               const std::vector<LightInstance>& lights)
        {
            std::vector<ShsbLightProperties> props(lights.size());
            for (size_t i = 0; i < lights.size(); ++i) props[i] = to_shsb(lights[i]);
            return flush(rt, depth_buffer, canvas_w, canvas_h, [&](shsb_rt canvas, shsb_rt depth) {
                return shsb_flat_draw_multi_light(ctx_, draws_.data(), (uint32_t)draws_.size(), &vp[0][0], &camera_pos[0], props.data(), (uint32_t)props.size(), canvas, depth);
            });
        }

    private:
        bool record(const DebugMesh& mesh, const glm::mat4& model, const glm::vec3& base_color, const LightSelection* selection)
        {
            const auto key = std::make_tuple((const void*)&mesh, mesh.vertices.size(), mesh.indices.size());
            auto it = meshes_.find(key);
            if (it == meshes_.end())
            {
                static_assert(sizeof(glm::vec3) == 12, "DebugMesh::vertices is a packed float3 array");
                shsb_mesh h = 0;
                if (shsb_mesh_upload(ctx_, mesh.vertices.empty() ? nullptr : &mesh.vertices[0][0], (uint32_t)mesh.vertices.size(), nullptr, 0, nullptr, 0,
                                     mesh.indices.empty() ? nullptr : mesh.indices.data(), (uint32_t)mesh.indices.size(), &h) != SHSB_OK)
                    return false;
                it = meshes_.emplace(key, h).first;
            }
            ShsbFlatDraw d{};
            d.mesh = it->second;
            std::memcpy(d.model, &model[0][0], 64);
            std::memcpy(d.base_color, &base_color[0], 12);
            if (selection)
            {
                d.selection_count = selection->count < kLightSelectionCapacity ? selection->count : kLightSelectionCapacity;
                for (uint32_t k = 0; k < kLightSelectionCapacity; ++k) d.selection[k] = selection->indices[k];
            }
            draws_.push_back(d);
            return true;
        }

        void release_targets()
        {
            if (canvas_) shsb_rt_destroy(ctx_, canvas_);
            if (depth_) shsb_rt_destroy(ctx_, depth_);
            canvas_ = depth_ = 0;
        }

        template <typename Submit>
        bool flush(RT_ColorLDR& rt, std::span<float> depth_buffer, int w, int h, Submit&& submit)
        {
            const bool ok = [&] {
                if (draws_.empty()) return true;
                if (w <= 0 || h <= 0 || rt.w != w || rt.h != h || depth_buffer.size() < (size_t)w * (size_t)h) return false; // the reference indexes both with rt.w
                if (w != tw_ || h != th_)
                {
                    release_targets();
                    if (shsb_rt_create(ctx_, SHSB_RT_COLOR_LDR, w, h, 0.1f, 1000.0f, &canvas_) != SHSB_OK) return false;
                    if (shsb_rt_create(ctx_, SHSB_RT_SHADOW, w, h, 0.1f, 1000.0f, &depth_) != SHSB_OK) return false;
                    tw_ = w; th_ = h;
                }
                static_assert(sizeof(Color) == 4, "RT_ColorLDR holds RGBA8 texels");
                const size_t n = (size_t)w * (size_t)h;
                if (shsb_rt_upload(ctx_, canvas_, SHSB_PLANE_COLOR, &rt.color.at(0, 0), n * 4) != SHSB_OK) return false;
                if (shsb_rt_upload(ctx_, depth_, SHSB_PLANE_DEPTH, depth_buffer.data(), n * 4) != SHSB_OK) return false;
                if (submit(canvas_, depth_) != SHSB_OK) return false;
                std::vector<Color> colour(n);
                std::vector<float> depth(n);
                if (shsb_rt_download(ctx_, canvas_, SHSB_PLANE_COLOR, colour.data(), n * 4) != SHSB_OK) return false;
                if (shsb_rt_download(ctx_, depth_, SHSB_PLANE_DEPTH, depth.data(), n * 4) != SHSB_OK) return false;
                std::memcpy(&rt.color.at(0, 0), colour.data(), n * 4); // nothing is touched unless every step succeeded
                std::memcpy(depth_buffer.data(), depth.data(), n * 4);
                return true;
            }();
            draws_.clear();
            return ok;
        }

        shsb_ctx ctx_;
        std::map<std::tuple<const void*, size_t, size_t>, shsb_mesh> meshes_;
        std::vector<ShsbFlatDraw> draws_;
        shsb_rt canvas_ = 0, depth_ = 0;
        int tw_ = 0, th_ = 0;
    };
}
