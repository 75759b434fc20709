// shs_b200/legacy_drop_in.hpp -- the reference-side binding of the LEGACY tile-job rasterizer (C++20, header-only): what a
// maintainer adds to cpp-folders/src/hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp (BASELINE configs[0] as shipped)
// to run its RendererSystem::process on a B200.  It speaks the demo's own types from hello-shs-renderer/shs_renderer.hpp --
// shs::Canvas, shs::ZBuffer, shs::ModelGeometry-style triangle / normal streams, glm matrices -- and forwards to the C-ABI
// (include/shsb.h: shsb_legacy_draw_blinn_phong).  The demo's Uniforms struct and shaders live in the demo's .cpp, not in a header,
// so the uniforms are passed field by field.
//
//     // RendererSystem::process(), :244-313 -- was: one job per 80x80 tile on the job system, draw_triangle_tile per triangle
//     shs::b200::legacy::Renderer gpu;                                  // once
//     gpu.begin_frame(*scene->canvas, *z_buffer);                       // uploads the (cleared) canvas and z-buffer
//     for (MonkeyObject* monkey : objects)
//         gpu.draw(monkey->geometry->triangles, monkey->geometry->normals, proj * view * monkey->get_world_matrix(),
//                  monkey->get_world_matrix(), scene->light_direction, viewer->position, monkey->color);
//     gpu.end_frame(*scene->canvas, *z_buffer);                         // downloads both; the SDL blit that follows is unchanged
//
// Needs shs_renderer.hpp (and therefore GLM, SDL2, Assimp) on the include path: it is compiled only inside the reference tree -- in
// this repository by tests/cpp/Makefile against oracle/glm_shim + oracle/legacy_shim (tests/cpp/legacy_drop_in_test.cpp).
// No CPU fallback: without a device every call is a no-op that reports false.
#pragma once

#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>

#include "shs_renderer.hpp"

#include "shsb.h"

namespace shs::b200::legacy
{
    class Renderer
    {
    public:
        explicit Renderer(int cuda_device = 0, int job_tile_w = 80, int job_tile_h = 80) : job_w_(job_tile_w), job_h_(job_tile_h)
        {
            ok_ = shsb_context_create(cuda_device, &ctx_) == SHSB_OK;
        }
        ~Renderer() { if (ctx_) shsb_context_destroy(ctx_); }
        Renderer(const Renderer&) = delete;
        Renderer& operator=(const Renderer&) = delete;

        bool valid() const { return ok_; }
        const char* last_error() const { return ctx_ ? shsb_last_error_string(ctx_) : "no CUDA device (there is no CPU fallback)"; }

        // Mirrors the host canvas / z-buffer into their device twins (created on first use, re-created on a size change).
        bool begin_frame(shs::Canvas& canvas, shs::ZBuffer& zbuf)
        {
            if (!ok_) return false;
            const int w = canvas.get_width(), h = canvas.get_height();
            if (w <= 0 || h <= 0 || zbuf.get_width() != w || zbuf.get_height() != h) return false;
            if (w != w_ || h != h_)
            {
                if (canvas_rt_) shsb_rt_destroy(ctx_, canvas_rt_);
                if (z_rt_) shsb_rt_destroy(ctx_, z_rt_);
                canvas_rt_ = z_rt_ = 0;
                if (shsb_rt_create(ctx_, SHSB_RT_COLOR_LDR, w, h, 0.1f, 1000.0f, &canvas_rt_) != SHSB_OK) return false;
                if (shsb_rt_create(ctx_, SHSB_RT_SHADOW, w, h, 0.1f, 1000.0f, &z_rt_) != SHSB_OK) return false;
                w_ = w; h_ = h;
            }
            static_assert(sizeof(shs::Color) == 4, "Canvas texels are RGBA8");
            const size_t n = (size_t)w * (size_t)h;
            return shsb_rt_upload(ctx_, canvas_rt_, SHSB_PLANE_COLOR, canvas.buffer().raw(), n * 4) == SHSB_OK &&
                   shsb_rt_upload(ctx_, z_rt_, SHSB_PLANE_DEPTH, zbuf.buffer().raw(), n * 4) == SHSB_OK;
        }

        // One object of RendererSystem::process: every triangle of the (triangles, normals) streams through the legacy vertex /
        // fragment shaders.  The streams are uploaded once per distinct `triangles` vector (keyed by its data pointer, like the demo
        // keeps one ModelGeometry per object).
        bool draw(const std::vector<glm::vec3>& triangles, const std::vector<glm::vec3>& normals, const glm::mat4& mvp, const glm::mat4& model,
                  const glm::vec3& light_dir, const glm::vec3& camera_pos, shs::Color color)
        {
            if (!ok_ || !canvas_rt_ || triangles.size() < 3 || normals.size() < triangles.size()) return false;
            static_assert(sizeof(glm::vec3) == 12, "vertex streams are tightly packed");
            shsb_mesh mesh = 0;
            const auto it = meshes_.find(triangles.data());
            if (it != meshes_.end()) mesh = it->second;
            else
            {
                if (shsb_mesh_upload(ctx_, &triangles[0].x, (uint32_t)triangles.size(), &normals[0].x, (uint32_t)normals.size(), nullptr, 0, nullptr, 0, &mesh) != SHSB_OK) return false;
                meshes_[triangles.data()] = mesh;
            }
            ShsbLegacyUniforms u{};
            std::memcpy(u.mvp, &mvp, 64);
            std::memcpy(u.model, &model, 64);
            u.light_dir[0] = light_dir.x; u.light_dir[1] = light_dir.y; u.light_dir[2] = light_dir.z;
            u.camera_pos[0] = camera_pos.x; u.camera_pos[1] = camera_pos.y; u.camera_pos[2] = camera_pos.z;
            u.color[0] = color.r; u.color[1] = color.g; u.color[2] = color.b; u.color[3] = color.a;
            u.job_tile_w = job_w_; u.job_tile_h = job_h_;
            return shsb_legacy_draw_blinn_phong(ctx_, mesh, &u, canvas_rt_, z_rt_) == SHSB_OK;
        }

        bool end_frame(shs::Canvas& canvas, shs::ZBuffer& zbuf)
        {
            if (!ok_ || !canvas_rt_ || canvas.get_width() != w_ || canvas.get_height() != h_) return false;
            const size_t n = (size_t)w_ * (size_t)h_;
            return shsb_rt_download(ctx_, canvas_rt_, SHSB_PLANE_COLOR, canvas.buffer().raw(), n * 4) == SHSB_OK &&
                   shsb_rt_download(ctx_, z_rt_, SHSB_PLANE_DEPTH, zbuf.buffer().raw(), n * 4) == SHSB_OK;
        }

    private:
        shsb_ctx ctx_ = nullptr;
        bool ok_ = false;
        int job_w_, job_h_;
        int w_ = 0, h_ = 0;
        shsb_rt canvas_rt_ = 0, z_rt_ = 0;
        std::unordered_map<const void*, shsb_mesh> meshes_{};
    };
}
