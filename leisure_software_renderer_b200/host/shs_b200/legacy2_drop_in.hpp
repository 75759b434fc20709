// shs_b200/legacy2_drop_in.hpp -- the reference-side binding of the LEGACY RENDER-TARGET demos (C++20, header-only): what a
// maintainer adds to cpp-folders/src/hello-render-target/hello_shadow_mapping_soft.cpp (2048^2 shadow map + PCSS lit pass; the
// config-3 flavour) or hello_pbr.cpp (Cook-Torrance + IBL lit pass with motion vectors; the config-4
// flavour, row L3) to run RendererSystem::process's two raster passes on a B200.  It speaks the demos' own types from
// hello-shs-renderer/shs_renderer.hpp -- shs::Canvas, shs::ZBuffer, shs::Buffer<glm::vec2>, shs::Texture2D, triangle / normal / uv
// streams, glm matrices -- and forwards to the C-ABI (include/shsb.h: shsb_legacy2_* / shsb_legacy3_*).  `struct Uniforms` lives in
// each demo's .cpp, not in a header, so the draw calls are templates over it and read the members both demos name alike.
//
//     // RendererSystem::process() of hello_shadow_mapping_soft.cpp, :1100-1370 -- was: two job-system fan-outs over 160x160 tiles
//     shs::b200::legacy2::Renderer gpu;                                          // once
//     gpu.begin_shadow(SHADOW_MAP_SIZE, SHADOW_MAP_SIZE);                        // PASS0: ShadowMap::clear on the device
//     gpu.shadow_draw(floor->verts, floor_model, light_vp);                      // was draw_triangle_tile_shadow per tile per triangle
//     gpu.shadow_draw(monkey->geometry->triangles, monkey_model, light_vp);
//     gpu.begin_frame(*rt_color, *rt_depth);                                     // PASS1: uploads the cleared canvas and z-buffer
//     gpu.draw_softshadow(floor->verts, floor->norms, floor->uvs, u_floor);      // u_*: the demo's own Uniforms, filled as before
//     gpu.draw_softshadow(monkey->geometry->triangles, monkey->geometry->normals, monkey->geometry->uvs, u_monkey);
//     gpu.end_frame(*rt_color, *rt_depth);                                       // downloads both; the blit that follows is unchanged
//
//     // hello_pbr.cpp: the same with gpu.set_ibl(env_ibl) once, begin_frame(rt.color, rt.depth, &rt.velocity),
//     // draw_pbr(..., u) and end_frame(rt.color, rt.depth, &rt.velocity); the skybox and motion-blur passes that follow are unchanged.
//
// `Uniforms::shadow != nullptr` selects the device shadow map of the last begin_shadow (the host ShadowMap object is not read);
// `Uniforms::albedo` textures are uploaded once per distinct Texture2D; `Uniforms::ibl != nullptr` selects the set_ibl environment.
// Needs shs_renderer.hpp (GLM, SDL2, Assimp) on the include path: compiled only inside the reference tree -- in this repository by
// tests/cpp/Makefile against the declaration shims (tests/cpp/legacy2_drop_in_test.cpp).  No CPU fallback: without a device every
// call is a no-op that reports false.
#pragma once

#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>

#include "shs_renderer.hpp"

#include "shsb.h"

namespace shs::b200::legacy2
{
    class Renderer
    {
    public:
        explicit Renderer(int cuda_device = 0, int job_tile_w = 160, int job_tile_h = 160) : job_w_(job_tile_w), job_h_(job_tile_h)
        {
            ok_ = shsb_context_create(cuda_device, &ctx_) == SHSB_OK;
        }
        ~Renderer() { if (ctx_) shsb_context_destroy(ctx_); }
        Renderer(const Renderer&) = delete;
        Renderer& operator=(const Renderer&) = delete;

        bool valid() const { return ok_; }
        const char* last_error() const { return ctx_ ? shsb_last_error_string(ctx_) : "no CUDA device (there is no CPU fallback)"; }

        // ---- PASS0
        bool begin_shadow(int w, int h) // ShadowMap(w, h) + ShadowMap::clear
        {
            if (!ok_ || w <= 0 || h <= 0) return false;
            if (w != sm_w_ || h != sm_h_)
            {
                if (shadow_rt_) shsb_rt_destroy(ctx_, shadow_rt_);
                shadow_rt_ = 0;
                if (shsb_rt_create(ctx_, SHSB_RT_SHADOW, w, h, 0.1f, 1000.0f, &shadow_rt_) != SHSB_OK) return false;
                sm_w_ = w; sm_h_ = h;
            }
            const float far_ = std::numeric_limits<float>::max();
            return shsb_rt_clear(ctx_, shadow_rt_, SHSB_PLANE_DEPTH, &far_) == SHSB_OK;
        }

        bool shadow_draw(const std::vector<glm::vec3>& triangles, const glm::mat4& model, const glm::mat4& light_vp)
        {
            if (!ok_ || !shadow_rt_ || triangles.size() < 3) return false;
            const shsb_mesh mesh = mesh_of(triangles, nullptr, nullptr);
            return mesh && shsb_legacy2_shadow_draw(ctx_, mesh, &model[0][0], &light_vp[0][0], job_w_, job_h_, shadow_rt_) == SHSB_OK;
        }

        // The device shadow map back into a host map (ShadowMap keeps its buffer behind test_and_set: call this on a cleared map).
        template <class ShadowMapT>
        bool download_shadow(ShadowMapT& sm)
        {
            if (!ok_ || !shadow_rt_ || sm.w != sm_w_ || sm.h != sm_h_) return false;
            std::vector<float> tmp((size_t)sm_w_ * (size_t)sm_h_);
            if (shsb_rt_download(ctx_, shadow_rt_, SHSB_PLANE_DEPTH, tmp.data(), tmp.size() * 4) != SHSB_OK) return false;
            for (int y = 0; y < sm_h_; ++y)
                for (int x = 0; x < sm_w_; ++x) sm.test_and_set(x, y, tmp[(size_t)y * sm_w_ + x]);
            return true;
        }

        // ---- PASS1: mirrors the host canvas / z-buffer (/ velocity buffer) into their device twins
        bool begin_frame(shs::Canvas& canvas, shs::ZBuffer& zbuf, shs::Buffer<glm::vec2>* velocity = nullptr)
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
                if (shsb_rt_create(ctx_, SHSB_RT_DEPTH_MOTION, w, h, 0.1f, 1000.0f, &z_rt_) != SHSB_OK) return false;
                w_ = w; h_ = h;
            }
            static_assert(sizeof(shs::Color) == 4 && sizeof(glm::vec2) == 8, "Canvas texels are RGBA8, velocity texels two floats");
            const size_t n = (size_t)w * (size_t)h;
            if (shsb_rt_upload(ctx_, canvas_rt_, SHSB_PLANE_COLOR, canvas.buffer().raw(), n * 4) != SHSB_OK) return false;
            if (shsb_rt_upload(ctx_, z_rt_, SHSB_PLANE_DEPTH, zbuf.buffer().raw(), n * 4) != SHSB_OK) return false;
            return !velocity || shsb_rt_upload(ctx_, z_rt_, SHSB_PLANE_MOTION, velocity->raw(), n * 8) == SHSB_OK;
        }

        // EnvIBL of hello_pbr.cpp (:455-461): anything with env_irradiance (shs::CubeMapLinear) and env_prefiltered_spec.mip
        template <class EnvIblT>
        bool set_ibl(const EnvIblT& env)
        {
            if (!ok_) return false;
            if (ibl_) { shsb_legacy3_ibl_destroy(ctx_, ibl_); ibl_ = 0; }
            std::vector<float> irr, pre;
            std::vector<int32_t> sizes;
            auto append = [](std::vector<float>& out, const auto& cube) {
                for (int f = 0; f < 6; ++f)
                    for (const glm::vec3& c : cube.face[f]) { out.push_back(c.x); out.push_back(c.y); out.push_back(c.z); }
            };
            if (!env.env_irradiance.valid() || !env.env_prefiltered_spec.valid()) return false;
            append(irr, env.env_irradiance);
            for (const auto& m : env.env_prefiltered_spec.mip) { if (!m.valid()) return false; append(pre, m); sizes.push_back(m.size); }
            return shsb_legacy3_ibl_upload(ctx_, irr.data(), env.env_irradiance.size, pre.data(), sizes.data(), (int32_t)sizes.size(), &ibl_) == SHSB_OK;
        }

        // One object of the lit pass of hello_shadow_mapping_soft.cpp: draw_triangle_tile_color_depth_softshadow over every job tile
        template <class UniformsT>
        bool draw_softshadow(const std::vector<glm::vec3>& triangles, const std::vector<glm::vec3>& normals, const std::vector<glm::vec2>& uvs, const UniformsT& u)
        {
            ShsbLegacy2Uniforms cu{};
            shsb_mesh mesh = 0;
            if (!prepare(triangles, normals, uvs, u, u.base_color, cu, mesh)) return false;
            return shsb_legacy2_draw_softshadow(ctx_, mesh, &cu, u.shadow ? shadow_rt_ : 0, canvas_rt_, z_rt_) == SHSB_OK;
        }

        // One object of the lit pass of hello_pbr.cpp: draw_triangle_tile_color_depth_motion over every job tile
        template <class UniformsT>
        bool draw_pbr(const std::vector<glm::vec3>& triangles, const std::vector<glm::vec3>& normals, const std::vector<glm::vec2>& uvs, const UniformsT& u)
        {
            ShsbLegacy2Uniforms cu{};
            shsb_mesh mesh = 0;
            if (!prepare(triangles, normals, uvs, u, u.mat.baseColor_srgb, cu, mesh)) return false;
            std::memcpy(cu.prev_mvp, &u.prev_mvp, 64);
            cu.metallic = u.mat.metallic; cu.roughness = u.mat.roughness; cu.ao = u.mat.ao;
            cu.ibl_diffuse_intensity = u.ibl_diffuse_intensity;
            cu.ibl_specular_intensity = u.ibl_specular_intensity;
            cu.ibl_reflection_strength = u.ibl_reflection_strength;
            return shsb_legacy3_draw_pbr(ctx_, mesh, &cu, u.shadow ? shadow_rt_ : 0, u.ibl ? ibl_ : 0, canvas_rt_, z_rt_) == SHSB_OK;
        }

        bool end_frame(shs::Canvas& canvas, shs::ZBuffer& zbuf, shs::Buffer<glm::vec2>* velocity = nullptr)
        {
            if (!ok_ || !canvas_rt_ || canvas.get_width() != w_ || canvas.get_height() != h_) return false;
            const size_t n = (size_t)w_ * (size_t)h_;
            if (shsb_rt_download(ctx_, canvas_rt_, SHSB_PLANE_COLOR, canvas.buffer().raw(), n * 4) != SHSB_OK) return false;
            if (shsb_rt_download(ctx_, z_rt_, SHSB_PLANE_DEPTH, zbuf.buffer().raw(), n * 4) != SHSB_OK) return false;
            return !velocity || shsb_rt_download(ctx_, z_rt_, SHSB_PLANE_MOTION, velocity->raw(), n * 8) == SHSB_OK;
        }

    private:
        // streams are uploaded once per distinct `triangles` vector (keyed by its data pointer and which streams came with it)
        shsb_mesh mesh_of(const std::vector<glm::vec3>& triangles, const std::vector<glm::vec3>* normals, const std::vector<glm::vec2>* uvs)
        {
            static_assert(sizeof(glm::vec3) == 12, "vertex streams are tightly packed");
            auto& table = normals ? lit_meshes_ : shadow_meshes_;
            const auto it = table.find(triangles.data());
            if (it != table.end()) return it->second;
            shsb_mesh mesh = 0;
            const bool has_uv = uvs && uvs->size() >= triangles.size();
            if (shsb_mesh_upload(ctx_, &triangles[0].x, (uint32_t)triangles.size(), normals ? &(*normals)[0].x : nullptr, normals ? (uint32_t)normals->size() : 0,
                                 has_uv ? &(*uvs)[0].x : nullptr, has_uv ? (uint32_t)uvs->size() : 0, nullptr, 0, &mesh) != SHSB_OK)
                return 0;
            table[triangles.data()] = mesh;
            return mesh;
        }

        template <class UniformsT>
        bool prepare(const std::vector<glm::vec3>& triangles, const std::vector<glm::vec3>& normals, const std::vector<glm::vec2>& uvs, const UniformsT& u,
                     shs::Color base, ShsbLegacy2Uniforms& cu, shsb_mesh& mesh)
        {
            if (!ok_ || !canvas_rt_ || triangles.size() < 3 || normals.size() < triangles.size()) return false;
            mesh = mesh_of(triangles, &normals, &uvs);
            if (!mesh) return false;
            std::memcpy(cu.mvp, &u.mvp, 64);
            std::memcpy(cu.model, &u.model, 64);
            std::memcpy(cu.mv, &u.mv, 64);
            static_assert(sizeof(glm::mat3) == 36, "mat3 is nine packed floats");
            std::memcpy(cu.normal_mat, &u.normal_mat, 36);
            std::memcpy(cu.light_vp, &u.light_vp, 64);
            cu.light_dir_world[0] = u.light_dir_world.x; cu.light_dir_world[1] = u.light_dir_world.y; cu.light_dir_world[2] = u.light_dir_world.z;
            cu.camera_pos[0] = u.camera_pos.x; cu.camera_pos[1] = u.camera_pos.y; cu.camera_pos[2] = u.camera_pos.z;
            cu.base_color[0] = base.r; cu.base_color[1] = base.g; cu.base_color[2] = base.b; cu.base_color[3] = base.a;
            cu.use_texture = (u.use_texture && u.albedo && u.albedo->valid()) ? 1 : 0; // the fragment shaders' own condition
            if (cu.use_texture)
            {
                const auto it = textures_.find(u.albedo);
                if (it != textures_.end()) cu.albedo = it->second;
                else
                {
                    if (shsb_texture_upload(ctx_, reinterpret_cast<const uint8_t*>(u.albedo->texels.raw()), u.albedo->w, u.albedo->h, &cu.albedo) != SHSB_OK) return false;
                    textures_[u.albedo] = cu.albedo;
                }
            }
            cu.job_tile_w = job_w_; cu.job_tile_h = job_h_;
            return true;
        }

        shsb_ctx ctx_ = nullptr;
        bool ok_ = false;
        int job_w_, job_h_;
        int w_ = 0, h_ = 0, sm_w_ = 0, sm_h_ = 0;
        shsb_rt canvas_rt_ = 0, z_rt_ = 0, shadow_rt_ = 0;
        shsb_ibl ibl_ = 0;
        std::unordered_map<const void*, shsb_mesh> lit_meshes_{}, shadow_meshes_{};
        std::unordered_map<const void*, shsb_tex> textures_{};
    };
}
