// oracle/ref_legacy_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with the product).
//
// The reference's LEGACY tile-job rasterizer -- BASELINE.json configs[0] "as shipped", SURVEY.md section 8a row L1 -- compiled from
// the reference's own sources where they lie under /root/reference:
//     cpp-folders/src/hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp   (the demo program: Uniforms, the Blinn-Phong
//         vertex / fragment shaders :48-96, RendererSystem::draw_triangle_tile :189-242, the tile fan-out of ::process :244-313)
//     cpp-folders/src/hello-shs-renderer/shs_renderer.hpp                          (Canvas, ZBuffer, barycentric_coordinate,
//         clip_to_screen :802-831, Camera3D / Viewer :1210-1355)
// The demo's `main` is renamed and never called; SDL2 / SDL2_image / Assimp / GLM are third-party dependencies of the reference
// that are absent from this container: oracle/legacy_shim declares the names the two sources mention (no behaviour), this file
// defines them as aborting stubs so that the library loads, and oracle/glm_shim states GLM's arithmetic.
// The raster functions pinned here touch none of them.  No reference source is copied: the .cpp is #included from its own tree.
//
// Built by oracle/Makefile (`make ref`) into oracle/_ref/libshs_legacy_ref.so (git-ignored; travels to the GPU box as a binary).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define main shs_legacy_demo_main
#include "hello_pipeline_blinn_phong_shading.cpp"
#undef main

#include "legacy_shim/stubs.inc" // aborting definitions of the declared-only third-party names (ctypes loads with RTLD_NOW)

namespace
{
    glm::mat4 load_mat4(const float* m) { glm::mat4 r; std::memcpy(&r, m, 64); return r; }
    void store_mat4(const glm::mat4& m, float* out) { std::memcpy(out, &m, 64); }
}

extern "C"
{
    // Viewer(position, speed, width, height) + Camera3D::update (shs_renderer.hpp:1210-1236, 1322-1346) as the demo builds them
    // (hello_pipeline_blinn_phong_shading.cpp:384): fov 60 deg, aspect hard-coded 4/3, z 0.1 .. 1000, LH.
    void shsref_legacy_camera(const float position[3], float horizontal_angle_deg, float vertical_angle_deg, float out_view[16], float out_proj[16])
    {
        shs::Viewer viewer(glm::vec3(position[0], position[1], position[2]), 50.0f, 10.0f, 10.0f);
        viewer.horizontal_angle = horizontal_angle_deg;
        viewer.vertical_angle = vertical_angle_deg;
        viewer.update();
        store_mat4(viewer.camera->view_matrix, out_view);
        store_mat4(viewer.camera->projection_matrix, out_proj);
    }

    // MonkeyObject::get_world_matrix (:122-128): t * r * s from three separate identity-based matrices.
    void shsref_legacy_world_matrix(const float position[3], const float scale[3], float rotation_angle_deg, float out_model[16])
    {
        const glm::mat4 t = glm::translate(glm::mat4(1.0f), glm::vec3(position[0], position[1], position[2]));
        const glm::mat4 r = glm::rotate(glm::mat4(1.0f), glm::radians(rotation_angle_deg), glm::vec3(0.0f, 1.0f, 0.0f));
        const glm::mat4 s = glm::scale(glm::mat4(1.0f), glm::vec3(scale[0], scale[1], scale[2]));
        store_mat4(t * r * s, out_model);
    }

    // uniforms.mvp = proj * view * uniforms.model (:273)
    void shsref_legacy_mvp(const float proj[16], const float view[16], const float model[16], float out_mvp[16])
    {
        store_mat4(load_mat4(proj) * load_mat4(view) * load_mat4(model), out_mvp);
    }

    // One object through RendererSystem::process's tile fan-out (:244-313), serially: for every (tile_w x tile_h) tile, for every
    // triangle, the reference's own static RendererSystem::draw_triangle_tile with the reference's own shaders.  `canvas_rgba` is
    // shs::Canvas's buffer (row 0 = bottom of the screen, draw_pixel_screen_space flips y), `zbuffer` is shs::ZBuffer's buffer
    // (row = screen y, top-down; cleared to FLT_MAX by the demo); both are read and written.
    int32_t shsref_legacy_draw(const float* positions, const float* normals, uint32_t n_vertices, const float mvp[16], const float model[16],
                               const float light_dir[3], const float camera_pos[3], const uint8_t color[4], int32_t width, int32_t height,
                               int32_t tile_w, int32_t tile_h, uint8_t* canvas_rgba, float* zbuffer)
    {
        if (!positions || !normals || !canvas_rgba || !zbuffer || width <= 0 || height <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        shs::Canvas canvas(width, height);
        shs::ZBuffer zbuf(width, height, 0.1f, 1000.0f);
        static_assert(sizeof(shs::Color) == 4, "Canvas texels are RGBA8");
        std::memcpy(canvas.buffer().raw(), canvas_rgba, (size_t)width * height * 4);
        std::memcpy(zbuf.buffer().raw(), zbuffer, (size_t)width * height * 4);

        Uniforms uniforms;
        uniforms.model = load_mat4(model);
        uniforms.mvp = load_mat4(mvp);
        uniforms.light_dir = glm::vec3(light_dir[0], light_dir[1], light_dir[2]);
        uniforms.camera_pos = glm::vec3(camera_pos[0], camera_pos[1], camera_pos[2]);
        uniforms.color = shs::Color{color[0], color[1], color[2], color[3]};

        const int cols = (width + tile_w - 1) / tile_w, rows = (height + tile_h - 1) / tile_h;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const glm::ivec2 t_min(tx * tile_w, ty * tile_h);
                const glm::ivec2 t_max(std::min((tx + 1) * tile_w, width) - 1, std::min((ty + 1) * tile_h, height) - 1);
                for (uint32_t i = 0; i + 2 < n_vertices; i += 3)
                {
                    const std::vector<glm::vec3> tri_verts = {glm::vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]),
                                                              glm::vec3(positions[3 * i + 3], positions[3 * i + 4], positions[3 * i + 5]),
                                                              glm::vec3(positions[3 * i + 6], positions[3 * i + 7], positions[3 * i + 8])};
                    const std::vector<glm::vec3> tri_norms = {glm::vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]),
                                                              glm::vec3(normals[3 * i + 3], normals[3 * i + 4], normals[3 * i + 5]),
                                                              glm::vec3(normals[3 * i + 6], normals[3 * i + 7], normals[3 * i + 8])};
                    RendererSystem::draw_triangle_tile(
                        canvas, zbuf, tri_verts, tri_norms,
                        [&uniforms](const glm::vec3& p, const glm::vec3& n) { return blinn_phong_vertex_shader(p, n, uniforms); },
                        [&uniforms](const shs::Varyings& v) { return blinn_phong_fragment_shader(v, uniforms); },
                        t_min, t_max);
                }
            }
        std::memcpy(canvas_rgba, canvas.buffer().raw(), (size_t)width * height * 4);
        std::memcpy(zbuffer, zbuf.buffer().raw(), (size_t)width * height * 4);
        return 0;
    }

    // The CPU baseline of SURVEY.md section 8d (B2): RendererSystem::process as shipped -- z-buffer clear, then ONE JOB PER TILE on the
    // reference's own shs::Job::ThreadedPriorityJobSystem with `n_threads` workers and its WaitGroup (:244-313) -- repeated `frames`
    // times on one object.  Returns the mean milliseconds per frame; the last frame is left in canvas_rgba / zbuffer.
    double shsref_legacy_frames_threaded(const float* positions, const float* normals, uint32_t n_vertices, const float mvp[16], const float model[16],
                                         const float light_dir[3], const float camera_pos[3], const uint8_t color[4], int32_t width, int32_t height,
                                         int32_t tile_w, int32_t tile_h, int32_t n_threads, int32_t frames, uint8_t* canvas_rgba, float* zbuffer)
    {
        if (!positions || !normals || !canvas_rgba || !zbuffer || width <= 0 || height <= 0 || tile_w <= 0 || tile_h <= 0 || n_threads <= 0 || frames <= 0) return -1.0;
        shs::Canvas canvas(width, height);
        shs::ZBuffer zbuf(width, height, 0.1f, 1000.0f);
        std::vector<glm::vec3> verts(n_vertices), norms(n_vertices);
        for (uint32_t i = 0; i < n_vertices; ++i)
        {
            verts[i] = glm::vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
            norms[i] = glm::vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        }
        Uniforms uniforms;
        uniforms.model = load_mat4(model);
        uniforms.mvp = load_mat4(mvp);
        uniforms.light_dir = glm::vec3(light_dir[0], light_dir[1], light_dir[2]);
        uniforms.camera_pos = glm::vec3(camera_pos[0], camera_pos[1], camera_pos[2]);
        uniforms.color = shs::Color{color[0], color[1], color[2], color[3]};
        auto* jobs = new shs::Job::ThreadedPriorityJobSystem(n_threads);
        shs::Job::WaitGroup wait_group;
        const int cols = (width + tile_w - 1) / tile_w, rows = (height + tile_h - 1) / tile_h;
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; ++f)
        {
            shs::Canvas::fill_pixel(canvas, 0, 0, width, height, shs::Color::black()); // the demo's main loop (:452)
            zbuf.clear();
            wait_group.reset();
            for (int ty = 0; ty < rows; ++ty)
                for (int tx = 0; tx < cols; ++tx)
                {
                    wait_group.add(1);
                    jobs->submit({[&, tx, ty]() {
                        const glm::ivec2 t_min(tx * tile_w, ty * tile_h);
                        const glm::ivec2 t_max(std::min((tx + 1) * tile_w, width) - 1, std::min((ty + 1) * tile_h, height) - 1);
                        for (size_t i = 0; i + 2 < verts.size(); i += 3)
                        {
                            const std::vector<glm::vec3> tri_verts = {verts[i], verts[i + 1], verts[i + 2]};
                            const std::vector<glm::vec3> tri_norms = {norms[i], norms[i + 1], norms[i + 2]};
                            RendererSystem::draw_triangle_tile(
                                canvas, zbuf, tri_verts, tri_norms,
                                [&uniforms](const glm::vec3& p, const glm::vec3& n) { return blinn_phong_vertex_shader(p, n, uniforms); },
                                [&uniforms](const shs::Varyings& v) { return blinn_phong_fragment_shader(v, uniforms); },
                                t_min, t_max);
                        }
                        wait_group.done();
                    }, shs::Job::PRIORITY_HIGH});
                }
            wait_group.wait();
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / frames;
        delete jobs;
        std::memcpy(canvas_rgba, canvas.buffer().raw(), (size_t)width * height * 4);
        std::memcpy(zbuffer, zbuf.buffer().raw(), (size_t)width * height * 4);
        return ms;
    }
}
