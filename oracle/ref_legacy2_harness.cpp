// oracle/ref_legacy2_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with the product).
//
// The reference's legacy SOFT-SHADOW demo -- the config-3 flavour, SURVEY.md section 8a row L2 -- compiled from its own source where
// it lies under /root/reference: cpp-folders/src/hello-render-target/hello_shadow_mapping_soft.cpp (shadow-map raster :796-839,
// camera raster with near-plane clipping and perspective-correct varyings :845-986, PCSS :333-445, fragment shader :991-1040) with
// the helpers of hello-shs-renderer/shs_renderer.hpp.  Same recipe as oracle/ref_legacy_harness.cpp: the demo's `main` is renamed and
// never called, SDL2 / Assimp names come from oracle/legacy_shim (declarations + aborting stubs), GLM from oracle/glm_shim.
// Built by oracle/Makefile (`make ref`) into oracle/_ref/libshs_legacy2_ref.so.
#include <cstdint>
#include <cstring>

#define main shs_legacy2_demo_main
#include "hello_shadow_mapping_soft.cpp"
#undef main
#include "legacy_shim/stubs.inc"

namespace
{
    glm::mat4 load_mat4(const float* m) { glm::mat4 r; std::memcpy(&r, m, 64); return r; }
    glm::mat3 load_mat3(const float* m) { glm::mat3 r; std::memcpy(&r, m, 36); return r; }
}

extern "C"
{
    struct ShsoL2Uniforms // struct Uniforms, hello_shadow_mapping_soft.cpp:714-732, as plain data
    {
        float mvp[16], model[16], mv[16], normal_mat[9], light_vp[16];
        float light_dir_world[3], camera_pos[3];
        uint8_t base_color[4];
        int32_t use_texture;
    };

    // PASS0 of RendererSystem::process (:1123-1200): one object's triangles into the shadow map, tile by tile.
    int32_t shsref_l2_shadow_draw(const float* positions, uint32_t n_vertices, const float model[16], const float light_vp[16],
                                  int32_t sm_w, int32_t sm_h, int32_t tile_w, int32_t tile_h, float* shadow_depth)
    {
        if (!positions || !shadow_depth || sm_w <= 0 || sm_h <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        ShadowMap sm(sm_w, sm_h);
        std::memcpy(sm.depth.raw(), shadow_depth, (size_t)sm_w * sm_h * 4);
        Uniforms u;
        u.model = load_mat4(model);
        u.light_vp = load_mat4(light_vp);
        const int cols = (sm_w + tile_w - 1) / tile_w, rows = (sm_h + tile_h - 1) / tile_h;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const glm::ivec2 t_min(tx * tile_w, ty * tile_h);
                const glm::ivec2 t_max(std::min((tx + 1) * tile_w, sm_w) - 1, std::min((ty + 1) * tile_h, sm_h) - 1);
                for (uint32_t i = 0; i + 2 < n_vertices; i += 3)
                {
                    const std::vector<glm::vec3> tri = {glm::vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]),
                                                        glm::vec3(positions[3 * i + 3], positions[3 * i + 4], positions[3 * i + 5]),
                                                        glm::vec3(positions[3 * i + 6], positions[3 * i + 7], positions[3 * i + 8])};
                    draw_triangle_tile_shadow(sm, tri, [&u](const glm::vec3& p) { return shadow_vertex_shader(p, u); }, t_min, t_max);
                }
            }
        std::memcpy(shadow_depth, sm.depth.raw(), (size_t)sm_w * sm_h * 4);
        return 0;
    }

    // PASS1 (:1205-1390): one object's triangles through draw_triangle_tile_color_depth_softshadow with vertex_shader_full and
    // fragment_shader_softshadow.  texture_rgba / shadow_depth may be null.  canvas in shs::Canvas order, zbuffer in shs::ZBuffer
    // order (note: this raster tests depth through test_and_set_depth_SCREEN_SPACE, i.e. the z-buffer rows are flipped too).
    int32_t shsref_l2_camera_draw(const float* positions, const float* normals, const float* uvs, uint32_t n_vertices, const ShsoL2Uniforms* un,
                                  const uint8_t* texture_rgba, int32_t tex_w, int32_t tex_h, const float* shadow_depth, int32_t sm_w, int32_t sm_h,
                                  int32_t width, int32_t height, int32_t tile_w, int32_t tile_h, uint8_t* canvas_rgba, float* zbuffer)
    {
        if (!positions || !normals || !uvs || !un || !canvas_rgba || !zbuffer || width <= 0 || height <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        shs::Canvas canvas(width, height);
        shs::ZBuffer zbuf(width, height, 0.1f, 1000.0f);
        std::memcpy(canvas.buffer().raw(), canvas_rgba, (size_t)width * height * 4);
        std::memcpy(zbuf.buffer().raw(), zbuffer, (size_t)width * height * 4);
        shs::Texture2D tex;
        if (texture_rgba && tex_w > 0 && tex_h > 0)
        {
            tex = shs::Texture2D(tex_w, tex_h);
            std::memcpy(tex.texels.raw(), texture_rgba, (size_t)tex_w * tex_h * 4);
        }
        ShadowMap sm;
        if (shadow_depth && sm_w > 0 && sm_h > 0)
        {
            sm.init(sm_w, sm_h);
            std::memcpy(sm.depth.raw(), shadow_depth, (size_t)sm_w * sm_h * 4);
        }
        Uniforms u;
        u.mvp = load_mat4(un->mvp);
        u.model = load_mat4(un->model);
        u.mv = load_mat4(un->mv);
        u.view = glm::mat4(1.0f); // not read by the shaders
        u.normal_mat = load_mat3(un->normal_mat);
        u.light_vp = load_mat4(un->light_vp);
        u.light_dir_world = glm::vec3(un->light_dir_world[0], un->light_dir_world[1], un->light_dir_world[2]);
        u.camera_pos = glm::vec3(un->camera_pos[0], un->camera_pos[1], un->camera_pos[2]);
        u.base_color = shs::Color{un->base_color[0], un->base_color[1], un->base_color[2], un->base_color[3]};
        u.albedo = tex.valid() ? &tex : nullptr;
        u.use_texture = un->use_texture != 0;
        u.shadow = (shadow_depth && sm_w > 0 && sm_h > 0) ? &sm : nullptr;

        const int cols = (width + tile_w - 1) / tile_w, rows = (height + tile_h - 1) / tile_h;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const glm::ivec2 t_min(tx * tile_w, ty * tile_h);
                const glm::ivec2 t_max(std::min((tx + 1) * tile_w, width) - 1, std::min((ty + 1) * tile_h, height) - 1);
                for (uint32_t i = 0; i + 2 < n_vertices; i += 3)
                {
                    std::vector<glm::vec3> tv(3), tn(3);
                    std::vector<glm::vec2> tu(3);
                    for (int k = 0; k < 3; ++k)
                    {
                        tv[k] = glm::vec3(positions[3 * (i + k)], positions[3 * (i + k) + 1], positions[3 * (i + k) + 2]);
                        tn[k] = glm::vec3(normals[3 * (i + k)], normals[3 * (i + k) + 1], normals[3 * (i + k) + 2]);
                        tu[k] = glm::vec2(uvs[2 * (i + k)], uvs[2 * (i + k) + 1]);
                    }
                    draw_triangle_tile_color_depth_softshadow(
                        canvas, zbuf, tv, tn, tu,
                        [&u](const glm::vec3& p, const glm::vec3& n, const glm::vec2& uv) { return vertex_shader_full(p, n, uv, u); },
                        [&u](const VaryingsFull& v, int px, int py) { return fragment_shader_softshadow(v, u, px, py); },
                        t_min, t_max);
                }
            }
        std::memcpy(canvas_rgba, canvas.buffer().raw(), (size_t)width * height * 4);
        std::memcpy(zbuffer, zbuf.buffer().raw(), (size_t)width * height * 4);
        return 0;
    }
}
