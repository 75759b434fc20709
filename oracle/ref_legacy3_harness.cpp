// oracle/ref_legacy3_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with the product).
//
// The reference's legacy PBR / IBL demo -- the config-4 flavour, SURVEY.md section 8a row L3 -- compiled from its own source where it
// lies under /root/reference: cpp-folders/src/hello-render-target/hello_pbr.cpp (shadow-map raster :827-875, camera raster with
// near-plane clipping, perspective-correct varyings and motion vectors :883-1045, 2x2 PCF :599-621, Cook-Torrance + IBL fragment
// shader with in-shader tonemap :627-727) with hello-shs-renderer/shs_renderer.hpp and the library's IBL sampling
// (shs-renderer-lib/include/shs/resources/ibl.hpp:215-287).  Same recipe as oracle/ref_legacy_harness.cpp.  The skybox background
// pass and the motion-blur pass of the demo are separate full-frame passes and are not part of this checker.
// Built by oracle/Makefile (`make ref`) into oracle/_ref/libshs_legacy3_ref.so.
#include <cstdint>
#include <cstring>

#define main shs_legacy3_demo_main
#include "hello_pbr.cpp"
#undef main
#include "legacy_shim/stubs.inc"

namespace
{
    glm::mat4 load_mat4(const float* m) { glm::mat4 r; std::memcpy(&r, m, 64); return r; }
    glm::mat3 load_mat3(const float* m) { glm::mat3 r; std::memcpy(&r, m, 36); return r; }
    shs::CubeMapLinear load_cube(const float* data, int size)
    {
        shs::CubeMapLinear c;
        c.size = size;
        for (int f = 0; f < 6; ++f)
        {
            c.face[f].resize((size_t)size * size);
            std::memcpy(c.face[f].data(), data + (size_t)f * size * size * 3, (size_t)size * size * 12);
        }
        return c;
    }
}

extern "C"
{
    struct ShsoL3Uniforms // struct Uniforms + MaterialPBR, hello_pbr.cpp:474-519, as plain data
    {
        float mvp[16], prev_mvp[16], model[16], mv[16], normal_mat[9], light_vp[16];
        float light_dir_world[3], camera_pos[3];
        uint8_t base_color_srgb[4];
        float metallic, roughness, ao;
        int32_t use_texture;
        float ibl_diffuse_intensity, ibl_specular_intensity, ibl_reflection_strength;
    };

    int32_t shsref_l3_shadow_draw(const float* positions, uint32_t n_vertices, const float model[16], const float light_vp[16],
                                  int32_t sm_w, int32_t sm_h, int32_t tile_w, int32_t tile_h, float* shadow_depth)
    {
        if (!positions || !shadow_depth || sm_w <= 0 || sm_h <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        ShadowMap sm(sm_w, sm_h);
        // shs::ShadowMap keeps its buffer private: replay the incoming depths through test_and_set (the map is FLT_MAX after clear,
        // so every finite incoming value is stored as is)
        for (int y = 0; y < sm_h; ++y)
            for (int x = 0; x < sm_w; ++x) sm.test_and_set(x, y, shadow_depth[(size_t)y * sm_w + x]);
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
        for (int y = 0; y < sm_h; ++y)
            for (int x = 0; x < sm_w; ++x) shadow_depth[(size_t)y * sm_w + x] = sm.sample(x, y);
        return 0;
    }

    // irradiance: 6 faces x irr_size^2 x RGB floats; prefiltered: n_mips cube maps, mip m of size spec_sizes[m], concatenated.
    // velocity: width x height x 2 floats in shs::Buffer order (row 0 = bottom; set_screen_space flips y), read and written.
    int32_t shsref_l3_camera_draw(const float* positions, const float* normals, const float* uvs, uint32_t n_vertices, const ShsoL3Uniforms* un,
                                  const uint8_t* texture_rgba, int32_t tex_w, int32_t tex_h, const float* shadow_depth, int32_t sm_w, int32_t sm_h,
                                  const float* irradiance, int32_t irr_size, const float* prefiltered, const int32_t* spec_sizes, int32_t n_mips,
                                  int32_t width, int32_t height, int32_t tile_w, int32_t tile_h, uint8_t* canvas_rgba, float* zbuffer, float* velocity)
    {
        if (!positions || !normals || !uvs || !un || !canvas_rgba || !zbuffer || !velocity || width <= 0 || height <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        RT_ColorDepthMotion rt(width, height, 0.1f, 1000.0f);
        std::memcpy(rt.color.buffer().raw(), canvas_rgba, (size_t)width * height * 4);
        std::memcpy(rt.depth.buffer().raw(), zbuffer, (size_t)width * height * 4);
        static_assert(sizeof(glm::vec2) == 8, "velocity texels are two floats");
        std::memcpy(rt.velocity.raw(), velocity, (size_t)width * height * 8);
        shs::Texture2D tex;
        if (texture_rgba && tex_w > 0 && tex_h > 0)
        {
            tex = shs::Texture2D(tex_w, tex_h);
            std::memcpy(tex.texels.raw(), texture_rgba, (size_t)tex_w * tex_h * 4);
        }
        ShadowMap sm(std::max(1, sm_w), std::max(1, sm_h));
        const bool has_shadow = shadow_depth && sm_w > 0 && sm_h > 0;
        if (has_shadow)
            for (int y = 0; y < sm_h; ++y)
                for (int x = 0; x < sm_w; ++x) sm.test_and_set(x, y, shadow_depth[(size_t)y * sm_w + x]);
        EnvIBL ibl;
        const bool has_ibl = irradiance && irr_size > 0 && prefiltered && spec_sizes && n_mips > 0;
        if (has_ibl)
        {
            ibl.env_irradiance = load_cube(irradiance, irr_size);
            size_t off = 0;
            for (int m = 0; m < n_mips; ++m)
            {
                ibl.env_prefiltered_spec.mip.push_back(load_cube(prefiltered + off, spec_sizes[m]));
                off += (size_t)6 * spec_sizes[m] * spec_sizes[m] * 3;
            }
        }
        Uniforms u;
        u.mvp = load_mat4(un->mvp);
        u.prev_mvp = load_mat4(un->prev_mvp);
        u.model = load_mat4(un->model);
        u.view = glm::mat4(1.0f);
        u.mv = load_mat4(un->mv);
        u.normal_mat = load_mat3(un->normal_mat);
        u.light_vp = load_mat4(un->light_vp);
        u.light_dir_world = glm::vec3(un->light_dir_world[0], un->light_dir_world[1], un->light_dir_world[2]);
        u.camera_pos = glm::vec3(un->camera_pos[0], un->camera_pos[1], un->camera_pos[2]);
        u.mat.baseColor_srgb = shs::Color{un->base_color_srgb[0], un->base_color_srgb[1], un->base_color_srgb[2], un->base_color_srgb[3]};
        u.mat.metallic = un->metallic;
        u.mat.roughness = un->roughness;
        u.mat.ao = un->ao;
        u.albedo = tex.valid() ? &tex : nullptr;
        u.use_texture = un->use_texture != 0;
        u.shadow = has_shadow ? &sm : nullptr;
        u.sky = nullptr;
        u.ibl = has_ibl ? &ibl : nullptr;
        u.ibl_diffuse_intensity = un->ibl_diffuse_intensity;
        u.ibl_specular_intensity = un->ibl_specular_intensity;
        u.ibl_reflection_strength = un->ibl_reflection_strength;

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
                    draw_triangle_tile_color_depth_motion(
                        rt, tv, tn, tu,
                        [&u](const glm::vec3& p, const glm::vec3& n, const glm::vec2& uv) { return vertex_shader_full(p, n, uv, u); },
                        [&u](const VaryingsFull& v) { return fragment_shader_pbr(v, u); },
                        t_min, t_max);
                }
            }
        std::memcpy(canvas_rgba, rt.color.buffer().raw(), (size_t)width * height * 4);
        std::memcpy(zbuffer, rt.depth.buffer().raw(), (size_t)width * height * 4);
        std::memcpy(velocity, rt.velocity.raw(), (size_t)width * height * 8);
        return 0;
    }
}
