// tests/cpp/legacy2_emul.cpp -- TEST INFRASTRUCTURE ONLY.  The device functions of leisure_software_renderer_b200/csrc/legacy2_core.cuh
// compiled by g++ (-ffp-contract=off == nvcc --fmad=false) and driven exactly like legacy2.cu drives them: set-up per source triangle
// into slot arrays, then -- per PIXEL, not per triangle -- a walk over the slots in draw order keeping (running minimum depth, last
// shadeable prefix minimum) and one shading call at the end.  It exists so that the order-independent reformulation of the demos'
// serial tile-job loops is checked against the pinned oracle on the CPU box, where no GPU is available; nothing in the product
// links or loads it (the product path launches the kernels of legacy2.cu and fails without the CUDA library).
// Same C signatures as oracle/oracle_legacy.cpp's shso_l2_* / shso_l3_* with the prefix shsemu_.
// Two builds: -DEMUL_LIBM (sinf / cosf of the host libm in the PCSS rotation: expected to equal the oracle bit for bit) and the
// default (the device's flavour: double-precision sin / cos rounded to float; counts how often that shows).
#include <cmath>
#include <cstring>
#include <vector>
#ifdef EMUL_LIBM
#define L2_SINCOSF(a, s, c) do { (s) = sinf(a); (c) = cosf(a); } while (0)
#endif
#include "legacy2_core.cuh"

using namespace shsb::l2;

namespace
{
    void mat_mul(const float* a, const float* b, float* out) // glm: column i of a * b
    {
        float r[16];
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k) r[i * 4 + k] = a[k] * b[i * 4] + a[4 + k] * b[i * 4 + 1] + a[8 + k] * b[i * 4 + 2] + a[12 + k] * b[i * 4 + 3];
        std::memcpy(out, r, 64);
    }

    void run(const Draw& d, unsigned char* canvas, float* zbuf, float* velocity)
    {
        const uint32_t n_slots = d.mode == MODE_SHADOW ? d.n_tris : 2u * d.n_tris;
        std::vector<RasterRec> rr(n_slots);
        std::vector<BoxRec> bb(n_slots);
        std::vector<ShadeRec> ss(d.mode == MODE_SHADOW ? 1 : n_slots);
        for (uint32_t t = 0; t < d.n_tris; ++t)
        {
            if (d.mode == MODE_SHADOW) setup_shadow(d, t, rr[t], bb[t]);
            else setup_camera(d, t, &rr[2 * (size_t)t], &bb[2 * (size_t)t], &ss[2 * (size_t)t]);
        }
        for (int py = 0; py < d.H; ++py)
            for (int px = 0; px < d.W; ++px)
            {
                const int jx0 = (px / d.job_w) * d.job_w, jx1 = std::min(jx0 + d.job_w, d.W) - 1;
                const int jy0 = (py / d.job_h) * d.job_h, jy1 = std::min(jy0 + d.job_h, d.H) - 1;
                const size_t row = (d.mode == MODE_SHADOW) ? (size_t)py : (size_t)((d.H - 1) - py);
                const size_t at = row * (size_t)d.W + (size_t)px;
                PixelState st;
                st.best_z = zbuf[at];
                st.shade_slot = 0xFFFFFFFFu;
                st.wrote = false;
                for (uint32_t s = 0; s < n_slots; ++s)
                    if (box_valid(bb[s])) pixel_visit(d.mode, rr[s], bb[s], s, px, py, jx0, jx1, jy0, jy1, st);
                if (!st.wrote) continue;
                zbuf[at] = st.best_z;
                if (d.mode == MODE_SHADOW || st.shade_slot == 0xFFFFFFFFu) continue;
                float vel[2] = {0.0f, 0.0f};
                shade_pixel(d, rr[st.shade_slot], ss[st.shade_slot], px, py, canvas + at * 4, vel);
                if (d.mode == MODE_PBR && velocity) { velocity[at * 2] = vel[0]; velocity[at * 2 + 1] = vel[1]; }
            }
    }
}

extern "C"
{
    struct L2U { float mvp[16], model[16], mv[16], normal_mat[9], light_vp[16]; float light_dir_world[3], camera_pos[3]; uint8_t base_color[4]; int32_t use_texture; };
    struct L3U
    {
        float mvp[16], prev_mvp[16], model[16], mv[16], normal_mat[9], light_vp[16];
        float light_dir_world[3], camera_pos[3];
        uint8_t base_color_srgb[4];
        float metallic, roughness, ao;
        int32_t use_texture;
        float ibl_diffuse_intensity, ibl_specular_intensity, ibl_reflection_strength;
    };

    int32_t shsemu_l2_shadow_draw(const float* positions, uint32_t n_vertices, const float model[16], const float light_vp[16], int32_t sm_w, int32_t sm_h,
                                  int32_t tile_w, int32_t tile_h, float* shadow_depth)
    {
        Draw d{};
        d.positions = positions; d.n_positions = n_vertices; d.n_tris = n_vertices / 3;
        d.mode = MODE_SHADOW; d.W = sm_w; d.H = sm_h; d.job_w = tile_w; d.job_h = tile_h;
        mat_mul(light_vp, model, d.light_model);
        run(d, nullptr, shadow_depth, nullptr);
        return 0;
    }
    int32_t shsemu_l3_shadow_draw(const float* positions, uint32_t n_vertices, const float model[16], const float light_vp[16], int32_t sm_w, int32_t sm_h,
                                  int32_t tile_w, int32_t tile_h, float* shadow_depth)
    {
        return shsemu_l2_shadow_draw(positions, n_vertices, model, light_vp, sm_w, sm_h, tile_w, tile_h, shadow_depth);
    }

    int32_t shsemu_l2_camera_draw(const float* positions, const float* normals, const float* uvs, uint32_t n_vertices, const L2U* un, const uint8_t* texture_rgba,
                                  int32_t tex_w, int32_t tex_h, const float* shadow_depth, int32_t sm_w, int32_t sm_h, int32_t W, int32_t H, int32_t tile_w,
                                  int32_t tile_h, uint8_t* canvas_rgba, float* zbuffer)
    {
        Draw d{};
        d.positions = positions; d.normals = normals; d.uvs = uvs;
        d.n_positions = d.n_normals = d.n_uvs = n_vertices; d.n_tris = n_vertices / 3;
        d.mode = MODE_SOFTSHADOW; d.W = W; d.H = H; d.job_w = tile_w; d.job_h = tile_h;
        std::memcpy(d.mvp, un->mvp, 64); std::memcpy(d.model, un->model, 64); std::memcpy(d.mv, un->mv, 64);
        std::memcpy(d.normal_mat, un->normal_mat, 36); std::memcpy(d.light_vp, un->light_vp, 64);
        std::memcpy(d.light_dir, un->light_dir_world, 12); std::memcpy(d.camera_pos, un->camera_pos, 12); std::memcpy(d.color, un->base_color, 4);
        d.use_texture = un->use_texture;
        d.tex = (texture_rgba && tex_w > 0 && tex_h > 0) ? texture_rgba : nullptr; d.tex_w = tex_w; d.tex_h = tex_h;
        d.shadow = (shadow_depth && sm_w > 0 && sm_h > 0) ? shadow_depth : nullptr; d.sm_w = sm_w; d.sm_h = sm_h;
        run(d, canvas_rgba, zbuffer, nullptr);
        return 0;
    }

    int32_t shsemu_l3_camera_draw(const float* positions, const float* normals, const float* uvs, uint32_t n_vertices, const L3U* un, const uint8_t* texture_rgba,
                                  int32_t tex_w, int32_t tex_h, const float* shadow_depth, int32_t sm_w, int32_t sm_h, const float* irradiance, int32_t irr_size,
                                  const float* prefiltered, const int32_t* spec_sizes, int32_t n_mips, int32_t W, int32_t H, int32_t tile_w, int32_t tile_h,
                                  uint8_t* canvas_rgba, float* zbuffer, float* velocity)
    {
        Draw d{};
        d.positions = positions; d.normals = normals; d.uvs = uvs;
        d.n_positions = d.n_normals = d.n_uvs = n_vertices; d.n_tris = n_vertices / 3;
        d.mode = MODE_PBR; d.W = W; d.H = H; d.job_w = tile_w; d.job_h = tile_h;
        std::memcpy(d.mvp, un->mvp, 64); std::memcpy(d.prev_mvp, un->prev_mvp, 64); std::memcpy(d.model, un->model, 64); std::memcpy(d.mv, un->mv, 64);
        std::memcpy(d.normal_mat, un->normal_mat, 36); std::memcpy(d.light_vp, un->light_vp, 64);
        std::memcpy(d.light_dir, un->light_dir_world, 12); std::memcpy(d.camera_pos, un->camera_pos, 12); std::memcpy(d.color, un->base_color_srgb, 4);
        d.use_texture = un->use_texture;
        d.tex = (texture_rgba && tex_w > 0 && tex_h > 0) ? texture_rgba : nullptr; d.tex_w = tex_w; d.tex_h = tex_h;
        d.shadow = (shadow_depth && sm_w > 0 && sm_h > 0) ? shadow_depth : nullptr; d.sm_w = sm_w; d.sm_h = sm_h;
        d.metallic = un->metallic; d.roughness = un->roughness; d.ao = un->ao;
        d.ibl_diffuse = un->ibl_diffuse_intensity; d.ibl_specular = un->ibl_specular_intensity; d.ibl_reflection = un->ibl_reflection_strength;
        const bool has_ibl = irradiance && irr_size > 0 && prefiltered && spec_sizes && n_mips > 0 && n_mips <= MAX_SPEC_MIPS;
        if (has_ibl)
        {
            d.irradiance = irradiance; d.irr_size = irr_size; d.prefiltered = prefiltered; d.n_mips = n_mips;
            uint32_t off = 0;
            for (int m = 0; m < n_mips; ++m) { d.spec_size[m] = spec_sizes[m]; d.spec_off[m] = off; off += 6u * (uint32_t)spec_sizes[m] * (uint32_t)spec_sizes[m] * 3u; }
        }
        run(d, canvas_rgba, zbuffer, velocity);
        return 0;
    }
}
