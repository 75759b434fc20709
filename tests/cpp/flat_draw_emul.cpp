// tests/cpp/flat_draw_emul.cpp -- TEST INFRASTRUCTURE ONLY.  The device functions of leisure_software_renderer_b200/csrc/flat_draw_core.cuh
// (and the occlusion raster functions of scene_cull_core.cuh they share) compiled by g++ (-ffp-contract=off == nvcc --fmad=false)
// and driven like flat_draw.cu drives them: one set-up per (draw, triangle) giving the flat colour and the triangle record, then a
// MINIMUM on (depth bits << 32 | 1 + running triangle number) per texel -- triangles and texels visited here in REVERSE order to
// show that the order does not matter -- and a resolve that writes depth and colour of the texels a triangle won.
// Same C signature as oracle/oracle_flat_draw.cpp's shso_flat_draw under the prefix shsemu_.  Nothing in the product links or loads it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "flat_draw_core.cuh"

using namespace shsb;

extern "C" int32_t shsemu_flat_draw(int32_t mode, uint32_t n_draws, const uint32_t* draw_mesh, const float* models16, const float* base3, const uint32_t* sel_counts,
                                    const uint32_t* sel_idx8, const uint32_t* mesh_table3, uint32_t n_meshes, const float* vertices, uint32_t n_vertices, const uint32_t* indices,
                                    uint32_t n_indices, const float view_proj[16], const float camera3[3], const float light_dir3[3], const void* lights128, uint32_t n_lights,
                                    int32_t W, int32_t H, uint8_t* canvas_rgba, float* depth)
{
    const fd::LightProps* lights = static_cast<const fd::LightProps*>(lights128);
    fd::V3 L = fd::v3(0, 0, 0);
    if (mode == fd::MODE_BLINN_PHONG) L = fd::glm_normalize(-fd::v3(light_dir3)); // taken once per batch on the host by the product (api.cu)
    std::vector<sc::OccTri> tris;
    std::vector<uint32_t> colours;
    for (uint32_t d = 0; d < n_draws; ++d)
    {
        if (draw_mesh[d] >= n_meshes) return 1;
        const uint32_t first = mesh_table3[3 * draw_mesh[d]], count = mesh_table3[3 * draw_mesh[d] + 1], base_v = mesh_table3[3 * draw_mesh[d] + 2];
        if ((uint64_t)first + count > n_indices) return 1;
        // the product uploads each mesh on its own (shsb_mesh_upload): positions from base_v on, n_positions = what is left
        const float* positions = vertices + (size_t)base_v * 3;
        const uint32_t n_positions = n_vertices - base_v;
        const float* model = models16 + (size_t)d * 16;
        for (uint32_t t = 0; t < count / 3; ++t)
        {
            sc::OccTri rec;
            rec.valid = false;
            uint32_t colour = 0;
            const uint32_t i0 = indices[first + 3 * t], i1 = indices[first + 3 * t + 1], i2 = indices[first + 3 * t + 2];
            if (i0 < n_positions && i1 < n_positions && i2 < n_positions)
            {
                float w0[4], w1[4], w2[4], s0[2], s1[2], s2[2], z0, z1, z2;
                sc::mul4(model, positions[3 * i0], positions[3 * i0 + 1], positions[3 * i0 + 2], 1.0f, w0);
                sc::mul4(model, positions[3 * i1], positions[3 * i1 + 1], positions[3 * i1 + 2], 1.0f, w1);
                sc::mul4(model, positions[3 * i2], positions[3 * i2 + 1], positions[3 * i2 + 2], 1.0f, w2);
                fd::V3 n;
                if (fd::project_world(w0, view_proj, W, H, s0, z0) && fd::project_world(w1, view_proj, W, H, s1, z1) && fd::project_world(w2, view_proj, W, H, s2, z2) &&
                    fd::face_normal(fd::v3(w0), fd::v3(w1), fd::v3(w2), n))
                {
                    rec = sc::occ_setup_triangle(s0, z0, s1, z1, s2, z2, W, H);
                    if (rec.valid)
                        colour = (mode == fd::MODE_BLINN_PHONG)
                                     ? fd::blinn_phong_colour(fd::v3(w0), fd::v3(w1), fd::v3(w2), n, fd::v3(camera3), L, fd::v3(base3 + 3 * d))
                                     : fd::multi_light_colour(fd::v3(w0), fd::v3(w1), fd::v3(w2), n, fd::v3(camera3), fd::v3(base3 + 3 * d), lights, n_lights, sel_idx8 + (size_t)d * 8,
                                                              sel_counts[d] < 8u ? sel_counts[d] : 8u);
                }
            }
            tris.push_back(rec);
            colours.push_back(colour);
        }
    }
    const size_t n_px = (size_t)W * H;
    std::vector<uint64_t> zkey(n_px);
    for (size_t i = 0; i < n_px; ++i)
    {
        uint32_t bits;
        std::memcpy(&bits, depth + i, 4);
        zkey[i] = (uint64_t)bits << 32;
    }
    for (size_t g = tris.size(); g-- > 0;)
    {
        const sc::OccTri& t = tris[g];
        if (!t.valid) continue;
        for (int y = t.max_y; y >= t.min_y; --y)
            for (int x = t.max_x; x >= t.min_x; --x)
            {
                float z;
                if (!sc::occ_texel_depth(t, x, y, z)) continue;
                z = z + 0.0f;
                uint32_t bits;
                std::memcpy(&bits, &z, 4);
                const uint64_t key = ((uint64_t)bits << 32) | (uint64_t)(g + 1);
                uint64_t& slot = zkey[(size_t)y * W + x];
                if (key < slot) slot = key;
            }
    }
    for (size_t i = 0; i < n_px; ++i)
    {
        const uint32_t order = (uint32_t)zkey[i];
        if (!order) continue;
        const uint32_t bits = (uint32_t)(zkey[i] >> 32);
        std::memcpy(depth + i, &bits, 4);
        std::memcpy(canvas_rgba + i * 4, &colours[order - 1], 4);
    }
    return 0;
}
