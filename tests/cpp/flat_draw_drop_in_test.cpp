// tests/cpp/flat_draw_drop_in_test.cpp -- host/shs_b200/flat_draw_drop_in.hpp against the reference draws it stands in for, over the
// reference's own types (DebugMesh, RT_ColorLDR, LightInstance with the four light models, LightSelection): a frame of boxes, spheres
// and a floor is drawn once by the reference's per-object calls (debug_draw::draw_mesh_blinn_phong_transformed from its header;
// draw_mesh_multi_light_transformed from the demo's text, oracle/_ref/flat_draw_generated.inc) and once through the batch binding.
// Compiled with SHS_HAS_JOLT=1 against the JoltPhysics declaration shim.
// Exit code 0 = depth buffers equal bit for bit and canvases within 1 LSB, 1 = mismatch, 77 = no CUDA device (after checking that the
// binding refused and drew nothing).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#define SHS_HAS_JOLT 1
#include "shs_b200/flat_draw_drop_in.hpp"
#include "shs/geometry/volumes.hpp"

using namespace shs;
namespace
{
#include "flat_draw_generated.inc"

    DebugMesh box_mesh()
    {
        DebugMesh m;
        for (int z = 0; z < 2; ++z) for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x) m.vertices.push_back(glm::vec3(x - 0.5f, y - 0.5f, z - 0.5f));
        const uint32_t idx[36] = {0, 1, 3, 0, 3, 2, 4, 6, 7, 4, 7, 5, 0, 4, 5, 0, 5, 1, 2, 3, 7, 2, 7, 6, 0, 2, 6, 0, 6, 4, 1, 5, 7, 1, 7, 3};
        m.indices.assign(idx, idx + 36);
        return m;
    }
    DebugMesh sphere_mesh(int n_lon, int n_lat)
    {
        DebugMesh m;
        for (int j = 0; j <= n_lat; ++j)
            for (int i = 0; i < n_lon; ++i)
            {
                const float th = 3.14159265f * (float)j / (float)n_lat, ph = 6.2831853f * (float)i / (float)n_lon;
                m.vertices.push_back(glm::vec3(0.5f * std::sin(th) * std::cos(ph), 0.5f * std::cos(th), 0.5f * std::sin(th) * std::sin(ph)));
            }
        for (int j = 0; j < n_lat; ++j)
            for (int i = 0; i < n_lon; ++i)
            {
                const uint32_t a = j * n_lon + i, b = j * n_lon + (i + 1) % n_lon, c = (j + 1) * n_lon + (i + 1) % n_lon, d = (j + 1) * n_lon + i;
                const uint32_t q[6] = {a, b, c, a, c, d};
                m.indices.insert(m.indices.end(), q, q + 6);
            }
        return m;
    }
    DebugMesh floor_mesh()
    {
        DebugMesh m;
        m.vertices = {glm::vec3(-0.5f, 0, -0.5f), glm::vec3(0.5f, 0, -0.5f), glm::vec3(0.5f, 0, 0.5f), glm::vec3(-0.5f, 0, 0.5f)};
        m.indices = {0, 2, 1, 0, 3, 2};
        return m;
    }
}

int main()
{
    const int W = 400, H = 300;
    std::mt19937 rng(11);
    std::uniform_real_distribution<float> U(-1.0f, 1.0f);
    const DebugMesh meshes[3] = {box_mesh(), sphere_mesh(14, 9), floor_mesh()};
    static const PointLightModel point_model;
    static const SpotLightModel spot_model;
    static const RectAreaLightModel rect_model;
    static const TubeAreaLightModel tube_model;
    const ILightModel* models[4] = {&point_model, &spot_model, &rect_model, &tube_model};
    std::vector<LightInstance> lights(24);
    for (size_t i = 0; i < lights.size(); ++i)
    {
        LightInstance& l = lights[i];
        l.model = models[i % 4];
        l.props.color = glm::vec3(0.4f + 0.6f * std::abs(U(rng)), 0.4f + 0.6f * std::abs(U(rng)), 0.4f + 0.6f * std::abs(U(rng)));
        l.props.intensity = 1.0f + 3.0f * std::abs(U(rng));
        l.props.position_ws = glm::vec3(U(rng) * 9.0f, 1.5f + 2.5f * std::abs(U(rng)), U(rng) * 9.0f);
        l.props.range = 4.0f + 8.0f * std::abs(U(rng));
        l.props.direction_ws = glm::normalize(glm::vec3(U(rng) * 0.6f, -1.0f, U(rng) * 0.6f));
        l.props.right_ws = glm::normalize(glm::vec3(1.0f, U(rng) * 0.2f, U(rng)));
        l.props.up_ws = glm::normalize(glm::cross(l.props.right_ws, l.props.direction_ws));
        l.props.attenuation_model = (LightAttenuationModel)(i % 3);
        l.props.attenuation_power = (i % 5 == 0) ? 2.0f : 1.0f;
    }
    struct Object { int mesh; glm::mat4 model; glm::vec3 base; LightSelection sel; };
    std::vector<Object> objects;
    {
        Object f{2, glm::mat4(1.0f), glm::vec3(0.55f, 0.55f, 0.6f), {}};
        f.model[0][0] = 60.0f; f.model[2][2] = 60.0f;
        objects.push_back(f);
    }
    for (int i = 0; i < 80; ++i)
    {
        Object o{i % 2, glm::mat4(1.0f), glm::vec3(std::abs(U(rng)), std::abs(U(rng)), std::abs(U(rng))), {}};
        const float s = 0.5f + 1.2f * std::abs(U(rng));
        o.model[0][0] = s; o.model[1][1] = s * (0.6f + std::abs(U(rng))); o.model[2][2] = s;
        o.model[3] = glm::vec4(U(rng) * 10.0f, 0.6f + 1.5f * std::abs(U(rng)), U(rng) * 10.0f, 1.0f);
        objects.push_back(o);
    }
    for (Object& o : objects)
    {
        o.sel.count = 1u + (uint32_t)(std::abs(U(rng)) * 7.99f);
        for (uint32_t k = 0; k < kLightSelectionCapacity; ++k) o.sel.indices[k] = (uint32_t)(std::abs(U(rng)) * 25.99f); // 24, 25: stale entries
    }
    const glm::vec3 eye(3.0f, 7.0f, -16.0f);
    const glm::mat4 vp = glm::perspectiveLH_NO(glm::radians(60.0f), (float)W / (float)H, 0.05f, 300.0f) * glm::lookAtLH(eye, glm::vec3(0.0f, 1.0f, 0.0f), glm::vec3(0, 1, 0));
    const glm::vec3 sun(0.3f, -1.0f, 0.25f);

    // ---- reference side: its own per-object calls
    RT_ColorLDR ref_bp(W, H, Color{12, 13, 18, 255}), ref_ml(W, H, Color{12, 13, 18, 255});
    std::vector<float> ref_bp_z((size_t)W * H, 1.0f), ref_ml_z((size_t)W * H, 1.0f);
    for (const Object& o : objects)
    {
        debug_draw::draw_mesh_blinn_phong_transformed(ref_bp, std::span<float>(ref_bp_z.data(), ref_bp_z.size()), meshes[o.mesh], o.model, vp, W, H, eye, sun, o.base);
        draw_mesh_multi_light_transformed(ref_ml, ref_ml_z, meshes[o.mesh], o.model, vp, W, H, eye, o.base, lights, o.sel);
    }
    size_t covered = 0, lit_differs = 0;
    for (size_t i = 0; i < ref_bp_z.size(); ++i)
    {
        covered += ref_ml_z[i] < 1.0f;
        lit_differs += std::memcmp(&ref_bp.color.data[i], &ref_ml.color.data[i], 4) != 0;
    }
    std::printf("reference side: %zu of %d texels covered, %zu differ between the two draws\n", covered, W * H, lit_differs);

    // ---- binding side
    shsb_ctx ctx = nullptr;
    const int rc = shsb_context_create(0, &ctx);
    RT_ColorLDR got_bp(W, H, Color{12, 13, 18, 255}), got_ml(W, H, Color{12, 13, 18, 255});
    std::vector<float> got_bp_z((size_t)W * H, 1.0f), got_ml_z((size_t)W * H, 1.0f);
    if (rc != SHSB_OK)
    {
        // no device: the binding has nothing to fall back on
        b200::FlatDrawBatch batch(nullptr);
        bool refused = true;
        for (const Object& o : objects) refused = refused && !batch.draw_mesh_blinn_phong_transformed(meshes[o.mesh], o.model, o.base);
        refused = refused && batch.size() == 0;
        ShsbFlatDraw d{};
        float m16[16] = {0}, v3[3] = {0};
        refused = refused && shsb_flat_draw_blinn_phong(nullptr, &d, 1, m16, v3, v3, 1, 2) != SHSB_OK && shsb_flat_draw_multi_light(nullptr, &d, 1, m16, v3, nullptr, 0, 1, 2) != SHSB_OK;
        size_t touched = 0;
        for (size_t i = 0; i < got_bp_z.size(); ++i) touched += got_bp_z[i] != 1.0f;
        std::printf("%s; %zu texels drawn without a device\n", refused ? "every call refused" : "A CALL WAS ACCEPTED WITHOUT A DEVICE", touched);
        return (refused && touched == 0) ? 77 : 1;
    }
    int bad = 0;
    {
        b200::FlatDrawBatch batch(ctx);
        for (const Object& o : objects) if (!batch.draw_mesh_blinn_phong_transformed(meshes[o.mesh], o.model, o.base)) { std::printf("record failed: %s\n", shsb_last_error_string(ctx)); return 1; }
        if (!batch.flush_blinn_phong(got_bp, std::span<float>(got_bp_z.data(), got_bp_z.size()), vp, W, H, eye, sun)) { std::printf("flush failed: %s\n", shsb_last_error_string(ctx)); return 1; }
        for (const Object& o : objects) batch.draw_mesh_multi_light_transformed(meshes[o.mesh], o.model, o.base, o.sel);
        if (!batch.flush_multi_light(got_ml, std::span<float>(got_ml_z.data(), got_ml_z.size()), vp, W, H, eye, lights)) { std::printf("flush failed: %s\n", shsb_last_error_string(ctx)); return 1; }
        // a second frame over the first (no clear): the batch sees the depth buffer the first left
        for (size_t k = 0; k < objects.size(); k += 3) batch.draw_mesh_multi_light_transformed(meshes[objects[k].mesh], objects[k].model, glm::vec3(1.0f, 0.0f, 1.0f), objects[k].sel);
        if (!batch.flush_multi_light(got_ml, std::span<float>(got_ml_z.data(), got_ml_z.size()), vp, W, H, eye, lights)) return 1;
        for (size_t k = 0; k < objects.size(); k += 3)
            draw_mesh_multi_light_transformed(ref_ml, ref_ml_z, meshes[objects[k].mesh], objects[k].model, vp, W, H, eye, glm::vec3(1.0f, 0.0f, 1.0f), lights, objects[k].sel);
    }
    auto compare = [&](const char* what, const RT_ColorLDR& a, const std::vector<float>& az, const RT_ColorLDR& b, const std::vector<float>& bz) {
        size_t depth_bad = 0, colour_bad = 0, colour_off = 0;
        for (size_t i = 0; i < az.size(); ++i)
        {
            depth_bad += std::memcmp(&az[i], &bz[i], 4) != 0;
            const uint8_t* p = reinterpret_cast<const uint8_t*>(&a.color.data[i]);
            const uint8_t* q = reinterpret_cast<const uint8_t*>(&b.color.data[i]);
            for (int c = 0; c < 4; ++c)
            {
                const int d = std::abs((int)p[c] - (int)q[c]);
                colour_off += d == 1;
                colour_bad += d > 1;
            }
        }
        std::printf("%s: depth differs at %zu texels, colour channels off by 1: %zu, by more: %zu\n", what, depth_bad, colour_off, colour_bad);
        if (depth_bad || colour_bad) ++bad;
    };
    compare("blinn-phong", got_bp, got_bp_z, ref_bp, ref_bp_z);
    compare("multi-light (two frames)", got_ml, got_ml_z, ref_ml, ref_ml_z);
    shsb_context_destroy(ctx);
    std::printf(bad ? "MISMATCH\n" : "flat-draw binding == reference\n");
    return bad ? 1 : 0;
}
