// tests/cpp/legacy2_drop_in_test.cpp -- the binding of the legacy render-target demos (host/shs_b200/legacy2_drop_in.hpp) exercised
// with the demos' own code.  This translation unit CONTAINS the reference's hello_shadow_mapping_soft.cpp (default; row L2) or, with
// -DDEMO_PBR, hello_pbr.cpp (row L3) -- main renamed, never called -- so the CPU side below is the reference's own
// draw_triangle_tile_shadow / draw_triangle_tile_color_depth_* + shaders + Canvas + ZBuffer + ShadowMap, and the GPU side is
// shs::b200::legacy2::Renderer fed with the SAME Uniforms objects.
//
// Built only where /root/reference exists (tests/cpp/Makefile, against oracle/glm_shim + oracle/legacy_shim).
// Exit code 0 = shadow map / z-buffer (/ velocity) bit-equal and canvas within 1 LSB, 1 = mismatch, 77 = no CUDA device (after
// checking that every call refused and nothing ran).
#include <cmath>
#include <cstdio>
#include <cstring>

#define main shs_legacy2_demo_main
#ifdef DEMO_PBR
#include "hello_pbr.cpp"
#else
#include "hello_shadow_mapping_soft.cpp"
#endif
#undef main
#include "legacy_shim/stubs.inc"

#include "shs_b200/legacy2_drop_in.hpp"

struct Soup { std::vector<glm::vec3> tris, nrms; std::vector<glm::vec2> uvs; };

static void make_blob(int rings, int segs, float radius, Soup& s)
{
    auto point = [&](int r, int sg, glm::vec3& p, glm::vec3& n, glm::vec2& uv) {
        const float th = (float)r / rings * 3.14159265f, ph = (float)sg / segs * 6.2831853f;
        n = glm::vec3(std::sin(th) * std::cos(ph), std::cos(th), std::sin(th) * std::sin(ph));
        p = n * (radius * (1.0f + 0.25f * std::sin(3.0f * ph) * std::sin(2.0f * th)));
        uv = glm::vec2((float)sg / segs, (float)r / rings);
    };
    for (int r = 0; r < rings; ++r)
        for (int sg = 0; sg < segs; ++sg)
        {
            glm::vec3 p[4], n[4];
            glm::vec2 t[4];
            point(r, sg, p[0], n[0], t[0]); point(r, sg + 1, p[1], n[1], t[1]); point(r + 1, sg, p[2], n[2], t[2]); point(r + 1, sg + 1, p[3], n[3], t[3]);
            const int order[6] = {0, 1, 2, 1, 3, 2};
            for (int k : order) { s.tris.push_back(p[k]); s.nrms.push_back(n[k]); s.uvs.push_back(t[k]); }
        }
}

static void make_floor(float half, int cells, Soup& s) // a grid in the xz plane that reaches behind the camera (near-plane clipping)
{
    for (int z = 0; z < cells; ++z)
        for (int x = 0; x < cells; ++x)
        {
            const float x0 = -half + 2 * half * x / cells, x1 = -half + 2 * half * (x + 1) / cells, z0 = -half + 2 * half * z / cells, z1 = -half + 2 * half * (z + 1) / cells;
            const glm::vec3 p[4] = {{x0, 0, z0}, {x1, 0, z0}, {x1, 0, z1}, {x0, 0, z1}};
            const int order[6] = {0, 1, 2, 0, 2, 3};
            for (int k : order) { s.tris.push_back(p[k]); s.nrms.push_back(glm::vec3(0, 1, 0)); s.uvs.push_back(glm::vec2(p[k].x, p[k].z) * 0.25f); }
        }
}

int main()
{
    const int W = 200, H = 150, SM = 256;
    Soup blob, floor_;
    make_blob(14, 20, 1.0f, blob);
    make_floor(12.0f, 6, floor_);

    const glm::vec3 cam(0.5f, 2.2f, -6.0f);
    const glm::mat4 view = glm::lookAtLH(cam, glm::vec3(0.0f, 0.8f, 0.0f), glm::vec3(0, 1, 0));
    glm::mat4 proj(0.0f); // perspective, left-handed, z in 0..1 (the demos' Camera3D convention)
    {
        const float t = std::tan(glm::radians(60.0f) * 0.5f), zn = 0.1f, zf = 200.0f;
        proj[0][0] = 1.0f / (((float)W / H) * t); proj[1][1] = 1.0f / t; proj[2][2] = zf / (zf - zn); proj[2][3] = 1.0f; proj[3][2] = -(zf * zn) / (zf - zn);
    }
    const glm::mat4 prev_view = glm::lookAtLH(cam + glm::vec3(0.3f, 0.0f, 0.2f), glm::vec3(0.0f, 0.8f, 0.0f), glm::vec3(0, 1, 0));
    const glm::vec3 light_dir = glm::normalize(glm::vec3(0.4668f, -0.3487f, 0.8127f));
    const glm::mat4 light_view = glm::lookAtLH(-light_dir * 30.0f, glm::vec3(0.0f), glm::vec3(0, 1, 0));
    glm::mat4 light_proj(1.0f); // ortho, z in 0..1
    light_proj[0][0] = 1.0f / 9.0f; light_proj[1][1] = 1.0f / 9.0f; light_proj[2][2] = 1.0f / (80.0f - 0.1f); light_proj[3][2] = -0.1f / (80.0f - 0.1f);
    const glm::mat4 light_vp = light_proj * light_view;

    shs::Texture2D tex(8, 8);
    for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) tex.texels.at(x, y) = ((x + y) & 1) ? shs::Color{230, 220, 200, 255} : shs::Color{60, 90, 140, 255};

    struct Obj { const Soup* g; glm::mat4 model; shs::Color color; bool textured; };
    glm::mat4 blob_model = glm::translate(glm::mat4(1.0f), glm::vec3(0.2f, 1.3f, 0.5f)) * glm::rotate(glm::mat4(1.0f), 0.7f, glm::vec3(0, 1, 0)) * glm::scale(glm::mat4(1.0f), glm::vec3(1.2f, 0.9f, 1.2f));
    const Obj objs[2] = {{&floor_, glm::mat4(1.0f), shs::Color{150, 160, 150, 255}, true}, {&blob, blob_model, shs::Color{210, 120, 60, 255}, false}};

#ifdef DEMO_PBR
    EnvIBL env;
    auto fill_cube = [](shs::CubeMapLinear& c, int size, float k) {
        c.size = size;
        for (int f = 0; f < 6; ++f)
        {
            c.face[f].resize((size_t)size * size);
            for (int i = 0; i < size * size; ++i) c.face[f][(size_t)i] = glm::vec3(0.2f + 0.1f * f, 0.3f + 0.02f * (i % 7), 0.25f + 0.015f * (i % 11)) * k;
        }
    };
    fill_cube(env.env_irradiance, 8, 1.0f);
    env.env_prefiltered_spec.mip.resize(4);
    for (int m = 0; m < 4; ++m) fill_cube(env.env_prefiltered_spec.mip[(size_t)m], 16 >> m, 1.5f - 0.2f * m);
#endif

    auto uniforms_of = [&](const Obj& o, const ShadowMap* sm) {
        Uniforms u;
        u.model = o.model;
        u.view = view;
        u.mv = view * o.model;
        u.mvp = proj * u.mv;
        u.normal_mat = glm::transpose(glm::inverse(glm::mat3(u.model)));
        u.light_vp = light_vp;
        u.light_dir_world = light_dir;
        u.camera_pos = cam;
        u.albedo = &tex;
        u.use_texture = o.textured;
        u.shadow = sm;
#ifdef DEMO_PBR
        u.prev_mvp = proj * prev_view * o.model;
        u.mat.baseColor_srgb = o.color;
        u.mat.metallic = o.textured ? 0.0f : 0.8f;
        u.mat.roughness = o.textured ? 0.7f : 0.3f;
        u.ibl = &env;
#else
        u.base_color = o.color;
#endif
        return u;
    };

    // ---- reference: the two passes of RendererSystem::process, serially over the job tiles
    ShadowMap sm_ref(SM, SM);
    {
        const int cols = (SM + TILE_SIZE_X - 1) / TILE_SIZE_X, rows = (SM + TILE_SIZE_Y - 1) / TILE_SIZE_Y;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const glm::ivec2 t_min(tx * TILE_SIZE_X, ty * TILE_SIZE_Y), t_max(std::min((tx + 1) * TILE_SIZE_X, SM) - 1, std::min((ty + 1) * TILE_SIZE_Y, SM) - 1);
                for (const Obj& o : objs)
                {
                    const Uniforms u = uniforms_of(o, nullptr);
                    for (size_t i = 0; i + 2 < o.g->tris.size(); i += 3)
                    {
                        const std::vector<glm::vec3> tv = {o.g->tris[i], o.g->tris[i + 1], o.g->tris[i + 2]};
                        draw_triangle_tile_shadow(sm_ref, tv, [&u](const glm::vec3& p) { return shadow_vertex_shader(p, u); }, t_min, t_max);
                    }
                }
            }
    }
#ifdef DEMO_PBR
    RT_ColorDepthMotion rt_ref(W, H, 0.1f, 200.0f, shs::Color{20, 20, 25, 255}), rt_gpu(W, H, 0.1f, 200.0f, shs::Color{20, 20, 25, 255});
    shs::Canvas& canvas_ref = rt_ref.color; shs::Canvas& canvas_gpu = rt_gpu.color;
    shs::ZBuffer& z_ref = rt_ref.depth; shs::ZBuffer& z_gpu = rt_gpu.depth;
#else
    shs::Canvas canvas_ref(W, H, shs::Color{20, 20, 25, 255}), canvas_gpu(W, H, shs::Color{20, 20, 25, 255});
    shs::ZBuffer z_ref(W, H, 0.1f, 200.0f), z_gpu(W, H, 0.1f, 200.0f);
#endif
    z_ref.clear();
    z_gpu.clear();
    {
        const int cols = (W + TILE_SIZE_X - 1) / TILE_SIZE_X, rows = (H + TILE_SIZE_Y - 1) / TILE_SIZE_Y;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const glm::ivec2 t_min(tx * TILE_SIZE_X, ty * TILE_SIZE_Y), t_max(std::min((tx + 1) * TILE_SIZE_X, W) - 1, std::min((ty + 1) * TILE_SIZE_Y, H) - 1);
                for (const Obj& o : objs)
                {
                    const Uniforms u = uniforms_of(o, &sm_ref);
                    for (size_t i = 0; i + 2 < o.g->tris.size(); i += 3)
                    {
                        const std::vector<glm::vec3> tv = {o.g->tris[i], o.g->tris[i + 1], o.g->tris[i + 2]}, tn = {o.g->nrms[i], o.g->nrms[i + 1], o.g->nrms[i + 2]};
                        const std::vector<glm::vec2> tu = {o.g->uvs[i], o.g->uvs[i + 1], o.g->uvs[i + 2]};
#ifdef DEMO_PBR
                        draw_triangle_tile_color_depth_motion(rt_ref, tv, tn, tu,
                            [&u](const glm::vec3& p, const glm::vec3& n, const glm::vec2& uv) { return vertex_shader_full(p, n, uv, u); },
                            [&u](const VaryingsFull& v) { return fragment_shader_pbr(v, u); }, t_min, t_max);
#else
                        draw_triangle_tile_color_depth_softshadow(canvas_ref, z_ref, tv, tn, tu,
                            [&u](const glm::vec3& p, const glm::vec3& n, const glm::vec2& uv) { return vertex_shader_full(p, n, uv, u); },
                            [&u](const VaryingsFull& v, int px, int py) { return fragment_shader_softshadow(v, u, px, py); }, t_min, t_max);
#endif
                    }
                }
            }
    }

    size_t ref_covered = 0, ref_texels = 0;
    for (size_t i = 0; i < (size_t)W * H; ++i) ref_covered += z_ref.buffer().raw()[i] < std::numeric_limits<float>::max();
    for (int y = 0; y < SM; ++y)
        for (int x = 0; x < SM; ++x) ref_texels += sm_ref.sample(x, y) < std::numeric_limits<float>::max();
    std::printf("reference side: %zu of %d px covered, %zu of %d shadow texels written\n", ref_covered, W * H, ref_texels, SM * SM);

    // ---- B200
    shs::b200::legacy2::Renderer gpu(0, TILE_SIZE_X, TILE_SIZE_Y);
    auto lit_draw = [&](const Obj& o, const ShadowMap* sm) {
        const Uniforms u = uniforms_of(o, sm);
#ifdef DEMO_PBR
        return gpu.draw_pbr(o.g->tris, o.g->nrms, o.g->uvs, u);
#else
        return gpu.draw_softshadow(o.g->tris, o.g->nrms, o.g->uvs, u);
#endif
    };
    if (!gpu.valid())
    {
        const bool refused = !gpu.begin_shadow(SM, SM) && !gpu.shadow_draw(blob.tris, blob_model, light_vp) && !gpu.begin_frame(canvas_gpu, z_gpu) && !lit_draw(objs[1], nullptr) &&
                             !gpu.end_frame(canvas_gpu, z_gpu);
        std::printf("SKIP: %s (%s)\n", gpu.last_error(), refused ? "every call refused, nothing ran on the CPU" : "A CALL CLAIMED SUCCESS WITHOUT A DEVICE");
        return refused ? 77 : 1;
    }
    bool ok = gpu.begin_shadow(SM, SM);
    for (const Obj& o : objs) ok = ok && gpu.shadow_draw(o.g->tris, o.model, light_vp);
#ifdef DEMO_PBR
    ok = ok && gpu.set_ibl(env) && gpu.begin_frame(canvas_gpu, z_gpu, &rt_gpu.velocity);
#else
    ok = ok && gpu.begin_frame(canvas_gpu, z_gpu);
#endif
    for (const Obj& o : objs) ok = ok && lit_draw(o, &sm_ref); // a non-null Uniforms::shadow selects the DEVICE shadow map
#ifdef DEMO_PBR
    ok = ok && gpu.end_frame(canvas_gpu, z_gpu, &rt_gpu.velocity);
#else
    ok = ok && gpu.end_frame(canvas_gpu, z_gpu);
#endif
    ShadowMap sm_gpu(SM, SM);
    ok = ok && gpu.download_shadow(sm_gpu);
    if (!ok) { std::printf("FAIL: %s\n", gpu.last_error()); return 1; }

    size_t sm_diff = 0, sm_written = 0, z_diff = 0, covered = 0, vel_diff = 0, loose = 0;
    for (int y = 0; y < SM; ++y)
        for (int x = 0; x < SM; ++x)
        {
            const float a = sm_ref.sample(x, y), b = sm_gpu.sample(x, y);
            if (std::memcmp(&a, &b, 4) != 0) ++sm_diff;
            if (a < std::numeric_limits<float>::max()) ++sm_written;
        }
    int max_lsb = 0;
    const size_t n = (size_t)W * H;
    for (size_t i = 0; i < n; ++i)
    {
        if (std::memcmp(&z_ref.buffer().raw()[i], &z_gpu.buffer().raw()[i], 4) != 0) ++z_diff;
        if (z_ref.buffer().raw()[i] < std::numeric_limits<float>::max()) ++covered;
#ifdef DEMO_PBR
        if (std::memcmp(&rt_ref.velocity.raw()[i], &rt_gpu.velocity.raw()[i], 8) != 0) ++vel_diff;
#endif
        const shs::Color a = canvas_ref.buffer().raw()[i], b = canvas_gpu.buffer().raw()[i];
        const int d = std::max(std::abs(a.r - b.r), std::max(std::abs(a.g - b.g), std::max(std::abs(a.b - b.b), std::abs(a.a - b.a))));
        if (d > 1) ++loose;
        max_lsb = std::max(max_lsb, d);
    }
    const bool pass = sm_diff == 0 && z_diff == 0 && vel_diff == 0 && loose <= 2 && covered > 5000 && sm_written > 500;
    std::printf("legacy2 drop-in (%s): shadow map %zu texels written / %zu differing, %zu of %zu px covered, z-buffer differing %zu, velocity differing %zu, "
                "canvas <= %d LSB (%zu px beyond 1) | %s\n",
#ifdef DEMO_PBR
                "hello_pbr",
#else
                "hello_shadow_mapping_soft",
#endif
                sm_written, sm_diff, covered, n, z_diff, vel_diff, max_lsb, loose, pass ? "OK" : "MISMATCH");
    return pass ? 0 : 1;
}
