// tests/cpp/legacy_drop_in_test.cpp -- the legacy binding (host/shs_b200/legacy_drop_in.hpp) exercised with the demo's own code:
// this translation unit CONTAINS the reference's hello_pipeline_blinn_phong_shading.cpp (its main renamed, never called), so the
// CPU side below is the reference's own RendererSystem::draw_triangle_tile + Blinn-Phong shaders + Canvas + ZBuffer + Viewer, and
// the GPU side is shs::b200::legacy::Renderer over the same canvas / z-buffer / geometry / uniforms.
//
// Built only where /root/reference exists (tests/cpp/Makefile, against oracle/glm_shim + oracle/legacy_shim).
// Exit code 0 = z-buffer bit-equal and canvas within 1 LSB, 1 = mismatch, 77 = no CUDA device (after checking that nothing ran).
#include <cmath>
#include <cstdio>
#include <cstring>

#define main shs_legacy_demo_main
#include "hello_pipeline_blinn_phong_shading.cpp"
#undef main
#include "legacy_shim/stubs.inc"

#include "shs_b200/legacy_drop_in.hpp"

// a triangle-soup blob in ModelGeometry's layout (three corners per triangle, a normal per corner)
static void make_blob(int rings, int segs, float radius, std::vector<glm::vec3>& tris, std::vector<glm::vec3>& nrms)
{
    auto point = [&](int r, int s, glm::vec3& p, glm::vec3& n) {
        const float th = (float)r / rings * 3.14159265f, ph = (float)s / segs * 6.2831853f;
        n = glm::vec3(std::sin(th) * std::cos(ph), std::cos(th), std::sin(th) * std::sin(ph));
        p = n * (radius * (1.0f + 0.25f * std::sin(3.0f * ph) * std::sin(2.0f * th)));
    };
    for (int r = 0; r < rings; ++r)
        for (int s = 0; s < segs; ++s)
        {
            glm::vec3 p[4], n[4];
            point(r, s, p[0], n[0]); point(r, s + 1, p[1], n[1]); point(r + 1, s, p[2], n[2]); point(r + 1, s + 1, p[3], n[3]);
            const int order[6] = {0, 1, 2, 1, 3, 2};
            for (int k : order) { tris.push_back(p[k]); nrms.push_back(n[k]); }
        }
}

int main()
{
    const int W = 320, H = 240;
    std::vector<glm::vec3> tris, nrms;
    make_blob(18, 24, 1.0f, tris, nrms);

    // the demo's own scene set-up (:152-153, 384): Viewer, light direction, two objects with MonkeyObject's world matrix
    Viewer viewer(glm::vec3(0.0f, 1.5f, -7.0f), 50.0f, (float)W, (float)H);
    const glm::vec3 light_direction = glm::normalize(glm::vec3(-1.0f, -0.4f, 1.0f));
    struct Obj { glm::vec3 pos, scale; float angle; shs::Color color; };
    const Obj objs[2] = {{glm::vec3(-1.2f, 0.0f, 1.0f), glm::vec3(1.4f), 30.0f, shs::Color{60, 100, 200, 255}},
                         {glm::vec3(1.0f, 0.4f, -0.5f), glm::vec3(1.0f, 1.6f, 1.0f), -75.0f, shs::Color{220, 140, 40, 255}}};
    auto world = [](const Obj& o) {
        const glm::mat4 t = glm::translate(glm::mat4(1.0f), o.pos);
        const glm::mat4 r = glm::rotate(glm::mat4(1.0f), glm::radians(o.angle), glm::vec3(0.0f, 1.0f, 0.0f));
        const glm::mat4 s = glm::scale(glm::mat4(1.0f), o.scale);
        return t * r * s;
    };
    const glm::mat4 view = viewer.camera->view_matrix, proj = viewer.camera->projection_matrix;

    // ---- reference: RendererSystem::process's loops (:262-305), serially
    shs::Canvas canvas_ref(W, H), canvas_gpu(W, H);
    shs::ZBuffer z_ref(W, H, viewer.camera->z_near, viewer.camera->z_far), z_gpu(W, H, viewer.camera->z_near, viewer.camera->z_far);
    shs::Canvas::fill_pixel(canvas_ref, 0, 0, W, H, shs::Color::black());
    shs::Canvas::fill_pixel(canvas_gpu, 0, 0, W, H, shs::Color::black());
    z_ref.clear();
    z_gpu.clear();
    const int cols = (W + TILE_SIZE_X - 1) / TILE_SIZE_X, rows = (H + TILE_SIZE_Y - 1) / TILE_SIZE_Y;
    for (int ty = 0; ty < rows; ++ty)
        for (int tx = 0; tx < cols; ++tx)
        {
            const glm::ivec2 t_min(tx * TILE_SIZE_X, ty * TILE_SIZE_Y);
            const glm::ivec2 t_max(std::min((tx + 1) * TILE_SIZE_X, W) - 1, std::min((ty + 1) * TILE_SIZE_Y, H) - 1);
            for (const Obj& o : objs)
            {
                Uniforms uniforms;
                uniforms.model = world(o);
                uniforms.mvp = proj * view * uniforms.model;
                uniforms.light_dir = light_direction;
                uniforms.camera_pos = viewer.position;
                uniforms.color = o.color;
                for (size_t i = 0; i < tris.size(); i += 3)
                {
                    const std::vector<glm::vec3> tv = {tris[i], tris[i + 1], tris[i + 2]}, tn = {nrms[i], nrms[i + 1], nrms[i + 2]};
                    RendererSystem::draw_triangle_tile(canvas_ref, z_ref, tv, tn,
                        [&uniforms](const glm::vec3& p, const glm::vec3& n) { return blinn_phong_vertex_shader(p, n, uniforms); },
                        [&uniforms](const shs::Varyings& v) { return blinn_phong_fragment_shader(v, uniforms); }, t_min, t_max);
                }
            }
        }

    // ---- B200
    shs::b200::legacy::Renderer gpu(0, TILE_SIZE_X, TILE_SIZE_Y);
    if (!gpu.valid())
    {
        const bool refused = !gpu.begin_frame(canvas_gpu, z_gpu) && !gpu.draw(tris, nrms, proj, view, light_direction, viewer.position, objs[0].color) && !gpu.end_frame(canvas_gpu, z_gpu);
        std::printf("SKIP: %s (%s)\n", gpu.last_error(), refused ? "every call refused, nothing ran on the CPU" : "A CALL CLAIMED SUCCESS WITHOUT A DEVICE");
        return refused ? 77 : 1;
    }
    bool ok = gpu.begin_frame(canvas_gpu, z_gpu);
    for (const Obj& o : objs)
    {
        const glm::mat4 model = world(o);
        ok = ok && gpu.draw(tris, nrms, proj * view * model, model, light_direction, viewer.position, o.color);
    }
    ok = ok && gpu.end_frame(canvas_gpu, z_gpu);
    if (!ok) { std::printf("FAIL: %s\n", gpu.last_error()); return 1; }

    size_t z_diff = 0, covered = 0;
    int max_lsb = 0;
    const size_t n = (size_t)W * H;
    for (size_t i = 0; i < n; ++i)
    {
        if (std::memcmp(&z_ref.buffer().raw()[i], &z_gpu.buffer().raw()[i], 4) != 0) ++z_diff;
        if (z_ref.buffer().raw()[i] < std::numeric_limits<float>::max()) ++covered;
        const shs::Color a = canvas_ref.buffer().raw()[i], b = canvas_gpu.buffer().raw()[i];
        max_lsb = std::max(max_lsb, std::max(std::abs(a.r - b.r), std::max(std::abs(a.g - b.g), std::max(std::abs(a.b - b.b), std::abs(a.a - b.a)))));
    }
    const bool pass = z_diff == 0 && max_lsb <= 1 && covered > 2000;
    std::printf("legacy drop-in: %zu of %zu px covered, z-buffer differing %zu, canvas <= %d LSB | %s\n", covered, n, z_diff, max_lsb, pass ? "OK" : "MISMATCH");
    return pass ? 0 : 1;
}
