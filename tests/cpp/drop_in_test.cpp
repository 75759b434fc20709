// tests/cpp/drop_in_test.cpp -- the drop-in binding exercised the way the reference's own callers use the path
// (exp-plumbing/hello_pass_basics.cpp:629-912, exp-plumbing/hello_software_triangle.cpp:116-198): build a Scene with
// the reference's types, run the reference's CPU passes and the B200 passes, compare.
//
// Built only where /root/reference exists (tests/cpp/Makefile); the binary travels to the GPU box.
// Exit code 0 = parity within the north_star gates, 1 = mismatch, 77 = no CUDA device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "shs_b200/drop_in.hpp"

static shs::MeshData make_blob(int rings, int segs, float radius)
{
    shs::MeshData m{};
    for (int r = 0; r <= rings; ++r)
        for (int s = 0; s <= segs; ++s)
        {
            const float v = (float)r / rings, u = (float)s / segs;
            const float th = v * 3.14159265f, ph = u * 6.2831853f;
            const glm::vec3 n(std::sin(th) * std::cos(ph), std::cos(th), std::sin(th) * std::sin(ph));
            m.positions.push_back(n * (radius * (1.0f + 0.15f * std::sin(5.0f * ph) * std::sin(3.0f * th))));
            m.normals.push_back(n);
            m.uvs.push_back(glm::vec2(u * 3.0f, v * 2.0f));
        }
    for (int r = 0; r < rings; ++r)
        for (int s = 0; s < segs; ++s)
        {
            const uint32_t a = r * (segs + 1) + s, b = a + 1, c = a + segs + 1, d = c + 1;
            m.indices.insert(m.indices.end(), {a, b, c, b, d, c});
        }
    return m;
}

static int ulp(float a, float b)
{
    int32_t x, y;
    std::memcpy(&x, &a, 4);
    std::memcpy(&y, &b, 4);
    return std::abs(x - y);
}

int main()
{
    shs::b200::Device dev(0);
    if (!dev.valid()) { std::printf("SKIP: %s\n", dev.last_error()); return 77; }

    const int W = 320, H = 200;
    shs::ResourceRegistry resources{};
    const shs::MeshAssetHandle blob = resources.add_mesh(make_blob(24, 32, 1.0f));
    shs::Texture2DData tex(16, 16);
    for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) tex.at(x, y) = shs::Color{(uint8_t)(x * 16), (uint8_t)(y * 16), (uint8_t)(255 - x * 8), 255};
    const shs::TextureAssetHandle tex_h = resources.add_texture(tex);
    const shs::MaterialAssetHandle gold = resources.add_material(shs::MaterialData{"gold", glm::vec3(0.94f, 0.76f, 0.29f), 0.95f, 0.2f, 1.0f});
    const shs::MaterialAssetHandle textured = resources.add_material(shs::MaterialData{"tex", glm::vec3(1.0f), 0.2f, 0.5f, 1.0f, glm::vec3(0.0f), 0.0f, tex_h, 0, 0, 0});

    shs::Scene scene{};
    scene.resources = &resources;
    scene.cam.pos = glm::vec3(0.0f, 2.0f, -6.0f);
    scene.cam.view = shs::look_at_lh(scene.cam.pos, glm::vec3(0.0f, 0.3f, 0.0f), glm::vec3(0, 1, 0));
    scene.cam.proj = shs::perspective_lh_no(glm::radians(60.0f), (float)W / (float)H, 0.1f, 100.0f);
    scene.cam.viewproj = scene.cam.proj * scene.cam.view;
    scene.sun.dir_ws = glm::normalize(glm::vec3(-0.35f, -1.0f, -0.25f));
    scene.sun.color = glm::vec3(1.0f, 0.97f, 0.92f);
    scene.sun.intensity = 2.2f;
    for (int i = 0; i < 5; ++i)
    {
        shs::RenderItem it{};
        it.tr.pos = glm::vec3(-3.0f + 1.5f * i, 0.2f * i, 0.5f * (i % 2));
        it.tr.rot_euler = glm::vec3(0.1f * i, 0.7f * i, 0.0f);
        it.tr.scl = glm::vec3(0.6f + 0.1f * i);
        it.mesh = blob;
        it.mat = (i % 3 == 0) ? 0 : ((i % 2) ? gold : textured);
        it.object_id = 100u + (uint64_t)i; // explicit motion keys (the derived key mixes in the material handle, pass_pbr_forward.hpp:143-148)
        scene.items.push_back(it);
    }
    shs::FrameParams fp{};
    fp.w = W; fp.h = H;
    fp.pass.shadow.enable = true;

    // ---- reference CPU path
    shs::RT_ColorHDR hdr_ref(W, H), hdr_gpu(W, H);
    shs::RT_ColorDepthMotion dm_ref(W, H, 0.1f, 100.0f), dm_gpu(W, H, 0.1f, 100.0f);
    shs::RT_ShadowDepth sm_ref(256, 256), sm_gpu(256, 256);
    shs::RT_ColorLDR ldr_ref(W, H), ldr_gpu(W, H);
    auto run = [&](bool gpu, shs::RT_ColorHDR& hdr, shs::RT_ColorDepthMotion& dm, shs::RT_ShadowDepth& sm, shs::RT_ColorLDR& ldr, shs::Context& ctx) {
        shs::RTRegistry rtr{};
        const shs::RTHandle h_hdr = rtr.reg<shs::RTHandle>(&hdr);
        const shs::RTHandle h_dm = rtr.reg<shs::RTHandle>(&dm);
        const shs::RT_Shadow h_sm = rtr.reg<shs::RT_Shadow>(&sm);
        const shs::RTHandle h_ldr = rtr.reg<shs::RTHandle>(&ldr);
        shs::PassShadowMap::Inputs si{&scene, &fp, &rtr, h_sm};
        shs::PassPBRForward::Inputs fi{};
        fi.scene = &scene; fi.fp = &fp; fi.rtr = &rtr; fi.rt_hdr = h_hdr; fi.rt_motion = h_dm; fi.rt_shadow = shs::RTHandle{h_sm.id};
        shs::PassTonemap::Inputs ti{&fp, &rtr, h_hdr, h_ldr};
        if (gpu)
        {
            shs::b200::PassShadowMap(dev).execute(ctx, si);
            shs::b200::PassPBRForward(dev).execute(ctx, fi);
            shs::b200::PassTonemap(dev).execute(ctx, ti);
        }
        else
        {
            shs::PassShadowMap().execute(ctx, si);
            shs::PassPBRForward().execute(ctx, fi);
            shs::PassTonemap().execute(ctx, ti);
        }
    };
    shs::Context ctx_ref{}, ctx_gpu{};
    run(false, hdr_ref, dm_ref, sm_ref, ldr_ref, ctx_ref);
    run(true, hdr_gpu, dm_gpu, sm_gpu, ldr_gpu, ctx_gpu);

    int bad = 0;
    int max_depth_ulp = 0, max_shadow_ulp = 0, max_lsb = 0;
    double se = 0.0, peak = 1.0;
    for (size_t i = 0; i < dm_ref.depth.data.size(); ++i) max_depth_ulp = std::max(max_depth_ulp, ulp(dm_ref.depth.data[i], dm_gpu.depth.data[i]));
    for (size_t i = 0; i < sm_ref.depth.size(); ++i) max_shadow_ulp = std::max(max_shadow_ulp, ulp(sm_ref.depth[i], sm_gpu.depth[i]));
    for (size_t i = 0; i < hdr_ref.color.data.size(); ++i)
    {
        const shs::ColorF a = hdr_ref.color.data[i], b = hdr_gpu.color.data[i];
        se += (a.r - b.r) * (double)(a.r - b.r) + (a.g - b.g) * (double)(a.g - b.g) + (a.b - b.b) * (double)(a.b - b.b);
        peak = std::max(peak, (double)std::max(a.r, std::max(a.g, a.b)));
        const shs::Color p = ldr_ref.color.data[i], q = ldr_gpu.color.data[i];
        max_lsb = std::max(max_lsb, std::max(std::abs(p.r - q.r), std::max(std::abs(p.g - q.g), std::abs(p.b - q.b))));
    }
    const double mse = se / (3.0 * hdr_ref.color.data.size());
    const double psnr = mse == 0.0 ? 999.0 : 10.0 * std::log10(peak * peak / mse);
    if (ctx_ref.debug.tri_input != ctx_gpu.debug.tri_input || ctx_ref.debug.tri_after_clip != ctx_gpu.debug.tri_after_clip || ctx_ref.debug.tri_raster != ctx_gpu.debug.tri_raster) ++bad;
    if (std::memcmp(&ctx_ref.shadow.light_viewproj, &ctx_gpu.shadow.light_viewproj, 64) != 0) ++bad;
    if (max_depth_ulp > 1 || max_shadow_ulp > 1 || max_lsb > 1 || psnr < 60.0) ++bad;

    // ---- second frame: objects and camera moved, ProceduralSky background, motion vectors against the first frame
    // (FrameParams::pass.motion_vectors.enable defaults to true; Context::history carries the previous model matrices)
    shs::ProceduralSky sky(glm::vec3(0.2f, -0.35f, 0.9f));
    dev.register_sky(&sky, glm::vec3(0.2f, -0.35f, 0.9f));
    scene.sky = &sky;
    scene.cam.prev_viewproj = scene.cam.viewproj;
    scene.cam.pos = glm::vec3(0.25f, 2.1f, -6.1f);
    scene.cam.view = shs::look_at_lh(scene.cam.pos, glm::vec3(0.0f, 0.3f, 0.0f), glm::vec3(0, 1, 0));
    scene.cam.viewproj = scene.cam.proj * scene.cam.view;
    for (size_t i = 0; i < scene.items.size(); ++i)
    {
        scene.items[i].tr.pos += glm::vec3(0.15f * (float)i, 0.05f, -0.1f * (float)i);
        scene.items[i].tr.rot_euler.y += 0.2f;
    }
    run(false, hdr_ref, dm_ref, sm_ref, ldr_ref, ctx_ref);
    run(true, hdr_gpu, dm_gpu, sm_gpu, ldr_gpu, ctx_gpu);
    size_t motion_diff = 0, motion_nonzero = 0;
    for (size_t i = 0; i < dm_ref.motion.data.size(); ++i)
    {
        if (std::memcmp(&dm_ref.motion.data[i], &dm_gpu.motion.data[i], sizeof(shs::Motion2f)) != 0) ++motion_diff;
        if (dm_ref.motion.data[i].x != 0.0f || dm_ref.motion.data[i].y != 0.0f) ++motion_nonzero;
    }
    double se2 = 0.0, peak2 = 1.0;
    int max_depth_ulp2 = 0;
    for (size_t i = 0; i < hdr_ref.color.data.size(); ++i)
    {
        const shs::ColorF a = hdr_ref.color.data[i], b = hdr_gpu.color.data[i];
        se2 += (a.r - b.r) * (double)(a.r - b.r) + (a.g - b.g) * (double)(a.g - b.g) + (a.b - b.b) * (double)(a.b - b.b);
        peak2 = std::max(peak2, (double)std::max(a.r, std::max(a.g, a.b)));
        max_depth_ulp2 = std::max(max_depth_ulp2, ulp(dm_ref.depth.data[i], dm_gpu.depth.data[i]));
    }
    const double mse2 = se2 / (3.0 * hdr_ref.color.data.size());
    const double psnr2 = mse2 == 0.0 ? 999.0 : 10.0 * std::log10(peak2 * peak2 / mse2);
    if (motion_diff != 0 || motion_nonzero == 0 || psnr2 < 60.0 || max_depth_ulp2 > 1) ++bad;
    std::printf("frame 2 (sky + motion): motion vectors differing %zu of %zu (non-zero %zu)  depth<=%d ULP  HDR PSNR %.1f dB\n",
                motion_diff, dm_ref.motion.data.size(), motion_nonzero, max_depth_ulp2, psnr2);

    // ---- post passes on frame 2: PassLightShafts -> PassMotionBlur (reference order, hello_pass_basics.cpp) on the reference's
    // own LDR / depth / motion (uploaded, so that both sides start from identical bytes); RGBA8 must be bit-equal
    {
        fp.dt = 1.0f / 30.0f;
        fp.pass.motion_blur.enable = true;
        fp.pass.motion_blur.samples = 12;
        fp.pass.light_shafts.enable = true;
        // look towards the sun so that it projects inside the frame (pass_light_shafts.hpp:77-93)
        shs::Scene sc2 = scene;
        sc2.cam.view = shs::look_at_lh(sc2.cam.pos, sc2.cam.pos - sc2.sun.dir_ws * 10.0f + glm::vec3(0.4f, -0.3f, 0.0f), glm::vec3(0, 1, 0));
        sc2.cam.viewproj = sc2.cam.proj * sc2.cam.view;
        shs::RT_ColorLDR shafts_ref(W, H), shafts_gpu(W, H), blur_ref(W, H), blur_gpu(W, H);
        shs::RTRegistry rtr{};
        const shs::RTHandle h_in = rtr.reg<shs::RTHandle>(&ldr_ref);
        const shs::RTHandle h_dm = rtr.reg<shs::RTHandle>(&dm_ref);
        const shs::RTHandle h_sr = rtr.reg<shs::RTHandle>(&shafts_ref), h_sg = rtr.reg<shs::RTHandle>(&shafts_gpu);
        const shs::RTHandle h_br = rtr.reg<shs::RTHandle>(&blur_ref), h_bg = rtr.reg<shs::RTHandle>(&blur_gpu);
        shs::PassLightShafts::Inputs li{}; li.scene = &sc2; li.fp = &fp; li.rtr = &rtr; li.rt_input_ldr = h_in; li.rt_depth_like = h_dm;
        shs::PassMotionBlur::Inputs mi{}; mi.fp = &fp; mi.rtr = &rtr; mi.rt_motion = h_dm;
        li.rt_output_ldr = h_sr; shs::PassLightShafts().execute(ctx_ref, li);
        mi.rt_input_ldr = h_sr; mi.rt_output_ldr = h_br; shs::PassMotionBlur().execute(ctx_ref, mi);
        li.rt_output_ldr = h_sg; shs::b200::PassLightShafts(dev, true, false).execute(ctx_gpu, li);
        mi.rt_input_ldr = h_sg; mi.rt_output_ldr = h_bg; shs::b200::PassMotionBlur(dev, true, false).execute(ctx_gpu, mi);
        size_t shafts_diff = 0, blur_diff = 0, shafts_changed = 0, blur_changed = 0;
        for (size_t i = 0; i < shafts_ref.color.data.size(); ++i)
        {
            if (std::memcmp(&shafts_ref.color.data[i], &shafts_gpu.color.data[i], 4) != 0) ++shafts_diff;
            if (std::memcmp(&blur_ref.color.data[i], &blur_gpu.color.data[i], 4) != 0) ++blur_diff;
            if (std::memcmp(&shafts_ref.color.data[i], &ldr_ref.color.data[i], 3) != 0) ++shafts_changed;
            if (std::memcmp(&blur_ref.color.data[i], &shafts_ref.color.data[i], 3) != 0) ++blur_changed;
        }
        if (shafts_diff || blur_diff || !shafts_changed || !blur_changed) ++bad;
        std::printf("post passes: light shafts differing %zu (changed %zu px)  motion blur differing %zu (changed %zu px)\n", shafts_diff, shafts_changed, blur_diff, blur_changed);
    }

    // ---- rasterize_mesh called directly, like exp-plumbing/hello_software_triangle.cpp:187
    shs::RT_ColorHDR h2_ref(W, H), h2_gpu(W, H);
    shs::RT_ColorDepthMotion d2_ref(W, H, 0.1f, 100.0f), d2_gpu(W, H, 0.1f, 100.0f);
    shs::ShaderUniforms u{};
    u.model = glm::rotate(glm::translate(glm::mat4(1.0f), glm::vec3(0.2f, 0.1f, 0.0f)), 0.6f, glm::vec3(0, 1, 0));
    u.viewproj = scene.cam.viewproj;
    u.light_dir_ws = scene.sun.dir_ws; u.light_color = scene.sun.color; u.light_intensity = 2.0f;
    u.camera_pos = scene.cam.pos;
    u.base_color = glm::vec3(0.3f, 0.6f, 0.9f); u.metallic = 0.1f; u.roughness = 0.4f;
    const shs::RasterizerStats s_ref = shs::rasterize_mesh(*resources.get_mesh(blob), shs::make_blinn_phong_program(), u, shs::RasterizerTarget{&h2_ref, &d2_ref});
    const shs::RasterizerStats s_gpu = shs::b200::rasterize_mesh(dev, *resources.get_mesh(blob), shs::b200::BuiltinProgram::BlinnPhong, u, shs::RasterizerTarget{&h2_gpu, &d2_gpu});
    int max_d2 = 0;
    for (size_t i = 0; i < d2_ref.depth.data.size(); ++i) max_d2 = std::max(max_d2, ulp(d2_ref.depth.data[i], d2_gpu.depth.data[i]));
    if (s_ref.tri_input != s_gpu.tri_input || s_ref.tri_after_clip != s_gpu.tri_after_clip || s_ref.tri_raster != s_gpu.tri_raster || max_d2 > 1) ++bad;

    // ---- the usual caller pattern: the host clears / refills its targets between draws.  The motion plane of BOTH sides is
    // filled with a marker the draw must leave alone on uncovered pixels and overwrite on covered ones (rasterizer.hpp:388-411
    // touches covered pixels only); a device twin that kept a stale motion plane would bring old vectors back here.
    {
        for (size_t i = 0; i < d2_ref.motion.data.size(); ++i)
        {
            d2_ref.motion.data[i] = shs::Motion2f{3.0f + (float)(i % 7), -2.0f};
            d2_gpu.motion.data[i] = d2_ref.motion.data[i];
            d2_ref.depth.data[i] = 1.0f; d2_gpu.depth.data[i] = 1.0f;
        }
        shs::ShaderUniforms u2 = u;
        u2.enable_motion_vectors = true;
        u2.prev_model = glm::translate(u.model, glm::vec3(0.05f, -0.02f, 0.0f));
        u2.prev_viewproj = scene.cam.viewproj;
        shs::rasterize_mesh(*resources.get_mesh(blob), shs::make_blinn_phong_program(), u2, shs::RasterizerTarget{&h2_ref, &d2_ref});
        shs::b200::rasterize_mesh(dev, *resources.get_mesh(blob), shs::b200::BuiltinProgram::BlinnPhong, u2, shs::RasterizerTarget{&h2_gpu, &d2_gpu});
        size_t mdiff = 0, kept = 0, written = 0;
        for (size_t i = 0; i < d2_ref.motion.data.size(); ++i)
        {
            if (std::memcmp(&d2_ref.motion.data[i], &d2_gpu.motion.data[i], sizeof(shs::Motion2f)) != 0) ++mdiff;
            const bool marker = d2_ref.motion.data[i].x == 3.0f + (float)(i % 7) && d2_ref.motion.data[i].y == -2.0f;
            if (marker) ++kept; else ++written;
        }
        if (mdiff != 0 || kept == 0 || written == 0) ++bad;
        std::printf("rasterize_mesh with motion vectors over a pre-filled plane: %zu differing, %zu pixels kept, %zu written\n", mdiff, kept, written);
    }

    std::printf("passes: tris %llu/%llu/%llu  depth<=%d ULP  shadow<=%d ULP  LDR<=%d LSB  HDR PSNR %.1f dB | rasterize_mesh: depth<=%d ULP tris %llu | %s\n",
                (unsigned long long)ctx_gpu.debug.tri_input, (unsigned long long)ctx_gpu.debug.tri_after_clip, (unsigned long long)ctx_gpu.debug.tri_raster,
                max_depth_ulp, max_shadow_ulp, max_lsb, psnr, max_d2, (unsigned long long)s_gpu.tri_raster, bad ? "MISMATCH" : "OK");
    return bad ? 1 : 0;
}
