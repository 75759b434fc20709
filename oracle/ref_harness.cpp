// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (builds into oracle/_ref/libshs_ref.so).
//
// C-ABI wrapper around the reference's OWN raster-path headers, compiled where they lie under
// /root/reference (nothing is copied) against oracle/glm_shim (GLM is the one missing
// dependency of these headers).  It exists to (a) pin the CPU restatement
// in oracle/oracle.cpp against the reference's real control flow, (b) generate the golden
// fixtures under tests/golden/, and (c) serve as the "reference" CPU baseline of bench.py.
//
// Only marshalling lives here: every pixel is produced by shs::rasterize_mesh,
// shs::PassPBRForward::execute, shs::PassShadowMap::execute, shs::PassTonemap::execute and the
// light packers of shs/lighting/light_types.hpp.
//
// Build: see oracle/Makefile (g++ -std=c++20 -O2 -ffp-contract=off, no -march, no fast-math:
// the reference's effective x86-64 semantics, cpp-folders/src/exp-plumbing/CMakeLists.txt:13).

#include <cstring>
#include <memory>
#include <vector>

#include "shs/core/context.hpp"
#include "shs/job/thread_pool_job_system.hpp"
#include "shs/lighting/light_types.hpp"
#include "shs/passes/pass_light_shafts.hpp"
#include "shs/passes/pass_motion_blur.hpp"
#include "shs/passes/pass_pbr_forward.hpp"
#include "shs/passes/pass_shadow_map.hpp"
#include "shs/passes/pass_tonemap.hpp"
#include "shs/shader/builtin_shaders.hpp"
#include "shs/sky/cubemap_sky.hpp"
#include "shs/sky/procedural_sky.hpp"
#include "shs/sw_render/rasterizer.hpp"

#include "oracle_abi.h"

namespace
{
    glm::mat4 load_mat4(const float* m)
    {
        glm::mat4 r;
        std::memcpy(&r, m, 64);
        return r;
    }

    void store_mat4(const glm::mat4& m, float* out) { std::memcpy(out, &m, 64); }

    glm::vec3 load_vec3(const float* v) { return glm::vec3(v[0], v[1], v[2]); }

    shs::MeshData make_mesh(const ShsoMesh& m)
    {
        shs::MeshData out{};
        out.positions.resize(m.n_positions);
        if (m.n_positions) std::memcpy(out.positions.data(), m.positions, (size_t)m.n_positions * 12);
        out.normals.resize(m.n_normals);
        if (m.n_normals) std::memcpy(out.normals.data(), m.normals, (size_t)m.n_normals * 12);
        out.uvs.resize(m.n_uvs);
        if (m.n_uvs) std::memcpy(out.uvs.data(), m.uvs, (size_t)m.n_uvs * 8);
        out.indices.assign(m.indices, m.indices + m.n_indices);
        return out;
    }

    shs::Texture2DData make_texture(const ShsoTexture& t)
    {
        shs::Texture2DData out(t.w, t.h);
        std::memcpy(out.texels.data(), t.rgba, (size_t)t.w * (size_t)t.h * 4);
        return out;
    }

    struct Assets
    {
        shs::ResourceRegistry reg{};
        explicit Assets(const ShsoAssets* a)
        {
            for (uint32_t i = 0; a && i < a->n_meshes; ++i) reg.add_mesh(make_mesh(a->meshes[i]));
            for (uint32_t i = 0; a && i < a->n_textures; ++i) reg.add_texture(make_texture(a->textures[i]));
        }
    };

    // Assets are rebuilt only when their CONTENT changes, so that a timed loop (bench.py --impl reference) does not re-copy
    // meshes every frame.  The key is a hash of every byte the descriptor block points at, not the block's address: a test
    // that builds one scene after another gets freed addresses handed back with different meshes behind them.
    struct AssetCache
    {
        uint64_t key = 0;
        std::unique_ptr<Assets> assets{};
        // arrays up to 1 MiB are hashed whole; larger ones (4K cubemap faces in a timed loop) by their size, their first and
        // last 64 KiB and every 4099th 8-byte word in between
        static void mix(uint64_t& h, const void* p, size_t n)
        {
            const unsigned char* b = static_cast<const unsigned char*>(p);
            if (n > (1u << 20))
            {
                mix(h, b, 65536);
                mix(h, b + n - 65536, 65536);
                for (size_t i = 65536; i + 8 <= n - 65536; i += 8 * 4099) { uint64_t w; std::memcpy(&w, b + i, 8); h = (h ^ w) * 0x100000001b3ull; h ^= h >> 29; }
                h = (h ^ (uint64_t)n) * 0x100000001b3ull;
                return;
            }
            size_t i = 0;
            for (; i + 8 <= n; i += 8) { uint64_t w; std::memcpy(&w, b + i, 8); h = (h ^ w) * 0x100000001b3ull; h ^= h >> 29; }
            for (; i < n; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
            h = (h ^ (uint64_t)n) * 0x100000001b3ull;
        }
        static uint64_t fingerprint(const ShsoAssets* a)
        {
            uint64_t h = 0xcbf29ce484222325ull;
            if (!a) return h;
            mix(h, &a->n_meshes, 4);
            mix(h, &a->n_textures, 4);
            for (uint32_t i = 0; i < a->n_meshes; ++i)
            {
                const ShsoMesh& m = a->meshes[i];
                mix(h, m.positions, (size_t)m.n_positions * 12);
                mix(h, m.normals, (size_t)m.n_normals * 12);
                mix(h, m.uvs, (size_t)m.n_uvs * 8);
                mix(h, m.indices, (size_t)m.n_indices * 4);
            }
            for (uint32_t i = 0; i < a->n_textures; ++i)
            {
                const ShsoTexture& t = a->textures[i];
                mix(h, &t.w, sizeof(t.w));
                mix(h, &t.h, sizeof(t.h));
                mix(h, t.rgba, (size_t)t.w * (size_t)t.h * 4);
            }
            return h;
        }
        Assets& get(const ShsoAssets* a)
        {
            const uint64_t k = fingerprint(a);
            if (!assets || key != k)
            {
                assets = std::make_unique<Assets>(a);
                key = k;
            }
            return *assets;
        }
    };
    AssetCache g_assets;

    std::unique_ptr<shs::ThreadPoolJobSystem> g_jobs;

    shs::ShaderProgram program_for(int32_t shader_id)
    {
        switch (shader_id)
        {
        case SHSB_SHADER_PBR_MR: return shs::make_pbr_mr_program();
        case SHSB_SHADER_BLINN_PHONG: return shs::make_blinn_phong_program();
        case SHSB_SHADER_DEBUG_ALBEDO: return shs::make_debug_view_shader_program(shs::DebugViewMode::Albedo);
        case SHSB_SHADER_DEBUG_NORMAL: return shs::make_debug_view_shader_program(shs::DebugViewMode::Normal);
        case SHSB_SHADER_DEBUG_DEPTH: return shs::make_debug_view_shader_program(shs::DebugViewMode::Depth);
        default: return shs::ShaderProgram{};
        }
    }

    void fill_frame_params(const ShsbFrameParams* fp, shs::FrameParams& out)
    {
        out.shading_model = (fp->shading_model == SHSB_SHADING_BLINN_PHONG) ? shs::ShadingModel::BlinnPhong : shs::ShadingModel::PBRMetalRough;
        out.debug_view = (shs::DebugViewMode)fp->debug_view;
        out.cull_mode = (shs::CullMode)fp->cull_mode;
        out.front_face_ccw = fp->front_face_ccw != 0;
        out.pass.shadow.enable = fp->shadow_enable != 0;
        out.pass.shadow.bias_const = fp->shadow_bias_const;
        out.pass.shadow.bias_slope = fp->shadow_bias_slope;
        out.pass.shadow.pcf_radius = fp->shadow_pcf_radius;
        out.pass.shadow.pcf_step = fp->shadow_pcf_step;
        out.pass.shadow.strength = fp->shadow_strength;
        out.pass.tonemap.exposure = fp->exposure;
        out.pass.tonemap.gamma = fp->gamma;
        out.pass.motion_vectors.enable = fp->motion_vectors_enable != 0;
    }

    // Scene::sky from the POD scene: the reference's own ProceduralSky / CubemapSky objects (kept alive in `holder`)
    void fill_sky(const ShsbScene* s, shs::Scene& scene, const shs::ResourceRegistry& reg, std::unique_ptr<shs::ISkyModel>& holder)
    {
        holder.reset();
        if (s->sky_kind == SHSB_SKY_PROCEDURAL) holder = std::make_unique<shs::ProceduralSky>(load_vec3(s->sky_sun_dir_ws));
        else if (s->sky_kind == SHSB_SKY_CUBEMAP)
        {
            shs::CubemapData cm{};
            for (int i = 0; i < 6; ++i)
            {
                const shs::Texture2DData* t = reg.get_texture((shs::TextureAssetHandle)s->sky_faces[i]);
                if (t) cm.face[(size_t)i] = *t;
            }
            holder = std::make_unique<shs::CubemapSky>(std::move(cm), s->sky_intensity);
        }
        scene.sky = holder.get();
    }

    void fill_scene(const ShsbScene* s, shs::Scene& scene, shs::ResourceRegistry& reg)
    {
        scene.resources = &reg;
        scene.cam.viewproj = load_mat4(s->cam_viewproj);
        scene.cam.prev_viewproj = load_mat4(s->cam_prev_viewproj);
        scene.cam.pos = load_vec3(s->cam_pos);
        scene.sun.dir_ws = load_vec3(s->sun_dir_ws);
        scene.sun.color = load_vec3(s->sun_color);
        scene.sun.intensity = s->sun_intensity;
        scene.sky = nullptr;
        scene.items.clear();
        scene.items.reserve(s->n_items);
        for (uint32_t i = 0; i < s->n_items; ++i)
        {
            const ShsbRenderItem& it = s->items[i];
            shs::RenderItem ri{};
            ri.tr.pos = load_vec3(it.tr.pos);
            ri.tr.rot_euler = load_vec3(it.tr.rot_euler);
            ri.tr.scl = load_vec3(it.tr.scl);
            ri.mesh = it.mesh;
            ri.mat = 0;
            if (it.has_material)
            {
                shs::MaterialData md{};
                md.base_color = load_vec3(it.base_color);
                md.metallic = it.metallic;
                md.roughness = it.roughness;
                md.ao = it.ao;
                md.base_color_tex = it.base_color_tex;
                ri.mat = reg.add_material(md);
            }
            ri.casts_shadow = it.casts_shadow != 0;
            ri.visible = it.visible != 0;
            ri.object_id = it.object_id;
            scene.items.push_back(ri);
        }
    }
}

extern "C" {

int32_t shsref_set_threads(int32_t n)
{
    // ThreadPoolJobSystem as in exp-plumbing/hello_pass_basics.cpp:629-630.
    g_jobs.reset();
    if (n > 1) g_jobs = std::make_unique<shs::ThreadPoolJobSystem>((size_t)n);
    return n;
}

void shsref_model_from_transform(const ShsbTransform* tr, float out_model[16])
{
    // passes/pass_pbr_forward.hpp:136-141
    glm::mat4 model(1.0f);
    model = glm::translate(model, load_vec3(tr->pos));
    model = glm::rotate(model, tr->rot_euler[0], glm::vec3(1.0f, 0.0f, 0.0f));
    model = glm::rotate(model, tr->rot_euler[1], glm::vec3(0.0f, 1.0f, 0.0f));
    model = glm::rotate(model, tr->rot_euler[2], glm::vec3(0.0f, 0.0f, 1.0f));
    model = glm::scale(model, load_vec3(tr->scl));
    store_mat4(model, out_model);
}

void shsref_camera_viewproj(const float eye[3], const float target[3], const float up[3],
                            float fovy_radians, float aspect, float znear, float zfar, float out_viewproj[16])
{
    // camera/convention.hpp:19-27; viewproj = proj * view as every reference demo does.
    const glm::mat4 view = shs::look_at_lh(load_vec3(eye), load_vec3(target), load_vec3(up));
    const glm::mat4 proj = shs::perspective_lh_no(fovy_radians, aspect, znear, zfar);
    store_mat4(proj * view, out_viewproj);
}

void shsref_pack_point_light(const float pos[3], float range, const float color[3], float intensity,
                             uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff,
                             int32_t jolt_bounds, void* out_record160)
{
    shs::PointLight p{};
    p.common.position_ws = load_vec3(pos);
    p.common.range = range;
    p.common.color = load_vec3(color);
    p.common.intensity = intensity;
    p.common.attenuation_model = (shs::LightAttenuationModel)atten_model;
    p.common.attenuation_power = atten_power;
    p.common.attenuation_bias = atten_bias;
    p.common.attenuation_cutoff = atten_cutoff;
    shs::CullingLightGPU rec = shs::make_point_culling_light(p);
    if (jolt_bounds)
    {
        // geometry/scene_shape.hpp:56-81 for a Jolt sphere shape: AABB = c +- r, sphere radius = |extent|.
        shs::AABB box{};
        box.minv = p.common.position_ws - glm::vec3(range);
        box.maxv = p.common.position_ws + glm::vec3(range);
        shs::Sphere s{};
        s.center = box.center();
        s.radius = glm::length(box.extent());
        shs::assign_light_cull_bounds(rec, s, box);
    }
    static_assert(sizeof(shs::CullingLightGPU) == SHSB_LIGHT_RECORD_BYTES, "record size");
    std::memcpy(out_record160, &rec, sizeof(rec));
}

void shsref_pack_spot_light(const float pos[3], float range, const float color[3], float intensity,
                            const float dir[3], float inner_rad, float outer_rad,
                            uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff,
                            void* out_record160)
{
    shs::SpotLight s{};
    s.common.position_ws = load_vec3(pos);
    s.common.range = range;
    s.common.color = load_vec3(color);
    s.common.intensity = intensity;
    s.common.attenuation_model = (shs::LightAttenuationModel)atten_model;
    s.common.attenuation_power = atten_power;
    s.common.attenuation_bias = atten_bias;
    s.common.attenuation_cutoff = atten_cutoff;
    s.direction_ws = load_vec3(dir);
    s.inner_angle_rad = inner_rad;
    s.outer_angle_rad = outer_rad;
    const shs::CullingLightGPU rec = shs::make_spot_culling_light(s);
    std::memcpy(out_record160, &rec, sizeof(rec));
}

// Area lights: the reference's own packers (lighting/light_types.hpp:379-436).  Reference-only entry points: the restatement has
// no packer of its own for these, the records they produce are committed as a fixture (tests/golden/golden_area_lights.npz).
void shsref_pack_rect_light(const float pos[3], float range, const float color[3], float intensity,
                            const float dir[3], const float right[3], float half_x, float half_y,
                            uint32_t flags, uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff, void* out_record160)
{
    shs::RectAreaLight r{};
    r.common.position_ws = load_vec3(pos);
    r.common.range = range;
    r.common.color = load_vec3(color);
    r.common.intensity = intensity;
    r.common.flags = flags;
    r.common.attenuation_model = (shs::LightAttenuationModel)atten_model;
    r.common.attenuation_power = atten_power;
    r.common.attenuation_bias = atten_bias;
    r.common.attenuation_cutoff = atten_cutoff;
    r.direction_ws = load_vec3(dir);
    r.right_ws = load_vec3(right);
    r.half_extents = glm::vec2(half_x, half_y);
    const shs::CullingLightGPU rec = shs::make_rect_area_culling_light(r);
    std::memcpy(out_record160, &rec, sizeof(rec));
}

void shsref_pack_tube_light(const float pos[3], float range, const float color[3], float intensity,
                            const float axis[3], float half_length, float radius,
                            uint32_t flags, uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff, void* out_record160)
{
    shs::TubeAreaLight t{};
    t.common.position_ws = load_vec3(pos);
    t.common.range = range;
    t.common.color = load_vec3(color);
    t.common.intensity = intensity;
    t.common.flags = flags;
    t.common.attenuation_model = (shs::LightAttenuationModel)atten_model;
    t.common.attenuation_power = atten_power;
    t.common.attenuation_bias = atten_bias;
    t.common.attenuation_cutoff = atten_cutoff;
    t.axis_ws = load_vec3(axis);
    t.half_length = half_length;
    t.radius = radius;
    const shs::CullingLightGPU rec = shs::make_tube_area_culling_light(t);
    std::memcpy(out_record160, &rec, sizeof(rec));
}

int32_t shsref_rasterize_mesh(const ShsoAssets* assets, shsb_mesh mesh, int32_t shader_id,
                              const ShsbUniforms* u, const ShsoTarget* tgt, const ShsbRasterCfg* cfg,
                              uint32_t key_base, ShsbStats* out_stats)
{
    Assets& A = g_assets.get(assets);
    const shs::MeshData* md = A.reg.get_mesh(mesh);
    if (!md || !tgt || !tgt->hdr) return SHSB_E_INVALID_ARGUMENT;
    const int W = tgt->w, H = tgt->h;

    shs::RT_ColorHDR hdr(W, H);
    std::memcpy(hdr.color.data.data(), tgt->hdr, (size_t)W * H * 16);
    std::unique_ptr<shs::RT_ColorDepthMotion> dm{};
    if (tgt->depth)
    {
        dm = std::make_unique<shs::RT_ColorDepthMotion>(W, H, tgt->zn, tgt->zf);
        std::memcpy(dm->depth.data.data(), tgt->depth, (size_t)W * H * 4);
    }
    std::unique_ptr<shs::RT_ShadowDepth> sm{};
    if (tgt->shadow && u->shadow_map)
    {
        sm = std::make_unique<shs::RT_ShadowDepth>(tgt->shadow_w, tgt->shadow_h);
        std::memcpy(sm->depth.data(), tgt->shadow, (size_t)tgt->shadow_w * tgt->shadow_h * 4);
    }

    shs::ShaderUniforms su{};
    su.model = load_mat4(u->model);
    su.viewproj = load_mat4(u->viewproj);
    su.prev_model = su.model;
    su.prev_viewproj = su.viewproj;
    su.light_dir_ws = load_vec3(u->light_dir_ws);
    su.light_color = load_vec3(u->light_color);
    su.light_intensity = u->light_intensity;
    su.camera_pos = load_vec3(u->camera_pos);
    su.base_color = load_vec3(u->base_color);
    su.metallic = u->metallic;
    su.roughness = u->roughness;
    su.ao = u->ao;
    su.base_color_tex = A.reg.get_texture(u->base_color_tex);
    su.shadow_map = sm.get();
    su.light_viewproj = load_mat4(u->light_viewproj);
    su.shadow_bias_const = u->shadow_bias_const;
    su.shadow_bias_slope = u->shadow_bias_slope;
    su.shadow_pcf_radius = u->shadow_pcf_radius;
    su.shadow_pcf_step = u->shadow_pcf_step;
    su.shadow_strength = u->shadow_strength;
    su.enable_motion_vectors = u->enable_motion_vectors != 0;
    su.prev_model = load_mat4(u->prev_model);
    su.prev_viewproj = load_mat4(u->prev_viewproj);

    shs::ShaderProgram prog = program_for(shader_id);
    if (!prog.valid()) return SHSB_E_UNSUPPORTED_SHADER;

    // AOVs through the reference's own callback API: the VS wrapper counts corner invocations
    // (3 per index-valid triangle, rasterizer.hpp:222-226), the FS wrapper records which
    // triangle produced each fragment the rasterizer decided to shade.
    std::vector<uint32_t> valid_tris{};
    if (tgt->tri_id || tgt->coverage)
    {
        const bool indexed = !md->indices.empty();
        const size_t tri_count = indexed ? (md->indices.size() / 3) : (md->positions.size() / 3);
        for (size_t ti = 0; ti < tri_count; ++ti)
        {
            uint32_t i0 = indexed ? md->indices[ti * 3 + 0] : (uint32_t)(ti * 3 + 0);
            uint32_t i1 = indexed ? md->indices[ti * 3 + 1] : (uint32_t)(ti * 3 + 1);
            uint32_t i2 = indexed ? md->indices[ti * 3 + 2] : (uint32_t)(ti * 3 + 2);
            if (i0 >= md->positions.size() || i1 >= md->positions.size() || i2 >= md->positions.size()) continue;
            valid_tris.push_back((uint32_t)ti);
        }
        uint64_t* vs_calls = new uint64_t(0);
        std::shared_ptr<uint64_t> counter(vs_calls);
        const shs::VertexShaderFn base_vs = prog.vs;
        const shs::FragmentShaderFn base_fs = prog.fs;
        prog.vs = [counter, base_vs](const shs::ShaderVertex& v, const shs::ShaderUniforms& uu) {
            ++(*counter);
            return base_vs(v, uu);
        };
        uint32_t* tri_id = tgt->tri_id;
        uint32_t* coverage = tgt->coverage;
        const std::vector<uint32_t>* vt = &valid_tris;
        prog.fs = [counter, base_fs, tri_id, coverage, vt, W, key_base](const shs::FragmentIn& fin, const shs::ShaderUniforms& uu) {
            const uint64_t valid_index = (*counter) / 3 - 1;
            const size_t px = (size_t)fin.py * (size_t)W + (size_t)fin.px;
            if (tri_id) tri_id[px] = key_base + (*vt)[(size_t)valid_index] * 8u;
            if (coverage) coverage[px] += 1u;
            return base_fs(fin, uu);
        };
    }

    shs::RasterizerTarget rt{};
    rt.hdr = &hdr;
    rt.depth_motion = dm.get();
    shs::RasterizerConfig rc{};
    rc.cull_mode = (shs::RasterizerCullMode)cfg->cull_mode;
    rc.front_face_ccw = cfg->front_face_ccw != 0;
    rc.job_system = (tgt->tri_id || tgt->coverage) ? nullptr : g_jobs.get();

    const shs::RasterizerStats st = shs::rasterize_mesh(*md, prog, su, rt, rc);

    std::memcpy(tgt->hdr, hdr.color.data.data(), (size_t)W * H * 16);
    if (dm) std::memcpy(tgt->depth, dm->depth.data.data(), (size_t)W * H * 4);
    if (dm && tgt->motion) std::memcpy(tgt->motion, dm->motion.data.data(), (size_t)W * H * 8);
    if (out_stats)
    {
        out_stats->tri_input += st.tri_input;
        out_stats->tri_after_clip += st.tri_after_clip;
        out_stats->tri_raster += st.tri_raster;
    }
    return SHSB_OK;
}

int32_t shsref_pass_pbr_forward(const ShsoAssets* assets, const ShsbScene* s, const ShsbFrameParams* fp,
                                const ShsoTarget* tgt, const float* shadow_light_viewproj,
                                int32_t preserve_existing_depth, ShsbStats* out_stats)
{
    return shsref_pass_pbr_forward_history(assets, s, fp, tgt, shadow_light_viewproj, preserve_existing_depth, nullptr, out_stats);
}

int32_t shsref_pass_pbr_forward_history(const ShsoAssets* assets, const ShsbScene* s, const ShsbFrameParams* fp,
                                        const ShsoTarget* tgt, const float* shadow_light_viewproj,
                                        int32_t preserve_existing_depth, const float* prev_models16, ShsbStats* out_stats)
{
    if (!s || !fp || !tgt || !tgt->hdr) return SHSB_E_INVALID_ARGUMENT;
    Assets base(nullptr);
    Assets& A = g_assets.get(assets);
    // Materials are per call: copy the registry value (meshes/textures are shared by value too,
    // but this happens outside any timed region except for material adds).
    shs::ResourceRegistry reg = A.reg;

    const int W = tgt->w, H = tgt->h;
    shs::RT_ColorHDR hdr(W, H);
    shs::RT_ColorDepthMotion dm(W, H, tgt->zn, tgt->zf);
    if (tgt->depth && preserve_existing_depth) std::memcpy(dm.depth.data.data(), tgt->depth, (size_t)W * H * 4);
    shs::RT_ShadowDepth sm{};
    const bool have_shadow = tgt->shadow != nullptr && shadow_light_viewproj != nullptr;
    if (have_shadow)
    {
        sm.resize(tgt->shadow_w, tgt->shadow_h);
        std::memcpy(sm.depth.data(), tgt->shadow, (size_t)tgt->shadow_w * tgt->shadow_h * 4);
    }

    shs::RTRegistry rtr{};
    const shs::RTHandle h_hdr = rtr.reg<shs::RTHandle>(&hdr);
    shs::RTHandle h_motion{};
    if (tgt->depth) h_motion = rtr.reg<shs::RTHandle>(&dm);
    shs::RTHandle h_shadow{};
    if (have_shadow) h_shadow = rtr.reg<shs::RTHandle>(&sm);

    shs::Context ctx{};
    ctx.job_system = g_jobs.get();
    if (have_shadow)
    {
        ctx.shadow.map = &sm;
        ctx.shadow.light_viewproj = load_mat4(shadow_light_viewproj);
        ctx.shadow.valid = true;
    }

    shs::Scene scene{};
    fill_scene(s, scene, reg);
    std::unique_ptr<shs::ISkyModel> sky_holder;
    fill_sky(s, scene, reg, sky_holder);
    shs::FrameParams f{};
    f.w = W;
    f.h = H;
    fill_frame_params(fp, f);
    if (prev_models16)
    {
        // Context::history as the previous PassPBRForward::execute left it (pass_pbr_forward.hpp:143-155, 212-213)
        for (size_t i = 0; i < scene.items.size(); ++i)
        {
            const auto& item = scene.items[i];
            uint64_t key = item.object_id;
            if (key == 0)
            {
                key = ((uint64_t)item.mesh << 32) ^ (uint64_t)item.mat ^ ((uint64_t)i + 1u);
                if (key == 0) key = 1;
            }
            ctx.history.prev_model_by_object[key] = load_mat4(prev_models16 + i * 16);
        }
        ctx.history.has_prev_frame = true;
    }

    shs::PassPBRForward pass{};
    shs::PassPBRForward::Inputs in{};
    in.scene = &scene;
    in.fp = &f;
    in.rtr = &rtr;
    in.rt_hdr = h_hdr;
    in.rt_motion = h_motion;
    in.rt_shadow = h_shadow;
    in.preserve_existing_depth = preserve_existing_depth != 0;
    pass.execute(ctx, in);

    std::memcpy(tgt->hdr, hdr.color.data.data(), (size_t)W * H * 16);
    if (tgt->depth) std::memcpy(tgt->depth, dm.depth.data.data(), (size_t)W * H * 4);
    if (tgt->depth && tgt->motion) std::memcpy(tgt->motion, dm.motion.data.data(), (size_t)W * H * 8);
    if (out_stats)
    {
        out_stats->tri_input = ctx.debug.tri_input;
        out_stats->tri_after_clip = ctx.debug.tri_after_clip;
        out_stats->tri_raster = ctx.debug.tri_raster;
    }
    return SHSB_OK;
}

int32_t shsref_pass_shadow_map(const ShsoAssets* assets, const ShsbScene* s, const ShsbFrameParams* fp,
                               float* shadow, int32_t sw, int32_t sh, float out_light_viewproj[16])
{
    if (!s || !fp || !shadow) return SHSB_E_INVALID_ARGUMENT;
    Assets& A = g_assets.get(assets);
    shs::ResourceRegistry reg = A.reg;
    shs::RT_ShadowDepth sm(sw, sh);
    shs::RTRegistry rtr{};
    const shs::RT_Shadow h = rtr.reg<shs::RT_Shadow>(&sm);
    shs::Context ctx{};
    ctx.job_system = g_jobs.get();
    shs::Scene scene{};
    fill_scene(s, scene, reg);
    shs::FrameParams f{};
    fill_frame_params(fp, f);

    shs::PassShadowMap pass{};
    shs::PassShadowMap::Inputs in{};
    in.scene = &scene;
    in.fp = &f;
    in.rtr = &rtr;
    in.rt_shadow = h;
    pass.execute(ctx, in);
    std::memcpy(shadow, sm.depth.data(), (size_t)sw * sh * 4);
    if (out_light_viewproj) store_mat4(ctx.shadow.light_viewproj, out_light_viewproj);
    return ctx.shadow.valid ? SHSB_OK : SHSB_E_INVALID_ARGUMENT;
}

int32_t shsref_pass_tonemap(const float* hdr_in, int32_t w, int32_t h, float exposure, float gamma, uint8_t* out_ldr)
{
    shs::RT_ColorHDR hdr(w, h);
    std::memcpy(hdr.color.data.data(), hdr_in, (size_t)w * h * 16);
    shs::RT_ColorLDR ldr(w, h);
    shs::RTRegistry rtr{};
    const shs::RTHandle h_hdr = rtr.reg<shs::RTHandle>(&hdr);
    const shs::RTHandle h_ldr = rtr.reg<shs::RTHandle>(&ldr);
    shs::Context ctx{};
    ctx.job_system = g_jobs.get();
    shs::FrameParams f{};
    f.pass.tonemap.exposure = exposure;
    f.pass.tonemap.gamma = gamma;
    shs::PassTonemap pass{};
    shs::PassTonemap::Inputs in{};
    in.fp = &f;
    in.rtr = &rtr;
    in.rt_hdr = h_hdr;
    in.rt_ldr = h_ldr;
    pass.execute(ctx, in);
    std::memcpy(out_ldr, ldr.color.data.data(), (size_t)w * h * 4);
    return SHSB_OK;
}

int32_t shsref_pass_motion_blur(const ShsbMotionBlurParams* p, const uint8_t* src_ldr, const float* motion, const float* depth,
                                int32_t w, int32_t h, uint8_t* out_ldr)
{
    shs::RT_ColorLDR src(w, h), dst(w, h);
    std::memcpy(src.color.data.data(), src_ldr, (size_t)w * h * 4);
    shs::RT_ColorDepthMotion dm(w, h, 0.1f, 1000.0f);
    std::memcpy(dm.depth.data.data(), depth, (size_t)w * h * 4);
    std::memcpy(dm.motion.data.data(), motion, (size_t)w * h * 8);
    shs::RTRegistry rtr{};
    shs::Context ctx{};
    ctx.job_system = g_jobs.get();
    shs::FrameParams f{};
    f.dt = p->dt;
    f.pass.motion_blur.enable = p->enable != 0;
    f.pass.motion_blur.samples = p->samples;
    f.pass.motion_blur.strength = p->strength;
    f.pass.motion_blur.max_velocity_px = p->max_velocity_px;
    f.pass.motion_blur.min_velocity_px = p->min_velocity_px;
    f.pass.motion_blur.depth_reject = p->depth_reject;
    shs::PassMotionBlur::Inputs in{};
    in.fp = &f;
    in.rtr = &rtr;
    in.rt_input_ldr = rtr.reg<shs::RTHandle>(&src);
    in.rt_output_ldr = rtr.reg<shs::RTHandle>(&dst);
    in.rt_motion = rtr.reg<shs::RTHandle>(&dm);
    shs::PassMotionBlur{}.execute(ctx, in);
    std::memcpy(out_ldr, dst.color.data.data(), (size_t)w * h * 4);
    return SHSB_OK;
}

int32_t shsref_pass_light_shafts(const ShsbLightShaftsParams* p, const uint8_t* src_ldr, const float* depth, int32_t w, int32_t h, uint8_t* out_ldr)
{
    shs::RT_ColorLDR src(w, h), dst(w, h);
    std::memcpy(src.color.data.data(), src_ldr, (size_t)w * h * 4);
    shs::RT_ColorDepthMotion dm(w, h, 0.1f, 1000.0f);
    if (depth) std::memcpy(dm.depth.data.data(), depth, (size_t)w * h * 4);
    shs::RTRegistry rtr{};
    shs::Context ctx{};
    ctx.job_system = g_jobs.get();
    shs::FrameParams f{};
    f.pass.light_shafts.enable = p->enable != 0;
    f.pass.light_shafts.steps = p->steps;
    f.pass.light_shafts.density = p->density;
    f.pass.light_shafts.weight = p->weight;
    f.pass.light_shafts.decay = p->decay;
    shs::Scene scene{};
    scene.cam.pos = load_vec3(p->cam_pos);
    scene.cam.viewproj = load_mat4(p->cam_viewproj);
    scene.sun.dir_ws = load_vec3(p->sun_dir_ws);
    shs::PassLightShafts::Inputs in{};
    in.scene = &scene;
    in.fp = &f;
    in.rtr = &rtr;
    in.rt_input_ldr = rtr.reg<shs::RTHandle>(&src);
    in.rt_output_ldr = rtr.reg<shs::RTHandle>(&dst);
    if (depth) in.rt_depth_like = rtr.reg<shs::RTHandle>(&dm);
    shs::PassLightShafts{}.execute(ctx, in);
    std::memcpy(out_ldr, dst.color.data.data(), (size_t)w * h * 4);
    return SHSB_OK;
}

} // extern "C"
