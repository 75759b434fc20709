// shs_b200/drop_in.hpp -- the reference-side binding of the B200 raster path (C++20, header-only).
//
// A maintainer of leisure-software-renderer adds this header to shs-renderer-lib, links libshsb.so and swaps
//     shs::rasterize_mesh(...)              ->  shs::b200::rasterize_mesh(dev, ...)
//     shs::PassPBRForward::execute(...)     ->  shs::b200::PassPBRForward(dev).execute(...)
//     shs::PassShadowMap::execute(...)      ->  shs::b200::PassShadowMap(dev).execute(...)
//     shs::PassTonemap::execute(...)        ->  shs::b200::PassTonemap(dev).execute(...)
//     shs::PassMotionBlur::execute(...)     ->  shs::b200::PassMotionBlur(dev).execute(...)
//     shs::PassLightShafts::execute(...)    ->  shs::b200::PassLightShafts(dev).execute(...)
//     PassTemporalAAAdapter (taa)           ->  shs::b200::PassTemporalAA(dev).execute(ctx, rtr, rt_ldr)
// Same argument types (the reference's own MeshData / ShaderUniforms / RasterizerTarget / Scene / FrameParams /
// RTRegistry), same results, same error convention (invalid input => silent return with zero stats,
// sw_render/rasterizer.hpp:190-194, passes/pass_pbr_forward.hpp:51-58).
//
// It needs the reference's headers (and therefore GLM) on the include path, so it is compiled only inside the
// reference tree -- in this repository by tests/cpp/Makefile against oracle/glm_shim.
//
// Host <-> device: the reference's render targets are std::vector-backed (gfx/rt_types.hpp:35-59).  Device keeps a
// device twin per host RT object; passes run on the twins and `Device::download(rt)` (or the `sync_host` flag of the
// pass wrappers, default true) copies results back, so existing code that reads hdr->color.at(x, y) keeps working.
#pragma once

#include <cstring>
#include <unordered_map>
#include <vector>

#include "shs/core/context.hpp"
#include "shs/passes/pass_light_shafts.hpp"
#include "shs/passes/pass_motion_blur.hpp"
#include "shs/passes/pass_pbr_forward.hpp"
#include "shs/passes/pass_shadow_map.hpp"
#include "shs/passes/pass_tonemap.hpp"
#include "shs/sky/cubemap_sky.hpp"
#include "shs/sky/procedural_sky.hpp"
#include "shs/sw_render/rasterizer.hpp"

#include "shsb.h"

namespace shs::b200
{
    // The builtin ShaderProgram factories are the only programs that exist on the device (std::function bodies cannot
    // be shipped to a GPU): shader/builtin_shaders.hpp:105,154,221 and pipeline/pass_adapters.hpp:335.
    enum class BuiltinProgram : int32_t
    {
        PbrMetalRough = SHSB_SHADER_PBR_MR,
        BlinnPhong = SHSB_SHADER_BLINN_PHONG,
        DebugAlbedo = SHSB_SHADER_DEBUG_ALBEDO,
        DebugNormal = SHSB_SHADER_DEBUG_NORMAL,
        DebugDepth = SHSB_SHADER_DEBUG_DEPTH,
        DepthOnly = SHSB_SHADER_DEPTH_ONLY
    };

    class Device
    {
    public:
        explicit Device(int cuda_device = 0) { ok_ = shsb_context_create(cuda_device, &ctx_) == SHSB_OK; }
        ~Device() { if (ctx_) shsb_context_destroy(ctx_); }
        Device(const Device&) = delete;
        Device& operator=(const Device&) = delete;

        bool valid() const { return ok_; }
        shsb_ctx ctx() const { return ctx_; }
        const char* last_error() const { return ctx_ ? shsb_last_error_string(ctx_) : "no CUDA device (there is no CPU fallback)"; }

        // ---- assets (uploaded once, keyed by the host object's address like ShadowRuntimeState::mesh_bounds_cache)
        shsb_mesh mesh(const MeshData& m)
        {
            auto it = meshes_.find(&m);
            if (it != meshes_.end()) return it->second;
            shsb_mesh h = 0;
            static_assert(sizeof(glm::vec3) == 12 && sizeof(glm::vec2) == 8, "MeshData streams are tightly packed");
            shsb_mesh_upload(ctx_, m.positions.empty() ? nullptr : &m.positions[0].x, (uint32_t)m.positions.size(),
                             m.normals.empty() ? nullptr : &m.normals[0].x, (uint32_t)m.normals.size(),
                             m.uvs.empty() ? nullptr : &m.uvs[0].x, (uint32_t)m.uvs.size(),
                             m.indices.empty() ? nullptr : m.indices.data(), (uint32_t)m.indices.size(), &h);
            meshes_[&m] = h;
            return h;
        }

        // A MeshData / Texture2DData whose CONTENT changed in place (procedural or streamed assets) must be announced: the device
        // copy is keyed by the host object's address.
        void invalidate(const MeshData& m)
        {
            const auto it = meshes_.find(&m);
            if (it == meshes_.end()) return;
            shsb_mesh_destroy(ctx_, it->second);
            meshes_.erase(it);
        }
        void invalidate(const Texture2DData* t)
        {
            const auto it = textures_.find(t);
            if (it == textures_.end()) return;
            shsb_texture_destroy(ctx_, it->second);
            textures_.erase(it);
        }

        shsb_tex texture(const Texture2DData* t)
        {
            if (!t || !t->valid()) return 0;
            auto it = textures_.find(t);
            if (it != textures_.end()) return it->second;
            shsb_tex h = 0;
            shsb_texture_upload(ctx_, &t->texels[0].r, t->w, t->h, &h);
            textures_[t] = h;
            return h;
        }

        // ---- render-target twins
        shsb_rt twin(RT_ColorHDR* rt) { return rt ? twin_of(rt, SHSB_RT_COLOR_HDR, rt->w, rt->h, 0.1f, 1000.0f) : 0; }
        shsb_rt twin(RT_ColorLDR* rt) { return rt ? twin_of(rt, SHSB_RT_COLOR_LDR, rt->w, rt->h, 0.1f, 1000.0f) : 0; }
        shsb_rt twin(RT_ColorDepthMotion* rt) { return rt ? twin_of(rt, SHSB_RT_DEPTH_MOTION, rt->w, rt->h, rt->zn, rt->zf) : 0; }
        shsb_rt twin(RT_ShadowDepth* rt) { return rt ? twin_of(rt, SHSB_RT_SHADOW, rt->w, rt->h, 0.1f, 1000.0f) : 0; }

        void upload(RT_ColorHDR* rt) { if (rt) shsb_rt_upload(ctx_, twin(rt), SHSB_PLANE_COLOR, rt->color.data.data(), rt->color.data.size() * sizeof(ColorF)); }
        void upload(RT_ColorLDR* rt) { if (rt) shsb_rt_upload(ctx_, twin(rt), SHSB_PLANE_COLOR, rt->color.data.data(), rt->color.data.size() * sizeof(Color)); }
        void upload_motion(RT_ColorDepthMotion* rt) { if (rt) shsb_rt_upload(ctx_, twin(rt), SHSB_PLANE_MOTION, rt->motion.data.data(), rt->motion.data.size() * sizeof(Motion2f)); }
        void upload(RT_ColorDepthMotion* rt) { if (rt) shsb_rt_upload(ctx_, twin(rt), SHSB_PLANE_DEPTH, rt->depth.data.data(), rt->depth.data.size() * sizeof(float)); }
        void upload(RT_ShadowDepth* rt) { if (rt) shsb_rt_upload(ctx_, twin(rt), SHSB_PLANE_DEPTH, rt->depth.data(), rt->depth.size() * sizeof(float)); }
        void download(RT_ColorHDR* rt) { if (rt) shsb_rt_download(ctx_, twin(rt), SHSB_PLANE_COLOR, rt->color.data.data(), rt->color.data.size() * sizeof(ColorF)); }
        void download(RT_ColorLDR* rt) { if (rt) shsb_rt_download(ctx_, twin(rt), SHSB_PLANE_COLOR, rt->color.data.data(), rt->color.data.size() * sizeof(Color)); }
        void download(RT_ColorDepthMotion* rt)
        {
            if (!rt) return;
            shsb_rt_download(ctx_, twin(rt), SHSB_PLANE_DEPTH, rt->depth.data.data(), rt->depth.data.size() * sizeof(float));
            shsb_rt_download(ctx_, twin(rt), SHSB_PLANE_MOTION, rt->motion.data.data(), rt->motion.data.size() * sizeof(Motion2f));
        }

        // ---- sky models.  ISkyModel implementations keep their parameters private, so a Scene::sky pointer has to be
        // introduced once: the reference's two models exist on the device, any other ISkyModel is refused (no fallback).
        struct SkyDesc { int32_t kind = SHSB_SKY_NONE; float intensity = 1.0f; glm::vec3 sun_dir{0.0f}; shsb_tex faces[6]{}; };
        void register_sky(const ProceduralSky* sky, const glm::vec3& sun_dir_ws)
        {
            SkyDesc d; d.kind = SHSB_SKY_PROCEDURAL; d.sun_dir = sun_dir_ws;
            skies_[sky] = d;
        }
        void register_sky(const CubemapSky* sky, const CubemapData& cubemap, float intensity = 1.0f)
        {
            SkyDesc d; d.kind = SHSB_SKY_CUBEMAP; d.intensity = intensity;
            for (int i = 0; i < 6; ++i) d.faces[i] = texture(&cubemap.face[(size_t)i]);
            skies_[sky] = d;
        }
        const SkyDesc* sky(const ISkyModel* s) const { const auto it = skies_.find(s); return it == skies_.end() ? nullptr : &it->second; }
        void download(RT_ShadowDepth* rt) { if (rt) shsb_rt_download(ctx_, twin(rt), SHSB_PLANE_DEPTH, rt->depth.data(), rt->depth.size() * sizeof(float)); }

    private:
        // A host RT object that was resized (RTRegistry::ensure_*, a window resize: same address, new extent) or whose depth
        // range changed gets a fresh twin; the old contents are gone on the host side as well (PixelBuffer2D reallocates).
        struct Twin { shsb_rt rt = 0; int w = 0, h = 0; float zn = 0.0f, zf = 0.0f; };
        shsb_rt twin_of(const void* key, int32_t kind, int w, int h, float zn, float zf)
        {
            auto it = rts_.find(key);
            if (it != rts_.end())
            {
                const Twin& t = it->second;
                if (t.w == w && t.h == h && t.zn == zn && t.zf == zf) return t.rt;
                if (t.rt) shsb_rt_destroy(ctx_, t.rt);
            }
            Twin t{};
            t.w = w; t.h = h; t.zn = zn; t.zf = zf;
            if (w > 0 && h > 0) shsb_rt_create(ctx_, kind, w, h, zn, zf, &t.rt);
            rts_[key] = t;
            return t.rt;
        }

        shsb_ctx ctx_ = nullptr;
        bool ok_ = false;
        std::unordered_map<const void*, shsb_mesh> meshes_{};
        std::unordered_map<const void*, shsb_tex> textures_{};
        std::unordered_map<const void*, Twin> rts_{};
        std::unordered_map<const ISkyModel*, SkyDesc> skies_{};
    };

    namespace detail
    {
        inline void copy_mat(const glm::mat4& m, float out[16]) { std::memcpy(out, &m, 64); }
        inline void copy_vec(const glm::vec3& v, float out[3]) { out[0] = v.x; out[1] = v.y; out[2] = v.z; }

        inline ShsbFrameParams frame_params(const FrameParams& fp)
        {
            ShsbFrameParams o{};
            o.shading_model = (fp.shading_model == ShadingModel::BlinnPhong) ? SHSB_SHADING_BLINN_PHONG : SHSB_SHADING_PBR_METAL_ROUGH;
            o.debug_view = (int32_t)fp.debug_view;
            o.cull_mode = (int32_t)fp.cull_mode;
            o.front_face_ccw = fp.front_face_ccw ? 1 : 0;
            o.shadow_enable = fp.pass.shadow.enable ? 1 : 0;
            o.shadow_bias_const = fp.pass.shadow.bias_const;
            o.shadow_bias_slope = fp.pass.shadow.bias_slope;
            o.shadow_pcf_radius = fp.pass.shadow.pcf_radius;
            o.shadow_pcf_step = fp.pass.shadow.pcf_step;
            o.shadow_strength = fp.pass.shadow.strength;
            o.exposure = fp.pass.tonemap.exposure;
            o.gamma = fp.pass.tonemap.gamma;
            o.light_culling = fp.technique.light_culling ? 1 : 0;
            o.tile_size = fp.technique.tile_size;
            o.max_lights_per_tile = fp.technique.max_lights_per_tile;
            o.motion_vectors_enable = fp.pass.motion_vectors.enable ? 1 : 0;
            return o;
        }

        // Scene + ResourceRegistry -> ShsbScene (items keep Scene::items order; materials resolved like
        // pass_pbr_forward.hpp:166-184).
        inline ShsbScene scene(Device& dev, const Scene& s, std::vector<ShsbRenderItem>& items)
        {
            ShsbScene o{};
            copy_mat(s.cam.viewproj, o.cam_viewproj);
            copy_mat(s.cam.prev_viewproj, o.cam_prev_viewproj);
            copy_vec(s.cam.pos, o.cam_pos);
            if (s.sky)
            {
                if (const Device::SkyDesc* d = dev.sky(s.sky))
                {
                    o.sky_kind = d->kind;
                    o.sky_intensity = d->intensity;
                    copy_vec(d->sun_dir, o.sky_sun_dir_ws);
                    for (int i = 0; i < 6; ++i) o.sky_faces[i] = d->faces[i];
                }
                else o.sky_kind = -1; // unknown ISkyModel: the pass refuses it
            }
            copy_vec(s.sun.dir_ws, o.sun_dir_ws);
            copy_vec(s.sun.color, o.sun_color);
            o.sun_intensity = s.sun.intensity;
            items.clear();
            items.reserve(s.items.size());
            for (const RenderItem& it : s.items)
            {
                ShsbRenderItem r{};
                copy_vec(it.tr.pos, r.tr.pos);
                copy_vec(it.tr.rot_euler, r.tr.rot_euler);
                copy_vec(it.tr.scl, r.tr.scl);
                const MeshData* mesh = s.resources ? s.resources->get_mesh((MeshAssetHandle)it.mesh) : nullptr;
                r.mesh = (mesh && !mesh->empty()) ? dev.mesh(*mesh) : 0;
                const MaterialData* mat = s.resources ? s.resources->get_material((MaterialAssetHandle)it.mat) : nullptr;
                r.has_material = mat ? (uint32_t)it.mat : 0u; // the handle itself: it enters the derived motion key (pass_pbr_forward.hpp:146)
                if (mat)
                {
                    copy_vec(mat->base_color, r.base_color);
                    r.metallic = mat->metallic;
                    r.roughness = mat->roughness;
                    r.ao = mat->ao;
                    r.base_color_tex = mat->base_color_tex ? dev.texture(s.resources->get_texture(mat->base_color_tex)) : 0;
                }
                r.casts_shadow = it.casts_shadow ? 1u : 0u;
                r.visible = it.visible ? 1u : 0u;
                r.object_id = it.object_id;
                items.push_back(r);
            }
            o.items = items.data();
            o.n_items = (uint32_t)items.size();
            return o;
        }
    }

    // rasterize_mesh, sw_render/rasterizer.hpp:181-187.  `program` replaces the std::function pair; the job system
    // fields of RasterizerConfig are accepted and ignored (the device is the job system).
    inline RasterizerStats rasterize_mesh(Device& dev, const MeshData& mesh, BuiltinProgram program, const ShaderUniforms& u,
                                          RasterizerTarget target, const RasterizerConfig& config = {}, bool sync_host = true)
    {
        RasterizerStats stats{};
        if (!dev.valid() || !target.hdr) return stats;                 // rasterizer.hpp:190
        if (mesh.positions.empty()) return stats;                       // :191
        if (target.hdr->w <= 0 || target.hdr->h <= 0) return stats;     // :194
        ShsbUniforms su{};
        detail::copy_mat(u.model, su.model);
        detail::copy_mat(u.viewproj, su.viewproj);
        detail::copy_mat(u.light_viewproj, su.light_viewproj);
        detail::copy_vec(u.light_dir_ws, su.light_dir_ws);
        detail::copy_vec(u.light_color, su.light_color);
        su.light_intensity = u.light_intensity;
        detail::copy_vec(u.camera_pos, su.camera_pos);
        detail::copy_vec(u.base_color, su.base_color);
        su.metallic = u.metallic; su.roughness = u.roughness; su.ao = u.ao;
        su.base_color_tex = dev.texture(u.base_color_tex);
        if (u.shadow_map)
        {
            RT_ShadowDepth* sm = const_cast<RT_ShadowDepth*>(u.shadow_map);
            su.shadow_map = dev.twin(sm);
            dev.upload(sm);
        }
        su.shadow_bias_const = u.shadow_bias_const; su.shadow_bias_slope = u.shadow_bias_slope;
        su.shadow_pcf_radius = u.shadow_pcf_radius; su.shadow_pcf_step = u.shadow_pcf_step; su.shadow_strength = u.shadow_strength;
        su.enable_motion_vectors = u.enable_motion_vectors ? 1 : 0;    // rasterizer.hpp:295
        detail::copy_mat(u.prev_model, su.prev_model);
        detail::copy_mat(u.prev_viewproj, su.prev_viewproj);
        ShsbRasterCfg cfg{};
        cfg.cull_mode = (int32_t)config.cull_mode;
        cfg.front_face_ccw = config.front_face_ccw ? 1 : 0;
        // rasterize_mesh composites into whatever the targets hold: mirror the host contents first
        dev.upload(target.hdr);
        // depth AND motion: the draw touches only covered pixels of both planes (rasterizer.hpp:359-411), and the read-back below
        // copies both planes whole, so a twin with a stale motion plane would overwrite the host's vectors on uncovered pixels
        if (target.depth_motion) { dev.upload(target.depth_motion); dev.upload_motion(target.depth_motion); }
        ShsbStats st{};
        if (shsb_rasterize_mesh(dev.ctx(), dev.mesh(mesh), (int32_t)program, &su, dev.twin(target.hdr),
                                target.depth_motion ? dev.twin(target.depth_motion) : 0, &cfg, &st) != SHSB_OK) return stats;
        if (sync_host)
        {
            dev.download(target.hdr);
            if (target.depth_motion) dev.download(target.depth_motion);
        }
        stats.tri_input = st.tri_input; stats.tri_after_clip = st.tri_after_clip; stats.tri_raster = st.tri_raster;
        return stats;
    }

    // PassPBRForward, passes/pass_pbr_forward.hpp:31-214 (same Inputs struct, same Context side effects on ctx.debug).
    // `inputs_on_device`: the shadow map and (with preserve_existing_depth) the depth plane were produced by the b200 passes
    // and already live in the twins; false (default) mirrors the host copies first.  `local_lights`: -1 = as FrameParams says
    // (technique.light_culling), 0 = sun only like the reference's CPU lit pass (quirk Q2, pass_pbr_forward.hpp:157-195 never
    // reads tile lists), 1 = walk the tile lists of the last shsb_light_cull (fp_stress_scene.frag:644-678).
    class PassPBRForward
    {
    public:
        explicit PassPBRForward(Device& dev, bool sync_host = true, bool inputs_on_device = false, int local_lights = -1)
            : dev_(dev), sync_host_(sync_host), on_device_(inputs_on_device), local_lights_(local_lights) {}
        using Inputs = shs::PassPBRForward::Inputs;

        void execute(Context& ctx, const Inputs& in)
        {
            if (!in.scene || !in.fp || !in.rtr) return;
            if (!in.rt_hdr.valid() || !dev_.valid()) return;
            ctx.debug.tri_input = 0; ctx.debug.tri_after_clip = 0; ctx.debug.tri_raster = 0;
            auto* hdr = static_cast<RT_ColorHDR*>(in.rtr->get(in.rt_hdr));
            if (!hdr || hdr->w <= 0 || hdr->h <= 0) return;
            auto* motion = in.rt_motion.valid() ? static_cast<RT_ColorDepthMotion*>(in.rtr->get(in.rt_motion)) : nullptr;
            auto* shadow = in.rt_shadow.valid() ? static_cast<RT_ShadowDepth*>(in.rtr->get(in.rt_shadow)) : nullptr;
            std::vector<ShsbRenderItem> items;
            const ShsbScene s = detail::scene(dev_, *in.scene, items);
            if (s.sky_kind < 0) return; // an ISkyModel that was not introduced with Device::register_sky: no fallback
            ShsbFrameParams fp = detail::frame_params(*in.fp);
            if (local_lights_ >= 0) fp.light_culling = local_lights_;
            if (!ctx.history.has_prev_frame) shsb_history_reset(dev_.ctx()); // Context::history owns the notion of "first frame"
            if (motion && in.preserve_existing_depth && !on_device_) dev_.upload(motion);
            const bool use_shadow = in.fp->pass.shadow.enable && shadow && ctx.shadow.valid;
            float lvp[16];
            if (use_shadow) { detail::copy_mat(ctx.shadow.light_viewproj, lvp); if (!on_device_) dev_.upload(shadow); }
            ShsbStats st{};
            if (shsb_pass_pbr_forward(dev_.ctx(), &s, &fp, dev_.twin(hdr), motion ? dev_.twin(motion) : 0, use_shadow ? dev_.twin(shadow) : 0,
                                      use_shadow ? lvp : nullptr, in.preserve_existing_depth ? 1 : 0, &st) != SHSB_OK) return;
            if (sync_host_) { dev_.download(hdr); if (motion) dev_.download(motion); }
            ctx.debug.tri_input = st.tri_input; ctx.debug.tri_after_clip = st.tri_after_clip; ctx.debug.tri_raster = st.tri_raster;
            ctx.history.has_prev_frame = true;
        }

    private:
        Device& dev_;
        bool sync_host_, on_device_;
        int local_lights_;
    };

    // PassShadowMap, passes/pass_shadow_map.hpp:27-205.
    class PassShadowMap
    {
    public:
        explicit PassShadowMap(Device& dev, bool sync_host = true) : dev_(dev), sync_host_(sync_host) {}
        using Inputs = shs::PassShadowMap::Inputs;

        void execute(Context& ctx, const Inputs& in)
        {
            ctx.shadow.reset();
            if (!in.scene || !in.fp || !in.rtr) return;
            if (!in.rt_shadow.valid() || !dev_.valid()) return;
            if (!in.fp->pass.shadow.enable) return;
            auto* shadow = static_cast<RT_ShadowDepth*>(in.rtr->get(in.rt_shadow));
            if (!shadow || shadow->w <= 0 || shadow->h <= 0) return;
            std::vector<ShsbRenderItem> items;
            const ShsbScene s = detail::scene(dev_, *in.scene, items);
            const ShsbFrameParams fp = detail::frame_params(*in.fp);
            float lvp[16];
            if (shsb_pass_shadow_map(dev_.ctx(), &s, &fp, dev_.twin(shadow), lvp) != SHSB_OK) return;
            if (sync_host_) dev_.download(shadow);
            ctx.shadow.map = shadow;
            std::memcpy(&ctx.shadow.light_viewproj, lvp, 64);
            ctx.shadow.valid = true;
        }

    private:
        Device& dev_;
        bool sync_host_;
    };

    // PassTonemap, passes/pass_tonemap.hpp:21-84.
    class PassTonemap
    {
    public:
        explicit PassTonemap(Device& dev, bool sync_host = true, bool hdr_on_device = true) : dev_(dev), sync_host_(sync_host), hdr_on_device_(hdr_on_device) {}
        using Inputs = shs::PassTonemap::Inputs;

        void execute(Context& ctx, const Inputs& in)
        {
            (void)ctx;
            if (!in.fp || !in.rtr || !dev_.valid()) return;
            if (!in.rt_hdr.valid() || !in.rt_ldr.valid()) return;
            auto* hdr = static_cast<RT_ColorHDR*>(in.rtr->get(in.rt_hdr));
            auto* ldr = static_cast<RT_ColorLDR*>(in.rtr->get(in.rt_ldr));
            if (!hdr || !ldr || hdr->w <= 0 || hdr->h <= 0 || ldr->w <= 0 || ldr->h <= 0) return;
            if (!hdr_on_device_) dev_.upload(hdr);
            if (shsb_pass_tonemap(dev_.ctx(), dev_.twin(hdr), dev_.twin(ldr), in.fp->pass.tonemap.exposure, in.fp->pass.tonemap.gamma) != SHSB_OK) return;
            if (sync_host_) dev_.download(ldr);
        }

    private:
        Device& dev_;
        bool sync_host_, hdr_on_device_;
    };
    // PassMotionBlur, passes/pass_motion_blur.hpp:25-200.  `inputs_on_device` = the LDR frame and the depth / motion planes
    // were produced by the b200 passes and already live in the twins (the normal case); false uploads the host copies first.
    class PassMotionBlur
    {
    public:
        explicit PassMotionBlur(Device& dev, bool sync_host = true, bool inputs_on_device = true) : dev_(dev), sync_host_(sync_host), on_device_(inputs_on_device) {}
        using Inputs = shs::PassMotionBlur::Inputs;

        void execute(Context& ctx, const Inputs& in)
        {
            (void)ctx;
            if (!in.fp || !in.rtr || !dev_.valid()) return;
            if (!in.rt_input_ldr.valid() || !in.rt_output_ldr.valid()) return;
            auto* src = static_cast<RT_ColorLDR*>(in.rtr->get(in.rt_input_ldr));
            auto* dst = static_cast<RT_ColorLDR*>(in.rtr->get(in.rt_output_ldr));
            auto* motion = in.rt_motion.valid() ? static_cast<RT_ColorDepthMotion*>(in.rtr->get(in.rt_motion)) : nullptr;
            if (!src || !dst || !motion) return; // rt_tmp is only scratch in the reference; the device owns its own
            if (!on_device_) { dev_.upload(src); dev_.upload(motion); dev_.upload_motion(motion); }
            const MotionBlurPassParams& mb = in.fp->pass.motion_blur;
            ShsbMotionBlurParams p{};
            p.enable = mb.enable ? 1 : 0;
            p.samples = mb.samples;
            p.strength = mb.strength;
            p.max_velocity_px = mb.max_velocity_px;
            p.min_velocity_px = mb.min_velocity_px;
            p.depth_reject = mb.depth_reject;
            p.dt = in.fp->dt;
            if (shsb_pass_motion_blur(dev_.ctx(), &p, dev_.twin(src), dev_.twin(dst), dev_.twin(motion)) != SHSB_OK) return;
            if (sync_host_) dev_.download(dst);
        }

    private:
        Device& dev_;
        bool sync_host_, on_device_;
    };

    // PassLightShafts, passes/pass_light_shafts.hpp:27-216.
    class PassLightShafts
    {
    public:
        explicit PassLightShafts(Device& dev, bool sync_host = true, bool inputs_on_device = true) : dev_(dev), sync_host_(sync_host), on_device_(inputs_on_device) {}
        using Inputs = shs::PassLightShafts::Inputs;

        void execute(Context& ctx, const Inputs& in)
        {
            (void)ctx;
            if (!in.scene || !in.fp || !in.rtr || !dev_.valid()) return;
            if (!in.rt_input_ldr.valid() || !in.rt_output_ldr.valid()) return;
            auto* src = static_cast<RT_ColorLDR*>(in.rtr->get(in.rt_input_ldr));
            auto* dst = static_cast<RT_ColorLDR*>(in.rtr->get(in.rt_output_ldr));
            if (!src || !dst || src->w <= 0 || src->h <= 0 || dst->w <= 0 || dst->h <= 0) return;
            auto* depth_like = in.rt_depth_like.valid() ? static_cast<RT_ColorDepthMotion*>(in.rtr->get(in.rt_depth_like)) : nullptr;
            if (!on_device_) { dev_.upload(src); if (depth_like) dev_.upload(depth_like); }
            const LightShaftsPassParams& ls = in.fp->pass.light_shafts;
            ShsbLightShaftsParams p{};
            p.enable = ls.enable ? 1 : 0;
            p.steps = ls.steps;
            p.density = ls.density;
            p.weight = ls.weight;
            p.decay = ls.decay;
            std::memcpy(p.cam_pos, &in.scene->cam.pos, 12);
            std::memcpy(p.sun_dir_ws, &in.scene->sun.dir_ws, 12);
            std::memcpy(p.cam_viewproj, &in.scene->cam.viewproj, 64);
            if (shsb_pass_light_shafts(dev_.ctx(), &p, dev_.twin(src), dev_.twin(dst), depth_like ? dev_.twin(depth_like) : 0) != SHSB_OK) return;
            if (sync_host_) dev_.download(dst);
        }

    private:
        Device& dev_;
        bool sync_host_, on_device_;
    };

    // PassTemporalAAAdapter::execute_resolved / reset_history, pipeline/pass_adapters.hpp:1402-1496.  The colour history
    // (Context::temporal_aa) lives on the device; ctx.temporal_aa only mirrors its size / validity flags.
    class PassTemporalAA
    {
    public:
        explicit PassTemporalAA(Device& dev, bool sync_host = true, bool inputs_on_device = true) : dev_(dev), sync_host_(sync_host), on_device_(inputs_on_device) {}

        void execute(Context& ctx, RTRegistry& rtr, RTHandle rt_ldr)
        {
            auto* ldr = static_cast<RT_ColorLDR*>(rtr.get(rt_ldr));
            if (!ldr || ldr->w <= 0 || ldr->h <= 0 || !dev_.valid()) return;
            if (!on_device_) dev_.upload(ldr);
            if (!ctx.temporal_aa.history_valid) shsb_taa_reset(dev_.ctx()); // someone called ctx.temporal_aa.reset()
            if (shsb_pass_taa(dev_.ctx(), dev_.twin(ldr)) != SHSB_OK) return;
            ctx.temporal_aa.history_w = ldr->w;
            ctx.temporal_aa.history_h = ldr->h;
            ctx.temporal_aa.history_valid = true;
            if (sync_host_) dev_.download(ldr);
        }

        void reset_history(Context& ctx)
        {
            ctx.temporal_aa.reset();
            shsb_taa_reset(dev_.ctx());
        }

    private:
        Device& dev_;
        bool sync_host_, on_device_;
    };
}
