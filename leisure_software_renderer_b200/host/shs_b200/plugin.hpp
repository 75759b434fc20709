// shs_b200/plugin.hpp -- the B200 raster path behind the reference's pass-plugin interface (C++20, header-only).
//
// The reference assembles a frame from IRenderPass objects (pipeline/render_pass.hpp:265-306) that a PassFactoryRegistry
// creates by id (pipeline/pass_registry.hpp:36-87) and that PluggablePipeline / PipelineRuntimeExecutor drive
// (pipeline/pluggable_pipeline.hpp:59-179, 743-1033): build_execution_request() on the planning side, execute_resolved() on
// the runtime side, runtime capabilities (depth_prepass_ready / light_culling_ready) threaded from pass to pass through
// PassExecutionResult.  A maintainer swaps
//     make_standard_pass_factory_registry(rt_shadow, rt_hdr, rt_motion, rt_ldr, rt_shafts_tmp, rt_motion_blur_tmp)   (pass_adapters.hpp:1497-1504)
// for
//     shs::b200::make_b200_pass_factory_registry(dev, rt_shadow, rt_hdr, rt_motion, rt_ldr, rt_shafts_tmp, rt_motion_blur_tmp)
// and keeps everything else: the same PassId / string ids, the same contracts and IO descriptions (so the frame graph, the
// planner's contract validation and the technique profiles of pipeline/technique_profile.hpp:42-110 see the same passes), the
// same Context side effects, the same error convention (invalid input => PassExecutionResult::not_executed(), never an
// exception).  Passes declare the Software backend and Software resource domain exactly like the adapters they replace: what
// they read and write are the reference's host-visible RT objects; the device twins behind them (drop_in.hpp) are an
// implementation detail the planner does not need to know about.
//
// Ids registered: shadow_map, depth_prepass, light_culling, cluster_build, cluster_light_assign, pbr_forward,
// pbr_forward_plus, pbr_forward_clustered, tonemap, light_shafts, motion_blur, depth_of_field, taa.
// Not registered (outside the raster hot path): gbuffer, ssao, deferred_lighting, deferred_lighting_tiled
// -- in the reference these adapters do no deferred shading either (pass_adapters.hpp:721-934 run the forward pass or nothing);
// a pipeline that asks for them gets them reported through configure_from_profile()'s out_missing_ids like any absent optional pass.
//
// There is no CPU fallback: a pass whose Device has no GPU returns not_executed().
#pragma once

#include <algorithm>
#include <memory>
#include <vector>

#include "shs/geometry/frustum_culling.hpp"
#include "shs/lighting/light_set.hpp"
#include "shs/pipeline/pass_contract_registry.hpp"
#include "shs/pipeline/pass_registry.hpp"
#include "shs/pipeline/render_pass.hpp"

#include "shs_b200/drop_in.hpp"

namespace shs::b200
{
    struct PluginOptions
    {
        // Every pass mirrors what it wrote into the reference's host RT objects (drop-in behaviour: code that reads
        // hdr->color.at(x, y) after pipeline.execute() keeps working).  false: results stay in the device twins until
        // Device::download(rt); the passes then never touch PCIe except for the per-frame draw list.
        bool sync_host = true;
        // The inputs a pass consumes were written by another b200 pass of the same pipeline and are already in the twins.
        // Set false when b200 passes are mixed with the reference's CPU passes in one pipeline: inputs are uploaded first.
        bool inputs_on_device = true;
        // The reference's CPU lit pass shades the sun only even in Forward+ (quirk Q2: pass_pbr_forward.hpp:157-195 never reads
        // tile lists; execute_generic_light_culling, pass_adapters.hpp:228-333, only counts).  false keeps that -- results
        // equal the reference's.  true makes pbr_forward_plus / pbr_forward_clustered walk the tile lists built by
        // light_culling / cluster_light_assign over Scene::local_lights the way the reference's GPU path does
        // (shaders/vulkan/fp_stress_scene.frag:644-678).  The lists are used only when the executor reports them ready
        // (PassExecutionRequest::light_culling_ready); with the reference's pipeline semantics that means an in-order frame,
        // and -- because an in-order frame WITH the depth pre-pass draws nothing (quirk Q1) -- technique.depth_prepass = false
        // with PluggablePipeline::set_strict_graph_validation(false) (tests/cpp/plugin_test.cpp).
        bool shade_local_lights = false;
        // Fill LightCullingRuntimePayload::tile_light_counts / visible_light_count (a device -> host read of 4 B per tile).
        bool fill_light_culling_payload = true;
    };

    namespace detail
    {
        // pass_adapters.hpp:219-226
        inline bool technique_uses_light_culling(const FrameParams& fp)
        {
            return fp.technique.light_culling || fp.technique.mode == TechniqueMode::ForwardPlus ||
                   fp.technique.mode == TechniqueMode::TiledDeferred || fp.technique.mode == TechniqueMode::ClusteredForward;
        }

        inline bool request_complete(const PassExecutionRequest& r, bool need_scene = true)
        {
            return r.valid && (!need_scene || r.inputs.scene) && r.inputs.frame && r.inputs.registry;
        }

        // Number of local lights the camera frustum does not reject (the broad phase of execute_generic_light_culling,
        // pass_adapters.hpp:270-290) on the CullingLightGPU proxy bounds: sphere first, AABB p-vertex refine when the sphere
        // straddles a plane; tolerances of geometry/jolt_culling.hpp:118-122 (outer epsilon 1e-5).
        inline uint32_t frustum_visible_lights(const std::vector<CullingLightGPU>& lights, const glm::mat4& view_proj)
        {
            const Frustum f = extract_frustum_planes(view_proj);
            const float eps = 1e-5f;
            uint32_t n = 0;
            for (const CullingLightGPU& l : lights)
            {
                const glm::vec3 c(l.cull_sphere.x, l.cull_sphere.y, l.cull_sphere.z);
                const float r = std::max(l.cull_sphere.w, 0.0f);
                bool outside = false, inside = true;
                for (const Plane& p : f.planes)
                {
                    const float d = p.signed_distance(c);
                    if (d < -(r + eps)) { outside = true; break; }
                    if (d < r + eps) inside = false;
                }
                if (outside) continue;
                if (!inside)
                {
                    const glm::vec3 mn(l.cull_aabb_min.x, l.cull_aabb_min.y, l.cull_aabb_min.z);
                    const glm::vec3 mx(l.cull_aabb_max.x, l.cull_aabb_max.y, l.cull_aabb_max.z);
                    for (const Plane& p : f.planes)
                    {
                        const glm::vec3 v((p.normal.x >= 0.0f) ? mx.x : mn.x, (p.normal.y >= 0.0f) ? mx.y : mn.y, (p.normal.z >= 0.0f) ? mx.z : mn.z);
                        if (p.signed_distance(v) < -eps) { outside = true; break; }
                    }
                    if (outside) continue;
                }
                ++n;
            }
            return n;
        }

        // execute_generic_light_culling (pass_adapters.hpp:228-333) on the device: Scene::local_lights -> the reference's own
        // CullingLightGPU packers (LightSet::flatten_cullable_gpu, lighting/light_set.hpp:52-70) -> shsb_lights_upload ->
        // shsb_light_cull.  The lists stay on the device for the lit pass; the payload receives what the CPU pass reports:
        // per-tile min(max_per_tile, directional + local) and the frustum-visible light count.
        // Bounds: the CullingLightGPU proxies (sphere(position, range) + its AABB), not the Jolt hulls of
        // append_local_light_shapes_from_set (pass_adapters.hpp:167-217) -- JoltPhysics is a third-party dependency of the
        // reference that is absent here; for point lights the proxy AABB equals the Jolt AABB.
        inline bool run_light_culling(Device& dev, const PluginOptions& opt, const Scene& scene, const FrameParams& fp, RTRegistry& rtr,
                                      RT_Motion rt_motion, LightCullingRuntimePayload* payload, bool depth_prepass_ready, bool force_enable)
        {
            if (!dev.valid()) return false;
            if (!(force_enable || technique_uses_light_culling(fp))) return false;
            if (!payload) return false;
            if (fp.technique.depth_prepass && !depth_prepass_ready) return false;
            int w = fp.w, h = fp.h;
            if (rt_motion.valid())
            {
                auto* motion = static_cast<RT_ColorDepthMotion*>(rtr.get(rt_motion));
                if (motion && motion->w > 0 && motion->h > 0) { w = motion->w; h = motion->h; }
            }
            if (w <= 0 || h <= 0) return false;
            const uint32_t tile_size = std::max<uint32_t>(1u, fp.technique.tile_size);
            const uint32_t tiles_x = (uint32_t)((w + (int)tile_size - 1) / (int)tile_size);
            const uint32_t tiles_y = (uint32_t)((h + (int)tile_size - 1) / (int)tile_size);
            const uint32_t max_per_tile = std::max<uint32_t>(1u, fp.technique.max_lights_per_tile);
            const uint32_t directional = (scene.sun.intensity > 0.0f) ? 1u : 0u;

            std::vector<CullingLightGPU> lights{};
            if (scene.local_lights) scene.local_lights->flatten_cullable_gpu(lights);
            static_assert(sizeof(CullingLightGPU) == 160, "shsb_lights_upload takes the reference's 160-byte records");
            if (shsb_lights_upload(dev.ctx(), lights.empty() ? nullptr : lights.data(), (uint32_t)lights.size()) != SHSB_OK) return false;
            if (!lights.empty()) // pass_adapters.hpp:304: no local lights => every tile just holds the directional count
            {
                float vp[16];
                copy_mat(scene.cam.viewproj, vp);
                if (shsb_light_cull(dev.ctx(), vp, (uint32_t)w, (uint32_t)h, tile_size, max_per_tile) != SHSB_OK) return false;
            }

            payload->tile_size = tile_size;
            payload->tile_count_x = tiles_x;
            payload->tile_count_y = tiles_y;
            payload->max_lights_per_tile = max_per_tile;
            const size_t n_tiles = (size_t)tiles_x * tiles_y;
            payload->tile_light_counts.assign(n_tiles, std::min(max_per_tile, directional));
            payload->visible_light_count = directional;
            if (opt.fill_light_culling_payload && !lights.empty())
            {
                payload->visible_light_count = directional + frustum_visible_lights(lights, scene.cam.viewproj);
                std::vector<uint32_t> counts(n_tiles, 0u);
                if (shsb_light_lists_download(dev.ctx(), counts.data(), counts.size(), nullptr, 0) != SHSB_OK) return false;
                for (size_t t = 0; t < n_tiles; ++t) // pass_adapters.hpp:317-328: early break at the cap == min(cap, sum)
                    payload->tile_light_counts[t] = (uint32_t)std::min<uint64_t>(max_per_tile, (uint64_t)directional + counts[t]);
            }
            return true;
        }
    }

    // ---- shadow_map: PassShadowMapAdapter, pass_adapters.hpp:356-399
    class PassShadowMapPlugin final : public IRenderPass
    {
    public:
        PassShadowMapPlugin(Device& dev, PluginOptions opt, RT_Shadow rt_shadow) : dev_(dev), opt_(opt), rt_shadow_(rt_shadow) {}
        const char* id() const override { return "shadow_map"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::Visibility;
            c.supported_modes_mask = technique_mode_mask_all();
            c.semantics = {write_semantic(PassSemantic::ShadowMap, ContractDomain::Software, "shadow")};
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.write(make_rt_resource_ref(static_cast<const RTHandle&>(rt_shadow_), PassResourceType::Shadow, "shadow", PassResourceDomain::Software));
            return io;
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            if (!detail::request_complete(request) || !dev_.valid()) return PassExecutionResult::not_executed();
            shs::PassShadowMap::Inputs in{};
            in.scene = request.inputs.scene;
            in.fp = request.inputs.frame;
            in.rtr = request.inputs.registry;
            in.rt_shadow = rt_shadow_;
            PassShadowMap(dev_, opt_.sync_host).execute(ctx, in);
            return PassExecutionResult::executed_no_outputs();
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        RT_Shadow rt_shadow_{};
    };

    // ---- depth_prepass: PassDepthPrepassAdapter, pass_adapters.hpp:401-528.  The scratch HDR target of the reference only
    // exists because rasterize_mesh insists on a colour target (sw_render/rasterizer.hpp:190); the device draws depth only.
    class PassDepthPrepassPlugin final : public IRenderPass
    {
    public:
        PassDepthPrepassPlugin(Device& dev, PluginOptions opt, RT_Motion rt_motion) : dev_(dev), opt_(opt), rt_motion_(rt_motion) {}
        const char* id() const override { return "depth_prepass"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::Visibility;
            c.supported_modes_mask = technique_mode_bit(TechniqueMode::ForwardPlus) | technique_mode_bit(TechniqueMode::TiledDeferred) |
                                     technique_mode_bit(TechniqueMode::ClusteredForward);
            c.semantics = {write_semantic(PassSemantic::Depth, ContractDomain::Software, "depth")};
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.write(make_named_resource_ref("technique.depth_prepass", PassResourceType::Temp, PassResourceDomain::Software));
            return io;
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            (void)ctx;
            if (!detail::request_complete(request) || !dev_.valid()) return PassExecutionResult::not_executed();
            const FrameParams& fp = *request.inputs.frame;
            if (!fp.technique.depth_prepass || !rt_motion_.valid()) return PassExecutionResult::not_executed();
            auto* motion = static_cast<RT_ColorDepthMotion*>(request.inputs.registry->get(rt_motion_));
            if (!motion || motion->w <= 0 || motion->h <= 0) return PassExecutionResult::not_executed();
            std::vector<ShsbRenderItem> items;
            const ShsbScene s = detail::scene(dev_, *request.inputs.scene, items);
            ShsbFrameParams sfp = detail::frame_params(fp);
            sfp.light_culling = 0;
            if (shsb_pass_depth_prepass(dev_.ctx(), &s, &sfp, dev_.twin(motion), nullptr) != SHSB_OK) return PassExecutionResult::not_executed();
            if (opt_.sync_host)
            {
                dev_.download(motion);
                motion->motion.clear(Motion2f{}); // pass_adapters.hpp:490: the pre-pass zeroes the motion plane on the host side
            }
            PassExecutionResult out = PassExecutionResult::executed_no_outputs();
            out.produced_depth = true;
            return out;
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        RT_Motion rt_motion_{};
    };

    // ---- light_culling / cluster_light_assign: PassLightCullingAdapter :530-589, PassClusterLightAssignAdapter :661-719
    class PassLightCullingPlugin final : public IRenderPass
    {
    public:
        PassLightCullingPlugin(Device& dev, PluginOptions opt, RT_Motion rt_motion, bool cluster_assign)
            : dev_(dev), opt_(opt), rt_motion_(rt_motion), cluster_assign_(cluster_assign) {}
        const char* id() const override { return cluster_assign_ ? "cluster_light_assign" : "light_culling"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        RHIQueueClass preferred_queue() const override { return RHIQueueClass::Compute; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::LightCulling;
            c.supported_modes_mask = cluster_assign_ ? technique_mode_bit(TechniqueMode::ClusteredForward)
                                                     : (technique_mode_bit(TechniqueMode::ForwardPlus) | technique_mode_bit(TechniqueMode::TiledDeferred) |
                                                        technique_mode_bit(TechniqueMode::ClusteredForward));
            c.requires_depth_prepass = true;
            c.prefer_async_compute = true;
            c.semantics = {read_semantic(PassSemantic::Depth, ContractDomain::Software, "depth")};
            if (cluster_assign_) c.semantics.push_back(read_semantic(PassSemantic::LightClusters, ContractDomain::Software, "clusters"));
            c.semantics.push_back(write_semantic(PassSemantic::LightGrid, ContractDomain::Software, "light_grid"));
            c.semantics.push_back(write_semantic(PassSemantic::LightIndexList, ContractDomain::Software, "light_index_list"));
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.read(make_named_resource_ref("technique.depth_prepass", PassResourceType::Temp, PassResourceDomain::Software));
            if (cluster_assign_) io.read(make_named_resource_ref("technique.cluster_grid", PassResourceType::Temp, PassResourceDomain::Software));
            io.write(make_named_resource_ref("technique.light_grid", PassResourceType::Temp, PassResourceDomain::Software));
            io.write(make_named_resource_ref("technique.light_index_list", PassResourceType::Temp, PassResourceDomain::Software));
            return io;
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            (void)ctx;
            if (!detail::request_complete(request)) return PassExecutionResult::not_executed();
            if (!detail::run_light_culling(dev_, opt_, *request.inputs.scene, *request.inputs.frame, *request.inputs.registry, rt_motion_,
                                           request.inputs.light_culling, request.depth_prepass_ready, cluster_assign_))
                return PassExecutionResult::not_executed();
            PassExecutionResult out = PassExecutionResult::executed_no_outputs();
            out.produced_light_grid = true;
            out.produced_light_index_list = true;
            return out;
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        RT_Motion rt_motion_{};
        bool cluster_assign_;
    };

    // ---- cluster_build: PassClusterBuildAdapter, pass_adapters.hpp:591-659 (grid dimensions into the payload).  On the
    // device it also reduces the pre-pass depth to per-tile [min, max] view depth (shsb_tile_depth_range), the input of the
    // depth-range / clustered bin builders (shsb_light_cull_ex).
    class PassClusterBuildPlugin final : public IRenderPass
    {
    public:
        PassClusterBuildPlugin(Device& dev, PluginOptions opt, RT_Motion rt_motion) : dev_(dev), opt_(opt), rt_motion_(rt_motion) {}
        const char* id() const override { return "cluster_build"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        RHIQueueClass preferred_queue() const override { return RHIQueueClass::Compute; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::LightCulling;
            c.supported_modes_mask = technique_mode_bit(TechniqueMode::ClusteredForward);
            c.requires_depth_prepass = true;
            c.prefer_async_compute = true;
            c.semantics = {read_semantic(PassSemantic::Depth, ContractDomain::Software, "depth"),
                           write_semantic(PassSemantic::LightClusters, ContractDomain::Software, "clusters")};
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.read(make_named_resource_ref("technique.depth_prepass", PassResourceType::Temp, PassResourceDomain::Software));
            io.write(make_named_resource_ref("technique.cluster_grid", PassResourceType::Temp, PassResourceDomain::Software));
            return io;
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            (void)ctx;
            if (!detail::request_complete(request, false) || !dev_.valid()) return PassExecutionResult::not_executed();
            const FrameParams& fp = *request.inputs.frame;
            RTRegistry& rtr = *request.inputs.registry;
            if (fp.technique.depth_prepass && !request.depth_prepass_ready) return PassExecutionResult::not_executed();
            int w = fp.w, h = fp.h;
            RT_ColorDepthMotion* motion = rt_motion_.valid() ? static_cast<RT_ColorDepthMotion*>(rtr.get(rt_motion_)) : nullptr;
            if (motion && motion->w > 0 && motion->h > 0) { w = motion->w; h = motion->h; }
            if (w <= 0 || h <= 0) return PassExecutionResult::not_executed();
            auto* payload = request.inputs.light_culling;
            if (!payload) return PassExecutionResult::not_executed();
            payload->tile_size = std::max<uint32_t>(1u, fp.technique.tile_size);
            payload->tile_count_x = (uint32_t)((w + (int)payload->tile_size - 1) / (int)payload->tile_size);
            payload->tile_count_y = (uint32_t)((h + (int)payload->tile_size - 1) / (int)payload->tile_size);
            const size_t n_tiles = (size_t)payload->tile_count_x * (size_t)payload->tile_count_y;
            if (payload->tile_light_counts.size() != n_tiles) payload->tile_light_counts.assign(n_tiles, 0u);
            if (motion && fp.technique.depth_prepass && request.depth_prepass_ready)
            {
                if (!opt_.inputs_on_device) dev_.upload(motion);
                (void)shsb_tile_depth_range(dev_.ctx(), dev_.twin(motion), payload->tile_size);
            }
            return PassExecutionResult::executed_no_outputs();
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        RT_Motion rt_motion_{};
    };

    // ---- pbr_forward / pbr_forward_plus / pbr_forward_clustered: pass_adapters.hpp:1005-1062, 1064-1132, 936-1003
    class PassPBRForwardPlugin final : public IRenderPass
    {
    public:
        enum class Flavor { Forward, ForwardPlus, Clustered };
        PassPBRForwardPlugin(Device& dev, PluginOptions opt, Flavor flavor, RTHandle rt_hdr, RT_Motion rt_motion, RTHandle rt_shadow)
            : dev_(dev), opt_(opt), flavor_(flavor), rt_hdr_(rt_hdr), rt_motion_(rt_motion), rt_shadow_(rt_shadow) {}
        const char* id() const override
        {
            return flavor_ == Flavor::Forward ? "pbr_forward" : (flavor_ == Flavor::ForwardPlus ? "pbr_forward_plus" : "pbr_forward_clustered");
        }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::ForwardOpaque;
            c.semantics = {read_semantic(PassSemantic::ShadowMap, ContractDomain::Software, "shadow")};
            if (flavor_ == Flavor::Forward)
            {
                c.supported_modes_mask = technique_mode_bit(TechniqueMode::Forward) | technique_mode_bit(TechniqueMode::ForwardPlus) |
                                         technique_mode_bit(TechniqueMode::ClusteredForward);
            }
            else
            {
                c.supported_modes_mask = technique_mode_bit(flavor_ == Flavor::ForwardPlus ? TechniqueMode::ForwardPlus : TechniqueMode::ClusteredForward);
                c.requires_depth_prepass = true;
                c.requires_light_culling = true;
                c.semantics.push_back(read_semantic(PassSemantic::Depth, ContractDomain::Software, "depth"));
                c.semantics.push_back(read_semantic(PassSemantic::LightGrid, ContractDomain::Software, "light_grid"));
                c.semantics.push_back(read_semantic(PassSemantic::LightIndexList, ContractDomain::Software, "light_index_list"));
            }
            c.semantics.push_back(write_semantic(PassSemantic::ColorHDR, ContractDomain::Software, "hdr"));
            c.semantics.push_back(write_semantic(PassSemantic::MotionVectors, ContractDomain::Software, "motion"));
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.read(make_rt_resource_ref(rt_shadow_, PassResourceType::Shadow, "shadow", PassResourceDomain::Software));
            if (flavor_ != Flavor::Forward)
            {
                io.read(make_named_resource_ref("technique.depth_prepass", PassResourceType::Temp, PassResourceDomain::Software));
                io.read(make_named_resource_ref("technique.light_grid", PassResourceType::Temp, PassResourceDomain::Software));
                io.read(make_named_resource_ref("technique.light_index_list", PassResourceType::Temp, PassResourceDomain::Software));
            }
            io.write(make_rt_resource_ref(rt_hdr_, PassResourceType::ColorHDR, "hdr", PassResourceDomain::Software));
            io.write(make_rt_resource_ref(static_cast<const RTHandle&>(rt_motion_), PassResourceType::Motion, "motion", PassResourceDomain::Software));
            return io;
        }
        void reset_history(Context& ctx, RTRegistry& rtr) override
        {
            (void)ctx; (void)rtr;
            if (dev_.valid()) shsb_history_reset(dev_.ctx());
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            if (!detail::request_complete(request) || !dev_.valid()) return PassExecutionResult::not_executed();
            const FrameParams& fp = *request.inputs.frame;
            shs::PassPBRForward::Inputs in{};
            in.scene = request.inputs.scene;
            in.fp = &fp;
            in.rtr = request.inputs.registry;
            in.rt_hdr = rt_hdr_;
            in.rt_motion = rt_motion_;
            in.rt_shadow = rt_shadow_;
            bool lists_ready = false;
            if (flavor_ != Flavor::Forward)
            {
                // pass_adapters.hpp:1108-1122 (Forward+) and :980-991 (clustered)
                const bool culling_enabled = (flavor_ == Flavor::ForwardPlus)
                    ? (fp.technique.light_culling || fp.technique.mode == TechniqueMode::ForwardPlus)
                    : detail::technique_uses_light_culling(fp);
                const bool depth_ready = (!fp.technique.depth_prepass) || request.depth_prepass_ready;
                const bool culling_ready = (!culling_enabled) || request.light_culling_ready;
                in.preserve_existing_depth = depth_ready && culling_ready && fp.technique.depth_prepass;
                lists_ready = culling_enabled && request.light_culling_ready;
            }
            const int local_lights = (opt_.shade_local_lights && lists_ready) ? 1 : 0;
            PassPBRForward(dev_, opt_.sync_host, opt_.inputs_on_device, local_lights).execute(ctx, in);
            return PassExecutionResult::executed_no_outputs();
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        Flavor flavor_;
        RTHandle rt_hdr_{};
        RT_Motion rt_motion_{};
        RTHandle rt_shadow_{};
    };

    // ---- tonemap: PassTonemapAdapter, pass_adapters.hpp:1134-1182
    class PassTonemapPlugin final : public IRenderPass
    {
    public:
        PassTonemapPlugin(Device& dev, PluginOptions opt, RTHandle rt_hdr, RTHandle rt_ldr) : dev_(dev), opt_(opt), rt_hdr_(rt_hdr), rt_ldr_(rt_ldr) {}
        const char* id() const override { return "tonemap"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::Composite;
            c.supported_modes_mask = technique_mode_mask_all();
            c.semantics = {read_semantic(PassSemantic::ColorHDR, ContractDomain::Software, "hdr"),
                           write_semantic(PassSemantic::ColorLDR, ContractDomain::Software, "ldr")};
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.read(make_rt_resource_ref(rt_hdr_, PassResourceType::ColorHDR, "hdr", PassResourceDomain::Software));
            io.write(make_rt_resource_ref(rt_ldr_, PassResourceType::ColorLDR, "ldr", PassResourceDomain::Software));
            return io;
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            if (!detail::request_complete(request, false) || !dev_.valid()) return PassExecutionResult::not_executed();
            shs::PassTonemap::Inputs in{};
            in.fp = request.inputs.frame;
            in.rtr = request.inputs.registry;
            in.rt_hdr = rt_hdr_;
            in.rt_ldr = rt_ldr_;
            PassTonemap(dev_, opt_.sync_host, opt_.inputs_on_device).execute(ctx, in);
            return PassExecutionResult::executed_no_outputs();
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        RTHandle rt_hdr_{}, rt_ldr_{};
    };

    // ---- light_shafts / motion_blur: PassLightShaftsAdapter :1184-1276, PassMotionBlurAdapter :1278-1365.  Both work in place
    // on rt_ldr through a temporary; the reference's tmp handle stays in describe_io (so the frame graph is unchanged) but the
    // device owns its own scratch plane and never touches the host tmp target.
    class PassLightShaftsPlugin final : public IRenderPass
    {
    public:
        PassLightShaftsPlugin(Device& dev, PluginOptions opt, RTHandle rt_ldr_inout, RTHandle rt_depth_like, RTHandle rt_shafts_tmp)
            : dev_(dev), opt_(opt), rt_ldr_(rt_ldr_inout), rt_depth_like_(rt_depth_like), rt_shafts_tmp_(rt_shafts_tmp) {}
        const char* id() const override { return "light_shafts"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::PostProcess;
            c.supported_modes_mask = technique_mode_mask_all();
            c.semantics = {read_write_semantic(PassSemantic::ColorLDR, ContractDomain::Software, "ldr"),
                           read_semantic(PassSemantic::MotionVectors, ContractDomain::Software, "depth_like")};
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.read_write(make_rt_resource_ref(rt_ldr_, PassResourceType::ColorLDR, "ldr", PassResourceDomain::Software));
            io.read(make_rt_resource_ref(rt_depth_like_, PassResourceType::Motion, "motion", PassResourceDomain::Software));
            if (rt_shafts_tmp_.valid()) io.write(make_rt_resource_ref(rt_shafts_tmp_, PassResourceType::Temp, "shafts_tmp", PassResourceDomain::Software));
            else io.write(make_named_resource_ref("light_shafts.auto_tmp", PassResourceType::Temp, PassResourceDomain::Software));
            return io;
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            if (!detail::request_complete(request) || !dev_.valid()) return PassExecutionResult::not_executed();
            shs::PassLightShafts::Inputs in{};
            in.scene = request.inputs.scene;
            in.fp = request.inputs.frame;
            in.rtr = request.inputs.registry;
            in.rt_input_ldr = rt_ldr_;
            in.rt_output_ldr = rt_ldr_;
            in.rt_depth_like = rt_depth_like_;
            in.rt_shafts_tmp = rt_shafts_tmp_;
            PassLightShafts(dev_, opt_.sync_host, opt_.inputs_on_device).execute(ctx, in);
            return PassExecutionResult::executed_no_outputs();
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        RTHandle rt_ldr_{}, rt_depth_like_{}, rt_shafts_tmp_{};
    };

    class PassMotionBlurPlugin final : public IRenderPass
    {
    public:
        PassMotionBlurPlugin(Device& dev, PluginOptions opt, RTHandle rt_ldr_inout, RTHandle rt_motion, RTHandle rt_tmp)
            : dev_(dev), opt_(opt), rt_ldr_(rt_ldr_inout), rt_motion_(rt_motion), rt_tmp_(rt_tmp) {}
        const char* id() const override { return "motion_blur"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::PostProcess;
            c.supported_modes_mask = technique_mode_mask_all();
            c.semantics = {read_write_semantic(PassSemantic::ColorLDR, ContractDomain::Software, "ldr"),
                           read_semantic(PassSemantic::MotionVectors, ContractDomain::Software, "motion")};
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.read_write(make_rt_resource_ref(rt_ldr_, PassResourceType::ColorLDR, "ldr", PassResourceDomain::Software));
            io.read(make_rt_resource_ref(rt_motion_, PassResourceType::Motion, "motion", PassResourceDomain::Software));
            if (rt_tmp_.valid()) io.write(make_rt_resource_ref(rt_tmp_, PassResourceType::Temp, "motion_tmp", PassResourceDomain::Software));
            else io.write(make_named_resource_ref("motion_blur.auto_tmp", PassResourceType::Temp, PassResourceDomain::Software));
            return io;
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            if (!detail::request_complete(request, false) || !dev_.valid()) return PassExecutionResult::not_executed();
            shs::PassMotionBlur::Inputs in{};
            in.fp = request.inputs.frame;
            in.rtr = request.inputs.registry;
            in.rt_input_ldr = rt_ldr_;
            in.rt_output_ldr = rt_ldr_;
            in.rt_motion = rt_motion_;
            in.rt_tmp = rt_tmp_;
            PassMotionBlur(dev_, opt_.sync_host, opt_.inputs_on_device).execute(ctx, in);
            return PassExecutionResult::executed_no_outputs();
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        RTHandle rt_ldr_{}, rt_motion_{}, rt_tmp_{};
    };

    // ---- depth_of_field: PassDepthOfFieldAdapter, pass_adapters.hpp:1367-1400 -- a no-op in the reference, kept so that the
    // Deferred / TiledDeferred technique profiles assemble the same pass list.
    class PassDepthOfFieldPlugin final : public IRenderPass
    {
    public:
        const char* id() const override { return "depth_of_field"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::PostProcess;
            c.supported_modes_mask = technique_mode_bit(TechniqueMode::Deferred) | technique_mode_bit(TechniqueMode::TiledDeferred);
            c.semantics = {read_write_semantic(PassSemantic::ColorLDR, ContractDomain::Software, "ldr"),
                           read_semantic(PassSemantic::Depth, ContractDomain::Software, "depth")};
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.read_write(make_named_resource_ref("technique.ldr", PassResourceType::ColorLDR, PassResourceDomain::Software));
            io.read(make_named_resource_ref("technique.depth_prepass", PassResourceType::Temp, PassResourceDomain::Software));
            return io;
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            (void)ctx;
            return request.valid ? PassExecutionResult::executed_no_outputs() : PassExecutionResult::not_executed();
        }
    };

    // ---- taa: PassTemporalAAAdapter, pass_adapters.hpp:1402-1496
    class PassTemporalAAPlugin final : public IRenderPass
    {
    public:
        PassTemporalAAPlugin(Device& dev, PluginOptions opt, RTHandle rt_ldr_inout) : dev_(dev), opt_(opt), rt_ldr_(rt_ldr_inout) {}
        const char* id() const override { return "taa"; }
        RenderBackendType preferred_backend() const override { return RenderBackendType::Software; }
        bool supports_backend(RenderBackendType backend) const override { return backend == RenderBackendType::Software; }
        TechniquePassContract describe_contract() const override
        {
            TechniquePassContract c{};
            c.role = TechniquePassRole::PostProcess;
            c.supported_modes_mask = technique_mode_mask_all();
            c.semantics = {read_write_semantic(PassSemantic::ColorLDR, ContractDomain::Software, "ldr"),
                           read_semantic(PassSemantic::HistoryColor, ContractDomain::Software, "history_in"),
                           write_semantic(PassSemantic::HistoryColor, ContractDomain::Software, "history_out")};
            return c;
        }
        PassIODesc describe_io() const override
        {
            PassIODesc io{};
            io.read_write(make_rt_resource_ref(rt_ldr_, PassResourceType::ColorLDR, "ldr", PassResourceDomain::Software));
            io.read(make_named_resource_ref("technique.history_color", PassResourceType::Temp, PassResourceDomain::Software));
            io.write(make_named_resource_ref("technique.history_color", PassResourceType::Temp, PassResourceDomain::Software));
            return io;
        }
        void reset_history(Context& ctx, RTRegistry& rtr) override
        {
            (void)rtr;
            if (dev_.valid()) PassTemporalAA(dev_).reset_history(ctx);
            else ctx.temporal_aa.reset();
        }
        PassExecutionResult execute_resolved(Context& ctx, const PassExecutionRequest& request) override
        {
            if (!request.valid || !request.inputs.registry || !dev_.valid()) return PassExecutionResult::not_executed();
            auto* ldr = static_cast<RT_ColorLDR*>(request.inputs.registry->get(rt_ldr_));
            if (!ldr || ldr->w <= 0 || ldr->h <= 0) return PassExecutionResult::not_executed();
            PassTemporalAA(dev_, opt_.sync_host, opt_.inputs_on_device).execute(ctx, *request.inputs.registry, rt_ldr_);
            return PassExecutionResult::executed_no_outputs();
        }

    private:
        Device& dev_;
        PluginOptions opt_;
        RTHandle rt_ldr_{};
    };

    // make_standard_pass_factory_registry (pass_adapters.hpp:1497-1569) with the raster hot path on the device.  `dev` must
    // outlive every pass the registry creates.  Descriptors are the reference's own standard contracts
    // (lookup_standard_pass_contract, pipeline/pass_contract_registry.hpp:22) with the Software-only backend mask.
    inline PassFactoryRegistry make_b200_pass_factory_registry(Device& dev, RT_Shadow rt_shadow, RTHandle rt_hdr, RT_Motion rt_motion, RTHandle rt_ldr,
                                                               RTHandle rt_shafts_tmp, RTHandle rt_motion_blur_tmp, PluginOptions opt = {})
    {
        PassFactoryRegistry reg{};
        const uint32_t sw_only = PassFactoryRegistry::backend_bit(RenderBackendType::Software);
        auto register_standard = [&](PassId pass_id, PassFactoryRegistry::Factory f) {
            reg.register_factory(pass_id, std::move(f));
            TechniquePassContract c{};
            if (lookup_standard_pass_contract(pass_id, c)) reg.register_descriptor(pass_id, c, sw_only, true);
        };
        Device* d = &dev;
        const RTHandle shadow_h{rt_shadow.id};
        using Fwd = PassPBRForwardPlugin;
        register_standard(PassId::ShadowMap, [=]() { return std::make_unique<PassShadowMapPlugin>(*d, opt, rt_shadow); });
        register_standard(PassId::PBRForward, [=]() { return std::make_unique<Fwd>(*d, opt, Fwd::Flavor::Forward, rt_hdr, rt_motion, shadow_h); });
        register_standard(PassId::DepthPrepass, [=]() { return std::make_unique<PassDepthPrepassPlugin>(*d, opt, rt_motion); });
        register_standard(PassId::LightCulling, [=]() { return std::make_unique<PassLightCullingPlugin>(*d, opt, rt_motion, false); });
        register_standard(PassId::ClusterBuild, [=]() { return std::make_unique<PassClusterBuildPlugin>(*d, opt, rt_motion); });
        register_standard(PassId::ClusterLightAssign, [=]() { return std::make_unique<PassLightCullingPlugin>(*d, opt, rt_motion, true); });
        register_standard(PassId::PBRForwardPlus, [=]() { return std::make_unique<Fwd>(*d, opt, Fwd::Flavor::ForwardPlus, rt_hdr, rt_motion, shadow_h); });
        register_standard(PassId::PBRForwardClustered, [=]() { return std::make_unique<Fwd>(*d, opt, Fwd::Flavor::Clustered, rt_hdr, rt_motion, shadow_h); });
        register_standard(PassId::Tonemap, [=]() { return std::make_unique<PassTonemapPlugin>(*d, opt, rt_hdr, rt_ldr); });
        reg.register_factory("light_shafts", [=]() { return std::make_unique<PassLightShaftsPlugin>(*d, opt, rt_ldr, rt_motion, rt_shafts_tmp); });
        register_standard(PassId::MotionBlur, [=]() { return std::make_unique<PassMotionBlurPlugin>(*d, opt, rt_ldr, rt_motion, rt_motion_blur_tmp); });
        register_standard(PassId::DepthOfField, [=]() { return std::make_unique<PassDepthOfFieldPlugin>(); });
        register_standard(PassId::TAA, [=]() { return std::make_unique<PassTemporalAAPlugin>(*d, opt, rt_ldr); });
        return reg;
    }
}
