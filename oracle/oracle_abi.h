/*
 * oracle/oracle_abi.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Shared C interface of the two CPU checkers:
 *   shso_*   : oracle/oracle.cpp      -- this repo's CPU restatement of the reference algorithm
 *   shsref_* : oracle/ref_harness.cpp -- the reference's OWN headers compiled from /root/reference
 *                                        against oracle/glm_shim (built into oracle/_ref/)
 * Both take host pointers and the POD structs of include/shsb.h, so a test feeds the same bytes to
 * the reference, the restatement and the CUDA path.  Nothing in the product may include this.
 */
#ifndef SHS_ORACLE_ABI_H
#define SHS_ORACLE_ABI_H

#include "../include/shsb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ShsoMesh /* MeshData, resources/mesh.hpp:23 */
{
    const float* positions;
    const float* normals;
    const float* uvs;
    const uint32_t* indices;
    uint32_t n_positions;
    uint32_t n_normals;
    uint32_t n_uvs;
    uint32_t n_indices;
} ShsoMesh;

typedef struct ShsoTexture /* Texture2DData, resources/texture.hpp:23 */
{
    const uint8_t* rgba;
    int32_t w;
    int32_t h;
} ShsoTexture;

typedef struct ShsoAssets /* what ResourceRegistry resolves handles against (1-based) */
{
    const ShsoMesh* meshes;
    const ShsoTexture* textures;
    uint32_t n_meshes;
    uint32_t n_textures;
} ShsoAssets;

typedef struct ShsoTarget /* RasterizerTarget + the RTs behind it, all host memory */
{
    float* hdr;          /* W*H*4, required                                                    */
    float* depth;        /* W*H or NULL (no depth target => painter)                           */
    const float* shadow; /* shadow map (sw*sh) or NULL                                         */
    uint32_t* tri_id;    /* optional AOV: draw-order key of the winning fragment               */
    uint32_t* coverage;  /* optional AOV: per-pixel count of fragments passing coverage + 1/w  */
    int32_t w, h;
    int32_t shadow_w, shadow_h;
    float zn, zf;
    float* motion;       /* optional W*H*2: RT_ColorDepthMotion::motion (needs depth)           */
} ShsoTarget;

#define SHSO_FN(ret, name, args) ret shso_##name args; ret shsref_##name args

/* glm::translate/rotate/scale chain, pass_pbr_forward.hpp:136-141 */
SHSO_FN(void, model_from_transform, (const ShsbTransform* tr, float out_model[16]));
/* look_at_lh / perspective_lh_no / proj*view, camera/convention.hpp:19-33 */
SHSO_FN(void, camera_viewproj, (const float eye[3], const float target[3], const float up[3],
                                float fovy_radians, float aspect, float znear, float zfar, float out_viewproj[16]));
/* make_point_culling_light / make_spot_culling_light, lighting/light_types.hpp:327-379;
 * jolt_bounds != 0 applies the Jolt sphere-shape bounds of scene_shape.hpp:56-81 (radius = |vec3(range)|). */
SHSO_FN(void, pack_point_light, (const float pos[3], float range, const float color[3], float intensity,
                                 uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff,
                                 int32_t jolt_bounds, void* out_record160));
SHSO_FN(void, pack_spot_light, (const float pos[3], float range, const float color[3], float intensity,
                                const float dir[3], float inner_rad, float outer_rad,
                                uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff,
                                void* out_record160));

/* rasterize_mesh, sw_render/rasterizer.hpp:181.  key_base: draw-order key of triangle 0, fan 0
 * (key = key_base + tri*8 + fan_k; the reference harness cannot see fan_k and reports fan 0). */
SHSO_FN(int32_t, rasterize_mesh, (const ShsoAssets* assets, shsb_mesh mesh, int32_t shader_id,
                                  const ShsbUniforms* u, const ShsoTarget* tgt, const ShsbRasterCfg* cfg,
                                  uint32_t key_base, ShsbStats* out_stats));

/* PassPBRForward::execute, passes/pass_pbr_forward.hpp:49 */
SHSO_FN(int32_t, pass_pbr_forward, (const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                                    const ShsoTarget* tgt, const float* shadow_light_viewproj,
                                    int32_t preserve_existing_depth, ShsbStats* out_stats));

/* The same pass with Context::history populated (core/context.hpp:84-94): prev_models16 = n_items model matrices of
 * the previous frame in scene->items order (has_prev_frame = true), or NULL for a first frame.  Writes tgt->motion when
 * fp->motion_vectors_enable and the target has depth + motion planes (rasterizer.hpp:295-307, 388-411). */
SHSO_FN(int32_t, pass_pbr_forward_history, (const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                                            const ShsoTarget* tgt, const float* shadow_light_viewproj,
                                            int32_t preserve_existing_depth, const float* prev_models16, ShsbStats* out_stats));

/* PassShadowMap::execute, passes/pass_shadow_map.hpp:44 */
SHSO_FN(int32_t, pass_shadow_map, (const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                                   float* shadow, int32_t sw, int32_t sh, float out_light_viewproj[16]));

/* PassTonemap::execute, passes/pass_tonemap.hpp:37 */
SHSO_FN(int32_t, pass_tonemap, (const float* hdr, int32_t w, int32_t h, float exposure, float gamma, uint8_t* out_ldr));

/* PassMotionBlur::execute, passes/pass_motion_blur.hpp:40 (input != output; all planes w x h; motion = 2 floats / px) */
SHSO_FN(int32_t, pass_motion_blur, (const ShsbMotionBlurParams* p, const uint8_t* src_ldr, const float* motion, const float* depth,
                                    int32_t w, int32_t h, uint8_t* out_ldr));

/* PassLightShafts::execute, passes/pass_light_shafts.hpp:43 (depth may be NULL: no rt_depth_like) */
SHSO_FN(int32_t, pass_light_shafts, (const ShsbLightShaftsParams* p, const uint8_t* src_ldr, const float* depth,
                                     int32_t w, int32_t h, uint8_t* out_ldr));

#undef SHSO_FN

/* ---- entry points without a twin in ref_harness.cpp: the light-list builders and TAA are pinned through the Jolt declaration
 * shim (ref_lightcull_harness.cpp / ref_taa_harness.cpp); the depth reduce and the Forward+ fragment loop follow a GLSL spec ---- */

/* cull_lights_tiled, lighting/jolt_light_culling.hpp:135-187 over CullingLightGPU cull_sphere/cull_aabb.
 * counts[T] uncapped; indices[T*max_per_tile] first max_per_tile entries, ascending. */
int32_t shso_light_cull(const void* records160, uint32_t n_lights, const float view_proj[16],
                        uint32_t viewport_w, uint32_t viewport_h, uint32_t tile_size, uint32_t max_per_tile,
                        uint32_t* counts, uint32_t* indices);

/* cull_lights_tiled_depth01_range / cull_lights_tiled_view_depth_range / cull_lights_clustered
 * (lighting/jolt_light_culling.hpp:196-412) and mode 0 = cull_lights_tiled.  counts[bins], indices[bins * max_per_bin],
 * bins = tiles (* depth_slices for clustered), bin = cz * tiles + ty * tiles_x + tx. */
int32_t shso_light_cull_ex(const void* records160, uint32_t n_lights, const ShsbLightCullDesc* desc,
                           const float* range_min, const float* range_max, uint32_t* counts, uint32_t* indices);

/* Per-tile [min, max] view depth of a z-buffer (software analogue of shaders/vulkan/fp_stress_depth_reduce.comp;
 * semantics defined in oracle.cpp). */
int32_t shso_tile_depth_range(const float* depth, int32_t w, int32_t h, uint32_t tile_size, float zn, float zf,
                              float* out_min, float* out_max);
/* shaders/vulkan/fp_stress_depth_reduce.comp itself, for a z-buffer that holds projective depth in [0, 1] (pinned against the shader's text) */
int32_t shso_tile_depth_range_ndc01(const float* depth, int32_t w, int32_t h, uint32_t tile_size, float z_near, float z_far,
                                    float* out_min, float* out_max);

/* Forward+ frame: PassPBRForward with the local-light loop of fp_stress_scene.frag:644-678 added to
 * the builtin fragment program.  counts/indices as produced by shso_light_cull. */
int32_t shso_pass_pbr_forward_plus(const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                                   const ShsoTarget* tgt, const float* shadow_light_viewproj,
                                   int32_t preserve_existing_depth,
                                   const void* records160, uint32_t n_lights,
                                   const uint32_t* counts, const uint32_t* indices,
                                   ShsbStats* out_stats);

/* PassDepthPrepassAdapter, pipeline/pass_adapters.hpp:401-528 */
int32_t shso_pass_depth_prepass(const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                                const ShsoTarget* tgt, ShsbStats* out_stats);

/* PassTemporalAAAdapter::execute_resolved, pipeline/pass_adapters.hpp:1438-1491 (that header needs Jolt types, so
 * restatement only).  history_valid == 0: seeds `history` from `ldr` and leaves the frame untouched. */
int32_t shso_pass_taa(uint8_t* ldr_inout, uint8_t* history_inout, int32_t history_valid, size_t n_pixels);

/* make_rect_area_culling_light / make_tube_area_culling_light, lighting/light_types.hpp:379-436 (reference only). */
void shsref_pack_rect_light(const float pos[3], float range, const float color[3], float intensity,
                            const float dir[3], const float right[3], float half_x, float half_y,
                            uint32_t flags, uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff, void* out_record160);
void shsref_pack_tube_light(const float pos[3], float range, const float color[3], float intensity,
                            const float axis[3], float half_length, float radius,
                            uint32_t flags, uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff, void* out_record160);

/* ThreadPoolJobSystem(n) handed to the reference passes as ctx.job_system / RasterizerConfig::job_system
 * (exp-plumbing/hello_pass_basics.cpp:629-630); n <= 1 means no job system (serial). */
int32_t shsref_set_threads(int32_t n_threads);

#ifdef __cplusplus
}
#endif

#endif
