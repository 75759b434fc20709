/*
 * shsb.h -- C-ABI of the B200-native rasterization hot path ("shsb" = SHS renderer, B200).
 *
 * This is the drop-in boundary: plain pointers, sizes and POD
 * structs, no C++ / torch types.  Every entry point names the reference interface it replaces
 * (paths relative to /root/reference/cpp-folders/src/shs-renderer-lib/include/shs/).
 *
 * Conventions kept from the reference:
 *   - matrices are column-major float[16] exactly like glm::mat4 (m[col*4+row]);
 *   - render targets are row-major, origin bottom-left, data[y*w+x]    (gfx/rt_types.hpp:35-59);
 *   - asset / RT handles are 1-based uint32_t, 0 = invalid             (gfx/rt_handle.hpp:19-20,
 *                                                                        resources/resource_registry.hpp:29);
 *   - no exceptions: the reference silently returns zero stats on bad input
 *     (sw_render/rasterizer.hpp:190-194); here every call returns an int32 status, SHSB_OK = 0,
 *     and shsb_last_error_string() explains a failure.
 *   - one host thread per context; calls are asynchronous on the context's CUDA streams until
 *     shsb_sync / a download.  Results are ordered as submitted: the library tracks hazards per render
 *     target, so asynchronous frames into DIFFERENT targets may overlap on the device while every
 *     operation on one target sees the operations submitted before it.
 *
 * There is NO CPU fallback: if no CUDA device is usable shsb_context_create fails with
 * SHSB_E_NO_DEVICE, and arbitrary host std::function shaders are rejected with
 * SHSB_E_UNSUPPORTED_SHADER (only the reference's builtin programs exist on the device).
 */
#ifndef SHSB_H
#define SHSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SHSB_API
#else
#define SHSB_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------ status codes */
enum
{
    SHSB_OK = 0,
    SHSB_E_INVALID_ARGUMENT = 1,
    SHSB_E_INVALID_HANDLE = 2,
    SHSB_E_NO_DEVICE = 3,
    SHSB_E_CUDA = 4,
    SHSB_E_UNSUPPORTED_SHADER = 5,
    SHSB_E_OUT_OF_MEMORY = 6,
    SHSB_E_SIZE_MISMATCH = 7,
    SHSB_E_UNSUPPORTED = 8,
    SHSB_E_TIMEOUT = 9,
    SHSB_E_OVERFLOW = 10 /* an earlier ASYNCHRONOUS submission did not fit its per-frame arena (see ShsbStats below) */
};

typedef struct shsb_context_t* shsb_ctx; /* opaque */
typedef uint32_t shsb_mesh;              /* MeshAssetHandle     resources/mesh.hpp:21          */
typedef uint32_t shsb_tex;               /* TextureAssetHandle  resources/texture.hpp:21       */
typedef uint32_t shsb_rt;                /* RTHandle::id        gfx/rt_handle.hpp:17           */

/* ------------------------------------------------------------------ enums mirrored from the reference */
enum /* RasterizerCullMode, sw_render/rasterizer.hpp:26-31 == CullMode, frame/frame_params.hpp */
{
    SHSB_CULL_NONE = 0,
    SHSB_CULL_BACK = 1,
    SHSB_CULL_FRONT = 2
};

enum /* device shader programs == the reference's builtin ShaderProgram factories */
{
    SHSB_SHADER_PBR_MR = 0,       /* make_pbr_mr_program        shader/builtin_shaders.hpp:154 */
    SHSB_SHADER_BLINN_PHONG = 1,  /* make_blinn_phong_program   shader/builtin_shaders.hpp:105 */
    SHSB_SHADER_DEBUG_ALBEDO = 2, /* make_debug_view_shader_program(Albedo)            :221   */
    SHSB_SHADER_DEBUG_NORMAL = 3, /* ...(Normal)                                               */
    SHSB_SHADER_DEBUG_DEPTH = 4,  /* ...(Depth)                                                */
    SHSB_SHADER_DEPTH_ONLY = 5,   /* detail::make_depth_prepass_program  pipeline/pass_adapters.hpp:335 */
    SHSB_SHADER_COUNT = 6
};

enum /* ShadingModel, frame/frame_params.hpp */
{
    SHSB_SHADING_PBR_METAL_ROUGH = 0,
    SHSB_SHADING_BLINN_PHONG = 1
};

enum /* DebugViewMode, frame/frame_params.hpp */
{
    SHSB_DEBUG_FINAL = 0,
    SHSB_DEBUG_ALBEDO = 1,
    SHSB_DEBUG_NORMAL = 2,
    SHSB_DEBUG_DEPTH = 3
};

enum /* render-target kinds, gfx/rt_types.hpp:61-157, gfx/rt_shadow.hpp:18 */
{
    SHSB_RT_COLOR_HDR = 1,    /* RT_ColorHDR: RGBA32F                                         */
    SHSB_RT_COLOR_LDR = 2,    /* RT_ColorLDR: RGBA8                                           */
    SHSB_RT_DEPTH_MOTION = 3, /* RT_ColorDepthVelocity: depth f32 (+ motion 2xf32), zn/zf     */
    SHSB_RT_SHADOW = 4        /* RT_ShadowDepth: f32                                          */
};

enum /* ISkyModel implementations of the reference (Scene::sky) */
{
    SHSB_SKY_NONE = 0,       /* background gradient, pass_pbr_forward.hpp:69-85 */
    SHSB_SKY_PROCEDURAL = 1, /* ProceduralSky, sky/procedural_sky.hpp:19-47     */
    SHSB_SKY_CUBEMAP = 2     /* CubemapSky, sky/cubemap_sky.hpp:62-116          */
};

enum /* planes of a render target for clear / upload / download / device_ptr */
{
    SHSB_PLANE_COLOR = 0,  /* HDR: float4; LDR: uchar4                                        */
    SHSB_PLANE_DEPTH = 1,  /* f32 (DEPTH_MOTION, SHADOW)                                      */
    SHSB_PLANE_MOTION = 2, /* 2 x f32 (DEPTH_MOTION)                                          */
    SHSB_PLANE_TRI_ID = 3, /* u32 AOV: draw-order key of the winning fragment, 0xFFFFFFFF none */
    SHSB_PLANE_COVERAGE = 4 /* u32 AOV: number of fragments that passed coverage+1/w test      */
};

#define SHSB_TRI_ID_NONE 0xFFFFFFFFu
#define SHSB_LIGHT_RECORD_BYTES 160u /* sizeof(CullingLightGPU), lighting/light_types.hpp:141-167 */

/* ------------------------------------------------------------------ POD mirrors of reference structs */

/* Passing `out_stats` makes a submission SYNCHRONOUS (the call returns when the frame is done, and re-runs the frame with larger
 * per-frame arenas if the set-up records / tile lists did not fit).  Passing NULL makes it asynchronous: such a frame cannot re-run
 * itself; if it overflows its arena it renders with geometry missing, raises a sticky flag, and the NEXT frame submission,
 * shsb_sync or shsb_rt_download returns SHSB_E_OVERFLOW once (capacities are grown x4 at that point: submit the frame again).
 * The arenas are sized from the submission (every source triangle one record, a quarter of them clipped into 7); a first frame
 * with statistics sizes them for good. */
typedef struct ShsbStats /* RasterizerStats, sw_render/rasterizer.hpp:48-53 (+ fragment counters) */
{
    uint64_t tri_input;
    uint64_t tri_after_clip;
    uint64_t tri_raster;
    uint64_t frag_covered; /* fragments passing coverage and 1/w tests (rasterizer.hpp:338,342) */
    uint64_t frag_shaded;  /* pixels whose final colour came from a fragment of this call       */
} ShsbStats;

typedef struct ShsbRasterCfg /* RasterizerConfig, sw_render/rasterizer.hpp:33-40 */
{
    int32_t cull_mode;      /* SHSB_CULL_*                                                    */
    int32_t front_face_ccw; /* bool                                                           */
    int32_t write_aovs;     /* extension: also fill the TRI_ID / COVERAGE planes of the HDR RT */
    int32_t reserved;       /* job_system / parallel_min_* have no meaning on the device      */
} ShsbRasterCfg;

typedef struct ShsbUniforms /* device-visible subset of ShaderUniforms, shader/types.hpp:82-116 */
{
    float model[16];
    float viewproj[16];
    float light_viewproj[16];
    float light_dir_ws[3];
    float light_intensity;
    float light_color[3];
    float metallic;
    float camera_pos[3];
    float roughness;
    float base_color[3];
    float ao;
    shsb_tex base_color_tex; /* 0 = none (sampler returns 1,1,1; builtin_shaders.hpp:35)     */
    shsb_rt shadow_map;      /* 0 = none (shadow_vis = 1)                                     */
    float shadow_bias_const;
    float shadow_bias_slope;
    int32_t shadow_pcf_radius;
    float shadow_pcf_step;
    float shadow_strength;
    int32_t enable_motion_vectors; /* ShaderUniforms::enable_motion_vectors (shader/types.hpp), rasterizer.hpp:295 */
    float prev_model[16];          /* ShaderUniforms::prev_model    (rasterizer.hpp:296-307)                      */
    float prev_viewproj[16];       /* ShaderUniforms::prev_viewproj (rasterizer.hpp:393)                          */
} ShsbUniforms;

typedef struct ShsbTransform /* Transform, scene/scene_types.hpp:26-31 */
{
    float pos[3];
    float rot_euler[3];
    float scl[3];
} ShsbTransform;

typedef struct ShsbRenderItem /* RenderItem (scene/scene_types.hpp:62-70) with its MaterialData resolved */
{
    ShsbTransform tr;
    shsb_mesh mesh;
    uint32_t has_material; /* RenderItem::mat, the material handle: 0 = none -> reference defaults (0.8,0.5,0.2) m=0.1 r=0.5 ao=1, pass_pbr_forward.hpp:179-184;
                              any other value = the MaterialData resolved below (the value itself enters the derived motion key like item.mat, :146) */
    float base_color[3];
    float metallic;
    float roughness;
    float ao;
    shsb_tex base_color_tex;
    uint32_t casts_shadow;
    uint32_t visible;
    uint64_t object_id;    /* RenderItem::object_id: key of the previous-frame model matrix; 0 = derive (pass_pbr_forward.hpp:143-148) */
} ShsbRenderItem;

typedef struct ShsbScene /* Scene, scene/scene_types.hpp:92-104 (camera + sun + items) */
{
    float cam_viewproj[16];
    float cam_pos[3];
    float sun_intensity;
    float sun_dir_ws[3];
    uint32_t n_items;
    float sun_color[3];
    uint32_t reserved;
    const ShsbRenderItem* items;
    float cam_prev_viewproj[16]; /* Camera::prev_viewproj (scene/scene_types.hpp:59); used once the context has a previous frame */
    /* Scene::sky (ISkyModel*, sky/sky_model.hpp:17): the reference's two sky models exist on the device */
    int32_t sky_kind;            /* SHSB_SKY_*                                                          */
    float sky_intensity;         /* CubemapSky intensity (sky/cubemap_sky.hpp:66)                       */
    float sky_sun_dir_ws[3];     /* ProceduralSky sun direction (sky/procedural_sky.hpp:22)             */
    shsb_tex sky_faces[6];       /* CubemapData::face: +X -X +Y -Y +Z -Z (sky/cubemap_sky.hpp:24)       */
    uint32_t reserved2;
} ShsbScene;

typedef struct ShsbFrameParams /* the FrameParams fields the path reads, frame/frame_params.hpp:117-171 */
{
    int32_t shading_model;  /* SHSB_SHADING_*                                                 */
    int32_t debug_view;     /* SHSB_DEBUG_*                                                   */
    int32_t cull_mode;      /* SHSB_CULL_*                                                    */
    int32_t front_face_ccw;
    int32_t shadow_enable;  /* pass.shadow.enable                                             */
    float shadow_bias_const;
    float shadow_bias_slope;
    int32_t shadow_pcf_radius;
    float shadow_pcf_step;
    float shadow_strength;
    float exposure;         /* pass.tonemap.exposure                                          */
    float gamma;            /* pass.tonemap.gamma                                             */
    int32_t light_culling;  /* technique.light_culling: Forward+ local-light loop on/off      */
    uint32_t tile_size;     /* technique.tile_size (16)                                       */
    uint32_t max_lights_per_tile; /* technique.max_lights_per_tile (128)                      */
    int32_t write_aovs;
    int32_t motion_vectors_enable; /* pass.motion_vectors.enable (frame/frame_params.hpp:46-49; reference default: true) */
    /* Sort-first screen partition (no reference counterpart -- the CPU renderer draws whole frames):
     * this submission rasterises, shades and writes ONLY the 16-px tile rows it owns, row ty (0 = top of the frame) being
     * owned iff ty >= own_row_first and (ty - own_row_first) % own_row_stride < own_row_count.  own_row_count = 0 means the
     * whole frame.  Pixels of other rows are left untouched in every plane, so the union of the ranks' submissions is
     * bit-identical to one whole-frame submission.  Draws whose projected bounds cannot reach an owned row are skipped on the host,
     * so triangle statistics cover the remaining draws; fragment statistics are per owner and add up to the whole frame's.
     * With 16-px light tiles the fused frame builds only the owned rows' light lists.  A partition that does not write motion
     * vectors leaves the motion history empty (it drops draws before their model matrix exists). */
    int32_t own_row_first;
    int32_t own_row_count;
    int32_t own_row_stride;
} ShsbFrameParams;

typedef struct ShsbMotionBlurParams /* MotionBlurPassParams (frame/frame_params.hpp:51-59) + FrameParams::dt */
{
    int32_t enable;        /* 0: the pass only copies input -> output (pass_motion_blur.hpp:56-60)   */
    int32_t samples;       /* clamped to 4..32 (pass_motion_blur.hpp:79)                              */
    float strength;
    float max_velocity_px;
    float min_velocity_px;
    float depth_reject;
    float dt;              /* FrameParams::dt, seconds (pass_motion_blur.hpp:84)                      */
    int32_t reserved;
} ShsbMotionBlurParams;

typedef struct ShsbLightShaftsParams /* LightShaftsPassParams (frame/frame_params.hpp:35-42) + the Scene fields the pass reads */
{
    int32_t enable;
    int32_t steps;         /* max(8, steps) (pass_light_shafts.hpp:135)                               */
    float density;
    float weight;
    float decay;
    float cam_pos[3];      /* Scene::cam.pos       (pass_light_shafts.hpp:81)                          */
    float sun_dir_ws[3];   /* Scene::sun.dir_ws                                                       */
    float reserved;
    float cam_viewproj[16];/* Scene::cam.viewproj  (pass_light_shafts.hpp:82)                          */
} ShsbLightShaftsParams;

enum /* LightCullingMode (lighting/light_culling_mode.hpp) as the bin builders of lighting/jolt_light_culling.hpp */
{
    SHSB_LIGHT_CULL_TILED = 0,            /* cull_lights_tiled                    :135-187 */
    SHSB_LIGHT_CULL_TILED_DEPTH01 = 1,    /* cull_lights_tiled_depth01_range      :196-258 */
    SHSB_LIGHT_CULL_TILED_VIEW_DEPTH = 2, /* cull_lights_tiled_view_depth_range   :261-324 */
    SHSB_LIGHT_CULL_CLUSTERED = 3         /* cull_lights_clustered                :341-412 */
};

typedef struct ShsbLightCullDesc /* arguments of the bin builders + LightBinCullingConfig (lighting/light_culling_runtime.hpp:29-36) */
{
    float view_proj[16];
    uint32_t viewport_w, viewport_h;
    uint32_t tile_size;
    uint32_t max_per_bin;   /* entries kept per bin in `indices` (counts stay uncapped)            */
    int32_t mode;           /* SHSB_LIGHT_CULL_*                                                    */
    uint32_t depth_slices;  /* clustered: cluster_depth_slices (16)                                 */
    float z_near, z_far;    /* view-depth and clustered modes                                       */
} ShsbLightCullDesc;

/* ------------------------------------------------------------------ context */

/* Creates a device context on CUDA device `device_ordinal`.  Replaces the reference's
 * Context + ThreadPoolJobSystem set-up (core/context.hpp:116, job/thread_pool_job_system.hpp:26):
 * the job-system fan-out becomes kernel launches on one CUDA stream. */
SHSB_API int32_t shsb_context_create(int32_t device_ordinal, shsb_ctx* out_ctx);
SHSB_API int32_t shsb_context_destroy(shsb_ctx ctx);
SHSB_API const char* shsb_last_error_string(shsb_ctx ctx);
SHSB_API int32_t shsb_sync(shsb_ctx ctx);
/* The context's main CUDA stream (cudaStream_t as void*) so callers can order their own work / events on it.
 * Asynchronous frames (out_stats == NULL) into different render targets may run on internal side streams; the
 * call therefore first does what shsb_fence does.  A caller that caches the handle calls shsb_fence before it
 * records an event that must cover frames submitted since, and after it has put work of its own on the stream
 * that later frames must not overtake. */
SHSB_API int32_t shsb_stream(shsb_ctx ctx, void** out_stream);
/* Two-way ordering point, no host synchronisation: the main stream waits for every frame submitted so far, and
 * every frame submitted afterwards starts behind whatever is on the main stream now.  Replaces nothing in the
 * reference (its passes run serially on the calling thread, pluggable_pipeline.hpp:125-136). */
SHSB_API int32_t shsb_fence(shsb_ctx ctx);
/* Number of render streams asynchronous frames are spread over (1..4, default 2; environment SHSB_TILE_STREAMS).
 * 1 = every tile kernel on the main stream, back to back (what bench.py uses to time the tile kernel alone). */
SHSB_API int32_t shsb_set_tile_streams(shsb_ctx ctx, int32_t n);
/* Number of kernels this context launched since creation (bench.py's gpu_launches). */
SHSB_API int32_t shsb_launch_count(shsb_ctx ctx, uint64_t* out_count);
/* Diagnostics: which instantiation of the tile kernel the last frame / draw of this context launched -- 0 the general one, else
 * PROGRAM * 10 + LIGHTS with PROGRAM 1 = PBR, 2 = Blinn-Phong and LIGHTS 1 = Forward+ over point / spot lights, 3 = Forward+ with area
 * lights possible, 2 = no local lights (csrc/tile_raster.cu); -1 before the first launch.  No reference counterpart. */
SHSB_API int32_t shsb_last_tile_kernel(shsb_ctx ctx, int32_t* out_mode);
/* Library build info string (arch, flags). */
SHSB_API const char* shsb_version(void);

/* ------------------------------------------------------------------ resources */

/* MeshData (resources/mesh.hpp:23-44): SoA positions(vec3) normals(vec3) uvs(vec2) + u32 indices.
 * n_normals / n_uvs may be shorter than n_positions (defaults (0,1,0) / (0,0), rasterizer.hpp:196-202);
 * n_indices == 0 means non-indexed (rasterizer.hpp:204-205). */
SHSB_API int32_t shsb_mesh_upload(shsb_ctx ctx,
                                  const float* positions, uint32_t n_positions,
                                  const float* normals, uint32_t n_normals,
                                  const float* uvs, uint32_t n_uvs,
                                  const uint32_t* indices, uint32_t n_indices,
                                  shsb_mesh* out_mesh);
SHSB_API int32_t shsb_mesh_destroy(shsb_ctx ctx, shsb_mesh mesh);

/* Texture2DData (resources/texture.hpp:23-50): RGBA8, row-major. */
SHSB_API int32_t shsb_texture_upload(shsb_ctx ctx, const uint8_t* rgba, int32_t w, int32_t h, shsb_tex* out_tex);
SHSB_API int32_t shsb_texture_destroy(shsb_ctx ctx, shsb_tex tex);

/* On-disk fixtures -> device.  The reference reads meshes through Assimp
 * (resources/loaders/mesh_loader_assimp.hpp:42-101) and textures through SDL_image (resources/loaders/texture_loader_sdl.hpp:21-56);
 * these two calls read the formats its fixtures use -- Wavefront OBJ (cpp-folders/src/assets/obj/...) and PNG -- with the same
 * conventions (one vertex per distinct v/vt/vn triple, file face order, fan triangulation, default normal (0,1,0) / uv (0,0);
 * RGBA8 with the vertical flip load_texture2d_sdl_image applies by default) and upload the result.  Other formats:
 * SHSB_E_UNSUPPORTED; unreadable / malformed files: SHSB_E_INVALID_ARGUMENT with the reason in shsb_last_error_string. */
SHSB_API int32_t shsb_mesh_load_obj(shsb_ctx ctx, const char* path, shsb_mesh* out_mesh);
SHSB_API int32_t shsb_texture_load_png(shsb_ctx ctx, const char* path, int32_t flip_y, shsb_tex* out_tex);
/* Sizes of a mesh (positions, normals, uvs, indices) / a texture (w, h), and their contents back on the host. */
SHSB_API int32_t shsb_mesh_info(shsb_ctx ctx, shsb_mesh mesh, uint32_t out_counts4[4]);
SHSB_API int32_t shsb_mesh_download(shsb_ctx ctx, shsb_mesh mesh, float* positions, float* normals, float* uvs, uint32_t* indices);
SHSB_API int32_t shsb_texture_info(shsb_ctx ctx, shsb_tex tex, int32_t out_wh2[2]);
SHSB_API int32_t shsb_texture_download(shsb_ctx ctx, shsb_tex tex, uint8_t* rgba, size_t bytes);

/* ------------------------------------------------------------------ render targets */

/* RT_ColorHDR / RT_ColorLDR / RT_ColorDepthVelocity / RT_ShadowDepth constructors
 * (gfx/rt_types.hpp:61-157, gfx/rt_shadow.hpp:18-40).  Storage is device-resident; initial
 * contents equal the reference constructors' (colour 0,0,0,1; depth 1.0; motion 0). */
SHSB_API int32_t shsb_rt_create(shsb_ctx ctx, int32_t kind, int32_t w, int32_t h, float zn, float zf, shsb_rt* out_rt);
SHSB_API int32_t shsb_rt_destroy(shsb_ctx ctx, shsb_rt rt);
/* PixelBuffer2D::clear (rt_types.hpp:53-56): `value` points at one pixel of the plane's type. */
SHSB_API int32_t shsb_rt_clear(shsb_ctx ctx, shsb_rt rt, int32_t plane, const void* value);
SHSB_API int32_t shsb_rt_upload(shsb_ctx ctx, shsb_rt rt, int32_t plane, const void* src, size_t bytes);
SHSB_API int32_t shsb_rt_download(shsb_ctx ctx, shsb_rt rt, int32_t plane, void* dst, size_t bytes);
/* Asynchronous download into PINNED host memory on the context's copy stream: the copy starts when everything
 * submitted so far has finished and overlaps with later submissions; a later pass that writes `rt` waits for
 * it.  The data is valid after shsb_sync(). */
SHSB_API int32_t shsb_rt_download_async(shsb_ctx ctx, shsb_rt rt, int32_t plane, void* dst_pinned, size_t bytes);
/* Raw device pointer of a plane (for NCCL frame gather through torch.distributed). */
SHSB_API int32_t shsb_rt_device_ptr(shsb_ctx ctx, shsb_rt rt, int32_t plane, void** out_ptr, size_t* out_bytes);

/* ------------------------------------------------------------------ sort-first frame assembly (multi-GPU, one process per GPU) */

/* The reference renders one frame on one CPU; it has no counterpart of this section (screen tiles / cameras are
 * partitioned over the GPUs of one box, scene replicated, and the only exchange step is the assembly of disjoint pixels on a
 * root GPU).  The assembly is a PUSH: every rank copies the planes (or row bands) it rendered straight into the root GPU's
 * assembly memory with copy-engine peer writes over NVLink -- no SMs besides one-thread flag kernels, no collective library,
 * no host synchronisation.  Memory is shared through CUDA IPC handles, which the caller ships to the other processes over
 * whatever host channel it has (MPI, torch.distributed, a pipe); ranks living in the root's own process use the raw pointers.
 *
 * Steps are numbered 1, 2, 3, ...; step s uses slot (s - 1) % slots of the assembly memory, so `slots` steps may be in flight.
 *   every rank : shsb_frame_gather(step, ...) once per plane / band, then shsb_gather_commit(step)
 *   root       : shsb_gather_wait(step) -> read the slot (shsb_gather_device_ptr / shsb_gather_download) -> shsb_gather_release(step)
 * All of it is stream-ordered on the context's gather stream: a push starts when the frame that wrote the render target has
 * finished, a later frame into that target waits for the push, a rank reusing a slot waits (on the device, bounded by a
 * time-out that surfaces as SHSB_E_TIMEOUT) for the root's release of the step that used it before. */
typedef uint32_t shsb_gather;
#define SHSB_GATHER_MAX_RANKS 16
typedef struct ShsbGatherExport
{
    unsigned char mem_handle[64]; /* cudaIpcMemHandle_t of the assembly memory (slots x slot_bytes) */
    unsigned char ctl_handle[64]; /* cudaIpcMemHandle_t of the control block (step counters)        */
    uint64_t mem_ptr, ctl_ptr;    /* the same two allocations as raw device pointers (same-process ranks) */
    uint64_t slot_bytes;
    uint32_t n_ranks, slots;
    int32_t root_device;
    int32_t root_pid;
} ShsbGatherExport;
/* Root (rank 0): allocates the assembly memory and fills `out_export` for the other ranks. */
SHSB_API int32_t shsb_gather_create(shsb_ctx ctx, uint32_t n_ranks, uint32_t slots, size_t slot_bytes, shsb_gather* out_gather, ShsbGatherExport* out_export);
/* Rank 1 .. n_ranks-1: maps the root's memory (peer access over NVLink is enabled on first use). */
SHSB_API int32_t shsb_gather_open(shsb_ctx ctx, const ShsbGatherExport* exp, uint32_t rank, shsb_gather* out_gather);
SHSB_API int32_t shsb_gather_destroy(shsb_ctx ctx, shsb_gather gather);
/* Pushes `bytes` of a plane of `rt`, from byte `src_offset` of the plane, to byte `dst_offset` of the step's slot.  Rows are
 * contiguous (row-major targets), so a screen band is one call; a camera of a batch is one call with src_offset 0. */
SHSB_API int32_t shsb_frame_gather(shsb_ctx ctx, shsb_gather gather, uint64_t step, shsb_rt rt, int32_t plane, size_t src_offset, size_t bytes, size_t dst_offset);
/* This rank has enqueued all its pushes of `step`. */
SHSB_API int32_t shsb_gather_commit(shsb_ctx ctx, shsb_gather gather, uint64_t step);
/* Root: orders the gather stream behind every rank's commit of `step` (asynchronous). */
SHSB_API int32_t shsb_gather_wait(shsb_ctx ctx, shsb_gather gather, uint64_t step);
/* Root: the slot of `step` may be overwritten (asynchronous, ordered behind whatever the root enqueued to read it). */
SHSB_API int32_t shsb_gather_release(shsb_ctx ctx, shsb_gather gather, uint64_t step);
/* Root: device pointer of the slot of `step`; synchronous / asynchronous (pinned destination, gather stream) read of it. */
SHSB_API int32_t shsb_gather_device_ptr(shsb_ctx ctx, shsb_gather gather, uint64_t step, void** out_ptr);
SHSB_API int32_t shsb_gather_download(shsb_ctx ctx, shsb_gather gather, uint64_t step, size_t offset, void* dst, size_t bytes);
SHSB_API int32_t shsb_gather_download_async(shsb_ctx ctx, shsb_gather gather, uint64_t step, size_t offset, void* dst_pinned, size_t bytes);
/* The context's gather stream (cudaStream_t as void*), e.g. to record timing events behind a wait. */
SHSB_API int32_t shsb_gather_stream(shsb_ctx ctx, void** out_stream);

/* ------------------------------------------------------------------ host helpers (bit-exact restatements) */

/* model = T * Rx * Ry * Rz * S via successive glm::translate/rotate/scale
 * (passes/pass_pbr_forward.hpp:136-141, passes/pass_shadow_map.hpp:57-65). */
SHSB_API int32_t shsb_model_from_transform(const ShsbTransform* tr, float out_model[16]);
/* viewproj = perspective_lh_no(fovy, aspect, zn, zf) * look_at_lh(eye, target, up)
 * (camera/convention.hpp:19-27; every reference demo builds Camera::viewproj this way). */
SHSB_API int32_t shsb_camera_viewproj(const float eye[3], const float target[3], const float up[3],
                                      float fovy_radians, float aspect, float znear, float zfar,
                                      float out_viewproj[16]);

/* ------------------------------------------------------------------ the hot path */

/* rasterize_mesh (sw_render/rasterizer.hpp:181-442): one mesh draw with a builtin program.
 * depth_motion_rt == 0 reproduces the "no depth target" painter behaviour (rasterizer.hpp:348). */
SHSB_API int32_t shsb_rasterize_mesh(shsb_ctx ctx, shsb_mesh mesh, int32_t shader_id,
                                     const ShsbUniforms* uniforms,
                                     shsb_rt hdr_rt, shsb_rt depth_motion_rt,
                                     const ShsbRasterCfg* cfg, ShsbStats* out_stats);

/* PassPBRForward::execute (passes/pass_pbr_forward.hpp:49-214): background gradient fill, depth
 * clear policy, per-item model matrix + uniforms, all draws in item order.  With
 * fp->light_culling != 0 the fragment program additionally walks the tile light lists built by
 * the last shsb_light_cull (Forward+). */
SHSB_API int32_t shsb_pass_pbr_forward(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp,
                                       shsb_rt hdr_rt, shsb_rt depth_motion_rt, shsb_rt shadow_rt,
                                       const float* shadow_light_viewproj, /* ctx.shadow.light_viewproj or NULL */
                                       int32_t preserve_existing_depth, ShsbStats* out_stats);

/* PassDepthPrepassAdapter (pipeline/pass_adapters.hpp:401-528): depth-only draw of all items. */
SHSB_API int32_t shsb_pass_depth_prepass(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp,
                                         shsb_rt depth_motion_rt, ShsbStats* out_stats);

/* PassShadowMap::execute (passes/pass_shadow_map.hpp:44-205): scene AABB -> texel-snapped ortho
 * light camera (camera/light_camera.hpp:33-99) -> depth-only, unclipped, uncull'd raster.
 * out_light_viewproj receives ctx.shadow.light_viewproj. */
SHSB_API int32_t shsb_pass_shadow_map(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp,
                                      shsb_rt shadow_rt, float out_light_viewproj[16]);

/* Context::history.reset() (core/context.hpp:84-94): forget the previous frame's model matrices, so that the next
 * PassPBRForward computes motion against the current matrices (zero object motion) like a first frame. */
SHSB_API int32_t shsb_history_reset(shsb_ctx ctx);

/* PassTonemap::execute (passes/pass_tonemap.hpp:37-84). */
SHSB_API int32_t shsb_pass_tonemap(shsb_ctx ctx, shsb_rt hdr_rt, shsb_rt ldr_rt, float exposure, float gamma);

/* ---- post passes that consume the path's outputs; RGBA8 in, RGBA8 out, bit-exact.
 * Targets of different sizes are cropped to the common minimum like the reference does. */

/* PassMotionBlur::execute (passes/pass_motion_blur.hpp:40-184): velocity-directed gather over the LDR frame with a
 * per-tap depth rejection; depth_motion_rt supplies RT_ColorDepthMotion::motion and ::depth.  input == output is
 * allowed (the reference's scratch path, :72-76,166-183). */
SHSB_API int32_t shsb_pass_motion_blur(shsb_ctx ctx, const ShsbMotionBlurParams* params, shsb_rt input_ldr, shsb_rt output_ldr,
                                       shsb_rt depth_motion_rt);

/* PassLightShafts::execute (passes/pass_light_shafts.hpp:43-214): screen-space march towards the projected sun over
 * the frame's luma, weighted by depth (depth_like_rt, 0 = none).  input == output is allowed. */
SHSB_API int32_t shsb_pass_light_shafts(shsb_ctx ctx, const ShsbLightShaftsParams* params, shsb_rt input_ldr, shsb_rt output_ldr,
                                        shsb_rt depth_like_rt);

/* PassTemporalAAAdapter::execute_resolved (pipeline/pass_adapters.hpp:1438-1491): blends the frame with the context's
 * colour history (Context::temporal_aa, core/context.hpp:99-113); the first frame after a reset only seeds it. */
SHSB_API int32_t shsb_pass_taa(shsb_ctx ctx, shsb_rt ldr_rt);
/* TemporalAARuntimeState::reset (core/context.hpp:106-112) == PassTemporalAAAdapter::reset_history. */
SHSB_API int32_t shsb_taa_reset(shsb_ctx ctx);

/* Local lights: n records of CullingLightGPU (lighting/light_types.hpp:141-167, 160 B each). */
SHSB_API int32_t shsb_lights_upload(shsb_ctx ctx, const void* records, uint32_t n_lights);

/* cull_lights_tiled (lighting/jolt_light_culling.hpp:135-187): per 2-D screen tile, the ascending
 * list of lights whose cull_sphere / cull_aabb is not Outside the tile's 6-plane cell.  Lists are
 * kept on the device for the Forward+ pass: counts[T] (uncapped) and indices[T*max_per_tile]
 * (first max_per_tile entries, LightCullingRuntimePayload pipeline/render_pass.hpp:32-50). */
SHSB_API int32_t shsb_light_cull(shsb_ctx ctx, const float view_proj[16], uint32_t viewport_w, uint32_t viewport_h,
                                 uint32_t tile_size, uint32_t max_per_tile);
SHSB_API int32_t shsb_light_lists_download(shsb_ctx ctx, uint32_t* counts, size_t n_counts,
                                           uint32_t* indices, size_t n_indices);

/* The other bin builders of lighting/jolt_light_culling.hpp.  Depth-range modes take the
 * per-tile ranges from host memory (range_min / range_max, one float per tile, tile (0,0) = top-left) or, when both are
 * NULL, from the context's last shsb_tile_depth_range result (view depth: use SHSB_LIGHT_CULL_TILED_VIEW_DEPTH).
 * Tiled modes (0-2) produce lists the Forward+ pass can use and shsb_light_lists_download returns; clustered bins
 * (bin = cz * tiles + ty * tiles_x + tx) are returned by shsb_cluster_lists_download. */
SHSB_API int32_t shsb_light_cull_ex(shsb_ctx ctx, const ShsbLightCullDesc* desc, const float* range_min, const float* range_max);
SHSB_API int32_t shsb_cluster_lists_download(shsb_ctx ctx, uint32_t* counts, size_t n_counts, uint32_t* indices, size_t n_indices);

/* Per-tile [min, max] linear view depth of a z-buffer written by the raster path (z01 = (view_z - zn) / (zf - zn),
 * sw_render/rasterizer.hpp:352-354); cleared pixels are skipped, empty tiles get [zn, zf]
 * (build_tile_view_depth_range_from_scene, lighting/light_culling_runtime.hpp:254-261).  The software analogue of
 * shaders/vulkan/fp_stress_depth_reduce.comp.  The result stays on the device for shsb_light_cull_ex. */
SHSB_API int32_t shsb_tile_depth_range(shsb_ctx ctx, shsb_rt depth_motion_rt, uint32_t tile_size);
/* shaders/vulkan/fp_stress_depth_reduce.comp itself (:40-81), for a depth plane that holds the HARDWARE's zero-to-one projective depth
 * of the reference's LH projection, d = f / (f - n) - n f / ((f - n) view_z) -- the depth attachment of the reference's Vulkan path, here
 * uploaded with shsb_rt_upload (the software rasteriser never writes this encoding: its fallback for zf <= zn is the interpolated clip z
 * * 0.5 + 0.5, sw_render/rasterizer.hpp:345-347): view_z = near * far / max(far - clamp(d, 0, 1) * (far - near), 1e-5) with
 * near = max(z_near, 0.001), far = max(z_far, near + 0.01) (depth01_to_view_lh_no, :31-38; z_near / z_far = ubo.depth_params.xy); texels
 * >= 1 are skipped, a tile without any other texel gets (0, 0) as in the shader.  Any target with a depth plane; the result stays on the
 * device like shsb_tile_depth_range's. */
SHSB_API int32_t shsb_tile_depth_range_ndc01(shsb_ctx ctx, shsb_rt depth_rt, uint32_t tile_size, float z_near, float z_far);
SHSB_API int32_t shsb_tile_depth_range_download(shsb_ctx ctx, float* out_min, float* out_max, size_t n_tiles);

/* Fused Forward+ frame = light cull + PassPBRForward (Forward+) + PassTonemap in one submission
 * with HDR, depth and LDR each written once (SURVEY.md section 8d B_frame). */
SHSB_API int32_t shsb_frame_forward_plus(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp,
                                         shsb_rt hdr_rt, shsb_rt depth_motion_rt, shsb_rt ldr_rt,
                                         ShsbStats* out_stats);

/* Last frame's per-stage device times in milliseconds (CUDA events on the context stream; recorded for synchronous
 * submissions -- out_stats != NULL -- and for frames sampled by shsb_timing_enable, zero otherwise):
 * [0] vertex+setup, [1] binning, [2] tile raster+shade, [3] light cull, [4] tonemap, [5] total. */
SHSB_API int32_t shsb_last_stage_ms(shsb_ctx ctx, float out_ms[8]);

/* Per-frame stage timing history for benchmarks: while enabled, frame submissions record five CUDA events on
 * their streams (no host synchronisation).  enable = n > 1 samples every n-th frame only: each event is one more
 * command for the GPU's host interface to fetch, and five per frame lengthen a 135 us frame by ~15 us
 * (profiles/r1_pcie_command_latency_s4.md).  shsb_timing_collect synchronises, writes
 * {vertex+setup, binning, tile raster+shade, total} milliseconds per frame and clears the history. */
SHSB_API int32_t shsb_timing_enable(shsb_ctx ctx, int32_t enable);
SHSB_API int32_t shsb_timing_collect(shsb_ctx ctx, float* out_ms, size_t cap_frames, size_t* out_frames);
/* Same history as absolute times: 5 floats per frame (front-end begin, after geometry, after binning, tile kernel
 * begin, tile kernel end), milliseconds since the first recorded event.  For pipeline timelines (tools/e2e_probe.py). */
SHSB_API int32_t shsb_timing_collect_abs(shsb_ctx ctx, float* out_ms, size_t cap_frames, size_t* out_frames);

/* Accumulated host-side submit cost in microseconds since the last reset: [0] scene -> draw list (model / normal
 * matrices), [1] staging copy, [2] arena checks, [3] stream capture / enqueue, [4] graph update + launch, [5] frames. */
SHSB_API int32_t shsb_host_submit_us(shsb_ctx ctx, double out_us[8], int32_t reset);

/* ------------------------------------------------------------------ legacy tile-job variant
 *
 * BASELINE.json configs[0] "as shipped": the demo cpp-folders/src/hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp with
 * the helpers of cpp-folders/src/hello-shs-renderer/shs_renderer.hpp -- a rasterizer of its own, not the library path: y-flipped
 * screen map, no clipping, `area <= 0` cull, dot-product barycentrics, affine NDC depth tested LESS against FLT_MAX, affine
 * varyings, RGBA8 by truncation (:189-242, shs_renderer.hpp:802-831).  Host helpers restate the demo's matrix set-up. */

typedef struct ShsbLegacyUniforms /* struct Uniforms, hello_pipeline_blinn_phong_shading.cpp:35-41 (+ the job-tile size :28-29) */
{
    float mvp[16];          /* proj * view * model (:273); build it with shsb_legacy_mvp for the reference's product order */
    float model[16];
    float light_dir[3];     /* HelloScene::light_direction (:152): the direction the light travels */
    float camera_pos[3];    /* Viewer::position */
    uint8_t color[4];       /* MonkeyObject::color, RGBA8 */
    int32_t job_tile_w;     /* TILE_SIZE_X / TILE_SIZE_Y (80): the demo's unit of work; 0 = 80.  It shows in the result only for   */
    int32_t job_tile_h;     /* ill-conditioned slivers at job-tile borders (see csrc/legacy.cu), and is honoured exactly.          */
} ShsbLegacyUniforms;

/* Viewer(position, speed, w, h) + Camera3D::update (shs_renderer.hpp:1210-1236, 1322-1346): fov 60 deg, aspect hard-coded 4/3,
 * z 0.1 .. 1000, left-handed; angles in degrees. */
SHSB_API int32_t shsb_legacy_camera(const float position[3], float horizontal_angle_deg, float vertical_angle_deg,
                                    float out_view[16], float out_proj[16]);
/* MonkeyObject::get_world_matrix (:122-128): translate * rotate(y, degrees) * scale, each built from the identity. */
SHSB_API int32_t shsb_legacy_world_matrix(const float position[3], const float scale[3], float rotation_angle_deg, float out_model[16]);
/* uniforms.mvp = proj * view * model (:273), evaluated left to right. */
SHSB_API int32_t shsb_legacy_mvp(const float proj[16], const float view[16], const float model[16], float out_mvp[16]);

/* One object of RendererSystem::process (:244-313) = draw_triangle_tile (:189-242) over every job tile and every triangle of
 * `mesh` (indexed meshes are expanded by their indices like ModelGeometry's loader does, shs_renderer.hpp:1262-1295; the mesh
 * needs a normal per position) with blinn_phong_vertex_shader / blinn_phong_fragment_shader (:48-96).
 * canvas_ldr: SHSB_RT_COLOR_LDR in shs::Canvas order (row 0 = bottom of the screen; uncovered pixels keep their content).
 * zbuffer:    the depth plane of a SHSB_RT_SHADOW or SHSB_RT_DEPTH_MOTION target in shs::ZBuffer order (row = screen y, top
 *             down), same size as the canvas; clear it to FLT_MAX with shsb_rt_clear like ZBuffer::clear (:671-674). */
SHSB_API int32_t shsb_legacy_draw_blinn_phong(shsb_ctx ctx, shsb_mesh mesh, const ShsbLegacyUniforms* uniforms,
                                              shsb_rt canvas_ldr, shsb_rt zbuffer);

/* ------------------------------------------------------------------ legacy render-target demos
 *
 * cpp-folders/src/hello-render-target/hello_shadow_mapping_soft.cpp (config-3 flavour: 2048^2 shadow map + PCSS soft-shadow lit
 * pass) and hello_pbr.cpp (config-4 flavour: Cook-Torrance + IBL lit pass with motion vectors).  Both are rasterizers of their own
 * on top of hello-shs-renderer/shs_renderer.hpp: y-flipped screen map, near-plane Sutherland-Hodgman clip + fan (:869-905), no
 * back-face cull, dot-product barycentrics, depth = affine VIEW-space z tested LESS against FLT_MAX with the z-buffer stored in
 * canvas order (row 0 = bottom), perspective-correct world position / uv, affine normal, RGBA8 by truncation.  One call = one object
 * of one pass = the demo's `for tile: for triangle: draw_triangle_tile_*` loops (soft :1130-1200 / :1223-1360). */

typedef uint32_t shsb_ibl; /* EnvIBL of hello_pbr.cpp:455-461: an irradiance cube + a prefiltered-specular mip chain (shs::CubeMapLinear) */

typedef struct ShsbLegacy2Uniforms /* struct Uniforms, hello_shadow_mapping_soft.cpp:714-732 / hello_pbr.cpp:474-519 (+ MaterialPBR, job tile) */
{
    float mvp[16];
    float prev_mvp[16];        /* pbr only (motion vectors) */
    float model[16];
    float mv[16];              /* view * model: its third row gives the depth that is tested */
    float normal_mat[9];       /* mat3, column-major: the demo sets transpose(inverse(mat3(model))) or the identity */
    float light_vp[16];
    float light_dir_world[3];  /* the direction the light travels */
    float camera_pos[3];
    uint8_t base_color[4];     /* Uniforms::base_color / MaterialPBR::baseColor_srgb */
    int32_t use_texture;
    shsb_tex albedo;           /* 0 = none; sampled with shs::sample_nearest (shs_renderer.hpp:367-377) */
    float metallic, roughness, ao;                                                      /* pbr only */
    float ibl_diffuse_intensity, ibl_specular_intensity, ibl_reflection_strength;       /* pbr only */
    int32_t job_tile_w;        /* TILE_SIZE_X / TILE_SIZE_Y (160); 0 = 160.  Honoured exactly, see csrc/legacy.cu */
    int32_t job_tile_h;
} ShsbLegacy2Uniforms;

/* draw_triangle_tile_shadow over every job tile and triangle of `mesh` (soft :796-839, pbr :827-875; identical) with
 * shadow_vertex_shader (light_vp * model taken first, :779).  shadow_map: SHSB_RT_SHADOW, any size; clear it to FLT_MAX with
 * shsb_rt_clear like ShadowMap::clear.  Rows are light-space screen rows (top down), as ShadowMap stores them. */
SHSB_API int32_t shsb_legacy2_shadow_draw(shsb_ctx ctx, shsb_mesh mesh, const float model[16], const float light_vp[16],
                                          int32_t job_tile_w, int32_t job_tile_h, shsb_rt shadow_map);

/* draw_triangle_tile_color_depth_softshadow (:845-986) with vertex_shader_full (:746-766) and fragment_shader_softshadow (:991-1040;
 * PCSS :333-445).  shadow_map: 0 = Uniforms::shadow == nullptr.  canvas_ldr: SHSB_RT_COLOR_LDR in shs::Canvas order; zbuffer: the
 * depth plane of a SHSB_RT_SHADOW / SHSB_RT_DEPTH_MOTION target of the same size, in CANVAS order (this demo's
 * ZBuffer::test_and_set_depth_screen_space flips the row), cleared to FLT_MAX. */
SHSB_API int32_t shsb_legacy2_draw_softshadow(shsb_ctx ctx, shsb_mesh mesh, const ShsbLegacy2Uniforms* uniforms, shsb_rt shadow_map,
                                              shsb_rt canvas_ldr, shsb_rt zbuffer);

/* EnvIBL as plain data: irradiance = 6 faces x irr_size^2 x RGB float32 (face order +X -X +Y -Y +Z -Z, shs/resources/ibl.hpp:237-262);
 * prefiltered = n_mips such cube maps of sizes spec_sizes[m], concatenated (n_mips <= 16). */
SHSB_API int32_t shsb_legacy3_ibl_upload(shsb_ctx ctx, const float* irradiance, int32_t irr_size, const float* prefiltered,
                                         const int32_t* spec_sizes, int32_t n_mips, shsb_ibl* out_ibl);
SHSB_API int32_t shsb_legacy3_ibl_destroy(shsb_ctx ctx, shsb_ibl ibl);

/* draw_triangle_tile_color_depth_motion (hello_pbr.cpp:883-1045) with vertex_shader_full (:535-556) and fragment_shader_pbr (:627-727;
 * shadow_factor_pcf_2x2 :599-621, cube-map sampling shs/resources/ibl.hpp:215-287).  ibl: 0 = Uniforms::ibl == nullptr.
 * depth_motion: SHSB_RT_DEPTH_MOTION of the canvas size = RT_ColorDepthMotion's depth + velocity planes, both in canvas order;
 * velocity in pixels, +y up, clamped to 22 px (:1018-1030). */
SHSB_API int32_t shsb_legacy3_draw_pbr(shsb_ctx ctx, shsb_mesh mesh, const ShsbLegacy2Uniforms* uniforms, shsb_rt shadow_map, shsb_ibl ibl,
                                       shsb_rt canvas_ldr, shsb_rt depth_motion);

/* ------------------------------------------------------------------ scene-level culling upstream of the path
 *
 * The two data-parallel steps the reference runs on the CPU right before draw submission.  Synchronous host-buffer calls (upload,
 * kernels, download): they sit outside the frame pipeline.  The software-occlusion pass (geometry/culling_software.hpp) is serial
 * in the order of objects only: shsb_software_occlusion walks that order on one persistent CTA. */

/* run_software_occlusion_pass (geometry/culling_software.hpp:253-333) as SceneCullingContext::run_software_occlusion drives it
 * (scene/scene_culling.hpp:186-222): the frustum-visible objects are sorted front to back by the view depth of their AABB centre
 * (std::sort with the reference's comparator, on the host: ties fall as they do there), then, in that order, an object whose
 * projected AABB rectangle is hidden at every texel of the occlusion depth buffer (z_near > depth + epsilon) is occluded, and every
 * other object rasterises its occluder mesh into the buffer (rasterize_mesh_depth_transformed :116-143, minimum depth per texel).
 *   object_aabbs6      n_objects x (min xyz, max xyz), world space          (get_world_aabb)
 *   frustum_visible    CullResult::visible_indices (indices >= n_objects are skipped, :294)
 *   object_mesh        per object: index into mesh_table, or 0xFFFFFFFF for "rasterises nothing" (the demos' guard, e.g.
 *                      exp-plumbing/hello_occlusion_culling_sw.cpp:361-363)
 *   object_models16    per object: the model matrix the occluder mesh is drawn with
 *   mesh_table3        per mesh: first index, index count, base vertex into occluder_indices / occluder_vertices (DebugMesh,
 *                      geometry/jolt_debug_draw.hpp:36-52, concatenated)
 *   enable_occlusion   0: every frustum-visible object is visible, in the given order (:270-284)
 * Outputs: out_occluded[n_objects] (0 / 1; objects outside frustum_visible keep 0), out_visible[<= n_visible] in the pass's order,
 * out_counts4 = CullingStats scene / frustum-visible / visible / occluded counts (culling_runtime.hpp:59-100), out_depth (optional,
 * occ_w x occ_h) = the occlusion depth buffer after the pass. */
SHSB_API int32_t shsb_software_occlusion(shsb_ctx ctx, const float* object_aabbs6, uint32_t n_objects, const uint32_t* frustum_visible, uint32_t n_visible,
                                         const uint32_t* object_mesh, const float* object_models16, const uint32_t* mesh_table3, uint32_t n_meshes,
                                         const float* occluder_vertices, uint32_t n_vertices, const uint32_t* occluder_indices, uint32_t n_indices,
                                         const float view[16], const float view_proj[16], int32_t occ_w, int32_t occ_h, float depth_epsilon, int32_t enable_occlusion,
                                         uint8_t* out_occluded, uint32_t* out_visible, uint32_t out_counts4[4], float* out_depth);

/* cull_vs_frustum over FastCullable + HasWorldAABB objects (geometry/jolt_culling.hpp:279-306; classify_vs_frustum :258-275) against
 * extract_frustum_planes(view_proj) (geometry/frustum_culling.hpp:47-65).
 * bounds10: n x (bounding-sphere centre xyz, radius, world AABB min xyz, max xyz) = SceneShape::bounding_sphere / world_aabb
 *           (geometry/scene_shape.hpp:56-81; the Jolt shape -> bounds step stays with the caller).
 * out_classes[n]: CullClass (0 outside, 1 intersecting, 2 inside); out_visible[<= n]: CullResult::visible_indices (ascending);
 * out_counts5: tested, outside, intersecting, inside, visible. */
SHSB_API int32_t shsb_cull_objects_frustum(shsb_ctx ctx, const float* bounds10, uint32_t n_objects, const float view_proj[16],
                                           uint8_t* out_classes, uint32_t* out_visible, uint32_t out_counts5[5]);

enum ShsbLightObjectCullMode { SHSB_LIGHT_OBJECT_CULL_NONE = 0, SHSB_LIGHT_OBJECT_CULL_SPHERE_AABB = 1, SHSB_LIGHT_OBJECT_CULL_VOLUME_AABB = 2 }; /* lighting/light_runtime.hpp:24-29 */

/* collect_object_lights (lighting/light_runtime.hpp:592-616) for n_objects objects at once: per object the up-to-8 nearest lights
 * (kLightSelectionCapacity) among those of `visible_lights` (visited in order; entries >= n_lights are skipped) that affect its
 * AABB under `cull_mode`, in the reference's slot order (add_light_candidate :263-289).
 * records160: CullingLightGPU records (LightInstance::packed; position_range.xyz is the light position).
 * out_counts[n_objects], out_indices8[n_objects * 8], out_dist2_8[n_objects * 8] (unused slots zero). */
SHSB_API int32_t shsb_collect_object_lights(shsb_ctx ctx, const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible_lights,
                                            uint32_t n_visible, const void* records160, uint32_t n_lights, int32_t cull_mode,
                                            uint32_t* out_counts, uint32_t* out_indices8, float* out_dist2_8);

/* build_tile_view_depth_range_from_scene (lighting/light_culling_runtime.hpp:188-264; project_aabb_bounds :92-153): per-tile
 * [min, max] view depth from the world AABBs of the visible objects -- the reference's CPU source of the ranges that
 * cull_lights_tiled_view_depth_range consumes.  visible_objects: scene indices (entries >= n_objects are skipped); view / view_proj:
 * the camera matrices.  The result stays on the device for shsb_light_cull_ex (range_min = range_max = NULL,
 * SHSB_LIGHT_CULL_TILED_VIEW_DEPTH) and is returned by shsb_tile_depth_range_download, like shsb_tile_depth_range's. */
SHSB_API int32_t shsb_tile_depth_range_from_scene(shsb_ctx ctx, const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible_objects,
                                                  uint32_t n_visible, const float view[16], const float view_proj[16], uint32_t viewport_w,
                                                  uint32_t viewport_h, uint32_t tile_size, float z_near, float z_far);

/* gather_light_scene_candidates_for_aabb (lighting/light_culling_runtime.hpp:373-449) followed by collect_object_lights, per object
 * -- the chain of exp-plumbing/hello_light_types_culling_sw.cpp:968-996 -- over the light bins the context built last:
 * clustered = 0: the tile lists of shsb_light_cull / shsb_light_cull_ex (tiled modes); clustered = 1: the cluster bins of
 * shsb_light_cull_ex(SHSB_LIGHT_CULL_CLUSTERED).  view / view_proj: the camera the bins were built for; z_near / z_far:
 * LightBinCullingConfig's.  records160 / n_lights: the records that were uploaded with shsb_lights_upload (the visible-light list
 * is the identity).  A bin contributes its first max_per_bin entries: build the bins with a cap no smaller than the largest count
 * to match the reference's uncapped lists.  out_candidates[n_objects]: the gathered list's length per object (the demo's
 * light_candidates statistic); the other outputs as shsb_collect_object_lights. */
SHSB_API int32_t shsb_select_object_lights_from_bins(shsb_ctx ctx, const float* object_aabbs6, uint32_t n_objects, const float view[16],
                                                     const float view_proj[16], int32_t clustered, float z_near, float z_far,
                                                     const void* records160, uint32_t n_lights, int32_t cull_mode, uint32_t* out_counts,
                                                     uint32_t* out_indices8, float* out_dist2_8, uint32_t* out_candidates);

/* ------------------------------------------------------------------ flat-shaded mesh draws: the consumer of the light selections
 *
 * The reference's software draws of a DebugMesh (geometry/jolt_debug_draw.hpp:36-52; vertices + indices) with ONE colour per
 * triangle, depth-tested into an RT_ColorLDR and a float depth buffer (sw_render/debug_draw.hpp:64-112: a fragment replaces the texel
 * when its depth is strictly below the stored one; depths outside [0, 1] are dropped).  The device takes a whole BATCH of draws per
 * call -- the demos issue one per visible object -- and resolves every texel to the minimum of (depth, draw order, triangle order),
 * which is what the reference's serial loop leaves.  Stream-ordered on the context's main stream like every other target writer;
 * the call returns after the set-up pass (one 4-byte read), the raster and resolve kernels are still in flight.
 *   canvas_ldr   SHSB_RT_COLOR_LDR, texel (x, y) = RT_ColorLDR::set_rgba(x, y): canvas_w / canvas_h of the reference are its size
 *   depth        depth plane of a SHSB_RT_SHADOW / SHSB_RT_DEPTH_MOTION target of the same size = the std::span<float> depth buffer,
 *                non-negative; clear it to 1.0f with shsb_rt_clear like the demos' std::fill
 * Meshes: shsb_mesh_upload(positions, indices) (no normals / uvs needed); an index beyond the vertices skips its triangle. */
typedef struct ShsbFlatDraw
{
    shsb_mesh mesh;
    uint32_t selection_count;  /* LightSelection::count (lighting/light_runtime.hpp:126-131), <= 8; multi-light draw only */
    float model[16];
    float base_color[3];
    uint32_t selection[8];     /* LightSelection::indices = one row of shsb_collect_object_lights / shsb_select_object_lights_from_bins */
} ShsbFlatDraw;

/* shs::LightProperties (lighting/light_runtime.hpp:52-71) as plain data, plus the LightType (light_types.hpp:24-32) of the
 * LightInstance's model, which selects the ILightModel::sample the draw evaluates: 1 Point, 2 Spot, 3 RectArea, 4 TubeArea (other
 * types contribute nothing: the demo registers no model for them). */
typedef struct ShsbLightProperties
{
    float color[3], intensity;
    float position_ws[3], range;
    float direction_ws[3], inner_angle_rad;
    float right_ws[3], outer_angle_rad;
    float up_ws[3], tube_half_length;
    float rect_half_extents[2], tube_radius, attenuation_power;
    float attenuation_bias, attenuation_cutoff;
    uint32_t attenuation_model;  /* LightAttenuationModel: 0 Linear, 1 Smooth, 2 InverseSquare */
    uint32_t flags;
    uint32_t light_type;
    uint32_t reserved[3];
} ShsbLightProperties; /* 128 bytes */

/* debug_draw::draw_mesh_blinn_phong_transformed (sw_render/debug_draw.hpp:153-203) for n_draws meshes in order: per triangle
 * ambient 0.18 + 0.72 N.L + 0.35 (N.H)^32 against one directional light, face normal cross(p2 - p0, p1 - p0). */
SHSB_API int32_t shsb_flat_draw_blinn_phong(shsb_ctx ctx, const ShsbFlatDraw* draws, uint32_t n_draws, const float view_proj[16], const float camera_pos[3],
                                            const float light_dir_ws[3], shsb_rt canvas_ldr, shsb_rt depth);

/* draw_mesh_multi_light_transformed (exp-plumbing/hello_light_types_culling_sw.cpp:366-422) for n_draws objects in order: per
 * triangle the ambient + hemisphere term, then light.model->sample(props, centroid, n, V) of every light of the draw's selection
 * (entries >= n_lights are skipped, :411): Point / Spot / RectArea / TubeArea models of lighting/light_runtime.hpp:291-520 with
 * eval_local_light_brdf / eval_distance_attenuation (:182-237).  std::pow / std::cos are evaluated in double and rounded once. */
SHSB_API int32_t shsb_flat_draw_multi_light(shsb_ctx ctx, const ShsbFlatDraw* draws, uint32_t n_draws, const float view_proj[16], const float camera_pos[3],
                                            const ShsbLightProperties* lights, uint32_t n_lights, shsb_rt canvas_ldr, shsb_rt depth);

#ifdef __cplusplus
}
#endif

#endif /* SHSB_H */
