// shsb_dev.cuh -- device-side data layout and exact-arithmetic helpers shared by all kernels.
//
// HBM layout (DESIGN.md "Data layout"):
//   DevMesh      SoA vertex streams exactly as MeshData (resources/mesh.hpp:23): float3 / float3 / float2 / u32
//   DevItem      one per draw (RenderItem + resolved material + host-computed model / normal matrix), 192 B
//   RasterRec    one per set-up fan triangle, 64 B = 4 x 16-B loads: everything the per-pixel coverage +
//                depth test needs (rasterizer.hpp:167-179, 336-361)
//   ShadeRec     one per set-up fan triangle, 128 B: attributes pre-multiplied by 1/w, read only for the
//                winning fragment of a pixel (rasterizer.hpp:309-328, 363-387)
//   tile lists   u32 RasterRec indices, CSR over 16x16-px screen tiles (tile rows are TOP-anchored so
//                that a raster tile is exactly a light tile of jolt_light_culling.hpp:103-107)
//
// Exactness: everything that decides coverage, depth, clip topology or light lists must reproduce the
// reference's IEEE-754 binary32, round-to-nearest, NO-FMA arithmetic (x86-64 -O3 without -mfma).  Those
// expressions are written with the xmul/xadd/xsub/xdiv helpers below (__fmul_rn & co are never contracted
// into FMA by nvcc) so they stay exact even in translation units compiled with FMA enabled.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include "legacy2_core.cuh"
#include "scene_cull_core.cuh"
#include "flat_draw_core.cuh"

namespace shsb
{
    constexpr int TILE = 16;           // raster tile WIDTH in pixels == Forward+ light tile of the reference default == unit of the own-row partition
#ifndef SHSB_TILE_H
#define SHSB_TILE_H 8
#endif
    // raster tile HEIGHT.  Measured on B200 (profiles/r2_tile_height_ab.md): 16 x 8 tiles -- two raster tiles per light tile, 4 instead of 8
    // warps waiting for the slowest 8x4 block of their tile -- beat 16 x 16 by 2.6 % (C2 frames/s) to 9-13 % (C3-C5 tile kernel); 16 x 4 loses on C2.
    constexpr int TILE_H = SHSB_TILE_H;
    static_assert(TILE_H == 16 || TILE_H == 8 || TILE_H == 4, "a raster tile is 16 x 16, 16 x 8 or 16 x 4 pixels");
    constexpr int TILE_PIXELS = TILE * TILE_H;
    constexpr uint32_t KEY_NONE = 0u;  // internal draw-order key = public key + 1; 0 = "no fragment yet"

    // ---------------------------------------------------------------- exact binary32 ops (never fused)
    __device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
    __device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
    __device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
    __device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
    __device__ __forceinline__ float xrcp(float a) { return __fdiv_rn(1.0f, a); }
    __device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
    // glm::max / std::max: (a < b) ? b : a      glm::min / std::min: (b < a) ? b : a
    __device__ __forceinline__ float gmax(float a, float b) { return (a < b) ? b : a; }
    __device__ __forceinline__ float gmin(float a, float b) { return (b < a) ? b : a; }
    __device__ __forceinline__ float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }          // glm::clamp
    __device__ __forceinline__ float sclamp(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); } // std::clamp

    struct F3 { float x, y, z; };
    __device__ __forceinline__ float xdot3(F3 a, F3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
    __device__ __forceinline__ F3 xnormalize3(F3 v)
    {
        const float k = xrcp(xsqrt(xdot3(v, v)));
        return F3{xmul(v.x, k), xmul(v.y, k), xmul(v.z, k)};
    }
    // column-major mat4 * (x,y,z,w): (m0*x + m1*y) + (m2*z + m3*w)   [GLM scalar order]
    __device__ __forceinline__ float4 xmat4_mul(const float* __restrict__ m, float x, float y, float z, float w)
    {
        float4 r;
        r.x = xadd(xadd(xmul(m[0], x), xmul(m[4], y)), xadd(xmul(m[8], z), xmul(m[12], w)));
        r.y = xadd(xadd(xmul(m[1], x), xmul(m[5], y)), xadd(xmul(m[9], z), xmul(m[13], w)));
        r.z = xadd(xadd(xmul(m[2], x), xmul(m[6], y)), xadd(xmul(m[10], z), xmul(m[14], w)));
        r.w = xadd(xadd(xmul(m[3], x), xmul(m[7], y)), xadd(xmul(m[11], z), xmul(m[15], w)));
        return r;
    }

    // ---------------------------------------------------------------- device records
    struct DevMesh
    {
        const float* positions; // n_positions * 3
        const float* normals;   // n_normals * 3
        const float* uvs;       // n_uvs * 2
        const uint32_t* indices;
        uint32_t n_positions, n_normals, n_uvs, n_indices;
    };

    struct __align__(16) DevItem
    {
        float model[16];
        float nrm[9];          // transpose(inverse(mat3(model))) or mat3(model), builtin_shaders.hpp:92-95 (host-computed)
        float base_color[3];
        float metallic, roughness, ao;
        uint32_t tex;          // 1-based, 0 = none
        uint32_t mesh;         // index into the context's DevMesh table
        uint32_t tri_offset;   // global index of this draw's triangle 0 (draw-order key = (tri_offset + ti) * 8 + fan)
        uint32_t tri_count;
        uint32_t pad[10];
        float c2p[16];         // curr_to_prev_model = prev_model * inverse(model), host-computed (rasterizer.hpp:296-307)
    };
    static_assert(sizeof(DevItem) == 256, "DevItem is 256 bytes");

    struct __align__(16) RasterRec
    {
        float ax, ay;          // s0 (screen position of corner 0)
        float v0x, v0y;        // s1 - s0
        float v1x, v1y;        // s2 - s0
        float inv_den;         // 1 / (v0x*v1y - v1x*v0y)
        float iw0, iw1, iw2;   // 1 / clip.w per corner
        float zw0, zw1, zw2;   // clip.z * (1/w) per corner   (shadow pass: NDC z per corner)
        uint32_t bbox_x;       // minx | maxx << 16   (clamped bbox, rasterizer.hpp:285-289)
        uint32_t bbox_y;       // miny | maxy << 16
        uint32_t key;          // draw-order key + 1
    };
    static_assert(sizeof(RasterRec) == 64, "RasterRec is 64 bytes");

    struct __align__(16) ShadeRec
    {
        float wp[3][3];        // world_pos * (1/w) per corner     (semantic varying 0)
        float n[3][3];         // normal_ws * (1/w) per corner     (semantic varying 1)
        float uv[3][2];        // uv * (1/w) per corner            (semantic varying 2)
        uint32_t item;
        uint32_t pad[7];
    };
    static_assert(sizeof(ShadeRec) == 128, "ShadeRec is 128 bytes");

    // Device mirror of ShsbStats.  Counters are SHARDED: a CTA adds to shard (blockIdx.x % STAT_SHARDS), each shard on
    // its own 128-byte line, so that millions of triangles do not serialise on one L2 atomic unit; the host sums.
    constexpr int STAT_SHARDS = 32;
    struct __align__(128) DevStats
    {
        unsigned long long tri_input, tri_after_clip, tri_raster, frag_covered, frag_shaded;
        unsigned int overflow_recs, overflow_lists, overflow_clipq, pad;
    };
    static_assert(sizeof(DevStats) == 128, "one stats shard per 128-byte line");

    struct DevLightRec // CullingLightGPU, lighting/light_types.hpp:141-167
    {
        float position_range[4], color_intensity[4], direction_spot[4], axis_spot_outer[4], up_shape_x[4], shape_attenuation[4];
        uint32_t type_shape_flags[4];
        float cull_sphere[4], cull_aabb_min[4], cull_aabb_max[4];
    };
    static_assert(sizeof(DevLightRec) == 160, "CullingLightGPU is 160 bytes");

    // Compact form of a local light (80 B) for the tile kernel's light loop, digested from the 160-byte record once
    // per upload: pos_r2 decides the range test, the rest is only read in range.
    struct SmLight
    {
        float4 pos_r2;        // xyz, w = range^2 (area lights: cull-sphere centre / radius^2; disabled lights: -1)
        float4 radiance_ir;   // color * intensity, w = 1 / range
        float4 atten;         // x = attenuation power, y = max(cutoff, 0), z = max(bias, 1e-5), w = 1 / max(inner_cos - outer_cos, 1e-6)
        float4 dir_outer;     // normalised spot direction, w = outer cos (clamped)
        uint32_t kind;        // bits 0-1 attenuation model, bit 2 spot, bit 3 area light (rect / tube -> eval_light_record), bit 4 power != 1
        uint32_t index;       // light index (area lights re-read their 160-B record)
        uint32_t pad0, pad1;
    };
    static_assert(sizeof(SmLight) == 80, "SmLight is 80 bytes");
    constexpr uint32_t KIND_SPOT = 4u, KIND_AREA = 8u, KIND_POW = 16u;

    // Everything a frame's kernels need, passed by value as a kernel parameter.
    struct FrameConst
    {
        float viewproj[16];
        float light_viewproj[16];
        float camera_pos[3];
        float sun_intensity;
        float sun_dir[3];
        float zn;
        float sun_color[3];
        float zf;
        int W, H;
        int tiles_x, tiles_y;
        int shader_id;
        int cull_mode;
        int front_face_ccw;
        int has_depth;          // depth target bound (rasterizer.hpp:348)
        int linear_depth;       // zf > zn + 1e-6 (rasterizer.hpp:354)
        int load_depth;         // 1: initial depth comes from the RT (preserve_existing_depth / separate draws); 0: clear to 1.0 in-tile
        int load_color;         // 1: keep RT colour where nothing is drawn; 0: write the background gradient (pass_pbr_forward.hpp:69-85)
        int write_aovs;
        int shadow_mode;        // PassShadowMap raster rules (pass_shadow_map.hpp:144-203)
        float* direct_depth;    // shadow mode: the depth plane, pre-filled with 1.0, that the set-up kernel rasterises small and medium
                                // triangles into directly (atomic minimum); null = every triangle goes through the binned tile path
        // shadow sampling (shadow_sample.hpp)
        const float* shadow_map;
        int shadow_w, shadow_h;
        float bias_const, bias_slope, pcf_step, shadow_strength;
        int pcf_radius;
        // Forward+
        int forward_plus;
        const DevLightRec* lights;
        const SmLight* sm_lights;   // digested copies of `lights`
        uint32_t n_lights;
        int area_lights;        // the uploaded set may hold rect / tube lights (unknown for device-resident records: assumed)
        const uint32_t* tile_counts;
        const uint32_t* tile_indices;
        uint32_t light_tile_size, max_per_tile, light_tiles_x, light_tiles_y;
        // fused tonemap (PassTonemap, pass_tonemap.hpp:49-81)
        int fuse_tonemap;
        float exposure, inv_gamma;
        // motion vectors (rasterizer.hpp:295-307, 388-411)
        int write_motion;       // winning fragments write their screen-space velocity
        int clear_motion;       // pixels no fragment reached get (0, 0): PassPBRForward clears the plane (pass_pbr_forward.hpp:87-98)
        float prev_viewproj[16];
        // sky background (Scene::sky, sky/skybox_renderer.hpp:25-57)
        int sky_kind;           // SHSB_SKY_*
        float inv_viewproj[16]; // host-computed glm::inverse(cam.viewproj)
        float sky_sun[3];       // normalised ProceduralSky sun direction
        float sky_intensity;
        uint32_t sky_faces[6];  // indices into the texture table (cubemap), all valid when sky_kind == cubemap
        // sort-first screen partition: the tile rows this submission owns (ShsbFrameParams::own_row_*); count 0 = all rows
        int own_first, own_count, own_stride;
        // hierarchical-Z early reject in the tile kernel (per 8x4 block: skip a staged triangle whose conservative nearest depth is
        // behind everything the block holds).  Only for submissions nobody reads fragment counters / AOVs of: a skipped triangle's
        // hidden fragments are not counted.
        int hiz;
    };

    // ty: RASTER tile row; the partition is stated in rows of TILE (16) pixels (ShsbFrameParams::own_row_*)
    __host__ __device__ __forceinline__ bool owned_row(const FrameConst& fc, int ty)
    {
        const int row = ty * TILE_H / TILE;
        return fc.own_count <= 0 || (row >= fc.own_first && ((row - fc.own_first) % fc.own_stride) < fc.own_count);
    }

    struct FrameBuffers
    {
        float4* hdr;              // W*H RGBA32F or null (depth-only)
        float* depth;             // W*H or null
        uchar4* ldr;              // W*H or null
        uint32_t* aov_tri_id;     // W*H or null
        uint32_t* aov_coverage;   // W*H or null
        float2* motion;           // W*H or null
    };

    struct Geometry // per-frame transient arena
    {
        const DevMesh* meshes;
        const DevItem* items;
        const uint2* block_table;   // per geometry CTA: (item index, first triangle)
        uint32_t n_blocks;
        RasterRec* rrecs;
        ShadeRec* srecs;
        uint32_t rec_capacity;
        uint32_t* rec_count;
        uint2* clip_queue;          // (item, triangle) pairs that need frustum clipping
        uint32_t clipq_capacity;
        uint32_t* clipq_count;
        uint32_t* tile_count;       // per tile: number of list entries (atomically counted by the geometry kernels)
        uint32_t* tile_offset;      // per tile: start of its segment in tile_list (allocated by atomicAdd on list_cursor)
        uint32_t* tile_fill;        // per tile write cursor
        uint32_t* tile_list;        // RasterRec indices
        uint32_t list_capacity;
        uint32_t* list_cursor;      // total entries allocated so far
        uint32_t* class_count;      // [4] tiles per scheduling class (heaviest first)
        uint32_t* tile_order;       // [4][n_tiles] tiles per class as (tx | ty << 16); CTA b of the tile kernel takes the b-th tile in class order
        DevStats* stats;            // [STAT_SHARDS]
        uint32_t* overflow_flag;    // mapped host words, sticky: [0] set by any kernel that had to drop a record / list entry / clip-queue entry,
                                    // [1] / [2] the tile-list entries / records the frame needed (written by the tile kernel when above capacity)
    };

    struct DevTexture
    {
        const uchar4* texels;
        int w, h;
    };

    // ---------------------------------------------------------------- post passes (post_passes.cu)
    struct PostMotionBlur // PassMotionBlur, passes/pass_motion_blur.hpp:40-184 (parameters already clamped on the host, :79-84)
    {
        const uchar4* src; int src_w;
        uchar4* dst; int dst_w;
        const float2* motion; const float* depth; int mot_w;
        int w, h;
        int samples;
        float strength, dt_scale, max_vel, min_vel, depth_eps;
    };

    struct PostLightShafts // PassLightShafts, passes/pass_light_shafts.hpp:43-214 (parameters clamped on the host, :135-138)
    {
        const uchar4* src; int src_w;
        uchar4* dst; int dst_w;
        const float* lumad; // luma * clamp(depth,0,1) plane, w * h
        int w, h;
        int steps;
        float density, weight, decay;
        float sun_u, sun_v;
    };

    // ---------------------------------------------------------------- legacy tile-job variant (legacy.cu)
    struct LegacyDraw // one object of RendererSystem::process, hello_pipeline_blinn_phong_shading.cpp:262-305
    {
        const float* positions;   // MeshData streams of the mesh handle (ModelGeometry::triangles / normals once expanded by `indices`)
        const float* normals;
        const uint32_t* indices;  // nullptr: non-indexed soup
        uint32_t n_positions, n_normals, n_tris;
        int W, H;
        int job_w, job_h;         // TILE_SIZE_X / TILE_SIZE_Y of the demo (80 x 80): see legacy.cu
        float mvp[16], model[16];
        float normal_matrix[9];   // mat3(transpose(inverse(model))), column-major, computed on the host
        float light_dir[3], camera_pos[3];
        unsigned char color[4];
    };

    struct LegacyTri // set-up output, one per source triangle
    {
        uint32_t valid;
        float ax, ay, v0x, v0y, v1x, v1y;   // screen-space A and the edge vectors B - A, C - A
        float d00, d01, d11, denom;         // the P-independent dot products of Canvas::barycentric_coordinate
        float z[3];                         // NDC z per corner (affine depth)
        float minx, maxx, miny, maxy;       // float bounding box of the three screen positions
        float normal[3][3], world[3][3];    // vertex-shader outputs, interpolated affinely
    };

    // ---------------------------------------------------------------- kernel launchers (one per .cu)
    void launch_cull_objects(const float* bounds10, uint32_t n, const float planes24[24], uint8_t* classes, uint32_t* visible, uint32_t* counts5, cudaStream_t s, uint64_t* launches);
    void launch_collect_object_lights(const float* boxes6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const float* records, uint32_t n_lights, int mode,
                                      uint32_t* out_counts, uint32_t* out_idx, float* out_d2, cudaStream_t s, uint64_t* launches);
    void launch_scene_tile_depth_range(const float* boxes6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const float view[16], const float view_proj[16],
                                       float z_near, float z_far, uint32_t tiles_x, uint32_t tiles_y, uint32_t* scratch3, float* out_min, float* out_max, cudaStream_t s, uint64_t* launches);
    void launch_select_object_lights(const float* boxes6, uint32_t n_objects, const float view[16], const float view_proj[16], const sc::BinGrid& grid, const uint32_t* bin_counts,
                                     const uint32_t* bin_indices, const float* records, uint32_t n_lights, int mode, uint32_t* seen, uint32_t words_per_object, uint32_t* out_counts,
                                     uint32_t* out_idx, float* out_d2, uint32_t* out_candidates, cudaStream_t s, uint64_t* launches);
    void launch_software_occlusion(const float* boxes6, uint32_t n_objects, const uint32_t* sorted, uint32_t n_sorted, const uint32_t* object_mesh, const float* object_models,
                                   const uint32_t* mesh_table, uint32_t n_meshes, const float* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices,
                                   const float view_proj[16], int width, int height, float epsilon, float* depth, uint8_t* occluded, uint32_t* visible, uint32_t* counts2,
                                   void* rect_scratch /* n_sorted x 32 bytes */, cudaStream_t s, uint64_t* launches);
    void launch_flat_draw_batch(const fd::BatchDesc& bd, const fd::DrawRec* draws, const fd::LightProps* lights, void* tris, uint32_t* colours, uint2* items, uint32_t item_cap,
                                uint32_t* item_total, unsigned long long* zkey, float* depth, bool init_keys, cudaStream_t s, uint64_t* launches);
    void launch_flat_raster_resolve(const fd::BatchDesc& bd, const void* tris, const uint32_t* colours, const uint2* items, uint32_t n_items, unsigned long long* zkey,
                                    float* depth, uchar4* canvas, cudaStream_t s, uint64_t* launches);
    size_t flat_tri_record_bytes();
    uint32_t legacy2_slots(const l2::Draw& d);
    void launch_legacy2_draw(const l2::Draw& d, l2::RasterRec* rr, l2::BoxRec* bb, l2::ShadeRec* ss, uchar4* canvas, float* zbuf, float2* velocity,
                             cudaStream_t s, uint64_t* launches);
    void launch_legacy_draw(const LegacyDraw& d, LegacyTri* tris, uchar4* canvas, float* zbuf, cudaStream_t s, uint64_t* launches);
    void launch_geometry(const FrameConst& fc, const Geometry& g, cudaStream_t s, uint64_t* launches);
    void launch_binning(const FrameConst& fc, const Geometry& g, cudaStream_t s, uint64_t* launches);
    void launch_tile_raster(const FrameConst& fc, const Geometry& g, const FrameBuffers& fb, const DevTexture* textures,
                            const float* srgb_lut, cudaStream_t s, uint64_t* launches, bool allow_fast = true, int* mode_out = nullptr);
    void launch_light_prep(const DevLightRec* src, DevLightRec* raw, SmLight* out, uint32_t n, cudaStream_t s, uint64_t* launches);
    void launch_upload(void* dst, const void* src_mapped, size_t bytes, cudaStream_t s, uint64_t* launches);
    void launch_tonemap(const float4* hdr, uchar4* ldr, int n_pixels, float exposure, float inv_gamma, cudaStream_t s, uint64_t* launches);
    void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t s, uint64_t* launches);
    void launch_fill_f4(float4* p, float4 v, size_t n, cudaStream_t s, uint64_t* launches);
    // frustum_planes24: the 6 normalised camera-frustum planes (nx, ny, nz, d) computed on the host
    void launch_light_cull(const DevLightRec* lights, uint32_t n_lights, const float* frustum_planes24, const float* inv_view_proj,
                           uint32_t vw, uint32_t vh, uint32_t ts, uint32_t max_per_tile,
                           uint32_t* scratch, uint32_t* counts, uint32_t* indices, cudaStream_t s, uint64_t* launches,
                           int own_first = 0, int own_count = 0, int own_stride = 1); // sort-first: only these tile rows' lists are built
    void launch_motion_blur(const PostMotionBlur& a, cudaStream_t s, uint64_t* launches);
    void launch_light_shafts(const PostLightShafts& a, const float* depth, int depth_w, float* lumad, cudaStream_t s, uint64_t* launches);
    void launch_taa(uchar4* ldr, uchar4* hist, size_t n_pixels, float keep, float blend, cudaStream_t s, uint64_t* launches);
    void launch_copy_ldr(const uchar4* src, int src_w, uchar4* dst, int dst_w, int w, int h, cudaStream_t s, uint64_t* launches);
    // depth-range / clustered modes (jolt_light_culling.hpp:196-412): mode 1 depth01 range, 2 view-depth range, 3 clustered;
    // vis_scratch holds n_lights + 1 words; counts[cells], indices[cells * max_per_bin], cells = tiles * n_slices
    void launch_light_cull_cells(const DevLightRec* lights, uint32_t n_lights, const float* frustum_planes24, const float* inv_view_proj,
                                 uint32_t vw, uint32_t vh, uint32_t ts, uint32_t max_per_bin, int mode, uint32_t n_slices,
                                 const float* range_min, const float* range_max, const float2* slice_ndc, float z_near, float z_far,
                                 uint32_t* vis_scratch, uint32_t* counts, uint32_t* indices, cudaStream_t s, uint64_t* launches);
    void launch_tile_depth_range(const float* depth, int W, int H, uint32_t ts, float zn, float zf, int ndc01, float* out_min, float* out_max, cudaStream_t s, uint64_t* launches);
    size_t light_cull_scratch_words(uint32_t n_lights, uint32_t vw, uint32_t vh, uint32_t ts); // u32 words of `scratch`
    // sort-first frame assembly (gather.cu): publish / await monotonic 64-bit step counters in (peer) device memory
    void launch_gather_signal(unsigned long long* flag, unsigned long long value, cudaStream_t s, uint64_t* launches);
    void launch_gather_wait(const unsigned long long* flags, uint32_t n_flags, unsigned long long value, unsigned long long timeout_ns, uint32_t* timed_out_mapped,
                            cudaStream_t s, uint64_t* launches);
}
