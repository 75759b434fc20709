// oracle/ref_glsl_a9_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Pins row A9 (the Forward+ per-fragment local-light loop).  The reference has that loop only as GLSL
// (shaders/vulkan/fp_stress_scene.frag:132-165, 421-523, 644-678; shaders/vulkan/common/light_math.glsl:44-78).  GLSL function
// bodies are C++ but for literals and swizzles, so oracle/extract_glsl_a9.py lifts the shader's own text into
// oracle/_ref/glsl_a9_generated.inc (lexical rewrites only; git-ignored) and this file compiles it against
// oracle/glsl_shim/glsl.hpp.  Built by `make -C oracle ref` into oracle/_ref/libshs_glsl_a9_ref.so; tests/test_a9_pinned_cpu.py
// holds oracle.cpp's eval_local_light / forward_plus_lights to it.
//
// What the harness itself supplies (everything else is the reference's text):
//   * the shader's interface variables as globals: v_world_pos, gl_FragCoord, light_buffer, tile_counts, tile_indices (ubo is
//     declared by the extracted CameraUBO block);
//   * eval_local_shadow == 1.0: no ShadowLightGPU buffer is bound on this path, which is the value the shader computes for a light
//     without SHS_LIGHT_FLAG_AFFECTS_SHADOWS or with meta.w == 0 (fp_stress_scene.frag:372-377);
//   * gl_FragCoord: Vulkan's window origin is the UPPER-left corner with pixel centres at +0.5, the software rasterizer's rows
//     run bottom-up (rasterizer.hpp:267-269), so pixel (px, py) of an H-row target is gl_FragCoord = (px + 0.5, (H - 1 - py) + 0.5).
#include <cstring>
#include <vector>

#include "glsl_shim/glsl.hpp"

namespace shs_glsl
{
    using namespace glsl;

    vec3 v_world_pos;
    vec4 gl_FragCoord;
    const uint* tile_counts = nullptr;
    const uint* tile_indices = nullptr;
    struct CullingLightGPU;
    struct LightBuffer { const CullingLightGPU* lights; } light_buffer{nullptr};

#include "_ref/glsl_a9_generated.inc"

    float eval_local_shadow(uint, vec3, vec3) { return 1.0f; }

    static_assert(sizeof(CullingLightGPU) == 160, "CullingLightGPU is 160 bytes (std430)");
}

extern "C" {

// eval_local_light(idx, ...) of the shader for one surface point; technique: 0 = PBR, 1 = Blinn (LIGHT_TECH_*)
void shsglsl_eval_local_light(const void* records160, uint32_t idx, const float P[3], const float N[3], const float V[3], const float albedo[3],
                              float metallic, float roughness, uint32_t technique, float out3[3])
{
    using namespace shs_glsl;
    light_buffer.lights = static_cast<const CullingLightGPU*>(records160);
    v_world_pos = vec3(P[0], P[1], P[2]);
    const vec3 r = eval_local_light(idx, vec3(N[0], N[1], N[2]), vec3(V[0], V[1], V[2]), vec3(albedo[0], albedo[1], albedo[2]), metallic, roughness, technique);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

// shs_eval_light_attenuation_quadratic, common/light_math.glsl:44-78
float shsglsl_attenuation_quadratic(float distance, float range, uint32_t model, float power, float bias, float cutoff)
{
    return shs_glsl::shs_eval_light_attenuation_quadratic(distance, range, model, power, bias, cutoff);
}

// The list walk of main() (fp_stress_scene.frag:644-678) for the fragment at framebuffer pixel (px, py) of a W x H target:
// culling_mode 0 = every light, 1 = tile lists (saturated lists fall back to every light), 3 = cluster bins.
void shsglsl_local_light_loop(const void* records160, uint32_t n_lights, const uint32_t* counts, const uint32_t* indices, uint32_t tiles_x, uint32_t tiles_y,
                              uint32_t max_per_tile, uint32_t tile_size, uint32_t culling_mode, uint32_t z_slices, const float view[16], float z_near, float z_far,
                              int32_t px, int32_t py, int32_t H, const float P[3], const float N[3], const float V[3], const float albedo[3], float metallic, float roughness,
                              uint32_t technique, float out3[3])
{
    using namespace shs_glsl;
    light_buffer.lights = static_cast<const CullingLightGPU*>(records160);
    tile_counts = counts;
    tile_indices = indices;
    ubo = CameraUBO{};
    if (view) std::memcpy(static_cast<void*>(&ubo.view), view, 64);
    ubo.screen_tile_lightcount.z = tiles_x; ubo.screen_tile_lightcount.w = n_lights;
    ubo.params.x = tiles_y; ubo.params.y = max_per_tile; ubo.params.z = tile_size; ubo.params.w = culling_mode;
    ubo.culling_params.x = z_slices;
    ubo.depth_params.x = z_near; ubo.depth_params.y = z_far;
    gl_FragCoord = vec4((float)px + 0.5f, (float)(H - 1 - py) + 0.5f, 0.0f, 1.0f);
    v_world_pos = vec3(P[0], P[1], P[2]);
    const vec3 r = shs_a9_local_light_loop(vec3(N[0], N[1], N[2]), vec3(V[0], V[1], V[2]), vec3(albedo[0], albedo[1], albedo[2]), metallic, roughness, technique);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------------------
// The per-tile depth reduce: shaders/vulkan/fp_stress_depth_reduce.comp, lifted the same way into
// oracle/_ref/glsl_depth_reduce_generated.inc (its CameraUBO block, depth01_to_view_lh_no, main() renamed shs_depth_reduce_main).
// The harness supplies the shader's interface: gl_GlobalInvocationID, the tile_depth_ranges buffer, and texelFetch on depth_tex --
// Vulkan image rows run top-down, the software z-buffer's rows bottom-up (gfx/rt_types.hpp:35-59), so texel (px, py) is
// depth[(H - 1 - py) * W + px], the same flip the light tiles use (jolt_light_culling.hpp:105-107).
namespace shs_glsl_reduce
{
    using namespace glsl;
    uvec3 gl_GlobalInvocationID;
    vec2* tile_depth_ranges = nullptr;
    struct DepthTex { const float* texels; int w, h; } depth_tex{nullptr, 0, 0};
    inline vec4 texelFetch(const DepthTex& t, ivec2 p, int) { return vec4(t.texels[(size_t)(t.h - 1 - p.y) * (size_t)t.w + (size_t)p.x], 0.0f, 0.0f, 1.0f); }

#include "_ref/glsl_depth_reduce_generated.inc"
}

extern "C" int32_t shsglsl_tile_depth_range_ndc01(const float* depth, int32_t w, int32_t h, uint32_t ts, float z_near, float z_far, float* out_min, float* out_max)
{
    using namespace shs_glsl_reduce;
    if (!depth || !out_min || !out_max || w <= 0 || h <= 0 || ts == 0) return 1;
    const uint32_t tiles_x = ((uint32_t)w + ts - 1) / ts, tiles_y = ((uint32_t)h + ts - 1) / ts;
    ubo = CameraUBO{};
    ubo.screen_tile_lightcount.x = (uint32_t)w; ubo.screen_tile_lightcount.y = (uint32_t)h; ubo.screen_tile_lightcount.z = tiles_x;
    ubo.params.x = tiles_y; ubo.params.z = ts;
    ubo.depth_params.x = z_near; ubo.depth_params.y = z_far;
    depth_tex = DepthTex{depth, w, h};
    std::vector<vec2> ranges((size_t)tiles_x * tiles_y);
    tile_depth_ranges = ranges.data();
    for (uint32_t ty = 0; ty < tiles_y; ++ty) // the dispatch: one invocation per tile (local_size 1 x 1 x 1)
        for (uint32_t tx = 0; tx < tiles_x; ++tx)
        {
            gl_GlobalInvocationID = uvec3{tx, ty, 0u};
            shs_depth_reduce_main();
        }
    for (size_t i = 0; i < ranges.size(); ++i) { out_min[i] = ranges[i].x; out_max[i] = ranges[i].y; }
    return 0;
}
