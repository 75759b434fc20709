// tile_raster.cu -- K3: per-tile scan conversion, depth test, deferred fragment shading, resolve.
//
// Replaces raster_rows (sw_render/rasterizer.hpp:330-422), the fragment programs it calls
// (shader/builtin_shaders.hpp:105-245), PCF lookup (lighting/shadow_sample.hpp:65-104), the background
// fill / depth clear of PassPBRForward (passes/pass_pbr_forward.hpp:64-98), the inline depth raster of
// PassShadowMap (passes/pass_shadow_map.hpp:191-201) and -- when fused -- PassTonemap
// (passes/pass_tonemap.hpp:52-81).
//
// One CTA of 256 threads owns one 16x16-pixel tile; one thread owns one pixel.  The colour / depth tile
// lives in registers (one pixel each), triangle records are staged through shared memory 256 at a time
// and broadcast to all pixels.  Per pixel the kernel keeps only the WINNING fragment:
//     depth target bound : minimum z01, ties broken by the smallest draw-order key
//                          (== the reference's strict "z01 >= zbuf -> skip", rasterizer.hpp:359)
//     no depth target    : the largest draw-order key (painter, rasterizer.hpp:348,419)
// and runs the fragment program ONCE for it.  This equals the reference's serial result because the
// builtin fragment programs are pure functions of (FragmentIn, uniforms) and never discard.
// Each output byte (HDR, depth, LDR) is written exactly once, with 128-bit / full-sector stores.
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        constexpr int TILE_THREADS = TILE_PIXELS; // 256 (16 x 16) or 128 (16 x 8): one thread per pixel
        constexpr int TILE_WARPS = TILE_THREADS / 32;
#ifdef SHSB_PHASE_CLOCKS
        // Debug build only (tools/phase_clocks.py): per scheduling class, cycles thread 0 of each tile CTA spends per phase.
        __device__ unsigned long long g_phase_clk[4][8];
#define PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_clk[cls][i], (unsigned long long)(t_ - phase_t0)); phase_t0 = t_; } } while (0)
#define PHASE_BEGIN() long long phase_t0 = clock64()
#define PHASE_COUNT() do { if (threadIdx.x == 0) atomicAdd(&g_phase_clk[cls][6], 1ull); } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#define PHASE_BEGIN() do { } while (0)
#define PHASE_COUNT() do { } while (0)
#endif
#ifndef TILE_MIN_CTAS
#define TILE_MIN_CTAS (1024 / TILE_PIXELS) /* 32 warps per SM either way (64 registers per thread) */
#endif
        // Debug build only (tools/gpu_store_bound.sh, libshsb_nostore.so): the tile kernel's render-target stores of tiles WITH geometry
        // are predicated on a value no computation produces, so that everything is still computed but nothing is written -- the time
        // difference to the normal build bounds what any other store mechanism (TMA bulk tensor stores from a shared-memory tile)
        // could gain.  Frames rendered by that build are garbage by construction.
#ifdef SHSB_NO_STORES
#define STORE_IF(v) if (__float_as_uint(v) == 0x7fc12345u)
#else
#define STORE_IF(v)
#endif
        constexpr float PI_F = 3.14159265358979323846f;

        struct V3 { float x, y, z; };
        __device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
        __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
        __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
        __device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
        __device__ __forceinline__ V3 operator*(V3 a, float k) { return V3{a.x * k, a.y * k, a.z * k}; }
        __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
        __device__ __forceinline__ V3 mix3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
        __device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
        __device__ __forceinline__ float sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
        // Single-MUFU approximations for colour math only (never for coverage / depth / list bits).  The CUDA
        // intrinsics (__fdividef, rsqrtf, __powf) wrap each MUFU in denormal scaling code when the translation unit is
        // not compiled with -ftz; the .ftz PTX forms are the bare instruction.
        __device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
        __device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
        __device__ __forceinline__ float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
        __device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
        // x^p for x >= 0, p > 0: lg2(0) = -inf -> ex2(-inf) = 0
        __device__ __forceinline__ float fast_pow(float x, float p) { return fast_ex2(p * fast_lg2(x)); }
        __device__ __forceinline__ V3 normalize_fast(V3 v) { return v * fast_rsqrt(dot(v, v)); }
        __device__ __forceinline__ V3 toV3(F3 v) { return V3{v.x, v.y, v.z}; }
        __device__ __forceinline__ float pow5(float x) { const float x2 = x * x; return x2 * x2 * x; }

        struct Surface
        {
            V3 P, N, V, albedo;
            float metallic, roughness;
            bool blinn;
            // per-pixel invariants of the local-light BRDF, hoisted out of the light loop by finish_surface()
            V3 F0, diffuse;        // mix(0.04, albedo, metallic); PBR: albedo*(1-metallic)/pi, Blinn: albedo/pi
            float a2, k, NdotV, g1v, spec_a, spec_b; // Blinn: spec_a = shininess, spec_b = spec_strength; PBR: g1v = a2 * G1(V)
        };

        __device__ __forceinline__ void finish_surface(Surface& s)
        {
            s.F0 = mix3(v3(0.04f, 0.04f, 0.04f), s.albedo, s.metallic);
            if (s.blinn)
            {
                const float smooth = 1.0f - sat(s.roughness);
                s.spec_a = mixf(10.0f, 96.0f, smooth);
                s.spec_b = mixf(0.15f, 0.65f, smooth);
                s.diffuse = s.albedo * (1.0f / PI_F);
                s.a2 = s.k = s.NdotV = s.g1v = 0.0f;
            }
            else
            {
                const float a = s.roughness * s.roughness;
                s.a2 = a * a;
                const float rr = s.roughness + 1.0f;
                s.k = rr * rr * 0.125f;
                s.NdotV = fmaxf(dot(s.N, s.V), 0.0f);
                s.g1v = s.a2 * s.NdotV * fast_rcp(fmaxf(s.NdotV * (1.0f - s.k) + s.k, 1e-6f));
                s.diffuse = s.albedo * ((1.0f - s.metallic) * (1.0f / PI_F));
                s.spec_a = s.spec_b = 0.0f;
            }
        }

        // ---- eval_pbr_light / eval_blinn_phong_light, shaders/vulkan/fp_stress_scene.frag:132-165
        // (same formulas; the three quotients NDF, G2 and 1/(4 NdotV NdotL) share one reciprocal).  NdotL > 0.
        __device__ __forceinline__ V3 eval_brdf_lit(const Surface& s, V3 L, float NdotL, V3 radiance)
        {
            const V3 H = normalize_fast(s.V + L);
            const float NdotH = fmaxf(dot(s.N, H), 0.0f);
            if (s.blinn)
            {
                const float spec = fast_pow(NdotH, s.spec_a);
                return radiance * (s.diffuse * NdotL + s.F0 * (spec * s.spec_b));
            }
            const float fres = pow5(1.0f - fmaxf(dot(H, s.V), 0.0f));
            const V3 F = s.F0 + (v3(1, 1, 1) - s.F0) * fres;
            const float dd = NdotH * NdotH * (s.a2 - 1.0f) + 1.0f;
            const float den = fmaxf(PI_F * dd * dd, 1e-6f) * fmaxf(NdotL * (1.0f - s.k) + s.k, 1e-6f) * fmaxf(4.0f * s.NdotV * NdotL, 1e-6f);
            const float spec = s.g1v * NdotL * fast_rcp(den);
            // kD = (1 - F) * (1 - metallic); kD * albedo / pi = (1 - F) * diffuse
            return ((v3(1, 1, 1) - F) * s.diffuse + F * spec) * radiance * NdotL;
        }

        __device__ __forceinline__ V3 eval_brdf(const Surface& s, V3 L, V3 radiance)
        {
            const float NdotL = fmaxf(dot(s.N, L), 0.0f);
            if (NdotL <= 0.0f) return v3(0, 0, 0);
            return eval_brdf_lit(s, L, NdotL, radiance);
        }

        // ---- shs_eval_light_attenuation_quadratic, shaders/vulkan/common/light_math.glsl:44-78
        __device__ __forceinline__ float attenuation_quadratic(float dist, float range, uint32_t model, float power, float bias, float cutoff)
        {
            const float safe_range = fmaxf(range, 1e-4f);
            const float t = sat(dist * fast_rcp(safe_range));
            const float edge = 1.0f - t;
            float falloff;
            if (model == 0u) falloff = edge;
            else if (model == 2u)
            {
                const float denom = fmaxf(dist * dist, fmaxf(bias, 1e-5f));
                falloff = safe_range * safe_range * fast_rcp(denom) * edge * edge;
            }
            else falloff = edge * edge;
            falloff = fmaxf(falloff, 0.0f);
            const float p = fmaxf(power, 0.001f);
            if (p != 1.0f) falloff = fast_pow(falloff, p);
            return (falloff <= fmaxf(cutoff, 0.0f)) ? 0.0f : falloff;
        }

        // One staged point / spot light against one surface -- eval_local_light, fp_stress_scene.frag:421-523.
        // The tests run cheapest-first; every early-out is a case where the GLSL adds exactly zero:
        // out of range (dist >= range), facing away (NdotL <= 0), attenuation at or below the cutoff, outside the cone.
        __device__ __forceinline__ void accumulate_point_spot(const Surface& s, const SmLight* __restrict__ lt, float dx, float dy, float dz, float dist2,
                                                              uint32_t kind, V3& sum)
        {
            const float nd = s.N.x * dx + s.N.y * dy + s.N.z * dz;
            if (!(nd > 0.0f)) return;
            const float4 ri = lt->radiance_ir, at = lt->atten;
            const float inv_dist = fast_rsqrt(dist2);
            const float dist = dist2 * inv_dist;
            const float edge = 1.0f - sat(dist * ri.w);
            const uint32_t model = kind & 3u;
            float falloff = (model == 0u) ? edge : edge * edge;
            if (model == 2u) falloff *= lt->pos_r2.w * fast_rcp(fmaxf(dist2, at.z));
            if (kind & KIND_POW) falloff = fast_pow(falloff, at.x);
            if (!(falloff > at.y)) return;
            if (kind & KIND_SPOT)
            {
                const float4 dc = lt->dir_outer;
                const float cone_cos = -(dc.x * dx + dc.y * dy + dc.z * dz) * inv_dist;
                const float t = sat((cone_cos - dc.w) * at.w);
                const float spot = t * t * (3.0f - 2.0f * t);
                if (!(spot > 0.0f)) return;
                falloff *= spot;
            }
            const V3 L = v3(dx * inv_dist, dy * inv_dist, dz * inv_dist);
            sum = sum + eval_brdf_lit(s, L, nd * inv_dist, v3(ri.x, ri.y, ri.z) * falloff);
        }

        // Any light type straight from the 160-B record (saturated tiles, non-16 light tiles, rect / tube lights).
        __device__ __noinline__ V3 eval_light_record(const Surface& s, const DevLightRec* __restrict__ rec)
        {
            const uint32_t type = rec->type_shape_flags[0], flags = rec->type_shape_flags[2], model = rec->type_shape_flags[3];
            const V3 zero = v3(0, 0, 0);
            if ((flags & 1u) == 0u || type < 1u || type > 4u) return zero;
            const V3 lpos = v3(rec->position_range[0], rec->position_range[1], rec->position_range[2]);
            const float range = fmaxf(rec->position_range[3], 0.001f);
            V3 to_light = lpos - s.P;
            float dist = sqrtf(dot(to_light, to_light));
            float rect_forward = 0.0f;
            if (type == 3u)
            {
                const V3 right = normalize_fast(v3(rec->axis_spot_outer[0], rec->axis_spot_outer[1], rec->axis_spot_outer[2]));
                const V3 up = normalize_fast(v3(rec->up_shape_x[0], rec->up_shape_x[1], rec->up_shape_x[2]));
                const V3 emit = normalize_fast(v3(rec->direction_spot[0], rec->direction_spot[1], rec->direction_spot[2]));
                const float hx = fmaxf(rec->up_shape_x[3], 1e-4f), hy = fmaxf(rec->shape_attenuation[0], 1e-4f);
                const V3 rel = s.P - lpos;
                rect_forward = dot(rel, emit);
                if (rect_forward <= 1e-4f || rect_forward >= range) return zero;
                const float lx = dot(rel, right), ly = dot(rel, up);
                const float x = fminf(fmaxf(lx, -hx), hx), y = fminf(fmaxf(ly, -hy), hy);
                const float dx = fmaxf(fabsf(lx) - hx, 0.0f), dy = fmaxf(fabsf(ly) - hy, 0.0f);
                if (sqrtf(dx * dx + dy * dy + rect_forward * rect_forward) >= range) return zero;
                to_light = (lpos + right * x + up * y) - s.P;
                dist = sqrtf(dot(to_light, to_light));
            }
            else if (type == 4u)
            {
                const V3 axis = normalize_fast(v3(rec->axis_spot_outer[0], rec->axis_spot_outer[1], rec->axis_spot_outer[2]));
                const float half_len = fmaxf(rec->up_shape_x[3], 1e-4f);
                const V3 p0 = lpos - axis * half_len, p1 = lpos + axis * half_len;
                const V3 seg = p1 - p0;
                const float u = sat(dot(s.P - p0, seg) / fmaxf(dot(seg, seg), 1e-6f));
                to_light = (p0 + seg * u) - s.P;
                dist = sqrtf(dot(to_light, to_light));
            }
            if (dist <= 1e-5f || dist >= range) return zero;
            const V3 L = to_light * (1.0f / fmaxf(dist, 1e-5f));
            float atten = attenuation_quadratic(dist, range, model, rec->shape_attenuation[1], rec->shape_attenuation[2], rec->shape_attenuation[3]);
            if (atten <= 0.0f) return zero;
            if (type == 2u)
            {
                const V3 sd = normalize_fast(v3(rec->direction_spot[0], rec->direction_spot[1], rec->direction_spot[2]));
                const float inner_cos = fminf(fmaxf(rec->direction_spot[3], -1.0f), 1.0f);
                const float outer_cos = fminf(fmaxf(rec->axis_spot_outer[3], -1.0f), inner_cos);
                const float t = sat((-dot(sd, L) - outer_cos) / fmaxf(inner_cos - outer_cos, 1e-6f));
                const float spot = t * t * (3.0f - 2.0f * t);
                if (spot <= 0.0f) return zero;
                atten *= spot;
            }
            else if (type == 3u)
            {
                const V3 emit = normalize_fast(v3(rec->direction_spot[0], rec->direction_spot[1], rec->direction_spot[2]));
                const float one_sided = fmaxf(-dot(emit, L), 0.0f);
                if (one_sided <= 0.0f) return zero;
                atten *= one_sided * sat(1.0f - rect_forward / fmaxf(range, 1e-4f));
            }
            else if (type == 4u)
            {
                const float edge_soften = fminf(fmaxf(fmaxf(rec->shape_attenuation[0], 1e-4f) / fmaxf(range, 1e-4f), 0.05f), 1.0f);
                atten *= mixf(0.65f, 1.0f, edge_soften);
            }
            const V3 radiance = v3(rec->color_intensity[0], rec->color_intensity[1], rec->color_intensity[2]) * (rec->color_intensity[3] * atten);
            return eval_brdf(s, L, radiance);
        }

        // ---- eval_fake_ibl, shader/builtin_shaders.hpp:57-85
        __device__ __forceinline__ V3 fake_ibl(V3 N, V3 V, V3 base_color, float metallic, float roughness, float ao)
        {
            const V3 n = normalize_fast(N), v = normalize_fast(V);
            const V3 mv = v * -1.0f;
            const V3 r = mv - n * (dot(n, mv) * 2.0f);
            const V3 zen = v3(0.32f, 0.46f, 0.72f), hor = v3(0.62f, 0.66f, 0.72f), gnd = v3(0.16f, 0.15f, 0.14f);
            const float up_n = sat(n.y * 0.5f + 0.5f), up_r = sat(r.y * 0.5f + 0.5f);
            const V3 env_n = mix3(gnd, mix3(hor, zen, up_n), up_n);
            const V3 env_r = mix3(gnd, mix3(hor, zen, up_r), up_r);
            const float m = sat(metallic), rgh = sat(roughness);
            const V3 F0 = mix3(v3(0.04f, 0.04f, 0.04f), v3(fmaxf(base_color.x, 0.0f), fmaxf(base_color.y, 0.0f), fmaxf(base_color.z, 0.0f)), m);
            const float fres = pow5(1.0f - fmaxf(0.0f, dot(n, v)));
            const V3 F = F0 + (v3(1, 1, 1) - F0) * fres;
            const V3 kd = (v3(1, 1, 1) - F) * (1.0f - m);
            const V3 diffuse_ibl = kd * base_color * env_n * 0.12f;
            const float spec_strength = 0.02f + (1.0f - rgh) * 0.18f;
            return (diffuse_ibl + env_r * F * spec_strength) * sat(ao);
        }

        // ---- sample_texture2d_bilinear_repeat_linear, shader/builtin_shaders.hpp:33-55.
        // Addressing is exact (it selects texels); srgb_lut[i] = powf(i/255, 2.2) computed on the host by libm.
        __device__ __forceinline__ V3 sample_texture(const DevTexture& tex, const float* __restrict__ lut, float uvx, float uvy)
        {
            const float u = xsub(uvx, floorf(uvx)), v = xsub(uvy, floorf(uvy));
            const float fx = xmul(u, (float)(tex.w - 1)), fy = xmul(v, (float)(tex.h - 1));
            const int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
            const int x1 = min(x0 + 1, tex.w - 1), y1 = min(y0 + 1, tex.h - 1);
            const float tx = xsub(fx, (float)x0), ty = xsub(fy, (float)y0);
            const uchar4 t00 = tex.texels[(size_t)y0 * tex.w + x0], t10 = tex.texels[(size_t)y0 * tex.w + x1];
            const uchar4 t01 = tex.texels[(size_t)y1 * tex.w + x0], t11 = tex.texels[(size_t)y1 * tex.w + x1];
            const V3 c00 = v3(lut[t00.x], lut[t00.y], lut[t00.z]), c10 = v3(lut[t10.x], lut[t10.y], lut[t10.z]);
            const V3 c01 = v3(lut[t01.x], lut[t01.y], lut[t01.z]), c11 = v3(lut[t11.x], lut[t11.y], lut[t11.z]);
            return mix3(mix3(c00, c10, tx), mix3(c01, c11, tx), ty);
        }

        // ---- shadow_visibility_dir, lighting/shadow_sample.hpp:65-104 (exact: it selects texels and compares depths)
        __device__ __forceinline__ float shadow_visibility(const FrameConst& fc, F3 pos_ws, float ndotl)
        {
            const float4 p = xmat4_mul(fc.light_viewproj, pos_ws.x, pos_ws.y, pos_ws.z, 1.0f);
            if (fabsf(p.w) < 1e-8f) return 1.0f;
            const float u = xadd(xmul(xdiv(p.x, p.w), 0.5f), 0.5f);
            const float v = xadd(xmul(xdiv(p.y, p.w), 0.5f), 0.5f);
            const float z = xadd(xmul(xdiv(p.z, p.w), 0.5f), 0.5f);
            if (u < 0.0f || u > 1.0f || v < 0.0f || v > 1.0f) return 1.0f;
            const float slope = xsub(1.0f, sclamp(ndotl, 0.0f, 1.0f));
            const float z_test = xsub(z, xadd(fc.bias_const, xmul(fc.bias_slope, slope)));
            const int cx = (int)roundf(xmul(u, (float)(fc.shadow_w - 1)));
            const int cy = (int)roundf(xmul(v, (float)(fc.shadow_h - 1)));
            const int r = max(0, fc.pcf_radius);
            if (r == 0)
            {
                const int x = min(max(cx, 0), fc.shadow_w - 1), y = min(max(cy, 0), fc.shadow_h - 1);
                return (z_test <= fc.shadow_map[(size_t)y * fc.shadow_w + x]) ? 1.0f : 0.0f;
            }
            const int step = max(1, (int)roundf(fmaxf(1.0f, fc.pcf_step)));
            int lit = 0;
            for (int oy = -r; oy <= r; ++oy)
            {
                const int y = min(max(cy + oy * step, 0), fc.shadow_h - 1);
                for (int ox = -r; ox <= r; ++ox)
                {
                    const int x = min(max(cx + ox * step, 0), fc.shadow_w - 1);
                    lit += (z_test <= fc.shadow_map[(size_t)y * fc.shadow_w + x]) ? 1 : 0;
                }
            }
            const int count = (2 * r + 1) * (2 * r + 1);
            return xdiv((float)lit, (float)count);
        }

        __device__ __forceinline__ uchar4 tonemap_pixel(float r, float g, float b, float exposure, float inv_gamma)
        {
            // PassTonemap, passes/pass_tonemap.hpp:58-80: max(0, c*exposure) -> x/(1+x) -> pow(x, 1/gamma) -> lround(x*255).
            // pow through MUFU lg2/ex2 (relative error ~1e-6): the 8-bit result can differ from libm's only when x*255
            // lands within ~3e-4 of a rounding boundary, i.e. by at most 1 LSB (the north_star colour gate).
            float c[3] = {r, g, b};
            unsigned char o[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
            {
                float v = fmaxf(0.0f, c[i] * exposure);
                v = v * fast_rcp(1.0f + v);
                v = (inv_gamma == 0.0f) ? 1.0f : fast_pow(v, inv_gamma);       // pow(x, 0) = 1; pow(0, p > 0) = 0
                const float q = floorf(v * 255.0f + 0.5f); // lround: half away from zero (v >= 0)
                o[i] = (unsigned char)fminf(fmaxf(q, 0.0f), 255.0f);
            }
            return make_uchar4(o[0], o[1], o[2], 255);
        }

        // ---- sky models (Scene::sky): exact arithmetic, they select texels and feed an HDR target compared at PSNR >= 60 dB
        __device__ __forceinline__ float xmix(float a, float b, float t) { return xadd(xmul(a, xsub(1.0f, t)), xmul(b, t)); } // glm::mix
        __device__ __forceinline__ F3 xmix3(F3 a, F3 b, float t) { return F3{xmix(a.x, b.x, t), xmix(a.y, b.y, t), xmix(a.z, b.z, t)}; }

        // ProceduralSky::sample, sky/procedural_sky.hpp:25-45
        __device__ __forceinline__ F3 sky_procedural(F3 dir, const float* __restrict__ sun)
        {
            const F3 d = xnormalize3(dir);
            const float t = gclamp(xadd(xmul(d.y, 0.5f), 0.5f), 0.0f, 1.0f);
            F3 sky = xmix3(F3{0.30f, 0.60f, 1.00f}, F3{0.05f, 0.20f, 0.50f}, t);
            const float sun_dot = xdot3(d, F3{-sun[0], -sun[1], -sun[2]});
            if (sun_dot > 0.9998f) sky = F3{15.0f, 15.0f, 15.0f};
            else if (sun_dot > 0.9990f)
            {
                const float glow = xdiv(xsub(sun_dot, 0.9990f), xsub(0.9998f, 0.9990f));
                sky = xmix3(sky, F3{10.0f, 8.0f, 4.0f}, glow);
            }
            return sky;
        }

        // sample_face_bilinear_linear, sky/cubemap_sky.hpp:39-60 (clamped addressing; lut[i] = powf(i/255, 2.2) from the host libm)
        __device__ __forceinline__ F3 sample_face_clamped(const DevTexture& tex, const float* __restrict__ lut, float u, float v)
        {
            u = gclamp(u, 0.0f, 1.0f);
            v = gclamp(v, 0.0f, 1.0f);
            const float fx = xmul(u, (float)(tex.w - 1)), fy = xmul(v, (float)(tex.h - 1));
            const int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
            const int x1 = min(x0 + 1, tex.w - 1), y1 = min(y0 + 1, tex.h - 1);
            const float tx = xsub(fx, (float)x0), ty = xsub(fy, (float)y0);
            const uchar4 t00 = tex.texels[(size_t)y0 * tex.w + x0], t10 = tex.texels[(size_t)y0 * tex.w + x1];
            const uchar4 t01 = tex.texels[(size_t)y1 * tex.w + x0], t11 = tex.texels[(size_t)y1 * tex.w + x1];
            const F3 v00{lut[t00.x], lut[t00.y], lut[t00.z]}, v10{lut[t10.x], lut[t10.y], lut[t10.z]};
            const F3 v01{lut[t01.x], lut[t01.y], lut[t01.z]}, v11{lut[t11.x], lut[t11.y], lut[t11.z]};
            return xmix3(xmix3(v00, v10, tx), xmix3(v01, v11, tx), ty);
        }

        // CubemapSky::sample, sky/cubemap_sky.hpp:69-109
        __device__ __forceinline__ F3 sky_cubemap(const FrameConst& fc, const DevTexture* __restrict__ textures, const float* __restrict__ lut, F3 d)
        {
            const float len = xsqrt(xdot3(d, d));
            if (len < 1e-8f) return F3{0.0f, 0.0f, 0.0f};
            d.x = xdiv(d.x, len); d.y = xdiv(d.y, len); d.z = xdiv(d.z, len);
            const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
            int face;
            float u, v;
            if (ax >= ay && ax >= az)
            {
                if (d.x > 0.0f) { face = 0; u = xdiv(-d.z, ax); v = xdiv(d.y, ax); }
                else { face = 1; u = xdiv(d.z, ax); v = xdiv(d.y, ax); }
            }
            else if (ay >= ax && ay >= az)
            {
                if (d.y > 0.0f) { face = 2; u = xdiv(d.x, ay); v = xdiv(-d.z, ay); }
                else { face = 3; u = xdiv(d.x, ay); v = xdiv(d.z, ay); }
            }
            else
            {
                if (d.z > 0.0f) { face = 4; u = xdiv(d.x, az); v = xdiv(d.y, az); }
                else { face = 5; u = xdiv(-d.x, az); v = xdiv(d.y, az); }
            }
            u = xmul(0.5f, xadd(u, 1.0f));
            v = xmul(0.5f, xadd(v, 1.0f));
            const F3 c = sample_face_clamped(textures[fc.sky_faces[face]], lut, u, v);
            return F3{xmul(c.x, fc.sky_intensity), xmul(c.y, fc.sky_intensity), xmul(c.z, fc.sky_intensity)};
        }

        // render_skybox_to_hdr, sky/skybox_renderer.hpp:25-57, for one pixel
        __device__ __forceinline__ F3 sky_pixel(const FrameConst& fc, const DevTexture* __restrict__ textures, const float* __restrict__ lut, int px, int py)
        {
            const float ndc_x = xsub(xdiv(xmul(2.0f, xadd((float)px, 0.5f)), (float)fc.W), 1.0f);
            const float ndc_y = xsub(xdiv(xmul(2.0f, xadd((float)py, 0.5f)), (float)fc.H), 1.0f);
            const float4 world = xmat4_mul(fc.inv_viewproj, ndc_x, ndc_y, 1.0f, 1.0f);
            if (fabsf(world.w) < 1e-8f) return F3{0.0f, 0.0f, 0.0f};
            const F3 dir = xnormalize3(F3{xsub(xdiv(world.x, world.w), fc.camera_pos[0]), xsub(xdiv(world.y, world.w), fc.camera_pos[1]),
                                          xsub(xdiv(world.z, world.w), fc.camera_pos[2])});
            return (fc.sky_kind == 1) ? sky_procedural(dir, fc.sky_sun) : sky_cubemap(fc, textures, lut, dir);
        }

        // Resolve of a pixel no fragment reached: motion (0, 0) when the pass clears the plane; colour = sky model or
        // background gradient (pass_pbr_forward.hpp:64-85) or, for a separate draw (load_color), the target's existing
        // colour; fused tonemap if an LDR target is bound.
        template <bool FAST>
        __device__ __forceinline__ void resolve_uncovered(const FrameConst& fc, const FrameBuffers& fb, const DevTexture* __restrict__ textures,
                                                          const float* __restrict__ lut, size_t pix, int px, int py)
        {
            const int F_clear_motion = FAST ? 0 : fc.clear_motion, F_load_color = FAST ? 0 : fc.load_color, F_sky_kind = FAST ? 0 : fc.sky_kind;
            if (F_clear_motion && fb.motion) fb.motion[pix] = make_float2(0.0f, 0.0f);
            if (F_load_color)
            {
                if (!(fc.fuse_tonemap && fb.ldr)) return;
                const float4 c = fb.hdr[pix];
                fb.ldr[pix] = tonemap_pixel(c.x, c.y, c.z, fc.exposure, fc.inv_gamma);
                return;
            }
            float r, g, b;
            if (F_sky_kind != 0)
            {
                const F3 c = sky_pixel(fc, textures, lut, px, py);
                r = c.x; g = c.y; b = c.z;
            }
            else
            {
                const float t = xdiv((float)py, (float)max(1, fc.H - 1));
                r = xadd(0.06f, xmul(0.08f, t)); g = xadd(0.08f, xmul(0.10f, t)); b = xadd(0.12f, xmul(0.12f, t));
            }
            STORE_IF(r) fb.hdr[pix] = make_float4(r, g, b, 1.0f);
            if (fc.fuse_tonemap && fb.ldr) { const uchar4 l = tonemap_pixel(r, g, b, fc.exposure, fc.inv_gamma); STORE_IF(__uint_as_float((uint32_t)l.x * 0x01010101u + 0x7fc12000u)) fb.ldr[pix] = l; }
        }

        constexpr int LIGHT_CAP = TILE_THREADS;       // staged lights per pass
#ifndef SHSB_CAND_PER_THREAD
#define SHSB_CAND_PER_THREAD 4
#endif
        constexpr int CAND_PER_THREAD = SHSB_CAND_PER_THREAD; // candidates filtered per thread per staging round (x TILE_THREADS per CTA)
        static_assert(CAND_PER_THREAD * (TILE_PIXELS / 32) <= 32, "the (slot, warp) ballots of a round are scanned by one warp");

        // FAST: the instantiation for the plain Forward+ frame (the benchmarked one) -- linear-depth target cleared by the pass, lit builtin
        // program, 16-pixel light tiles with lights, no shadow-map pass mode, no sky model, no motion plane traffic, no AOVs, no Hi-Z.  The
        // mode flags are compile-time constants there, so the other modes' code is not in the kernel at all (instruction cache, uniform
        // branches); launch_tile_raster picks the instantiation, every other frame runs the general one.  Same source, same arithmetic.
        // PROG: 0 = general (every mode flag read from fc), 1 = FAST with the PBR program, 2 = FAST with the Blinn-Phong program.
        // LIGHTS (FAST only): 1 = Forward+ over 16-pixel light tiles, point / spot lights only, no sun shadow map; 3 = the same with rect /
        // tube lights possible (they are evaluated from the 160-byte record by a function that is CALLED with the surface by reference, which
        // keeps the surface addressable: stack frame, spills in the staging code); 2 = no local lights (sun shadow map allowed).
        // The instantiations without local lights need 56 registers uncapped; capped at 48 (10 CTAs of 128 threads per SM) they do not
        // spill and hide more latency: -4 % (C3) / -7 % (C4) on the tile kernel against 8 CTAs, 12 CTAs (40 registers) give no more
        // (profiles/r2_tile_kernel_specialisation.md)
#ifndef SHSB_NOLIGHT_CTAS
#define SHSB_NOLIGHT_CTAS 10
#endif
        // The point / spot Forward+ instantiations at 10 CTAs per SM (48 registers, 104 bytes of spills in the STAGING code, none in the light
        // loop): 10 195 frames/s on C2 against 9610 at 8 CTAs (64 registers), 9870 at 9, 10 035 at 11-12, 9205 at 14 (same file).  The
        // general kernel and the area-light instantiations keep 8: round 1 measured more CTAs as a loss for the kernel that carries everything.
#ifndef SHSB_FPLUS_CTAS
#define SHSB_FPLUS_CTAS 10
#endif
        template <int PROG, int LIGHTS>
        __global__ void __launch_bounds__(TILE_THREADS, (LIGHTS == 2 ? SHSB_NOLIGHT_CTAS : (LIGHTS == 1 ? SHSB_FPLUS_CTAS : TILE_MIN_CTAS))) tile_kernel(const FrameConst fc, const Geometry g, const FrameBuffers fb,
                                                                    const DevTexture* __restrict__ textures,
                                                                    const float* __restrict__ srgb_lut)
        {
            constexpr bool FAST = PROG != 0;
            static_assert(FAST ? (LIGHTS >= 1 && LIGHTS <= 3) : LIGHTS == 0, "FAST instantiations fix the light mode");
            const int F_shadow_mode = FAST ? 0 : fc.shadow_mode, F_has_depth = FAST ? 1 : fc.has_depth, F_linear_depth = FAST ? 1 : fc.linear_depth;
            const int F_load_depth = FAST ? 0 : fc.load_depth, F_load_color = FAST ? 0 : fc.load_color, F_write_motion = FAST ? 0 : fc.write_motion;
            const int F_clear_motion = FAST ? 0 : fc.clear_motion, F_hiz = FAST ? 0 : fc.hiz, F_sky_kind = FAST ? 0 : fc.sky_kind;
            const int F_shader_id = PROG == 1 ? 0 : (PROG == 2 ? 1 : fc.shader_id), F_forward_plus = FAST ? (LIGHTS != 2 ? 1 : 0) : fc.forward_plus;
            const float* const F_shadow_map = (FAST && LIGHTS != 2) ? nullptr : fc.shadow_map;
            const uint32_t F_light_tile_size = FAST ? (uint32_t)TILE : fc.light_tile_size;
            uint32_t* const F_aov_tri_id = FAST ? nullptr : fb.aov_tri_id;
            uint32_t* const F_aov_coverage = FAST ? nullptr : fb.aov_coverage;
            float2* const F_motion = FAST ? nullptr : fb.motion;
            // the triangle staging buffer (raster phase) and the light staging buffer (shading phase) are never live at
            // the same time: they share one 20-KB allocation
            __shared__ __align__(16) unsigned char s_stage[LIGHT_CAP * sizeof(SmLight)];
            static_assert(sizeof(SmLight) >= sizeof(RasterRec), "staging buffer is sized by SmLight");
            RasterRec* s_rec = reinterpret_cast<RasterRec*>(s_stage);
            SmLight* s_light = reinterpret_cast<SmLight*>(s_stage);
            __shared__ uint32_t s_idx[TILE_THREADS];
            __shared__ float s_znear[TILE_THREADS];               // hierarchical Z: conservative nearest z01 of each staged triangle
            __shared__ unsigned s_wmask[TILE_THREADS / 32][TILE_THREADS / 32]; // [consumer warp][staging warp]
            __shared__ unsigned s_ballot[CAND_PER_THREAD * (TILE_THREADS / 32)]; // light staging: [candidate slot][warp] == ascending light order
            __shared__ float s_box[TILE_THREADS / 32][6];
            __shared__ unsigned long long s_frag[TILE_THREADS / 32][2];

            PHASE_BEGIN();
            // the frame's total demand, for the host to size the arenas from when a front-end kernel had to drop something
            if (blockIdx.x == 0 && threadIdx.x == 0)
            {
                const uint32_t need_lists = *g.list_cursor, need_recs = *g.rec_count;
                if (need_lists > g.list_capacity) g.overflow_flag[1] = need_lists;
                if (need_recs > g.rec_capacity) g.overflow_flag[2] = need_recs;
            }
            // heaviest scheduling class first (alloc_kernel, binning.cu): the cheap background tiles fill the tail
            const uint32_t n_tiles_total = (uint32_t)fc.tiles_x * (uint32_t)fc.tiles_y;
            const uint32_t cc0 = g.class_count[0], cc1 = g.class_count[1], cc2 = g.class_count[2];
            const uint32_t n_nonempty = cc0 + cc1 + cc2;
            if (blockIdx.x >= n_nonempty)
            {
                // ---------------- empty tiles (class 3; most of a frame): nothing to rasterise, no barriers, no counters.
                // One CTA resolves FOUR tiles, one thread 4 horizontally adjacent pixels with 128-bit stores; the
                // background colour depends on the row only, so it is evaluated once per 4 pixels.
                constexpr int Q = TILE_PIXELS / 4; // threads per empty tile: 4 per row
                const uint32_t first = (blockIdx.x - n_nonempty) * 4u + (threadIdx.x / Q);
                if (first >= g.class_count[3]) return; // owned empty tiles only (sort-first partitions leave other rows untouched)
                const uint32_t packed = g.tile_order[(size_t)3 * n_tiles_total + first];
                const int q = threadIdx.x % Q;
                const int x0 = (int)(packed & 0xffffu) * TILE + (q & 3) * 4;
                const int fy = (int)(packed >> 16) * TILE_H + (q >> 2);
                if (x0 >= fc.W || fy >= fc.H) return;
                const int py = fc.H - 1 - fy;
                const size_t pix = (size_t)py * (size_t)fc.W + (size_t)x0;
                const bool clear_depth = fb.depth && (F_has_depth || F_shadow_mode) && !F_load_depth;
                const bool shade = !(F_shadow_mode || F_shader_id == 5 || !fb.hdr);
                if (x0 - (q & 3) * 4 + TILE <= fc.W && (fc.W & 3) == 0 && !F_load_color && F_sky_kind == 0) // whole tile row inside, 16-byte aligned rows, row-constant colour
                {
                    if (clear_depth) *reinterpret_cast<float4*>(fb.depth + pix) = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                    if (F_aov_tri_id) *reinterpret_cast<uint4*>(F_aov_tri_id + pix) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                    if (F_aov_coverage) *reinterpret_cast<uint4*>(F_aov_coverage + pix) = make_uint4(0u, 0u, 0u, 0u);
                    if (!shade) return;
                    if (F_clear_motion && F_motion)
                    {
                        float4* m = reinterpret_cast<float4*>(F_motion + pix);
                        m[0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); m[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    }
                    // background gradient, pass_pbr_forward.hpp:73-81
                    const float t = xdiv((float)py, (float)max(1, fc.H - 1));
                    const float4 c = make_float4(xadd(0.06f, xmul(0.08f, t)), xadd(0.08f, xmul(0.10f, t)), xadd(0.12f, xmul(0.12f, t)), 1.0f);
                    // the 4 threads of a row write 64 contiguous bytes per store instruction (whole sectors)
                    float4* row = fb.hdr + (pix - (size_t)((q & 3) * 4)) + (size_t)(q & 3);
                    row[0] = c; row[4] = c; row[8] = c; row[12] = c;
                    if (fc.fuse_tonemap && fb.ldr)
                    {
                        const uchar4 l = tonemap_pixel(c.x, c.y, c.z, fc.exposure, fc.inv_gamma);
                        const uint32_t w = (uint32_t)l.x | ((uint32_t)l.y << 8) | ((uint32_t)l.z << 16) | ((uint32_t)l.w << 24);
                        *reinterpret_cast<uint4*>(fb.ldr + pix) = make_uint4(w, w, w, w);
                    }
                    return;
                }
                for (int i = 0; i < 4 && x0 + i < fc.W; ++i)
                {
                    if (clear_depth) fb.depth[pix + i] = 1.0f;
                    if (F_aov_tri_id) F_aov_tri_id[pix + i] = 0xFFFFFFFFu;
                    if (F_aov_coverage) F_aov_coverage[pix + i] = 0u;
                    if (shade) resolve_uncovered<FAST>(fc, fb, textures, srgb_lut, pix + i, x0 + i, py);
                }
                return;
            }
            uint32_t ord = blockIdx.x, cls = 0;
            if (ord >= cc0) { ord -= cc0; cls = 1; if (ord >= cc1) { ord -= cc1; cls = 2; } }
            const uint32_t packed = g.tile_order[(size_t)cls * n_tiles_total + ord]; // tx | ty << 16
            const int tx = (int)(packed & 0xffffu), ty = (int)(packed >> 16);
            const int tile = ty * fc.tiles_x + tx;
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            // a warp owns an 8x4-pixel block: 128 contiguous bytes of HDR per row
            const int px = tx * TILE + (warp & 1) * 8 + (lane & 7);
            const int fy = ty * TILE_H + (warp >> 1) * 4 + (lane >> 3);
            const int py = fc.H - 1 - fy;
            const bool valid = px < fc.W && fy < fc.H;
            const size_t pix = valid ? ((size_t)py * (size_t)fc.W + (size_t)px) : 0;
            const uint32_t tile_tris = g.tile_count[tile]; // > 0: classes 0-2 hold exactly the tiles with triangles
            PHASE_MARK(7); // prologue (class / tile lookup)
            PHASE_COUNT();

            float bz = 1.0f;
            if ((F_has_depth || F_shadow_mode) && F_load_depth && valid) bz = fb.depth[pix];
            uint32_t bkey = KEY_NONE, bidx = 0;
            uint32_t n_cov = 0;
            const float pxf = xadd((float)px, 0.5f), pyf = xadd((float)py, 0.5f);
            const float zrange = xsub(fc.zf, fc.zn);

            // hierarchical Z (north_star): the farthest depth the warp's 8x4 block holds; a staged triangle that cannot be nearer than
            // that anywhere is skipped before its bbox is even decoded.  Linear view-depth targets only (the bound below is for them).
            const bool hiz = F_hiz && F_has_depth && F_linear_depth && !F_shadow_mode;
            float blk_zmax = 1.0f;
            bool z_dirty = F_load_depth != 0; // loaded depths: take the block maximum before the first test

            const uint32_t off0 = g.tile_offset[tile];
            const uint32_t off1 = min(off0 + tile_tris, g.list_capacity);
            for (uint32_t base = off0; base < off1; base += TILE_THREADS)
            {
                if (base != off0) __syncthreads(); // the previous batch's readers are done with s_rec / s_wmask
                const uint32_t n = min((uint32_t)TILE_THREADS, off1 - base);
                int sminx = 0, smaxx = 0, sminy = 0, smaxy = 0;
                if (threadIdx.x < n)
                {
                    const uint32_t rec = g.tile_list[base + threadIdx.x];
                    const float4* src = reinterpret_cast<const float4*>(g.rrecs + rec);
                    float4* dst = reinterpret_cast<float4*>(&s_rec[threadIdx.x]);
                    const float4 q3 = __ldg(src + 3);
                    dst[0] = __ldg(src + 0); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2); dst[3] = q3;
                    s_idx[threadIdx.x] = rec;
                    if (hiz)
                    {
                        // z01 = clamp((1 / denom - zn) / (zf - zn)) with denom a (rounded) convex combination of the corners' 1 / w: it cannot
                        // exceed the largest of them by more than a few ULP, so 1 / max(1 / w) -- shrunk by 1e-5 relative, twice, and 1e-6
                        // absolute, orders of magnitude above any rounding in the per-pixel expression -- is a lower bound of every
                        // fragment's depth.  (All 1 / w are positive: records exist only for triangles inside / clipped to w > 0.)
                        const float4 q1 = dst[1], q2 = dst[2];
                        const float max_iw = fmaxf(fmaxf(q1.w, q2.x), q2.y);
                        const float zb = (max_iw > 0.0f) ? ((__fdividef(0.99999f, max_iw) - fc.zn) * __fdividef(0.99999f, fc.zf - fc.zn) - 1e-6f) : 0.0f;
                        s_znear[threadIdx.x] = zb;
                    }
                    const uint32_t bx = __float_as_uint(q3.y), by = __float_as_uint(q3.z);
                    sminx = (int)(bx & 0xffffu); smaxx = (int)(bx >> 16);
                    sminy = (int)(by & 0xffffu); smaxy = (int)(by >> 16);
                }
                // per consumer warp (8x4-pixel block) a 256-bit mask of the staged triangles whose bbox touches it
#pragma unroll
                for (int w = 0; w < TILE_THREADS / 32; ++w)
                {
                    const int rx0 = tx * TILE + (w & 1) * 8, rx1 = rx0 + 7;
                    const int ry_hi = fc.H - 1 - (ty * TILE_H + (w >> 1) * 4), ry_lo = ry_hi - 3;
                    const bool hit = threadIdx.x < n && !(smaxx < rx0 || sminx > rx1 || smaxy < ry_lo || sminy > ry_hi);
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (lane == 0) s_wmask[w][warp] = m;
                }
                __syncthreads();
                for (int k = 0; k < TILE_THREADS / 32; ++k)
                {
                  unsigned wm = s_wmask[warp][k];
                  while (wm)
                  {
                    const uint32_t j = (uint32_t)k * 32u + (uint32_t)(__ffs(wm) - 1);
                    wm &= wm - 1u;
                    if (hiz)
                    {
                        __syncwarp(); // wm is warp-uniform: every lane is back here once per staged triangle
                        if (__any_sync(0xffffffffu, z_dirty))
                        {
                            blk_zmax = __uint_as_float(__reduce_max_sync(0xffffffffu, valid ? __float_as_uint(bz) : 0u)); // depths are in [0, 1]: bit order == value order
                            z_dirty = false;
                        }
                        if (s_znear[j] > blk_zmax) continue; // strictly behind every pixel of the block: no fragment of it can win (ties need equality)
                    }
                    const RasterRec& r = s_rec[j];
                    const int minx = (int)(r.bbox_x & 0xffffu), maxx = (int)(r.bbox_x >> 16);
                    const int miny = (int)(r.bbox_y & 0xffffu), maxy = (int)(r.bbox_y >> 16);
                    if (!valid || px < minx || px > maxx || py < miny || py > maxy) continue;  // the reference's bbox loop bounds
                    // barycentric_2d, rasterizer.hpp:167-179
                    const float v2x = xsub(pxf, r.ax), v2y = xsub(pyf, r.ay);
                    const float bv = xmul(xsub(xmul(v2x, r.v1y), xmul(r.v1x, v2y)), r.inv_den);
                    const float bw = xmul(xsub(xmul(r.v0x, v2y), xmul(v2x, r.v0y)), r.inv_den);
                    const float bu = xsub(xsub(1.0f, bv), bw);
                    if (bu < 0.0f || bv < 0.0f || bw < 0.0f) continue;
                    if (F_shadow_mode)
                    {
                        // pass_shadow_map.hpp:197-200: affine NDC z, keep the minimum
                        const float z_ndc = xadd(xadd(xmul(bu, r.zw0), xmul(bv, r.zw1)), xmul(bw, r.zw2));
                        const float z01 = sclamp(xadd(xmul(z_ndc, 0.5f), 0.5f), 0.0f, 1.0f);
                        ++n_cov;
                        if (z01 < bz) { bz = z01; bkey = r.key; }
                        continue;
                    }
                    const float denom = xadd(xadd(xmul(bu, r.iw0), xmul(bv, r.iw1)), xmul(bw, r.iw2));
                    if (denom <= 1e-10f) continue;
                    ++n_cov;
                    if (F_has_depth)
                    {
                        float z01;
                        if (F_linear_depth)
                        {
                            const float view_z = xrcp(denom);
                            z01 = gclamp(xdiv(xsub(view_z, fc.zn), zrange), 0.0f, 1.0f);
                        }
                        else
                        {
                            const float z_clip = xadd(xadd(xmul(bu, r.zw0), xmul(bv, r.zw1)), xmul(bw, r.zw2));
                            z01 = gclamp(xadd(xmul(xmul(z_clip, xrcp(denom)), 0.5f), 0.5f), 0.0f, 1.0f);
                        }
                        if (z01 < bz || (z01 == bz && r.key < bkey)) { bz = z01; bkey = r.key; bidx = s_idx[j]; z_dirty = true; }
                    }
                    else if (r.key > bkey) { bkey = r.key; bidx = s_idx[j]; }
                  }
                }
            }

            PHASE_MARK(0); // raster (staging + scan)
            // ---------------- resolve: depth + AOVs
            if (valid)
            {
                if (fb.depth && (F_has_depth || F_shadow_mode) && (bkey != KEY_NONE || !F_load_depth)) { STORE_IF(bz) fb.depth[pix] = bz; }
                if (F_aov_tri_id) F_aov_tri_id[pix] = (bkey != KEY_NONE) ? (bkey - 1u) : 0xFFFFFFFFu;
                if (F_aov_coverage) F_aov_coverage[pix] = n_cov;
            }
            // fragment counters: warp reduce -> one shared slot per warp; thread 0 adds them up behind the next barrier
            {
                unsigned long long c = n_cov;
                uint32_t sh = (bkey != KEY_NONE) ? 1u : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { c += __shfl_down_sync(0xffffffffu, c, o); sh += __shfl_down_sync(0xffffffffu, sh, o); }
                if (FAST && LIGHTS == 2)
                {
                    // no local lights: nothing after this point needs the other warps of the tile, so the counters go out per warp and the
                    // tile has no barrier at all behind its raster loop
                    if (lane == 0)
                    {
                        DevStats* st = g.stats + (blockIdx.x & (STAT_SHARDS - 1));
                        if (c) atomicAdd(&st->frag_covered, c);
                        if (sh) atomicAdd(&st->frag_shaded, (unsigned long long)sh);
                    }
                }
                else if (lane == 0) { s_frag[warp][0] = c; s_frag[warp][1] = sh; }
            }
            const bool has = valid && bkey != KEY_NONE;
            const bool shade = !(F_shadow_mode || F_shader_id == 5 || !fb.hdr);

            // ---------------- phase A (per pixel): re-derive the winning fragment and run the builtin program
            const bool lit_shader = F_shader_id == 0 || F_shader_id == 1;
            float out_r = 0.0f, out_g = 0.0f, out_b = 0.0f;
            Surface surf;
            surf.P = v3(0, 0, 0); surf.N = v3(0, 1, 0); surf.V = v3(0, 1, 0); surf.albedo = v3(0, 0, 0);
            surf.metallic = 0.0f; surf.roughness = 1.0f; surf.blinn = F_shader_id == 1;
            surf.F0 = v3(0, 0, 0); surf.diffuse = v3(0, 0, 0);
            surf.a2 = surf.k = surf.NdotV = surf.g1v = surf.spec_a = surf.spec_b = 0.0f;
            if (has && shade)
            {
                // bit-identical to the coverage test above
                const float4* rsrc = reinterpret_cast<const float4*>(g.rrecs + bidx);
                const float4 r0 = __ldg(rsrc + 0), r1 = __ldg(rsrc + 1), r2 = __ldg(rsrc + 2), r3 = __ldg(rsrc + 3);
                const float v2x = xsub(pxf, r0.x), v2y = xsub(pyf, r0.y);
                const float bv = xmul(xsub(xmul(v2x, r1.y), xmul(r1.x, v2y)), r1.z);
                const float bw = xmul(xsub(xmul(r0.z, v2y), xmul(v2x, r0.w)), r1.z);
                const float bu = xsub(xsub(1.0f, bv), bw);
                const float denom = xadd(xadd(xmul(bu, r1.w), xmul(bv, r2.x)), xmul(bw, r2.y));
                const float inv_denom = xrcp(denom);
                float depth01 = bz;
                if (!F_has_depth)
                {
                    const float z_clip = xadd(xadd(xmul(bu, r2.z), xmul(bv, r2.w)), xmul(bw, r3.x));
                    depth01 = gclamp(xadd(xmul(xmul(z_clip, inv_denom), 0.5f), 0.5f), 0.0f, 1.0f);
                }
                const float4* ssrc = reinterpret_cast<const float4*>(g.srecs + bidx);
                float a[28];
#pragma unroll
                for (int i = 0; i < 7; ++i)
                {
                    const float4 q = __ldg(ssrc + i);
                    a[i * 4 + 0] = q.x; a[i * 4 + 1] = q.y; a[i * 4 + 2] = q.z; a[i * 4 + 3] = q.w;
                }
                // layout: wp[3][3] = a[0..8], n[3][3] = a[9..17], uv[3][2] = a[18..23], item = a[24]
                auto interp = [&](float c0, float c1, float c2) {
                    return xmul(xadd(xadd(xmul(bu, c0), xmul(bv, c1)), xmul(bw, c2)), inv_denom); // rasterizer.hpp:368
                };
                const F3 wpos{interp(a[0], a[3], a[6]), interp(a[1], a[4], a[7]), interp(a[2], a[5], a[8])};
                const F3 nrm_i{interp(a[9], a[12], a[15]), interp(a[10], a[13], a[16]), interp(a[11], a[14], a[17])};
                const float uvx = interp(a[18], a[20], a[22]), uvy = interp(a[19], a[21], a[23]);
                const DevItem& it = g.items[__float_as_uint(a[24])];
                const F3 n_ws = xnormalize3(nrm_i); // rasterizer.hpp:381
                if (F_write_motion && F_motion)
                {
                    // rasterizer.hpp:388-411: the winning fragment is the last one the serial loop lets past the depth test
                    const float4 pw = xmat4_mul(it.c2p, wpos.x, wpos.y, wpos.z, 1.0f);
                    const float4 cc = xmat4_mul(fc.viewproj, wpos.x, wpos.y, wpos.z, 1.0f);
                    const float4 pc = xmat4_mul(fc.prev_viewproj, pw.x, pw.y, pw.z, pw.w);
                    float mx = 0.0f, my = 0.0f;
                    if (fabsf(cc.w) > 1e-8f && fabsf(pc.w) > 1e-8f)
                    {
                        mx = xmul(xmul(xsub(xdiv(cc.x, cc.w), xdiv(pc.x, pc.w)), 0.5f), (float)fc.W);
                        my = xmul(xmul(xsub(xdiv(cc.y, cc.w), xdiv(pc.y, pc.w)), 0.5f), (float)fc.H);
                        const float len = xsqrt(xadd(xmul(mx, mx), xmul(my, my)));
                        if (len > 96.0f && len > 1e-6f) { const float k = xdiv(96.0f, len); mx = xmul(mx, k); my = xmul(my, k); }
                    }
                    F_motion[pix] = make_float2(mx, my);
                }

                if (F_shader_id == 2) { out_r = it.base_color[0]; out_g = it.base_color[1]; out_b = it.base_color[2]; }
                else if (F_shader_id == 3)
                {
                    const F3 n = xnormalize3(n_ws);
                    out_r = xadd(xmul(n.x, 0.5f), 0.5f); out_g = xadd(xmul(n.y, 0.5f), 0.5f); out_b = xadd(xmul(n.z, 0.5f), 0.5f);
                }
                else if (F_shader_id == 4) { const float d = sclamp(depth01, 0.0f, 1.0f); out_r = out_g = out_b = d; }
                else
                {
                    // ---- builtin lit programs.  N, L, NdotL stay exact: they feed the shadow bias and the NdotL > 0 branches.
                    const F3 Nx = xnormalize3(n_ws);
                    const F3 Lx = xnormalize3(F3{-fc.sun_dir[0], -fc.sun_dir[1], -fc.sun_dir[2]});
                    const F3 Vx = xnormalize3(F3{xsub(fc.camera_pos[0], wpos.x), xsub(fc.camera_pos[1], wpos.y), xsub(fc.camera_pos[2], wpos.z)});
                    const float NdotL = fmaxf(0.0f, xdot3(Nx, Lx));
                    const V3 N = toV3(Nx), L = toV3(Lx), V = toV3(Vx);
                    V3 albedo_tex = v3(1, 1, 1);
                    if (it.tex != 0u) albedo_tex = sample_texture(textures[it.tex - 1u], srgb_lut, uvx, uvy);
                    const V3 base = v3(it.base_color[0], it.base_color[1], it.base_color[2]);
                    const V3 albedo = v3(fmaxf(base.x * albedo_tex.x, 0.0f), fmaxf(base.y * albedo_tex.y, 0.0f), fmaxf(base.z * albedo_tex.z, 0.0f));
                    float shadow_vis = 1.0f;
                    if (F_shadow_map && NdotL > 0.0f)
                    {
                        shadow_vis = shadow_visibility(fc, wpos, NdotL);
                        shadow_vis = mixf(1.0f, shadow_vis, sat(fc.shadow_strength));
                    }
                    const V3 light_color = v3(fc.sun_color[0], fc.sun_color[1], fc.sun_color[2]);
                    const V3 H = normalize_fast(V + L);
                    V3 c;
                    surf.P = toV3(wpos); surf.N = N; surf.V = V; surf.albedo = albedo;
                    if (F_shader_id == 1)
                    {
                        // make_blinn_phong_program, builtin_shaders.hpp:105-152
                        const float NdotH = fmaxf(0.0f, dot(N, H));
                        const float rough = sat(it.roughness), metal = sat(it.metallic);
                        const float spec_pow = fmaxf(4.0f, 8.0f + (1.0f - rough) * 120.0f);
                        const float spec_norm = (spec_pow + 2.0f) / (2.0f * PI_F);
                        const float spec = powf(NdotH, spec_pow) * spec_norm * (0.04f + 0.96f * metal) * NdotL;
                        const V3 diffuse = albedo * ((1.0f - metal) * (NdotL / PI_F));
                        const V3 direct = (diffuse + v3(spec, spec, spec)) * light_color * (fc.sun_intensity * shadow_vis);
                        c = direct + fake_ibl(N, V, albedo, it.metallic, it.roughness, it.ao);
                        surf.metallic = metal; surf.roughness = rough;
                    }
                    else
                    {
                        // make_pbr_mr_program, builtin_shaders.hpp:154-214
                        const float NdotV = fmaxf(0.0f, dot(N, V)), NdotH = fmaxf(0.0f, dot(N, H)), VdotH = fmaxf(0.0f, dot(V, H));
                        const float rough = fminf(fmaxf(it.roughness, 0.04f), 1.0f), metal = sat(it.metallic);
                        const V3 F0 = mix3(v3(0.04f, 0.04f, 0.04f), albedo, metal);
                        const float aa = rough * rough, a2 = aa * aa;
                        const float denomD = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
                        const float D = a2 / (PI_F * denomD * denomD + 1e-7f);
                        const float k = (aa + 1.0f) * (aa + 1.0f) * 0.125f;
                        const float G = (NdotV / (NdotV * (1.0f - k) + k + 1e-7f)) * (NdotL / (NdotL * (1.0f - k) + k + 1e-7f));
                        const V3 F = F0 + (v3(1, 1, 1) - F0) * pow5(1.0f - VdotH);
                        const V3 spec = F * ((D * G) / fmaxf(4.0f * NdotL * NdotV, 1e-6f));
                        const V3 kd = (v3(1, 1, 1) - F) * (1.0f - metal);
                        const V3 diff = kd * albedo * (1.0f / PI_F);
                        V3 direct = v3(0, 0, 0);
                        if (NdotL > 0.0f && NdotV > 0.0f) direct = (diff + spec) * light_color * (fc.sun_intensity * NdotL * shadow_vis);
                        c = direct + fake_ibl(N, V, albedo, metal, rough, it.ao);
                        surf.metallic = metal; surf.roughness = rough;
                    }
                    out_r = c.x; out_g = c.y; out_b = c.z;
                    if (F_forward_plus) finish_surface(surf);
                }
            }

            PHASE_MARK(1); // resolve + phase A (surface)
            // the only barrier every non-empty tile passes: publishes the fragment counters (and tells phase B
            // whether any pixel of the tile is shaded at all)
            // (FAST Forward+ instantiations: the barrier of the position box below publishes the counters as well -- one barrier instead of two;
            // FAST without local lights: none, see above)
            int any_has = 1;
            if (!FAST) any_has = __syncthreads_or((has && shade) ? 1 : 0);
            if (!FAST && threadIdx.x == 0)
            {
                unsigned long long c = 0ull, sh = 0ull;
#pragma unroll
                for (int w = 0; w < TILE_THREADS / 32; ++w) { c += s_frag[w][0]; sh += s_frag[w][1]; }
                DevStats* st = g.stats + (blockIdx.x & (STAT_SHARDS - 1));
                if (c) atomicAdd(&st->frag_covered, c);
                if (sh) atomicAdd(&st->frag_shaded, sh);
            }
            if (!shade) return;

            PHASE_MARK(2); // barrier + stats
            // ---------------- phase B (whole CTA): Forward+ local lights, fp_stress_scene.frag:644-678.
            // Light tile == raster tile when the list tile size is 16.  The tile cell spans all depths, so most listed
            // lights cannot reach the surfaces actually visible in the tile: candidates are first tested against the
            // world-space AABB of the tile's shaded positions (conservative: a rejected light fails "dist < range"
            // at every pixel and would add exactly zero) and the survivors are compacted IN ASCENDING ORDER into
            // shared memory.  A staging round filters up to 1024 candidates (4 per thread) with two barriers; a
            // saturated list (count >= max_per_tile) walks ALL lights, as the GLSL does (:662-668).
            const bool use_lights = F_forward_plus && lit_shader && fc.n_lights > 0;
            if (use_lights && F_light_tile_size == (uint32_t)TILE)
            {
                if (any_has)
                {
                    // CTA-wide AABB of shaded world positions
                    const float INF = 3.0e38f;
                    float bx0 = has ? surf.P.x : INF, by0 = has ? surf.P.y : INF, bz0 = has ? surf.P.z : INF;
                    float bx1 = has ? surf.P.x : -INF, by1 = has ? surf.P.y : -INF, bz1 = has ? surf.P.z : -INF;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                    {
                        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o));
                        bz0 = fminf(bz0, __shfl_xor_sync(0xffffffffu, bz0, o)); bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
                        by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o)); bz1 = fmaxf(bz1, __shfl_xor_sync(0xffffffffu, bz1, o));
                    }
                    if (lane == 0) { s_box[warp][0] = bx0; s_box[warp][1] = by0; s_box[warp][2] = bz0; s_box[warp][3] = bx1; s_box[warp][4] = by1; s_box[warp][5] = bz1; }
                    __syncthreads();
                    if (FAST && threadIdx.x == 0)
                    {
                        unsigned long long c = 0ull, sh = 0ull;
#pragma unroll
                        for (int w = 0; w < TILE_THREADS / 32; ++w) { c += s_frag[w][0]; sh += s_frag[w][1]; }
                        DevStats* st = g.stats + (blockIdx.x & (STAT_SHARDS - 1));
                        if (c) atomicAdd(&st->frag_covered, c);
                        if (sh) atomicAdd(&st->frag_shaded, sh);
                    }
#pragma unroll
                    for (int w = 0; w < TILE_THREADS / 32; ++w)
                    {
                        bx0 = fminf(bx0, s_box[w][0]); by0 = fminf(by0, s_box[w][1]); bz0 = fminf(bz0, s_box[w][2]);
                        bx1 = fmaxf(bx1, s_box[w][3]); by1 = fmaxf(by1, s_box[w][4]); bz1 = fmaxf(bz1, s_box[w][5]);
                    }
                    if (FAST) any_has = bx0 <= bx1; // no shaded pixel in the tile: the box is empty (+INF .. -INF), CTA-uniform

                    const uint32_t list_id = (uint32_t)min(ty * TILE_H / TILE, (int)fc.light_tiles_y - 1) * fc.light_tiles_x + (uint32_t)min(tx, (int)fc.light_tiles_x - 1);
                    const uint32_t listed = min(fc.tile_counts[list_id], fc.max_per_tile);
                    const bool saturated = listed >= fc.max_per_tile;
                    const uint32_t n_src = !any_has ? 0u : (saturated ? fc.n_lights : listed); // (FAST: any_has is known only here)
                    const uint32_t* __restrict__ list = fc.tile_indices + (size_t)list_id * fc.max_per_tile;
                    V3 sum = v3(0, 0, 0);
                    for (uint32_t sbase = 0; sbase < n_src; sbase += CAND_PER_THREAD * TILE_THREADS)
                    {
                        if (sbase) __syncthreads(); // every warp has finished the previous round (s_ballot readers, light loop)
                        // ---- filter: candidate slot c of this thread is source position sbase + c*256 + tid
                        uint32_t cidx[CAND_PER_THREAD];
                        unsigned cmask[CAND_PER_THREAD];
#pragma unroll
                        for (int c = 0; c < CAND_PER_THREAD; ++c)
                        {
                            cidx[c] = 0xFFFFFFFFu;
                            cmask[c] = 0u;
                            if (sbase + (uint32_t)c * TILE_THREADS >= n_src) continue; // CTA-uniform
                            const uint32_t i = sbase + (uint32_t)c * TILE_THREADS + threadIdx.x;
                            bool keep = false;
                            if (i < n_src)
                            {
                                const uint32_t idx = saturated ? i : list[i];
                                if (idx < fc.n_lights)
                                {
                                    // pos_r2 = (centre, reach^2) of the digested record; reach^2 < 0 marks a disabled light
                                    const float4 q0 = __ldg(&fc.sm_lights[idx].pos_r2);
                                    const float dx = fmaxf(fmaxf(bx0 - q0.x, q0.x - bx1), 0.0f);
                                    const float dy = fmaxf(fmaxf(by0 - q0.y, q0.y - by1), 0.0f);
                                    const float dz = fmaxf(fmaxf(bz0 - q0.z, q0.z - bz1), 0.0f);
                                    keep = (dx * dx + dy * dy + dz * dz) <= q0.w * 1.001f + 1e-6f;
                                    if (keep) cidx[c] = idx;
                                }
                            }
                            cmask[c] = __ballot_sync(0xffffffffu, keep);
                            if (lane == 0) s_ballot[c * (TILE_THREADS / 32) + warp] = cmask[c];
                        }
                        __syncthreads(); // ballots visible
                        // ---- exclusive prefix over the 32 (slot, warp) ballots, computed redundantly by every warp
                        const bool slot_live = lane < CAND_PER_THREAD * TILE_WARPS && sbase + (uint32_t)(lane / TILE_WARPS) * TILE_THREADS < n_src;
                        const uint32_t cnt = slot_live ? (uint32_t)__popc(s_ballot[lane]) : 0u;
                        uint32_t incl = cnt;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1)
                        {
                            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += v;
                        }
                        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                        uint32_t dst[CAND_PER_THREAD];
#pragma unroll
                        for (int c = 0; c < CAND_PER_THREAD; ++c)
                        {
                            const uint32_t before = __shfl_sync(0xffffffffu, incl - cnt, c * (TILE_THREADS / 32) + warp);
                            dst[c] = before + (uint32_t)__popc(cmask[c] & ((1u << lane) - 1u));
                        }
                        for (uint32_t pass = 0; pass < total; pass += LIGHT_CAP)
                        {
                            if (pass) __syncthreads(); // the previous pass's light loop is done with s_light
#pragma unroll
                            for (int c = 0; c < CAND_PER_THREAD; ++c)
                            {
                                if (cidx[c] == 0xFFFFFFFFu || dst[c] < pass || dst[c] >= pass + LIGHT_CAP) continue;
                                // the 80-byte record was digested once at upload (light_prep_kernel): a straight copy
                                const float4* src = reinterpret_cast<const float4*>(fc.sm_lights + cidx[c]);
                                float4* dstp = reinterpret_cast<float4*>(s_light + (dst[c] - pass));
                                const float4 l0 = __ldg(src + 0), l1 = __ldg(src + 1), l2 = __ldg(src + 2), l3 = __ldg(src + 3), l4 = __ldg(src + 4);
                                dstp[0] = l0; dstp[1] = l1; dstp[2] = l2; dstp[3] = l3; dstp[4] = l4;
                            }
                            PHASE_MARK(3); // light staging (AABB, filter, compaction)
                            __syncthreads();
                            if (has)
                            {
                                const uint32_t kept = min(total - pass, (uint32_t)LIGHT_CAP);
                                const SmLight* lt = s_light;
                                for (uint32_t j = 0; j < kept; ++j, ++lt)
                                {
                                    const float4 q0 = lt->pos_r2;
                                    const float dx = q0.x - surf.P.x, dy = q0.y - surf.P.y, dz = q0.z - surf.P.z;
                                    const float dist2 = dx * dx + dy * dy + dz * dz;
                                    if (!(dist2 < q0.w) || !(dist2 > 1e-10f)) continue;
                                    const uint32_t kind = lt->kind;
                                    if (LIGHTS != 1 && (kind & KIND_AREA)) sum = sum + eval_light_record(surf, fc.lights + lt->index);
                                    else accumulate_point_spot(surf, lt, dx, dy, dz, dist2, kind, sum);
                                }
                            }
                        }
                    }
                    PHASE_MARK(4); // light loop (+ trailing staging rounds)
                    out_r += sum.x; out_g += sum.y; out_b += sum.z;
                }
            }
            else if (use_lights && has)
            {
                // generic light-tile size: per-pixel list lookup, tile_y counted from the top
                V3 sum = v3(0, 0, 0);
                const uint32_t ltx = min((uint32_t)px / F_light_tile_size, fc.light_tiles_x - 1u);
                const uint32_t lty = min((uint32_t)fy / F_light_tile_size, fc.light_tiles_y - 1u);
                const uint32_t list_id = lty * fc.light_tiles_x + ltx;
                const uint32_t cnt = min(fc.tile_counts[list_id], fc.max_per_tile);
                if (cnt >= fc.max_per_tile)
                {
                    for (uint32_t i = 0; i < fc.n_lights; ++i) sum = sum + eval_light_record(surf, fc.lights + i);
                }
                else
                {
                    for (uint32_t i = 0; i < cnt; ++i)
                    {
                        const uint32_t idx = fc.tile_indices[(size_t)list_id * fc.max_per_tile + i];
                        if (idx < fc.n_lights) sum = sum + eval_light_record(surf, fc.lights + idx);
                    }
                }
                out_r += sum.x; out_g += sum.y; out_b += sum.z;
            }

            // ---------------- phase C (per pixel): resolve colour (+ fused tonemap), each byte written once
            if (!valid) return;
            if (!has) { resolve_uncovered<FAST>(fc, fb, textures, srgb_lut, pix, px, py); return; }
            if (F_clear_motion && !F_write_motion && F_motion) F_motion[pix] = make_float2(0.0f, 0.0f); // plane cleared, vectors disabled
            PHASE_MARK(5);
            STORE_IF(out_r) fb.hdr[pix] = make_float4(out_r, out_g, out_b, 1.0f);
            if (fc.fuse_tonemap && fb.ldr) { const uchar4 l = tonemap_pixel(out_r, out_g, out_b, fc.exposure, fc.inv_gamma); STORE_IF(__uint_as_float((uint32_t)l.x * 0x01010101u + 0x7fc12000u)) fb.ldr[pix] = l; }
        }

        // Digests the 160-byte CullingLightGPU records into the 80-byte form the tile kernel's light loop reads
        // (run once per shsb_lights_upload, not per frame or per tile).
        // `src` may be pinned HOST memory (zero-copy): the upload is then done by the SMs, not by a copy engine, so it
        // never queues behind a multi-megabyte frame read-back on the same engine.  `raw` receives the verbatim record.
        __global__ void __launch_bounds__(128) light_prep_kernel(const DevLightRec* __restrict__ src, DevLightRec* __restrict__ raw, SmLight* __restrict__ out, uint32_t n)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n) return;
            {
                const uint4* s4 = reinterpret_cast<const uint4*>(src + i);
                uint4* d4 = reinterpret_cast<uint4*>(raw + i);
                uint4 w[10];
#pragma unroll
                for (int k = 0; k < 10; ++k) w[k] = s4[k];
#pragma unroll
                for (int k = 0; k < 10; ++k) d4[k] = w[k];
            }
            const DevLightRec* rec = raw + i; // this thread's own stores are visible to it
            const float4 pr = *reinterpret_cast<const float4*>(rec->position_range);
            const uint4 tf = *reinterpret_cast<const uint4*>(rec->type_shape_flags);
            const float4 ci = *reinterpret_cast<const float4*>(rec->color_intensity);
            const float4 sa = *reinterpret_cast<const float4*>(rec->shape_attenuation);
            SmLight sl;
            const float range = fmaxf(pr.w, 0.001f);
            const float power = fmaxf(sa.y, 0.001f);
            sl.pos_r2 = make_float4(pr.x, pr.y, pr.z, range * range);
            sl.radiance_ir = make_float4(ci.x * ci.w, ci.y * ci.w, ci.z * ci.w, 1.0f / range);
            sl.atten = make_float4(power, fmaxf(sa.w, 0.0f), fmaxf(sa.z, 1e-5f), 0.0f);
            sl.dir_outer = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            sl.kind = (tf.w == 0u ? 0u : (tf.w == 2u ? 2u : 1u)) | (power != 1.0f ? KIND_POW : 0u);
            sl.index = i; sl.pad0 = sl.pad1 = 0u;
            if (tf.x == 2u)
            {
                const float4 ds = *reinterpret_cast<const float4*>(rec->direction_spot);
                const float4 ax = *reinterpret_cast<const float4*>(rec->axis_spot_outer);
                const float dl = fast_rsqrt(ds.x * ds.x + ds.y * ds.y + ds.z * ds.z);
                const float inner_cos = fminf(fmaxf(ds.w, -1.0f), 1.0f);
                const float outer_cos = fminf(fmaxf(ax.w, -1.0f), inner_cos);
                sl.dir_outer = make_float4(ds.x * dl, ds.y * dl, ds.z * dl, outer_cos);
                sl.atten.w = 1.0f / fmaxf(inner_cos - outer_cos, 1e-6f);
                sl.kind |= KIND_SPOT;
            }
            else if (tf.x > 2u)
            {
                // Area lights light a point only if it is closer than `range` to the nearest point of the rectangle / segment
                // (eval_light_record: dist < range), so the pre-test sphere is (position, range + the emitter's half diagonal /
                // half length).  NOT the record's cull sphere: make_rect_area_culling_light stores direction_spot = -axis_z
                // (lighting/light_types.hpp:383) while its bounds extend along +axis_z, so the reference's cull bounds and the
                // side the GLSL lights do not coincide; the tile lists follow the bounds (like the reference), the fragment
                // loop follows the GLSL, and this filter must be conservative for the latter.
                const float4 ux = *reinterpret_cast<const float4*>(rec->up_shape_x);
                const float ext = (tf.x == 3u) ? sqrtf(fmaxf(ux.w, 1e-4f) * fmaxf(ux.w, 1e-4f) + fmaxf(sa.x, 1e-4f) * fmaxf(sa.x, 1e-4f)) : fmaxf(ux.w, 1e-4f);
                const float reach = (range + ext) * 1.001f + 1e-4f;
                sl.pos_r2 = make_float4(pr.x, pr.y, pr.z, reach * reach); // conservative pre-test; eval_light_record decides
                sl.kind |= KIND_AREA;
            }
            const bool enabled = (tf.z & 1u) != 0u && tf.x >= 1u && tf.x <= 4u;
            if (!enabled) sl.pos_r2.w = -1.0f; // fails every range test
            out[i] = sl;
        }

        __global__ void __launch_bounds__(256) tonemap_kernel(const float4* __restrict__ hdr, uchar4* __restrict__ ldr, int n, float exposure, float inv_gamma)
        {
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            {
                const float4 c = __ldg(hdr + i);
                ldr[i] = tonemap_pixel(c.x, c.y, c.z, exposure, inv_gamma);
            }
        }

        __global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n)
        {
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
        }

        __global__ void fill_f4_kernel(float4* p, float4 v, size_t n)
        {
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
        }
    }

#ifdef SHSB_PHASE_CLOCKS
    extern "C" __attribute__((visibility("default"))) int shsb_debug_phase_clocks(unsigned long long* out32, int reset)
    {
        cudaDeviceSynchronize();
        if (cudaMemcpyFromSymbol(out32, g_phase_clk, sizeof(unsigned long long) * 32) != cudaSuccess) return 1;
        if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z)); }
        return 0;
    }
#endif

    void launch_tile_raster(const FrameConst& fc, const Geometry& g, const FrameBuffers& fb, const DevTexture* textures,
                            const float* srgb_lut, cudaStream_t s, uint64_t* launches, bool allow_fast, int* mode_out)
    {
        const int n_tiles = fc.tiles_x * fc.tiles_y;
        if (n_tiles <= 0) return;
        // allow_fast = false (SHSB_NO_FAST_TILE=1 at context creation): always the general instantiation -- the A/B knob, and what a
        // caller sets who needs the LAST BITS of the colour planes to agree between frames of different modes: the instantiations share
        // their source, but the compiler contracts and schedules the colour arithmetic of each on its own, so HDR values may differ by
        // an ULP or two between them (depth, coverage, triangle ids and light lists are exact in both)
        const bool fast = allow_fast && !fc.shadow_mode && fc.has_depth && fc.linear_depth && !fc.load_depth && !fc.load_color && !fc.write_motion && !fc.clear_motion &&
                          !fc.hiz && fc.sky_kind == 0 && (fc.shader_id == 0 || fc.shader_id == 1) && !fb.aov_tri_id && !fb.aov_coverage && fb.hdr && fb.depth;
        const bool lights = fc.forward_plus && fc.n_lights > 0;
        const int light_mode = !fast ? 0 : ((lights && fc.light_tile_size == (uint32_t)TILE && !fc.shadow_map) ? 1 : (!lights ? 2 : 0));
        const bool blinn = fc.shader_id == 1;
        int prog = 0, lights_t = 0;
        if (light_mode == 1) { prog = blinn ? 2 : 1; lights_t = fc.area_lights ? 3 : 1; }
        else if (light_mode == 2) { prog = blinn ? 2 : 1; lights_t = 2; }
        if (mode_out) *mode_out = prog * 10 + lights_t;
        switch (prog * 10 + lights_t)
        {
            case 11: tile_kernel<1, 1><<<n_tiles, TILE_THREADS, 0, s>>>(fc, g, fb, textures, srgb_lut); break;
            case 21: tile_kernel<2, 1><<<n_tiles, TILE_THREADS, 0, s>>>(fc, g, fb, textures, srgb_lut); break;
            case 13: tile_kernel<1, 3><<<n_tiles, TILE_THREADS, 0, s>>>(fc, g, fb, textures, srgb_lut); break;
            case 23: tile_kernel<2, 3><<<n_tiles, TILE_THREADS, 0, s>>>(fc, g, fb, textures, srgb_lut); break;
            case 12: tile_kernel<1, 2><<<n_tiles, TILE_THREADS, 0, s>>>(fc, g, fb, textures, srgb_lut); break;
            case 22: tile_kernel<2, 2><<<n_tiles, TILE_THREADS, 0, s>>>(fc, g, fb, textures, srgb_lut); break;
            default: tile_kernel<0, 0><<<n_tiles, TILE_THREADS, 0, s>>>(fc, g, fb, textures, srgb_lut); break;
        }
        *launches += 1;
    }

    void launch_light_prep(const DevLightRec* src, DevLightRec* raw, SmLight* out, uint32_t n, cudaStream_t s, uint64_t* launches)
    {
        if (!n) return;
        light_prep_kernel<<<(n + 127) / 128, 128, 0, s>>>(src, raw, out, n);
        *launches += 1;
    }

    namespace
    {
        __global__ void __launch_bounds__(256) upload_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16)
        {
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
        }
    }

    // Small host -> device upload done by the SMs from mapped pinned memory (see light_prep_kernel).
    void launch_upload(void* dst, const void* src_mapped, size_t bytes, cudaStream_t s, uint64_t* launches)
    {
        const size_t n16 = (bytes + 15) / 16;
        if (!n16) return;
        const int grid = (int)min((n16 + 255) / 256, (size_t)64);
        upload_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src_mapped), n16);
        *launches += 1;
    }

    void launch_tonemap(const float4* hdr, uchar4* ldr, int n_pixels, float exposure, float inv_gamma, cudaStream_t s, uint64_t* launches)
    {
        if (n_pixels <= 0) return;
        const int grid = min((n_pixels + 255) / 256, 148 * 16);
        tonemap_kernel<<<grid, 256, 0, s>>>(hdr, ldr, n_pixels, exposure, inv_gamma);
        *launches += 1;
    }

    void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t s, uint64_t* launches)
    {
        if (!n) return;
        const int grid = (int)min((n + 255) / 256, (size_t)148 * 16);
        fill_u32_kernel<<<grid, 256, 0, s>>>(p, v, n);
        *launches += 1;
    }

    void launch_fill_f4(float4* p, float4 v, size_t n, cudaStream_t s, uint64_t* launches)
    {
        if (!n) return;
        const int grid = (int)min((n + 255) / 256, (size_t)148 * 16);
        fill_f4_kernel<<<grid, 256, 0, s>>>(p, v, n);
        *launches += 1;
    }
}
