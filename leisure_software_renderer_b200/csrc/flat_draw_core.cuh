// flat_draw_core.cuh -- the consumer of the per-object light selections (last link of the chain): the
// reference's flat-shaded software mesh draws, as host/device inline functions.  flat_draw.cu wraps them in kernels,
// tests/cpp/flat_draw_emul.cpp compiles the same functions with g++ and checks them against the pinned oracle on the CPU box (test
// code: the product launches kernels).
//   blinn_phong_colour   the per-triangle colour of debug_draw::draw_mesh_blinn_phong_transformed, sw_render/debug_draw.hpp:153-203
//   multi_light_colour   the per-triangle colour of draw_mesh_multi_light_transformed, exp-plumbing/hello_light_types_culling_sw.cpp:
//                        366-422: ambient + hemisphere term, then every light of the object's LightSelection through its light
//                        model's sample() (Point / Spot / RectArea / TubeArea, lighting/light_runtime.hpp:310-316, 358-382,
//                        430-457, 499-518) and eval_local_light_brdf / eval_distance_attenuation (:182-237)
// Projection, triangle setup and the per-texel depth are debug_draw::project_world_to_screen / draw_filled_triangle
// (sw_render/debug_draw.hpp:41-112), formula for formula those of geometry/culling_software.hpp: scene_cull_core.cuh's
// occ_project_vertex / occ_setup_triangle / occ_texel_depth serve both.
// IEEE binary32, unfused, GLM's scalar order (compiled --fmad=false / -ffp-contract=off).  std::pow / std::cos of float arguments are
// evaluated in double and rounded once: glibc's powf / cosf do the same internally, so the results agree except in the rare
// double-rounding case; a colour channel is floor(lit * 255) and moves only when the product crosses an integer there.
#pragma once
#include "scene_cull_core.cuh"

#ifndef FD_POWF
#define FD_POWF(x, y) ((float)pow((double)(x), (double)(y)))
#endif
#ifndef FD_COSF
#define FD_COSF(x) ((float)cos((double)(x)))
#endif

namespace shsb
{
    namespace fd
    {
        enum LightType { LIGHT_DIRECTIONAL = 0, LIGHT_POINT = 1, LIGHT_SPOT = 2, LIGHT_RECT_AREA = 3, LIGHT_TUBE_AREA = 4, LIGHT_PROBE = 5 }; // light_types.hpp:24-32
        enum Attenuation { ATT_LINEAR = 0, ATT_SMOOTH = 1, ATT_INVERSE_SQUARE = 2 };                                                    // light_types.hpp:49-54

        // ShsbLightProperties (include/shsb.h) = shs::LightProperties (light_runtime.hpp:52-71) + the LightType of the instance's model
        struct LightProps
        {
            float color[3], intensity;
            float position[3], range;
            float direction[3], inner_angle;
            float right[3], outer_angle;
            float up[3], tube_half_length;
            float rect_half_extents[2], tube_radius, attenuation_power;
            float attenuation_bias, attenuation_cutoff;
            uint32_t attenuation_model, flags;
            uint32_t light_type, reserved[3];
        };
        static_assert(sizeof(LightProps) == 128, "ShsbLightProperties is 128 bytes");

        enum Mode { MODE_BLINN_PHONG = 0, MODE_MULTI_LIGHT = 1 };

        // one draw of a batch: a DebugMesh (vertices + indices, geometry/jolt_debug_draw.hpp:36-52) with its model matrix, base colour
        // and, for the multi-light draw, the object's LightSelection (light_runtime.hpp:126-131)
        struct DrawRec
        {
            const float* positions;
            const uint32_t* indices;
            uint32_t n_positions, n_tris, tri_base, selection_count;
            float model[16];
            float base[3];
            uint32_t reserved;
            uint32_t selection[8];
        };

        struct BatchDesc
        {
            float view_proj[16];
            float camera[3];
            float L[3]; // glm::normalize(-light_dir_ws), Blinn-Phong draw only
            int W, H, mode;
            uint32_t n_draws, n_tris, n_lights;
        };

        struct V3 { float x, y, z; };
        SC_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
        SC_HD V3 v3(const float* p) { return v3(p[0], p[1], p[2]); }
        SC_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
        SC_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
        SC_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
        SC_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
        SC_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
        SC_HD V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
        SC_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                                   // glm: (x + y) + z
        SC_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); } // glm::cross
        SC_HD float length(V3 a) { return sqrtf(dot(a, a)); }
        SC_HD V3 glm_normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }                                        // v * inversesqrt(dot(v, v))
        SC_HD V3 max0(V3 a) { return v3(sc::gmax(a.x, 0.0f), sc::gmax(a.y, 0.0f), sc::gmax(a.z, 0.0f)); }
        SC_HD float clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }           // std::clamp
        SC_HD float clamp01_glm(float v) { return sc::gmin(sc::gmax(v, 0.0f), 1.0f); }                                // glm::clamp = min(max(x, lo), hi)

        SC_HD V3 normalize_or(V3 v, V3 fallback) // geometry/volumes.hpp:135-140
        {
            const float len2 = dot(v, v);
            if (len2 <= 1e-10f) return fallback;
            return v * (1.0f / sqrtf(len2));
        }

        struct Contribution { V3 diffuse, specular; };
        SC_HD Contribution no_light() { Contribution c; c.diffuse = v3(0, 0, 0); c.specular = v3(0, 0, 0); return c; }

        // eval_distance_attenuation, light_runtime.hpp:182-210
        SC_HD float distance_attenuation(const LightProps& p, float distance)
        {
            const float range = sc::gmax(p.range, 0.001f);
            if (distance >= range) return 0.0f;
            const float norm = clampf(1.0f - distance / range, 0.0f, 1.0f);
            float falloff = 0.0f;
            if (p.attenuation_model == ATT_LINEAR) falloff = norm;
            else if (p.attenuation_model == ATT_SMOOTH) falloff = norm * norm * (3.0f - 2.0f * norm);
            else if (p.attenuation_model == ATT_INVERSE_SQUARE)
            {
                const float denom = sc::gmax(distance * distance, p.attenuation_bias);
                const float inv = 1.0f / denom;
                const float range_norm = range * range;
                falloff = sc::gmin(1.0f, inv * range_norm) * (norm * norm);
            }
            falloff = FD_POWF(sc::gmax(falloff, 0.0f), sc::gmax(p.attenuation_power, 0.001f));
            if (p.attenuation_cutoff > 0.0f && falloff < p.attenuation_cutoff) return 0.0f;
            return sc::gmax(falloff, 0.0f);
        }

        // eval_local_light_brdf, light_runtime.hpp:212-237
        SC_HD Contribution local_light_brdf(const LightProps& p, V3 L, float distance, float shaping, float spec_power, float spec_scale, V3 n, V3 view_dir)
        {
            Contribution out = no_light();
            const float ndotl = sc::gmax(dot(n, L), 0.0f);
            if (ndotl <= 0.0f) return out;
            const float attenuation = distance_attenuation(p, distance) * sc::gmax(shaping, 0.0f);
            if (attenuation <= 0.0f) return out;
            const V3 radiance = max0(v3(p.color)) * sc::gmax(p.intensity, 0.0f) * attenuation;
            const V3 H = normalize_or(L + view_dir, L);
            const float ndoth = sc::gmax(dot(n, H), 0.0f);
            const float spec = spec_scale * FD_POWF(ndoth, spec_power);
            out.diffuse = radiance * ndotl;
            out.specular = radiance * spec;
            return out;
        }

        SC_HD V3 safe_forward(const LightProps& p) { return normalize_or(v3(p.direction), v3(0.0f, -1.0f, 0.0f)); } // :132-135

        // basis_from_forward_and_hint, light_runtime.hpp:137-151 (right_from_forward: camera/camera_math.hpp:28-31)
        SC_HD void basis_from_forward_and_hint(V3 forward, V3 up_hint, V3& right, V3& up, V3& fwd)
        {
            fwd = normalize_or(forward, v3(0.0f, 0.0f, 1.0f));
            const V3 up_ref = normalize_or(up_hint, v3(0.0f, 1.0f, 0.0f));
            right = cross(up_ref, fwd);
            right = normalize_or(right, glm_normalize(cross(up_ref, fwd)));
            up = normalize_or(cross(fwd, right), v3(0.0f, 1.0f, 0.0f));
            right = normalize_or(cross(up, fwd), right);
        }

        // ILightModel::sample of the four local light models
        SC_HD Contribution sample_light(const LightProps& p, V3 pos, V3 n, V3 view_dir)
        {
            if (p.light_type == LIGHT_POINT) // PointLightModel::sample, :310-316
            {
                const V3 to_light = v3(p.position) - pos;
                const float dist = length(to_light);
                if (dist <= 1e-4f || dist > p.range) return no_light();
                return local_light_brdf(p, to_light / dist, dist, 1.0f, 36.0f, 0.30f, n, view_dir);
            }
            if (p.light_type == LIGHT_SPOT) // SpotLightModel::sample, :358-382
            {
                const V3 to_light = v3(p.position) - pos;
                const float dist = length(to_light);
                if (dist <= 1e-4f || dist > p.range) return no_light();
                const V3 L = to_light / dist;
                const V3 dir = safe_forward(p);
                const float half_pi = 1.5707963267948966f;
                const float inner = clampf(p.inner_angle, 0.02f, half_pi - 0.02f);
                const float outer = clampf(sc::gmax(inner + 0.005f, p.outer_angle), inner + 0.005f, half_pi - 0.005f);
                const float cos_inner = FD_COSF(inner), cos_outer = FD_COSF(outer);
                const float cos_theta = dot(-L, dir);
                if (cos_theta <= cos_outer) return no_light();
                float t = (cos_theta - cos_outer) / sc::gmax(cos_inner - cos_outer, 1e-5f);
                t = clampf(t, 0.0f, 1.0f);
                const float shaping = t * t * (3.0f - 2.0f * t);
                return local_light_brdf(p, L, dist, shaping, 34.0f, 0.32f, n, view_dir);
            }
            if (p.light_type == LIGHT_RECT_AREA) // RectAreaLightModel::sample, :430-457
            {
                V3 right, up, fwd;
                basis_from_forward_and_hint(safe_forward(p), v3(p.up), right, up, fwd);
                const float hx = sc::gmax(p.rect_half_extents[0], 0.05f), hy = sc::gmax(p.rect_half_extents[1], 0.05f);
                const V3 d = pos - v3(p.position);
                const float ux = clampf(dot(d, right), -hx, hx);
                const float uy = clampf(dot(d, up), -hy, hy);
                const V3 emit_pt = v3(p.position) + right * ux + up * uy;
                const V3 to_light = emit_pt - pos;
                const float dist = length(to_light);
                if (dist <= 1e-4f || dist > p.range) return no_light();
                const V3 L = to_light / dist;
                const float emission_facing = sc::gmax(dot(fwd, -L), 0.0f);
                if (emission_facing <= 0.0f) return no_light();
                const float shape_gain = 0.65f + 0.55f * emission_facing;
                return local_light_brdf(p, L, dist, shape_gain, 26.0f, 0.26f, n, view_dir);
            }
            if (p.light_type == LIGHT_TUBE_AREA) // TubeAreaLightModel::sample, :499-518 (closest_point_on_segment :254-261)
            {
                const V3 axis = normalize_or(v3(p.right), v3(1.0f, 0.0f, 0.0f));
                const float half_len = sc::gmax(p.tube_half_length, 0.1f);
                const V3 a = v3(p.position) - axis * half_len;
                const V3 b = v3(p.position) + axis * half_len;
                const V3 ab = b - a;
                const float denom = dot(ab, ab);
                V3 emit_pt = a;
                if (!(denom <= 1e-8f))
                {
                    const float t = clampf(dot(pos - a, ab) / denom, 0.0f, 1.0f);
                    emit_pt = a + ab * t;
                }
                const V3 to_light = emit_pt - pos;
                const float dist = length(to_light);
                if (dist <= 1e-4f || dist > p.range) return no_light();
                const V3 L = to_light / dist;
                const float radial_softening = clampf(1.0f - dist / sc::gmax(p.range, 0.1f), 0.0f, 1.0f);
                const float shaping = 0.75f + 0.35f * radial_softening;
                return local_light_brdf(p, L, dist, shaping, 22.0f, 0.20f, n, view_dir);
            }
            return no_light(); // the demo registers no model for the other types
        }

        SC_HD uint32_t pack_colour(V3 lit) // glm::clamp(lit, 0, 1), then static_cast<uint8_t>(std::clamp(v * 255, 0, 255)); alpha 255
        {
            const float r = clampf(clamp01_glm(lit.x) * 255.0f, 0.0f, 255.0f), g = clampf(clamp01_glm(lit.y) * 255.0f, 0.0f, 255.0f),
                        b = clampf(clamp01_glm(lit.z) * 255.0f, 0.0f, 255.0f);
            return (uint32_t)(uint8_t)r | ((uint32_t)(uint8_t)g << 8) | ((uint32_t)(uint8_t)b << 16) | 0xFF000000u;
        }

        // project_world_to_screen, sw_render/debug_draw.hpp:41-58 (world: vec3(model * vec4(local, 1)))
        SC_HD bool project_world(const float* world, const float* view_proj, int width, int height, float xy[2], float& depth01)
        {
            float clip[4];
            sc::mul4(view_proj, world[0], world[1], world[2], 1.0f, clip);
            if (clip[3] <= 0.001f) return false;
            const float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
            if (nz < -1.0f || nz > 1.0f) return false;
            xy[0] = (nx + 1.0f) * 0.5f * (float)width;
            xy[1] = (ny + 1.0f) * 0.5f * (float)height;
            depth01 = nz * 0.5f + 0.5f;
            return true;
        }

        // world-space face normal shared by both draws; false = degenerate triangle (skipped by the reference)
        SC_HD bool face_normal(V3 p0, V3 p1, V3 p2, V3& n)
        {
            n = cross(p2 - p0, p1 - p0); // LH + clockwise front faces: the RH cross order flipped
            const float n2 = dot(n, n);
            if (n2 <= 1e-10f) return false;
            n = n * (1.0f / sqrtf(n2));
            return true;
        }

        // debug_draw.hpp:176-199.  L = glm::normalize(-light_dir_ws), taken once per draw by the caller.
        SC_HD uint32_t blinn_phong_colour(V3 p0, V3 p1, V3 p2, V3 n, V3 camera_pos, V3 L, V3 base)
        {
            const V3 centroid = (p0 + p1 + p2) * (1.0f / 3.0f);
            const V3 V = glm_normalize(camera_pos - centroid);
            const V3 H = glm_normalize(L + V);
            const float ndotl = sc::gmax(0.0f, dot(n, L));
            const float ndoth = sc::gmax(0.0f, dot(n, H));
            const float ambient = 0.18f;
            const float diffuse = 0.72f * ndotl;
            const float specular = (ndotl > 0.0f) ? (0.35f * FD_POWF(ndoth, 32.0f)) : 0.0f;
            const V3 lit = base * (ambient + diffuse) + v3(specular, specular, specular);
            return pack_colour(lit);
        }

        // hello_light_types_culling_sw.cpp:401-420.  selection: the object's LightSelection indices (entries >= n_lights are skipped).
        SC_HD uint32_t multi_light_colour(V3 p0, V3 p1, V3 p2, V3 n, V3 camera_pos, V3 base, const LightProps* lights, uint32_t n_lights,
                                          const uint32_t* selection, uint32_t selection_count)
        {
            const V3 centroid = (p0 + p1 + p2) * (1.0f / 3.0f);
            const V3 V = normalize_or(camera_pos - centroid, v3(0.0f, 0.0f, 1.0f));
            const float hemi = 0.5f + 0.5f * clampf(n.y, -1.0f, 1.0f);
            V3 lit = base * (0.22f + 0.12f * hemi); // kAmbientBase, kAmbientHemi (:55-56)
            for (uint32_t si = 0; si < selection_count; ++si)
            {
                const uint32_t li = selection[si];
                if (li >= n_lights) continue;
                const Contribution c = sample_light(lights[li], centroid, n, V);
                lit = lit + (base * c.diffuse + c.specular);
            }
            return pack_colour(lit);
        }
    }
}
