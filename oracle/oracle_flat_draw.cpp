// oracle/oracle_flat_draw.cpp -- TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's flat-shaded software mesh draws, the
// consumer of the per-object light selections, each function citing the reference lines it follows
// (/root/reference/cpp-folders/src/shs-renderer-lib/include/shs/ and cpp-folders/src/exp-plumbing/).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it; the product path never does.
// PINNED against the reference's own headers compiled with the JoltPhysics declaration shim (oracle/ref_flat_draw_harness.cpp:
// shsref_flat_draw drives debug_draw::draw_mesh_blinn_phong_transformed, debug_draw::draw_filled_triangle and the four
// ILightModel::sample implementations; tests/test_flat_draw_cpu.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace
{
    struct F3 { float x = 0, y = 0, z = 0; };
    inline F3 mk(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
    inline F3 add(F3 a, F3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
    inline F3 sub(F3 a, F3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
    inline F3 mul(F3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
    inline F3 mulv(F3 a, F3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
    inline F3 divs(F3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
    inline F3 neg(F3 a) { return mk(-a.x, -a.y, -a.z); }
    inline float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    inline F3 cross3(F3 a, F3 b) { return mk(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
    inline F3 glm_normalize(F3 a) { return mul(a, 1.0f / std::sqrt(dot3(a, a))); } // glm::normalize = v * inversesqrt(dot(v, v))
    inline F3 normalize_or(F3 v, F3 fb)                                             // geometry/volumes.hpp:135-140
    {
        const float l2 = dot3(v, v);
        if (l2 <= 1e-10f) return fb;
        return mul(v, 1.0f / std::sqrt(l2));
    }
    // glm's scalar mat4 * vec4(p, 1): (m0 * x + m1 * y) + (m2 * z + m3 * w), column-major m
    inline void xform(const float* m, F3 p, float out[4])
    {
        for (int r = 0; r < 4; ++r) out[r] = (m[r] * p.x + m[4 + r] * p.y) + (m[8 + r] * p.z + m[12 + r] * 1.0f);
    }

    // ShsbLightProperties (include/shsb.h): shs::LightProperties (lighting/light_runtime.hpp:52-71) + the model's LightType
    struct Props
    {
        float color[3], intensity, position[3], range, direction[3], inner, right[3], outer, up[3], tube_half_length, rect_half[2], tube_radius, att_power, att_bias, att_cutoff;
        uint32_t att_model, flags, type, reserved[3];
    };
    static_assert(sizeof(Props) == 128, "ShsbLightProperties is 128 bytes");

    struct Contribution { F3 diffuse, specular; };

    // eval_distance_attenuation, lighting/light_runtime.hpp:182-210
    float distance_attenuation(const Props& p, float distance)
    {
        const float range = std::max(p.range, 0.001f);
        if (distance >= range) return 0.0f;
        const float norm = std::clamp(1.0f - distance / range, 0.0f, 1.0f);
        float falloff = 0.0f;
        switch (p.att_model)
        {
            case 0u: falloff = norm; break;
            case 1u: falloff = norm * norm * (3.0f - 2.0f * norm); break;
            case 2u:
            {
                const float denom = std::max(distance * distance, p.att_bias);
                const float inv = 1.0f / denom;
                falloff = std::min(1.0f, inv * (range * range)) * (norm * norm);
                break;
            }
            default: break;
        }
        falloff = std::pow(std::max(falloff, 0.0f), std::max(p.att_power, 0.001f));
        if (p.att_cutoff > 0.0f && falloff < p.att_cutoff) return 0.0f;
        return std::max(falloff, 0.0f);
    }

    // eval_local_light_brdf, lighting/light_runtime.hpp:212-237
    Contribution brdf(const Props& p, F3 L, float distance, float shaping, float spec_power, float spec_scale, F3 n, F3 v)
    {
        Contribution out{};
        const float ndotl = std::max(dot3(n, L), 0.0f);
        if (ndotl <= 0.0f) return out;
        const float att = distance_attenuation(p, distance) * std::max(shaping, 0.0f);
        if (att <= 0.0f) return out;
        const F3 colour = mk(std::max(p.color[0], 0.0f), std::max(p.color[1], 0.0f), std::max(p.color[2], 0.0f));
        const F3 radiance = mul(mul(colour, std::max(p.intensity, 0.0f)), att);
        const F3 H = normalize_or(add(L, v), L);
        const float ndoth = std::max(dot3(n, H), 0.0f);
        const float spec = spec_scale * std::pow(ndoth, spec_power);
        out.diffuse = mul(radiance, ndotl);
        out.specular = mul(radiance, spec);
        return out;
    }

    // ILightModel::sample: Point :310-316, Spot :358-382, RectArea :430-457, TubeArea :499-518
    Contribution sample(const Props& p, F3 pos, F3 n, F3 v)
    {
        const F3 lp = mk(p.position[0], p.position[1], p.position[2]);
        const F3 fwd0 = normalize_or(mk(p.direction[0], p.direction[1], p.direction[2]), mk(0.0f, -1.0f, 0.0f)); // safe_forward :132-135
        if (p.type == 1u || p.type == 2u)
        {
            const F3 to_light = sub(lp, pos);
            const float dist = std::sqrt(dot3(to_light, to_light));
            if (dist <= 1e-4f || dist > p.range) return {};
            const F3 L = divs(to_light, dist);
            if (p.type == 1u) return brdf(p, L, dist, 1.0f, 36.0f, 0.30f, n, v);
            const float half_pi = 1.5707963267948966f;
            const float inner = std::clamp(p.inner, 0.02f, half_pi - 0.02f);
            const float outer = std::clamp(std::max(inner + 0.005f, p.outer), inner + 0.005f, half_pi - 0.005f);
            const float cos_inner = std::cos(inner), cos_outer = std::cos(outer);
            const float cos_theta = dot3(neg(L), fwd0);
            if (cos_theta <= cos_outer) return {};
            float t = (cos_theta - cos_outer) / std::max(cos_inner - cos_outer, 1e-5f);
            t = std::clamp(t, 0.0f, 1.0f);
            return brdf(p, L, dist, t * t * (3.0f - 2.0f * t), 34.0f, 0.32f, n, v);
        }
        if (p.type == 3u)
        {
            // basis_from_forward_and_hint :137-151 (right_from_forward, camera/camera_math.hpp:28-31)
            const F3 fwd = normalize_or(fwd0, mk(0.0f, 0.0f, 1.0f));
            const F3 up_ref = normalize_or(mk(p.up[0], p.up[1], p.up[2]), mk(0.0f, 1.0f, 0.0f));
            F3 right = cross3(up_ref, fwd);
            right = normalize_or(right, glm_normalize(cross3(up_ref, fwd)));
            const F3 up = normalize_or(cross3(fwd, right), mk(0.0f, 1.0f, 0.0f));
            right = normalize_or(cross3(up, fwd), right);
            const float hx = std::max(p.rect_half[0], 0.05f), hy = std::max(p.rect_half[1], 0.05f);
            const F3 d = sub(pos, lp);
            const float ux = std::clamp(dot3(d, right), -hx, hx), uy = std::clamp(dot3(d, up), -hy, hy);
            const F3 emit = add(add(lp, mul(right, ux)), mul(up, uy));
            const F3 to_light = sub(emit, pos);
            const float dist = std::sqrt(dot3(to_light, to_light));
            if (dist <= 1e-4f || dist > p.range) return {};
            const F3 L = divs(to_light, dist);
            const float facing = std::max(dot3(fwd, neg(L)), 0.0f);
            if (facing <= 0.0f) return {};
            return brdf(p, L, dist, 0.65f + 0.55f * facing, 26.0f, 0.26f, n, v);
        }
        if (p.type == 4u)
        {
            const F3 axis = normalize_or(mk(p.right[0], p.right[1], p.right[2]), mk(1.0f, 0.0f, 0.0f));
            const float half_len = std::max(p.tube_half_length, 0.1f);
            const F3 a = sub(lp, mul(axis, half_len)), b = add(lp, mul(axis, half_len));
            // closest_point_on_segment :254-261
            const F3 ab = sub(b, a);
            const float denom = dot3(ab, ab);
            F3 emit = a;
            if (!(denom <= 1e-8f)) emit = add(a, mul(ab, std::clamp(dot3(sub(pos, a), ab) / denom, 0.0f, 1.0f)));
            const F3 to_light = sub(emit, pos);
            const float dist = std::sqrt(dot3(to_light, to_light));
            if (dist <= 1e-4f || dist > p.range) return {};
            const F3 L = divs(to_light, dist);
            const float soft = std::clamp(1.0f - dist / std::max(p.range, 0.1f), 0.0f, 1.0f);
            return brdf(p, L, dist, 0.75f + 0.35f * soft, 22.0f, 0.20f, n, v);
        }
        return {};
    }

    // project_world_to_screen, sw_render/debug_draw.hpp:41-58
    bool project(const float* world4, const float* vp, int w, int h, float s[2], float& z)
    {
        float clip[4];
        xform(vp, mk(world4[0], world4[1], world4[2]), clip);
        if (clip[3] <= 0.001f) return false;
        const float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
        if (nz < -1.0f || nz > 1.0f) return false;
        s[0] = (nx + 1.0f) * 0.5f * (float)w;
        s[1] = (ny + 1.0f) * 0.5f * (float)h;
        z = nz * 0.5f + 0.5f;
        return true;
    }

    inline float edge(const float* a, const float* b, float px, float py) { return (px - a[0]) * (b[1] - a[1]) - (py - a[1]) * (b[0] - a[0]); } // edge_fn :36-39

    // draw_filled_triangle, sw_render/debug_draw.hpp:60-112
    void fill_triangle(uint8_t* canvas, float* depth_buffer, int W, int H, const float* p0, float z0, const float* p1, float z1, const float* p2, float z2, const uint8_t rgba[4])
    {
        const float area = edge(p0, p1, p2[0], p2[1]);
        if (std::abs(area) <= 1e-6f) return;
        const int min_x = std::max(0, (int)std::floor(std::min(p0[0], std::min(p1[0], p2[0]))));
        const int min_y = std::max(0, (int)std::floor(std::min(p0[1], std::min(p1[1], p2[1]))));
        const int max_x = std::min(W - 1, (int)std::ceil(std::max(p0[0], std::max(p1[0], p2[0]))));
        const int max_y = std::min(H - 1, (int)std::ceil(std::max(p0[1], std::max(p1[1], p2[1]))));
        if (min_x > max_x || min_y > max_y) return;
        const bool ccw = area > 0.0f;
        for (int y = min_y; y <= max_y; ++y)
            for (int x = min_x; x <= max_x; ++x)
            {
                const float px = (float)x + 0.5f, py = (float)y + 0.5f;
                const float w0 = edge(p1, p2, px, py), w1 = edge(p2, p0, px, py), w2 = edge(p0, p1, px, py);
                const bool inside = ccw ? (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) : (w0 <= 0.0f && w1 <= 0.0f && w2 <= 0.0f);
                if (!inside) continue;
                const float depth = (w0 / area) * z0 + (w1 / area) * z1 + (w2 / area) * z2;
                if (depth < 0.0f || depth > 1.0f) continue;
                const size_t di = (size_t)y * (size_t)W + (size_t)x;
                if (depth < depth_buffer[di])
                {
                    depth_buffer[di] = depth;
                    std::memcpy(canvas + di * 4, rgba, 4);
                }
            }
    }

    inline uint8_t to_u8(float v) { return (uint8_t)std::clamp(v * 255.0f, 0.0f, 255.0f); }
    inline float clamp01_glm(float v) { return std::min(std::max(v, 0.0f), 1.0f); } // glm::clamp = min(max(x, lo), hi)
}

extern "C"
{
    // mode 0: debug_draw::draw_mesh_blinn_phong_transformed (sw_render/debug_draw.hpp:153-203) per draw;
    // mode 1: draw_mesh_multi_light_transformed (exp-plumbing/hello_light_types_culling_sw.cpp:366-422) per draw.
    // Draw d uses mesh draw_mesh[d] of (mesh_table3: first index, index count, base vertex; vertices; indices), model models16 + 16 d,
    // base colour base3 + 3 d and, in mode 1, the light selection sel_counts[d], sel_idx8 + 8 d.  lights128: ShsbLightProperties.
    // canvas_rgba (W x H x 4) and depth (W x H) are read and updated in place.
    int32_t shso_flat_draw(int32_t mode, uint32_t n_draws, const uint32_t* draw_mesh, const float* models16, const float* base3, const uint32_t* sel_counts,
                           const uint32_t* sel_idx8, const uint32_t* mesh_table3, uint32_t n_meshes, const float* vertices, uint32_t n_vertices, const uint32_t* indices,
                           uint32_t n_indices, const float view_proj[16], const float camera3[3], const float light_dir3[3], const void* lights128, uint32_t n_lights,
                           int32_t W, int32_t H, uint8_t* canvas_rgba, float* depth)
    {
        if (mode < 0 || mode > 1 || W <= 0 || H <= 0 || !view_proj || !camera3 || !canvas_rgba || !depth || (n_draws && (!draw_mesh || !models16 || !base3))) return 1;
        if (mode == 1 && n_draws && (!sel_counts || !sel_idx8)) return 1;
        if (mode == 0 && !light_dir3) return 1;
        const Props* lights = static_cast<const Props*>(lights128);
        const F3 camera = mk(camera3[0], camera3[1], camera3[2]);
        F3 Ldir{};
        if (mode == 0) Ldir = glm_normalize(neg(mk(light_dir3[0], light_dir3[1], light_dir3[2]))); // :165
        for (uint32_t d = 0; d < n_draws; ++d)
        {
            if (draw_mesh[d] >= n_meshes) return 1;
            const uint32_t first = mesh_table3[3 * draw_mesh[d]], count = mesh_table3[3 * draw_mesh[d] + 1], base_v = mesh_table3[3 * draw_mesh[d] + 2];
            if ((uint64_t)first + count > n_indices) return 1;
            const float* model = models16 + (size_t)d * 16;
            const F3 base = mk(base3[3 * d], base3[3 * d + 1], base3[3 * d + 2]);
            for (uint32_t i = 0; i + 2 < count; i += 3)
            {
                float world[3][4], s[3][2], z[3];
                bool ok = true;
                for (int k = 0; k < 3; ++k)
                {
                    const uint64_t vi = (uint64_t)base_v + indices[first + i + k];
                    if (vi >= n_vertices) { ok = false; break; } // the reference would read out of bounds: such triangles are skipped here and on the device
                    xform(model, mk(vertices[vi * 3], vertices[vi * 3 + 1], vertices[vi * 3 + 2]), world[k]);
                }
                if (!ok) continue;
                for (int k = 0; k < 3 && ok; ++k) ok = project(world[k], view_proj, W, H, s[k], z[k]);
                if (!ok) continue;
                const F3 p0 = mk(world[0][0], world[0][1], world[0][2]), p1 = mk(world[1][0], world[1][1], world[1][2]), p2 = mk(world[2][0], world[2][1], world[2][2]);
                F3 n = cross3(sub(p2, p0), sub(p1, p0));
                const float n2 = dot3(n, n);
                if (n2 <= 1e-10f) continue;
                n = mul(n, 1.0f / std::sqrt(n2));
                const F3 centroid = mul(add(add(p0, p1), p2), 1.0f / 3.0f);
                F3 lit;
                if (mode == 0)
                {
                    const F3 V = glm_normalize(sub(camera, centroid));
                    const F3 Hh = glm_normalize(add(Ldir, V));
                    const float ndotl = std::max(0.0f, dot3(n, Ldir)), ndoth = std::max(0.0f, dot3(n, Hh));
                    const float ambient = 0.18f, diffuse = 0.72f * ndotl;
                    const float specular = (ndotl > 0.0f) ? (0.35f * std::pow(ndoth, 32.0f)) : 0.0f;
                    lit = add(mul(base, ambient + diffuse), mk(specular, specular, specular));
                }
                else
                {
                    const F3 V = normalize_or(sub(camera, centroid), mk(0.0f, 0.0f, 1.0f));
                    const float hemi = 0.5f + 0.5f * std::clamp(n.y, -1.0f, 1.0f);
                    lit = mul(base, 0.22f + 0.12f * hemi); // kAmbientBase, kAmbientHemi (:55-56)
                    for (uint32_t si = 0; si < sel_counts[d] && si < 8u; ++si)
                    {
                        const uint32_t li = sel_idx8[(size_t)d * 8 + si];
                        if (li >= n_lights) continue;
                        const Contribution c = sample(lights[li], centroid, n, V);
                        lit = add(lit, add(mulv(base, c.diffuse), c.specular));
                    }
                }
                const uint8_t rgba[4] = {to_u8(clamp01_glm(lit.x)), to_u8(clamp01_glm(lit.y)), to_u8(clamp01_glm(lit.z)), 255};
                fill_triangle(canvas_rgba, depth, W, H, s[0], z[0], s[1], z[1], s[2], z[2], rgba);
            }
        }
        return 0;
    }
}
