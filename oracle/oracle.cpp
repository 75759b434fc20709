// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
// product path (leisure_software_renderer_b200/, include/); only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load the library built from this file.
//
// CPU restatement of the reference's rasterization hot path, written from
// the algorithm, not from the reference's source text: plain float arrays, fixed-size polygons,
// explicit operation order.  Citations are relative to
// /root/reference/cpp-folders/src/shs-renderer-lib/include/shs/ .
//
// PINNING: the reference has no test, golden vector or fixture that touches a pixel
// (SURVEY.md section 4), so this restatement is pinned against outputs of the reference ITSELF:
// oracle/_ref/libshs_ref.so (oracle/ref_harness.cpp = the reference's own headers compiled
// from /root/reference) is run on the same inputs by tests/test_oracle_vs_reference.py
// (bit-exact HDR/depth/shadow/LDR/stats) and the resulting fixtures are committed under
// tests/golden/.  The tile / depth-range / clustered light lists (A11; lighting/jolt_light_culling.hpp +
// geometry/jolt_culling.hpp) and the temporal-AA adapter (pipeline/pass_adapters.hpp) are guarded by
// SHS_HAS_JOLT and need JoltPhysics v5.2.0 (absent): they are pinned against the reference's headers compiled
// with a JoltPhysics DECLARATION shim (oracle/jolt_shim, oracle/ref_lightcull_harness.cpp,
// oracle/ref_taa_harness.cpp; tests/test_light_cull_pinned_cpu.py, tests/test_post_passes_cpu.py) --
// the headers use Jolt only to turn a light's shape into bounds, and bounds are an input here.
// The Forward+ per-fragment local-light loop (A9) exists in the reference only as GLSL
// (shaders/vulkan/fp_stress_scene.frag:132-165, 421-523, 644-678; common/light_math.glsl:44-78): it is pinned against that
// text compiled as C++ (oracle/extract_glsl_a9.py lifts it with lexical rewrites only, oracle/ref_glsl_a9_harness.cpp compiles it
// against oracle/glsl_shim; tests/test_a9_pinned_cpu.py: per-light radiance, attenuation and the list walk bit for bit, plus the
// committed fixture tests/golden/golden_a9_glsl.npz).  What stays DEFINED here: how the CPU rasterizer's fragment (bottom-up rows)
// maps to the shader's gl_FragCoord (Vulkan: top-down) -- tile_y = (H-1-py) / tile_size -- and that the loop's radiance is ADDED to
// the CPU builtin program's colour (the shader's own ambient / sun terms are not the CPU path's).
// One part has NO compilable reference and stays "parity unpinned": the per-tile depth reduce (a GLSL compute shader on a
// different depth encoding in the reference, shaders/vulkan/fp_stress_depth_reduce.comp).
//
// All arithmetic is IEEE-754 binary32, round-to-nearest, no FMA (-ffp-contract=off), evaluated
// in the order the reference (and GLM's scalar path, see oracle/glm_shim/glm/glm.hpp) evaluates it.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle_abi.h"

namespace
{
    // ---------------------------------------------------------------- small float helpers
    struct V3 { float x, y, z; };
    struct V4 { float x, y, z, w; };

    inline float fmax_glm(float a, float b) { return (a < b) ? b : a; } // glm::max / std::max
    inline float fmin_glm(float a, float b) { return (b < a) ? b : a; } // glm::min / std::min
    inline float clampf(float x, float lo, float hi) { return fmin_glm(fmax_glm(x, lo), hi); } // glm::clamp
    inline float std_clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); } // std::clamp

    inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
    inline V3 add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
    inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
    inline V3 mul(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
    inline V3 scale(V3 a, float k) { return V3{a.x * k, a.y * k, a.z * k}; }
    inline V3 divs(V3 a, float k) { return V3{a.x / k, a.y / k, a.z / k}; }
    inline V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }
    inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // left to right
    inline V3 normalize3(V3 v) { return scale(v, 1.0f / std::sqrt(dot3(v, v))); }
    inline float length3(V3 v) { return std::sqrt(dot3(v, v)); }
    inline V3 cross3(V3 x, V3 y) { return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
    inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
    inline V3 mix3(V3 a, V3 b, float t) { return V3{mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)}; }
    inline V3 max3(V3 a, V3 b) { return V3{fmax_glm(a.x, b.x), fmax_glm(a.y, b.y), fmax_glm(a.z, b.z)}; }
    inline V3 load3(const float* p) { return V3{p[0], p[1], p[2]}; }

    // column-major mat4 * vec4, GLM scalar order: (m0*x + m1*y) + (m2*z + m3*w)
    inline V4 mat4_mul(const float* m, float x, float y, float z, float w)
    {
        V4 r;
        r.x = (m[0] * x + m[4] * y) + (m[8] * z + m[12] * w);
        r.y = (m[1] * x + m[5] * y) + (m[9] * z + m[13] * w);
        r.z = (m[2] * x + m[6] * y) + (m[10] * z + m[14] * w);
        r.w = (m[3] * x + m[7] * y) + (m[11] * z + m[15] * w);
        return r;
    }

    inline void mat4_mul_mat4(const float* a, const float* b, float* out)
    {
        float r[16];
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k)
                r[i * 4 + k] = a[0 + k] * b[i * 4 + 0] + a[4 + k] * b[i * 4 + 1] + a[8 + k] * b[i * 4 + 2] + a[12 + k] * b[i * 4 + 3];
        std::memcpy(out, r, sizeof(r));
    }

    // glm::inverse(mat4): cofactors, then * (1/det).  m[c*4+r].
    void mat4_inverse(const float* m, float* out)
    {
#define M(c, r) m[(c) * 4 + (r)]
        const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
        const float c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
        const float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
        const float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
        const float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
        const float c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
        const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
        const float c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
        const float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
        const float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
        const float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
        const float c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
        const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
        const float c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
        const float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
        const float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
        const float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
        const float c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
        const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
        const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
        const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)};
        const float v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
        const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)};
        const float v3_[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
        const float sa[4] = {+1.0f, -1.0f, +1.0f, -1.0f}, sb[4] = {-1.0f, +1.0f, -1.0f, +1.0f};
        float inv[16];
        for (int i = 0; i < 4; ++i)
        {
            inv[0 * 4 + i] = (v1[i] * f0[i] - v2[i] * f1[i] + v3_[i] * f2[i]) * sa[i];
            inv[1 * 4 + i] = (v0[i] * f0[i] - v2[i] * f3[i] + v3_[i] * f4[i]) * sb[i];
            inv[2 * 4 + i] = (v0[i] * f1[i] - v1[i] * f3[i] + v3_[i] * f5[i]) * sa[i];
            inv[3 * 4 + i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb[i];
        }
        const float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];
        const float det = (d0 + d1) + (d2 + d3);
        const float ood = 1.0f / det;
        for (int i = 0; i < 16; ++i) out[i] = inv[i] * ood;
#undef M
    }

    // glm::determinant(mat4): Laplace expansion along column 0 of the cofactors of rows 2,3 (func_matrix.inl compute_determinant<4,4>)
    float mat4_determinant(const float* m)
    {
#define M(c, r) m[(c) * 4 + (r)]
        const float s00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
        const float s01 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
        const float s02 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
        const float s03 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
        const float s04 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
        const float s05 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
        const float d0 = +(M(1, 1) * s00 - M(1, 2) * s01 + M(1, 3) * s02);
        const float d1 = -(M(1, 0) * s00 - M(1, 2) * s03 + M(1, 3) * s04);
        const float d2 = +(M(1, 0) * s01 - M(1, 1) * s03 + M(1, 3) * s05);
        const float d3 = -(M(1, 0) * s02 - M(1, 1) * s04 + M(1, 2) * s05);
        return M(0, 0) * d0 + M(0, 1) * d1 + M(0, 2) * d2 + M(0, 3) * d3;
#undef M
    }

    // normal matrix of make_default_vertex_out, shader/builtin_shaders.hpp:92-95:
    // mat3(model); if |det| > 1e-8 -> transpose(inverse(.)).  Returns column-major 3x3.
    void normal_matrix(const float* model, float* n9)
    {
        float m[9];
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) m[c * 3 + r] = model[c * 4 + r];
#define M(c, r) m[(c) * 3 + (r)]
        const float det =
            +M(0, 0) * (M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2))
            - M(1, 0) * (M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2))
            + M(2, 0) * (M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2));
        if (std::fabs(det) > 1e-8f)
        {
            const float ood = 1.0f / det;
            float inv[9];
#define I(c, r) inv[(c) * 3 + (r)]
            I(0, 0) = +(M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) * ood;
            I(1, 0) = -(M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2)) * ood;
            I(2, 0) = +(M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1)) * ood;
            I(0, 1) = -(M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2)) * ood;
            I(1, 1) = +(M(0, 0) * M(2, 2) - M(2, 0) * M(0, 2)) * ood;
            I(2, 1) = -(M(0, 0) * M(2, 1) - M(2, 0) * M(0, 1)) * ood;
            I(0, 2) = +(M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2)) * ood;
            I(1, 2) = -(M(0, 0) * M(1, 2) - M(1, 0) * M(0, 2)) * ood;
            I(2, 2) = +(M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1)) * ood;
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) n9[c * 3 + r] = I(r, c); // transpose
#undef I
        }
        else
        {
            std::memcpy(n9, m, sizeof(m));
        }
#undef M
    }

    // ---------------------------------------------------------------- vertex stage
    // What the builtin programs let the rasterizer see of a VertexOut (shader/types.hpp:54-64):
    // clip + the three semantic varyings that override the fixed fields (rasterizer.hpp:375-387).
    // a[0..2] = WorldPos.xyz, a[3..5] = NormalWS.xyz, a[6..7] = UV0.xy.
    enum { NATTR = 8 };
    struct Corner
    {
        float clip[4];
        float a[NATTR];
    };

    struct DrawConst
    {
        float model[16];
        float viewproj[16];
        float nrm[9];
        bool has_varyings; // false for the depth-prepass program (pass_adapters.hpp:335-353)
        // motion vectors (rasterizer.hpp:295-307): curr_to_prev_model = prev_model * inverse(model) when |det(model)| > 1e-10
        bool write_motion = false;
        float curr_to_prev_model[16];
        float prev_viewproj[16];
    };

    inline void set_motion(DrawConst& dc, bool enable, const float* prev_model, const float* prev_viewproj)
    {
        dc.write_motion = enable;
        if (!enable) return;
        const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        if (std::fabs(mat4_determinant(dc.model)) > 1e-10f)
        {
            float inv[16];
            mat4_inverse(dc.model, inv);
            mat4_mul_mat4(prev_model, inv, dc.curr_to_prev_model);
        }
        else std::memcpy(dc.curr_to_prev_model, ident, 64);
        std::memcpy(dc.prev_viewproj, prev_viewproj, 64);
    }

    inline Corner run_vs(const DrawConst& dc, const ShsoMesh& mesh, uint32_t idx)
    {
        // read_v, rasterizer.hpp:196-202
        const float* p = mesh.positions + (size_t)idx * 3;
        float nx = 0.0f, ny = 1.0f, nz = 0.0f, tu = 0.0f, tv = 0.0f;
        if (idx < mesh.n_normals) { nx = mesh.normals[(size_t)idx * 3 + 0]; ny = mesh.normals[(size_t)idx * 3 + 1]; nz = mesh.normals[(size_t)idx * 3 + 2]; }
        if (idx < mesh.n_uvs) { tu = mesh.uvs[(size_t)idx * 2 + 0]; tv = mesh.uvs[(size_t)idx * 2 + 1]; }
        Corner o;
        // make_default_vertex_out, builtin_shaders.hpp:87-103
        const V4 wp = mat4_mul(dc.model, p[0], p[1], p[2], 1.0f);
        const V4 cl = mat4_mul(dc.viewproj, wp.x, wp.y, wp.z, wp.w);
        o.clip[0] = cl.x; o.clip[1] = cl.y; o.clip[2] = cl.z; o.clip[3] = cl.w;
        if (dc.has_varyings)
        {
            const float* n = dc.nrm;
            V3 nn;
            nn.x = n[0] * nx + n[3] * ny + n[6] * nz;
            nn.y = n[1] * nx + n[4] * ny + n[7] * nz;
            nn.z = n[2] * nx + n[5] * ny + n[8] * nz;
            nn = normalize3(nn);
            o.a[0] = wp.x; o.a[1] = wp.y; o.a[2] = wp.z;
            o.a[3] = nn.x; o.a[4] = nn.y; o.a[5] = nn.z;
            o.a[6] = tu; o.a[7] = tv;
        }
        else
        {
            for (int i = 0; i < NATTR; ++i) o.a[i] = 0.0f;
        }
        return o;
    }

    // ---------------------------------------------------------------- clipping (rasterizer.hpp:69-164)
    inline bool corner_inside(const Corner& c)
    {
        const float x = c.clip[0], y = c.clip[1], z = c.clip[2], w = c.clip[3];
        if (!(w > 0.0f)) return false;
        return (x >= -w && x <= w) && (y >= -w && y <= w) && (z >= -w && z <= w);
    }

    inline float plane_dist(const Corner& c, int plane)
    {
        switch (plane)
        {
        case 0: return c.clip[0] + c.clip[3]; // left
        case 1: return c.clip[3] - c.clip[0]; // right
        case 2: return c.clip[1] + c.clip[3]; // bottom
        case 3: return c.clip[3] - c.clip[1]; // top
        case 4: return c.clip[2] + c.clip[3]; // near
        default: return c.clip[3] - c.clip[2]; // far
        }
    }

    inline Corner lerp_corner(const Corner& a, const Corner& b, float t)
    {
        Corner o;
        for (int i = 0; i < 4; ++i) o.clip[i] = mixf(a.clip[i], b.clip[i], t);
        for (int i = 0; i < NATTR; ++i) o.a[i] = mixf(a.a[i], b.a[i], t); // varyings are NOT renormalised (rasterizer.hpp:74)
        return o;
    }

    enum { MAX_POLY = 12 };

    int clip_polygon(const Corner* in, int n_in, Corner* out)
    {
        Corner bufA[MAX_POLY], bufB[MAX_POLY];
        Corner* src = bufA;
        Corner* dst = bufB;
        int n = n_in;
        for (int i = 0; i < n; ++i) src[i] = in[i];
        for (int plane = 0; plane < 6; ++plane)
        {
            int m = 0;
            for (int i = 0; i < n; ++i)
            {
                const Corner& cur = src[i];
                const Corner& nxt = src[(i + 1) % n];
                const float da = plane_dist(cur, plane);
                const float db = plane_dist(nxt, plane);
                const bool cur_in = da >= 0.0f;
                const bool nxt_in = db >= 0.0f;
                if (m + 2 > MAX_POLY) break; // a clipped triangle has at most 9 corners; guard only
                if (cur_in && nxt_in)
                {
                    dst[m++] = nxt;
                }
                else if (cur_in != nxt_in)
                {
                    const float denom = da - db;
                    if (std::fabs(denom) > 1e-8f) dst[m++] = lerp_corner(cur, nxt, da / denom);
                    if (nxt_in) dst[m++] = nxt;
                }
            }
            std::swap(src, dst);
            n = m;
            if (n == 0) break;
        }
        for (int i = 0; i < n; ++i) out[i] = src[i];
        return n;
    }

    // ---------------------------------------------------------------- fragment programs
    struct FsEnv
    {
        int shader_id;
        V3 light_dir_ws, light_color, camera_pos, base_color;
        float light_intensity, metallic, roughness, ao;
        const ShsoTexture* tex;
        const float* shadow;
        int shadow_w, shadow_h;
        float light_viewproj[16];
        float bias_const, bias_slope, pcf_step, shadow_strength;
        int pcf_radius;
        // Forward+ (A9)
        bool forward_plus;
        const uint8_t* lights;
        uint32_t n_lights;
        const uint32_t* tile_counts;
        const uint32_t* tile_indices;
        uint32_t tile_size, max_per_tile, tiles_x, tiles_y;
        int H;
    };

    struct Frag
    {
        V3 world_pos, normal_ws;
        float u, v, depth01;
        int px, py;
    };

    inline V3 srgb_to_linear(const uint8_t* t)
    {
        // builtin_shaders.hpp:25-31
        return V3{std::pow((float)t[0] / 255.0f, 2.2f), std::pow((float)t[1] / 255.0f, 2.2f), std::pow((float)t[2] / 255.0f, 2.2f)};
    }

    V3 sample_bilinear_repeat(const ShsoTexture* tex, float uvx, float uvy)
    {
        // builtin_shaders.hpp:33-55
        if (!tex || tex->w <= 0 || tex->h <= 0 || !tex->rgba) return V3{1.0f, 1.0f, 1.0f};
        const float u = uvx - std::floor(uvx);
        const float v = uvy - std::floor(uvy);
        const float fx = u * (float)(tex->w - 1);
        const float fy = v * (float)(tex->h - 1);
        const int x0 = (int)std::floor(fx);
        const int y0 = (int)std::floor(fy);
        const int x1 = std::min(x0 + 1, tex->w - 1);
        const int y1 = std::min(y0 + 1, tex->h - 1);
        const float tx = fx - (float)x0;
        const float ty = fy - (float)y0;
        auto at = [&](int x, int y) { return tex->rgba + ((size_t)y * (size_t)tex->w + (size_t)x) * 4; };
        const V3 c00 = srgb_to_linear(at(x0, y0));
        const V3 c10 = srgb_to_linear(at(x1, y0));
        const V3 c01 = srgb_to_linear(at(x0, y1));
        const V3 c11 = srgb_to_linear(at(x1, y1));
        return mix3(mix3(c00, c10, tx), mix3(c01, c11, tx), ty);
    }

    // ---------------------------------------------------------------- sky models (Scene::sky)
    // ProceduralSky::sample, sky/procedural_sky.hpp:25-45.  sun = normalize(ctor argument).
    V3 sky_procedural(V3 dir, V3 sun)
    {
        const V3 d = normalize3(dir);
        const float t = clampf(d.y * 0.5f + 0.5f, 0.0f, 1.0f);
        const V3 zenith{0.05f, 0.20f, 0.50f}, horizon{0.30f, 0.60f, 1.00f};
        V3 sky = mix3(horizon, zenith, t);
        const float sun_dot = dot3(d, V3{-sun.x, -sun.y, -sun.z});
        if (sun_dot > 0.9998f) sky = V3{15.0f, 15.0f, 15.0f};
        else if (sun_dot > 0.9990f)
        {
            const float glow = (sun_dot - 0.9990f) / (0.9998f - 0.9990f);
            sky = mix3(sky, V3{10.0f, 8.0f, 4.0f}, glow);
        }
        return sky;
    }

    // sample_face_bilinear_linear, sky/cubemap_sky.hpp:39-60 (clamped, not repeated; same per-tap pow as the material sampler)
    V3 sample_face_clamped(const ShsoTexture* tex, float u, float v)
    {
        if (!tex || tex->w <= 0 || tex->h <= 0 || !tex->rgba) return V3{0.0f, 0.0f, 0.0f};
        u = clampf(u, 0.0f, 1.0f);
        v = clampf(v, 0.0f, 1.0f);
        const float fx = u * (float)(tex->w - 1);
        const float fy = v * (float)(tex->h - 1);
        const int x0 = (int)std::floor(fx);
        const int y0 = (int)std::floor(fy);
        const int x1 = std::min(x0 + 1, tex->w - 1);
        const int y1 = std::min(y0 + 1, tex->h - 1);
        const float tx = fx - (float)x0;
        const float ty = fy - (float)y0;
        auto at = [&](int x, int y) { return tex->rgba + ((size_t)y * (size_t)tex->w + (size_t)x) * 4; };
        const V3 v00 = srgb_to_linear(at(x0, y0));
        const V3 v10 = srgb_to_linear(at(x1, y0));
        const V3 v01 = srgb_to_linear(at(x0, y1));
        const V3 v11 = srgb_to_linear(at(x1, y1));
        return mix3(mix3(v00, v10, tx), mix3(v01, v11, tx), ty);
    }

    // CubemapSky::sample, sky/cubemap_sky.hpp:69-109.  faces: +X -X +Y -Y +Z -Z; nullptr face => CubemapData::valid() false.
    V3 sky_cubemap(V3 dir, const ShsoTexture* const faces[6], float intensity)
    {
        for (int i = 0; i < 6; ++i) if (!faces[i] || faces[i]->w <= 0 || faces[i]->h <= 0 || !faces[i]->rgba) return V3{0.0f, 0.0f, 0.0f};
        V3 d = dir;
        const float len = std::sqrt(dot3(d, d));
        if (len < 1e-8f) return V3{0.0f, 0.0f, 0.0f};
        d.x /= len; d.y /= len; d.z /= len;
        const float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
        int face = 0;
        float u = 0.5f, v = 0.5f;
        if (ax >= ay && ax >= az)
        {
            if (d.x > 0.0f) { face = 0; u = (-d.z / ax); v = (d.y / ax); }
            else { face = 1; u = (d.z / ax); v = (d.y / ax); }
        }
        else if (ay >= ax && ay >= az)
        {
            if (d.y > 0.0f) { face = 2; u = (d.x / ay); v = (-d.z / ay); }
            else { face = 3; u = (d.x / ay); v = (d.z / ay); }
        }
        else
        {
            if (d.z > 0.0f) { face = 4; u = (d.x / az); v = (d.y / az); }
            else { face = 5; u = (-d.x / az); v = (d.y / az); }
        }
        u = 0.5f * (u + 1.0f);
        v = 0.5f * (v + 1.0f);
        return scale(sample_face_clamped(faces[face], u, v), intensity);
    }

    V3 fake_ibl(V3 N, V3 V, V3 base_color, float metallic, float roughness, float ao)
    {
        // eval_fake_ibl, builtin_shaders.hpp:57-85
        const V3 n = normalize3(N);
        const V3 v = normalize3(V);
        const V3 mv = neg(v);
        const V3 r = sub(mv, scale(scale(n, dot3(n, mv)), 2.0f)); // glm::reflect(-v, n) = I - N*dot(N,I)*2
        const V3 sky_zenith{0.32f, 0.46f, 0.72f}, sky_horizon{0.62f, 0.66f, 0.72f}, ground{0.16f, 0.15f, 0.14f};
        const float up_n = std_clampf(n.y * 0.5f + 0.5f, 0.0f, 1.0f);
        const float up_r = std_clampf(r.y * 0.5f + 0.5f, 0.0f, 1.0f);
        const V3 env_n = mix3(ground, mix3(sky_horizon, sky_zenith, up_n), up_n);
        const V3 env_r = mix3(ground, mix3(sky_horizon, sky_zenith, up_r), up_r);
        const float m = std_clampf(metallic, 0.0f, 1.0f);
        const float rgh = std_clampf(roughness, 0.0f, 1.0f);
        const V3 F0 = mix3(V3{0.04f, 0.04f, 0.04f}, max3(base_color, V3{0, 0, 0}), m);
        const float fres = std::pow(1.0f - std::max(0.0f, dot3(n, v)), 5.0f);
        const V3 one{1.0f, 1.0f, 1.0f};
        const V3 F = add(F0, scale(sub(one, F0), fres));
        const V3 kd = scale(sub(one, F), 1.0f - m);
        const V3 diffuse_ibl = scale(mul(mul(kd, base_color), env_n), 0.12f);
        const float spec_strength = 0.02f + (1.0f - rgh) * 0.18f;
        const V3 spec_ibl = scale(mul(env_r, F), spec_strength);
        return scale(add(diffuse_ibl, spec_ibl), std_clampf(ao, 0.0f, 1.0f));
    }

    float shadow_visibility(const FsEnv& e, V3 pos_ws, float ndotl)
    {
        // shadow_visibility_dir, lighting/shadow_sample.hpp:31-104
        const V4 p = mat4_mul(e.light_viewproj, pos_ws.x, pos_ws.y, pos_ws.z, 1.0f);
        if (std::fabs(p.w) < 1e-8f) return 1.0f;
        const float nx = p.x / p.w, ny = p.y / p.w, nz = p.z / p.w;
        const float u = nx * 0.5f + 0.5f, v = ny * 0.5f + 0.5f, z = nz * 0.5f + 0.5f;
        if (u < 0.0f || u > 1.0f || v < 0.0f || v > 1.0f) return 1.0f;
        const float slope = 1.0f - std_clampf(ndotl, 0.0f, 1.0f);
        const float bias = e.bias_const + e.bias_slope * slope;
        const float z_test = z - bias;
        const float fx = u * (float)(e.shadow_w - 1);
        const float fy = v * (float)(e.shadow_h - 1);
        const int cx = (int)std::round(fx);
        const int cy = (int)std::round(fy);
        // builtin_shaders.hpp:137-138: radius = max(0, r), step = max(1.0f, step); then shadow_sample.hpp:86,93
        const int r = std::max(0, std::max(0, e.pcf_radius));
        auto fetch = [&](int x, int y) {
            x = std::clamp(x, 0, e.shadow_w - 1);
            y = std::clamp(y, 0, e.shadow_h - 1);
            return e.shadow[(size_t)y * (size_t)e.shadow_w + (size_t)x];
        };
        if (r == 0) return (z_test <= fetch(cx, cy)) ? 1.0f : 0.0f;
        const int step = std::max(1, (int)std::round(std::max(1.0f, e.pcf_step)));
        int count = 0, lit = 0;
        for (int oy = -r; oy <= r; ++oy)
            for (int ox = -r; ox <= r; ++ox)
            {
                lit += (z_test <= fetch(cx + ox * step, cy + oy * step)) ? 1 : 0;
                ++count;
            }
        return (count > 0) ? (float)lit / (float)count : 1.0f;
    }

    // ---- Forward+ local lights (A9).  PINNED against the GLSL compiled as C++ (tests/test_a9_pinned_cpu.py): semantics follow
    // fp_stress_scene.frag:421-523 (eval_local_light), :132-165 (eval_pbr_light /
    // eval_blinn_phong_light), common/light_math.glsl:44-78 (attenuation_quadratic).
    struct LightRec // CullingLightGPU, lighting/light_types.hpp:141-167
    {
        float position_range[4], color_intensity[4], direction_spot[4], axis_spot_outer[4], up_shape_x[4], shape_attenuation[4];
        uint32_t type_shape_flags[4];
        float cull_sphere[4], cull_aabb_min[4], cull_aabb_max[4];
    };
    static_assert(sizeof(LightRec) == SHSB_LIGHT_RECORD_BYTES, "CullingLightGPU is 160 bytes");

    float attenuation_quadratic(float distance, float range, uint32_t model, float power, float bias, float cutoff)
    {
        const float safe_range = std::max(range, 1e-4f);
        const float t = std_clampf(distance / safe_range, 0.0f, 1.0f);
        const float edge = 1.0f - t;
        float falloff;
        if (model == 0u) falloff = edge; // Linear
        else if (model == 2u)            // InverseSquare
        {
            const float denom = std::max(distance * distance, std::max(bias, 1e-5f));
            const float inv = (safe_range * safe_range) / denom;
            falloff = inv * edge * edge;
        }
        else falloff = edge * edge;      // Smooth
        falloff = std::pow(std::max(falloff, 0.0f), std::max(power, 0.001f));
        if (falloff <= std::max(cutoff, 0.0f)) return 0.0f;
        return falloff;
    }

    V3 eval_pbr_light(V3 N, V3 V, V3 L, V3 radiance, V3 albedo, float metallic, float roughness)
    {
        const float NdotL = std::max(dot3(N, L), 0.0f);
        if (NdotL <= 0.0f) return V3{0, 0, 0};
        const V3 H = normalize3(add(V, L));
        const V3 F0 = mix3(V3{0.04f, 0.04f, 0.04f}, albedo, metallic);
        const float fres = std::pow(1.0f - std::max(dot3(H, V), 0.0f), 5.0f);
        const V3 one{1.0f, 1.0f, 1.0f};
        const V3 F = add(F0, scale(sub(one, F0), fres));
        const float a = roughness * roughness;
        const float a2 = a * a;
        const float NdotH = std::max(dot3(N, H), 0.0f);
        const float dd = (NdotH * NdotH) * (a2 - 1.0f) + 1.0f;
        const float NDF = a2 / std::max(3.14159265358979323846f * dd * dd, 1e-6f);
        const float NdotV = std::max(dot3(N, V), 0.0f);
        const float rr = roughness + 1.0f;
        const float k = (rr * rr) / 8.0f;
        const float g1 = NdotV / std::max(NdotV * (1.0f - k) + k, 1e-6f);
        const float g2 = NdotL / std::max(NdotL * (1.0f - k) + k, 1e-6f);
        const float G = g1 * g2;
        const float denom = std::max(4.0f * NdotV * NdotL, 1e-6f);
        const V3 specular = divs(scale(F, NDF * G), denom);
        const V3 kD = scale(sub(one, F), 1.0f - metallic);
        const V3 diff = divs(mul(kD, albedo), 3.14159265358979323846f);
        return scale(mul(add(diff, specular), radiance), NdotL);
    }

    V3 eval_blinn_light(V3 N, V3 V, V3 L, V3 radiance, V3 albedo, float metallic, float roughness)
    {
        const float NdotL = std::max(dot3(N, L), 0.0f);
        if (NdotL <= 0.0f) return V3{0, 0, 0};
        const V3 H = normalize3(add(V, L));
        const float smooth = 1.0f - std_clampf(roughness, 0.0f, 1.0f);
        const float shininess = mixf(10.0f, 96.0f, smooth);
        const float spec = std::pow(std::max(dot3(N, H), 0.0f), shininess);
        const V3 spec_color = mix3(V3{0.04f, 0.04f, 0.04f}, albedo, metallic);
        const float spec_strength = mixf(0.15f, 0.65f, smooth);
        const V3 d = scale(divs(albedo, 3.14159265358979323846f), NdotL);
        const V3 s = scale(scale(spec_color, spec), spec_strength);
        return mul(radiance, add(d, s));
    }

    V3 eval_local_light(const FsEnv& e, uint32_t idx, V3 P, V3 N, V3 V, V3 albedo, float metallic, float roughness, bool blinn)
    {
        LightRec lt;
        std::memcpy(&lt, e.lights + (size_t)idx * SHSB_LIGHT_RECORD_BYTES, sizeof(lt));
        const uint32_t type = lt.type_shape_flags[0], flags = lt.type_shape_flags[2], att_model = lt.type_shape_flags[3];
        const V3 zero{0, 0, 0};
        if ((flags & 1u) == 0u) return zero;
        if (type < 1u || type > 4u) return zero;
        const V3 lpos = load3(lt.position_range);
        V3 sample_pos = lpos;
        const float range = std::max(lt.position_range[3], 0.001f);
        V3 to_light = sub(sample_pos, P);
        float dist = length3(to_light);
        float rect_forward = 0.0f;
        if (type == 3u) // RectArea
        {
            const V3 right = normalize3(load3(lt.axis_spot_outer));
            const V3 up = normalize3(load3(lt.up_shape_x));
            const V3 emit = normalize3(load3(lt.direction_spot));
            const float hx = std::max(lt.up_shape_x[3], 1e-4f), hy = std::max(lt.shape_attenuation[0], 1e-4f);
            const V3 rel = sub(P, lpos);
            rect_forward = dot3(rel, emit);
            if (rect_forward <= 1e-4f || rect_forward >= range) return zero;
            const float lx = dot3(rel, right), ly = dot3(rel, up);
            const float x = std_clampf(lx, -hx, hx), y = std_clampf(ly, -hy, hy);
            const float dx = std::max(std::fabs(lx) - hx, 0.0f), dy = std::max(std::fabs(ly) - hy, 0.0f);
            if (length3(V3{dx, dy, rect_forward}) >= range) return zero;
            sample_pos = add(add(lpos, scale(right, x)), scale(up, y));
            to_light = sub(sample_pos, P);
            dist = length3(to_light);
        }
        else if (type == 4u) // TubeArea
        {
            const V3 axis = normalize3(load3(lt.axis_spot_outer));
            const float half_len = std::max(lt.up_shape_x[3], 1e-4f);
            const V3 p0 = sub(lpos, scale(axis, half_len)), p1 = add(lpos, scale(axis, half_len));
            const V3 seg = sub(p1, p0);
            const float seg_len2 = std::max(dot3(seg, seg), 1e-6f);
            const float uu = std_clampf(dot3(sub(P, p0), seg) / seg_len2, 0.0f, 1.0f);
            sample_pos = add(p0, scale(seg, uu));
            to_light = sub(sample_pos, P);
            dist = length3(to_light);
        }
        if (dist <= 1e-5f || dist >= range) return zero;
        const V3 L = divs(to_light, std::max(dist, 1e-5f));
        float atten = attenuation_quadratic(dist, range, att_model, lt.shape_attenuation[1], lt.shape_attenuation[2], lt.shape_attenuation[3]);
        if (atten <= 0.0f) return zero;
        if (type == 2u) // Spot
        {
            const V3 spot_dir = normalize3(load3(lt.direction_spot));
            const float inner_cos = std_clampf(lt.direction_spot[3], -1.0f, 1.0f);
            const float outer_cos = std_clampf(lt.axis_spot_outer[3], -1.0f, inner_cos);
            const float cone_cos = dot3(spot_dir, neg(L));
            const float t = std_clampf((cone_cos - outer_cos) / std::max(inner_cos - outer_cos, 1e-6f), 0.0f, 1.0f);
            const float spot = t * t * (3.0f - 2.0f * t);
            if (spot <= 0.0f) return zero;
            atten *= spot;
        }
        else if (type == 3u)
        {
            const V3 emit = normalize3(load3(lt.direction_spot));
            const float one_sided = std::max(dot3(emit, neg(L)), 0.0f);
            if (one_sided <= 0.0f) return zero;
            const float ff = std_clampf(1.0f - rect_forward / std::max(range, 1e-4f), 0.0f, 1.0f);
            atten *= one_sided * ff;
        }
        else if (type == 4u)
        {
            const float tube_radius = std::max(lt.shape_attenuation[0], 1e-4f);
            const float edge_soften = std_clampf(tube_radius / std::max(range, 1e-4f), 0.05f, 1.0f);
            atten *= mixf(0.65f, 1.0f, edge_soften);
        }
        const V3 radiance = scale(scale(load3(lt.color_intensity), lt.color_intensity[3]), atten);
        return blinn ? eval_blinn_light(N, V, L, radiance, albedo, metallic, roughness)
                     : eval_pbr_light(N, V, L, radiance, albedo, metallic, roughness);
    }

    V3 forward_plus_lights(const FsEnv& e, const Frag& f, V3 N, V3 V, V3 albedo, float metallic, float roughness, bool blinn)
    {
        // fp_stress_scene.frag:644-678 with the tile map fixed by SURVEY.md 8a A9:
        // tile_x = px / ts, tile_y = (H-1-py) / ts (tiles are top-origin, pixels bottom-origin).
        V3 sum{0, 0, 0};
        uint32_t tx = (uint32_t)f.px / e.tile_size;
        uint32_t ty = (uint32_t)(e.H - 1 - f.py) / e.tile_size;
        tx = std::min(tx, e.tiles_x - 1u);
        ty = std::min(ty, e.tiles_y - 1u);
        const uint32_t list_id = ty * e.tiles_x + tx;
        const uint32_t count = std::min(e.tile_counts[list_id], e.max_per_tile);
        const size_t base = (size_t)list_id * e.max_per_tile;
        if (count >= e.max_per_tile)
        {
            for (uint32_t i = 0; i < e.n_lights; ++i) sum = add(sum, eval_local_light(e, i, f.world_pos, N, V, albedo, metallic, roughness, blinn));
        }
        else
        {
            for (uint32_t i = 0; i < count; ++i)
            {
                const uint32_t idx = e.tile_indices[base + i];
                if (idx >= e.n_lights) continue;
                sum = add(sum, eval_local_light(e, idx, f.world_pos, N, V, albedo, metallic, roughness, blinn));
            }
        }
        return sum;
    }

    void run_fs(const FsEnv& e, const Frag& f, float* out4)
    {
        out4[3] = 1.0f;
        switch (e.shader_id)
        {
        case SHSB_SHADER_DEPTH_ONLY:
            out4[0] = out4[1] = out4[2] = 0.0f; // pass_adapters.hpp:345-351
            return;
        case SHSB_SHADER_DEBUG_ALBEDO:
            out4[0] = e.base_color.x; out4[1] = e.base_color.y; out4[2] = e.base_color.z;
            return;
        case SHSB_SHADER_DEBUG_NORMAL:
        {
            const V3 n = normalize3(f.normal_ws);
            out4[0] = n.x * 0.5f + 0.5f; out4[1] = n.y * 0.5f + 0.5f; out4[2] = n.z * 0.5f + 0.5f;
            return;
        }
        case SHSB_SHADER_DEBUG_DEPTH:
        {
            const float d = std_clampf(f.depth01, 0.0f, 1.0f);
            out4[0] = out4[1] = out4[2] = d;
            return;
        }
        default: break;
        }

        const V3 albedo_tex = sample_bilinear_repeat(e.tex, f.u, f.v);
        const V3 albedo = max3(mul(e.base_color, albedo_tex), V3{0, 0, 0});
        const V3 N = normalize3(f.normal_ws);
        const V3 L = normalize3(neg(e.light_dir_ws));
        const V3 V = normalize3(sub(e.camera_pos, f.world_pos));
        V3 c;
        float fp_metal, fp_rough;
        if (e.shader_id == SHSB_SHADER_BLINN_PHONG)
        {
            // make_blinn_phong_program, builtin_shaders.hpp:105-152
            const V3 H = normalize3(add(L, V));
            const float NdotL = std::max(0.0f, dot3(N, L));
            const float NdotH = std::max(0.0f, dot3(N, H));
            const float rough = std_clampf(e.roughness, 0.0f, 1.0f);
            const float metal = std_clampf(e.metallic, 0.0f, 1.0f);
            const float spec_pow = std::max(4.0f, 8.0f + (1.0f - rough) * 120.0f);
            const float spec_norm = (spec_pow + 2.0f) / (2.0f * 3.14159265358979323846f);
            const float spec_f0 = 0.04f + 0.96f * metal;
            const float spec = std::pow(NdotH, spec_pow) * spec_norm * spec_f0 * NdotL;
            const float kd = 1.0f - metal;
            const V3 diffuse = scale(scale(albedo, kd), NdotL / 3.14159265358979323846f); // kd*albedo: vec3(kd)*albedo
            float shadow_vis = 1.0f;
            if (e.shadow && NdotL > 0.0f)
            {
                shadow_vis = shadow_visibility(e, f.world_pos, NdotL);
                shadow_vis = mixf(1.0f, shadow_vis, std_clampf(e.shadow_strength, 0.0f, 1.0f));
            }
            const V3 ds = add(diffuse, V3{spec, spec, spec});
            const V3 direct = scale(scale(mul(ds, e.light_color), e.light_intensity), shadow_vis);
            const V3 ibl = fake_ibl(N, V, albedo, e.metallic, e.roughness, e.ao);
            c = add(direct, ibl);
            fp_metal = metal;
            fp_rough = rough;
        }
        else
        {
            // make_pbr_mr_program, builtin_shaders.hpp:154-214
            const V3 H = normalize3(add(V, L));
            const float NdotL = std::max(0.0f, dot3(N, L));
            const float NdotV = std::max(0.0f, dot3(N, V));
            const float NdotH = std::max(0.0f, dot3(N, H));
            const float VdotH = std::max(0.0f, dot3(V, H));
            const float rough = std_clampf(e.roughness, 0.04f, 1.0f);
            const float metal = std_clampf(e.metallic, 0.0f, 1.0f);
            const V3 F0 = mix3(V3{0.04f, 0.04f, 0.04f}, albedo, metal);
            const float a = rough * rough;
            const float a2 = a * a;
            const float denomD = (NdotH * NdotH) * (a2 - 1.0f) + 1.0f;
            const float D = a2 / (3.14159265358979323846f * denomD * denomD + 1e-7f);
            const float k = ((a + 1.0f) * (a + 1.0f)) * 0.125f;
            const float g1v = NdotV / (NdotV * (1.0f - k) + k + 1e-7f);
            const float g1l = NdotL / (NdotL * (1.0f - k) + k + 1e-7f);
            const float G = g1v * g1l;
            const V3 one{1.0f, 1.0f, 1.0f};
            const V3 F = add(F0, scale(sub(one, F0), std::pow(1.0f - VdotH, 5.0f)));
            const V3 spec = divs(scale(F, D * G), std::max(4.0f * NdotL * NdotV, 1e-6f)); // (D*G)*F / max(..)
            const V3 kd = scale(sub(one, F), 1.0f - metal);
            const V3 diff = scale(mul(kd, albedo), 1.0f / 3.14159265358979323846f);
            const V3 radiance = scale(e.light_color, e.light_intensity);
            float shadow_vis = 1.0f;
            if (e.shadow && NdotL > 0.0f)
            {
                shadow_vis = shadow_visibility(e, f.world_pos, NdotL);
                shadow_vis = mixf(1.0f, shadow_vis, std_clampf(e.shadow_strength, 0.0f, 1.0f));
            }
            V3 direct{0, 0, 0};
            if (NdotL > 0.0f && NdotV > 0.0f) direct = scale(scale(mul(add(diff, spec), radiance), NdotL), shadow_vis);
            const V3 ibl = fake_ibl(N, V, albedo, metal, rough, e.ao);
            c = add(direct, ibl);
            fp_metal = metal;
            fp_rough = rough;
        }
        if (e.forward_plus)
        {
            c = add(c, forward_plus_lights(e, f, N, V, albedo, fp_metal, fp_rough, e.shader_id == SHSB_SHADER_BLINN_PHONG));
        }
        out4[0] = c.x; out4[1] = c.y; out4[2] = c.z;
    }

    // ---------------------------------------------------------------- the draw (rasterizer.hpp:181-442)
    void draw_mesh(const ShsoMesh& mesh, const DrawConst& dc, const FsEnv& env, const ShsoTarget& tgt,
                   int cull_mode, bool front_face_ccw, uint32_t key_base, ShsbStats& st)
    {
        const int W = tgt.w, H = tgt.h;
        if (!tgt.hdr || W <= 0 || H <= 0 || mesh.n_positions == 0) return;
        const bool indexed = mesh.n_indices != 0;
        const size_t tri_count = indexed ? (mesh.n_indices / 3) : (mesh.n_positions / 3);
        const float fw1 = (float)(W - 1), fh1 = (float)(H - 1);
        const bool linear_depth = tgt.depth && (tgt.zf > tgt.zn + 1e-6f);

        for (size_t ti = 0; ti < tri_count; ++ti)
        {
            st.tri_input++;
            uint32_t i0, i1, i2;
            if (indexed) { i0 = mesh.indices[ti * 3]; i1 = mesh.indices[ti * 3 + 1]; i2 = mesh.indices[ti * 3 + 2]; }
            else { i0 = (uint32_t)(ti * 3); i1 = i0 + 1; i2 = i0 + 2; }
            if (i0 >= mesh.n_positions || i1 >= mesh.n_positions || i2 >= mesh.n_positions) continue;

            Corner poly[MAX_POLY];
            poly[0] = run_vs(dc, mesh, i0);
            poly[1] = run_vs(dc, mesh, i1);
            poly[2] = run_vs(dc, mesh, i2);
            int n = 3;
            if (!(corner_inside(poly[0]) && corner_inside(poly[1]) && corner_inside(poly[2])))
            {
                Corner in[3] = {poly[0], poly[1], poly[2]};
                n = clip_polygon(in, 3, poly);
            }
            if (n < 3) continue;

            for (int k = 1; k + 1 < n; ++k)
            {
                st.tri_after_clip++;
                const Corner* c[3] = {&poly[0], &poly[k], &poly[k + 1]};
                float sx[3], sy[3], ndcz[3];
                bool finite = true;
                for (int j = 0; j < 3; ++j)
                {
                    const float w = c[j]->clip[3];
                    const float nx = c[j]->clip[0] / w, ny = c[j]->clip[1] / w, nz = c[j]->clip[2] / w;
                    if (!std::isfinite(nx) || !std::isfinite(ny) || !std::isfinite(nz)) { finite = false; break; }
                    sx[j] = (nx * 0.5f + 0.5f) * fw1;
                    sy[j] = (ny * 0.5f + 0.5f) * fh1;
                    ndcz[j] = nz;
                }
                if (!finite) continue;
                (void)ndcz;

                const float e0x = sx[1] - sx[0], e0y = sy[1] - sy[0];
                const float e1x = sx[2] - sx[0], e1y = sy[2] - sy[0];
                const float area2 = e0x * e1y - e0y * e1x;
                if (std::fabs(area2) < 1e-10f) continue;
                const bool ccw = area2 > 0.0f;
                const bool is_front = (ccw == front_face_ccw);
                if (cull_mode == SHSB_CULL_BACK && !is_front) continue;
                if (cull_mode == SHSB_CULL_FRONT && is_front) continue;

                const float minxf = std::min({sx[0], sx[1], sx[2]}), maxxf = std::max({sx[0], sx[1], sx[2]});
                const float minyf = std::min({sy[0], sy[1], sy[2]}), maxyf = std::max({sy[0], sy[1], sy[2]});
                const int minx = std::max(0, (int)std::floor(minxf));
                const int maxx = std::min(W - 1, (int)std::ceil(maxxf));
                const int miny = std::max(0, (int)std::floor(minyf));
                const int maxy = std::min(H - 1, (int)std::ceil(maxyf));
                if (minx > maxx || miny > maxy) continue;
                st.tri_raster++;

                float iw[3], zw[3], aw[3][NATTR];
                for (int j = 0; j < 3; ++j)
                {
                    iw[j] = 1.0f / c[j]->clip[3];
                    zw[j] = c[j]->clip[2] * iw[j];
                    for (int a = 0; a < NATTR; ++a) aw[j][a] = c[j]->a[a] * iw[j];
                }

                // barycentric_2d constants (rasterizer.hpp:169-174): v0 = b-a, v1 = c-a
                const float v0x = sx[1] - sx[0], v0y = sy[1] - sy[0];
                const float v1x = sx[2] - sx[0], v1y = sy[2] - sy[0];
                const float den = v0x * v1y - v1x * v0y;
                const bool degenerate = std::fabs(den) < 1e-8f;
                const float inv_den = 1.0f / den;
                const uint32_t key = key_base + (uint32_t)ti * 8u + (uint32_t)(k - 1);

                for (int y = miny; y <= maxy; ++y)
                {
                    for (int x = minx; x <= maxx; ++x)
                    {
                        if (degenerate) continue; // bc = (-1,-1,-1) -> rejected
                        const float px = (float)x + 0.5f, py = (float)y + 0.5f;
                        const float v2x = px - sx[0], v2y = py - sy[0];
                        const float bv = (v2x * v1y - v1x * v2y) * inv_den;
                        const float bw = (v0x * v2y - v2x * v0y) * inv_den;
                        const float bu = 1.0f - bv - bw;
                        if (bu < 0.0f || bv < 0.0f || bw < 0.0f) continue;

                        const float denom = bu * iw[0] + bv * iw[1] + bw * iw[2];
                        if (denom <= 1e-10f) continue;
                        const float inv_denom = 1.0f / denom;
                        const size_t pix = (size_t)y * (size_t)W + (size_t)x;
                        if (tgt.coverage) tgt.coverage[pix] += 1u;
                        st.frag_covered++;

                        const float z_clip = bu * zw[0] + bv * zw[1] + bw * zw[2];
                        const float z_ndc = z_clip * inv_denom;
                        float z01 = clampf(z_ndc * 0.5f + 0.5f, 0.0f, 1.0f);
                        if (tgt.depth)
                        {
                            if (linear_depth)
                            {
                                const float view_z = 1.0f / denom;
                                z01 = clampf((view_z - tgt.zn) / (tgt.zf - tgt.zn), 0.0f, 1.0f);
                            }
                            float& zbuf = tgt.depth[pix];
                            if (z01 >= zbuf) continue;
                            zbuf = z01;
                        }
                        if (tgt.tri_id) tgt.tri_id[pix] = key;

                        Frag f;
                        float at[NATTR];
                        for (int a = 0; a < NATTR; ++a) at[a] = (bu * aw[0][a] + bv * aw[1][a] + bw * aw[2][a]) * inv_denom;
                        f.px = x; f.py = y; f.depth01 = z01;
                        if (dc.has_varyings)
                        {
                            f.world_pos = V3{at[0], at[1], at[2]};
                            f.normal_ws = normalize3(V3{at[3], at[4], at[5]});
                            f.u = at[6]; f.v = at[7];
                        }
                        else
                        {
                            f.world_pos = V3{0, 0, 0}; f.normal_ws = V3{0, 1, 0}; f.u = f.v = 0.0f;
                        }
                        if (dc.write_motion && tgt.motion && tgt.depth)
                        {
                            // rasterizer.hpp:388-411
                            const V4 pw = mat4_mul(dc.curr_to_prev_model, f.world_pos.x, f.world_pos.y, f.world_pos.z, 1.0f);
                            const V4 cc = mat4_mul(dc.viewproj, f.world_pos.x, f.world_pos.y, f.world_pos.z, 1.0f);
                            const V4 pc = mat4_mul(dc.prev_viewproj, pw.x, pw.y, pw.z, pw.w);
                            float mx = 0.0f, my = 0.0f;
                            if (std::fabs(cc.w) > 1e-8f && std::fabs(pc.w) > 1e-8f)
                            {
                                const float cnx = cc.x / cc.w, cny = cc.y / cc.w, pnx = pc.x / pc.w, pny = pc.y / pc.w;
                                mx = ((cnx - pnx) * 0.5f) * (float)W;
                                my = ((cny - pny) * 0.5f) * (float)H;
                                const float len = std::sqrt(mx * mx + my * my);
                                const float max_vel = 96.0f;
                                if (len > max_vel && len > 1e-6f) { const float k = max_vel / len; mx *= k; my *= k; }
                            }
                            tgt.motion[pix * 2 + 0] = mx;
                            tgt.motion[pix * 2 + 1] = my;
                        }
                        run_fs(env, f, tgt.hdr + pix * 4);
                    }
                }
            }
        }
    }

    void fill_env_from_uniforms(FsEnv& e, int shader_id, const ShsbUniforms* u, const ShsoAssets* assets, const ShsoTarget* tgt)
    {
        std::memset(&e, 0, sizeof(e));
        e.shader_id = shader_id;
        e.light_dir_ws = load3(u->light_dir_ws);
        e.light_color = load3(u->light_color);
        e.camera_pos = load3(u->camera_pos);
        e.base_color = load3(u->base_color);
        e.light_intensity = u->light_intensity;
        e.metallic = u->metallic;
        e.roughness = u->roughness;
        e.ao = u->ao;
        e.tex = (assets && u->base_color_tex >= 1 && u->base_color_tex <= assets->n_textures) ? &assets->textures[u->base_color_tex - 1] : nullptr;
        e.shadow = (u->shadow_map && tgt->shadow) ? tgt->shadow : nullptr;
        e.shadow_w = tgt->shadow_w;
        e.shadow_h = tgt->shadow_h;
        std::memcpy(e.light_viewproj, u->light_viewproj, 64);
        e.bias_const = u->shadow_bias_const;
        e.bias_slope = u->shadow_bias_slope;
        e.pcf_radius = u->shadow_pcf_radius;
        e.pcf_step = u->shadow_pcf_step;
        e.shadow_strength = u->shadow_strength;
        e.H = tgt->h;
    }

    const ShsoMesh* find_mesh(const ShsoAssets* a, shsb_mesh h)
    {
        if (!a || h == 0 || h > a->n_meshes) return nullptr;
        return &a->meshes[h - 1];
    }

    void model_from_transform(const ShsbTransform* tr, float* m)
    {
        // glm::translate(I, pos); rotate x, y, z; scale -- ext/matrix_transform.inl operation order.
        float cur[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        auto col = [&](float* mm, int c) { return mm + c * 4; };
        {
            // Result[3] = m[0]*v0 + m[1]*v1 + m[2]*v2 + m[3]
            float r3[4];
            for (int i = 0; i < 4; ++i) r3[i] = col(cur, 0)[i] * tr->pos[0] + col(cur, 1)[i] * tr->pos[1] + col(cur, 2)[i] * tr->pos[2] + col(cur, 3)[i];
            std::memcpy(col(cur, 3), r3, 16);
        }
        for (int ax = 0; ax < 3; ++ax)
        {
            const float a = tr->rot_euler[ax];
            const float c = std::cos(a), s = std::sin(a);
            V3 axis = normalize3(V3{ax == 0 ? 1.0f : 0.0f, ax == 1 ? 1.0f : 0.0f, ax == 2 ? 1.0f : 0.0f});
            const float av[3] = {axis.x, axis.y, axis.z};
            const float tp[3] = {(1.0f - c) * av[0], (1.0f - c) * av[1], (1.0f - c) * av[2]};
            float R[3][3];
            R[0][0] = c + tp[0] * av[0];
            R[0][1] = tp[0] * av[1] + s * av[2];
            R[0][2] = tp[0] * av[2] - s * av[1];
            R[1][0] = tp[1] * av[0] - s * av[2];
            R[1][1] = c + tp[1] * av[1];
            R[1][2] = tp[1] * av[2] + s * av[0];
            R[2][0] = tp[2] * av[0] + s * av[1];
            R[2][1] = tp[2] * av[1] - s * av[0];
            R[2][2] = c + tp[2] * av[2];
            float res[16];
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 4; ++i)
                    res[j * 4 + i] = col(cur, 0)[i] * R[j][0] + col(cur, 1)[i] * R[j][1] + col(cur, 2)[i] * R[j][2];
            std::memcpy(res + 12, col(cur, 3), 16);
            std::memcpy(cur, res, 64);
        }
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 4; ++i) cur[j * 4 + i] = cur[j * 4 + i] * tr->scl[j];
        std::memcpy(m, cur, 64);
    }

    struct ItemMaterial { V3 base_color; float metallic, roughness, ao; shsb_tex tex; };

    ItemMaterial resolve_material(const ShsbRenderItem& it)
    {
        // pass_pbr_forward.hpp:166-184
        ItemMaterial m;
        if (it.has_material) { m.base_color = load3(it.base_color); m.metallic = it.metallic; m.roughness = it.roughness; m.ao = it.ao; m.tex = it.base_color_tex; }
        else { m.base_color = V3{0.8f, 0.5f, 0.2f}; m.metallic = 0.1f; m.roughness = 0.5f; m.ao = 1.0f; m.tex = 0; }
        return m;
    }

    int32_t forward_impl(const ShsoAssets* assets, const ShsbScene* s, const ShsbFrameParams* fp, const ShsoTarget* tgt,
                         const float* shadow_lvp, int32_t preserve_depth, bool depth_only,
                         const void* lights, uint32_t n_lights, const uint32_t* counts, const uint32_t* indices,
                         ShsbStats* out_stats, const float* prev_models16 = nullptr)
    {
        if (!s || !fp || !tgt) return SHSB_E_INVALID_ARGUMENT;
        const int W = tgt->w, H = tgt->h;
        if (W <= 0 || H <= 0) return SHSB_E_INVALID_ARGUMENT;
        ShsbStats st{};
        std::vector<float> scratch;
        ShsoTarget T = *tgt;
        if (depth_only)
        {
            // PassDepthPrepassAdapter::execute_with_scratch, pass_adapters.hpp:474-528
            if (!T.depth) return SHSB_E_INVALID_ARGUMENT;
            scratch.assign((size_t)W * H * 4, 0.0f);
            T.hdr = scratch.data();
            for (size_t i = 0; i < (size_t)W * H; ++i) T.depth[i] = 1.0f;
        }
        else
        {
            if (!T.hdr) return SHSB_E_INVALID_ARGUMENT;
            if (s->sky_kind == SHSB_SKY_PROCEDURAL || s->sky_kind == SHSB_SKY_CUBEMAP)
            {
                // render_skybox_to_hdr, sky/skybox_renderer.hpp:25-57
                float inv_vp[16];
                mat4_inverse(s->cam_viewproj, inv_vp);
                const V3 cam = load3(s->cam_pos);
                const ShsoTexture* faces[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
                for (int i = 0; i < 6; ++i)
                    if (assets && s->sky_faces[i] >= 1 && s->sky_faces[i] <= assets->n_textures) faces[i] = &assets->textures[s->sky_faces[i] - 1];
                const V3 sun = normalize3(load3(s->sky_sun_dir_ws));
                for (int y = 0; y < H; ++y)
                {
                    const float ndc_y = (2.0f * ((float)y + 0.5f) / (float)H) - 1.0f;
                    float* row = T.hdr + (size_t)y * W * 4;
                    for (int x = 0; x < W; ++x)
                    {
                        const float ndc_x = (2.0f * ((float)x + 0.5f) / (float)W) - 1.0f;
                        V4 world = mat4_mul(inv_vp, ndc_x, ndc_y, 1.0f, 1.0f);
                        V3 c{0.0f, 0.0f, 0.0f};
                        if (!(std::fabs(world.w) < 1e-8f))
                        {
                            world.x /= world.w; world.y /= world.w; world.z /= world.w;
                            const V3 dir = normalize3(V3{world.x - cam.x, world.y - cam.y, world.z - cam.z});
                            c = (s->sky_kind == SHSB_SKY_PROCEDURAL) ? sky_procedural(dir, sun) : sky_cubemap(dir, faces, s->sky_intensity);
                        }
                        row[x * 4 + 0] = c.x; row[x * 4 + 1] = c.y; row[x * 4 + 2] = c.z; row[x * 4 + 3] = 1.0f;
                    }
                }
            }
            else
            // background gradient, pass_pbr_forward.hpp:69-85 (scene.sky == nullptr)
            for (int y = 0; y < H; ++y)
            {
                const float t = (float)y / (float)std::max(1, H - 1);
                const float c0 = 0.06f + 0.08f * t, c1 = 0.08f + 0.10f * t, c2 = 0.12f + 0.12f * t;
                float* row = T.hdr + (size_t)y * W * 4;
                for (int x = 0; x < W; ++x) { row[x * 4 + 0] = c0; row[x * 4 + 1] = c1; row[x * 4 + 2] = c2; row[x * 4 + 3] = 1.0f; }
            }
            // depth clear policy, pass_pbr_forward.hpp:87-98
            if (T.depth && !preserve_depth) for (size_t i = 0; i < (size_t)W * H; ++i) T.depth[i] = 1.0f;
            if (T.depth && T.motion) for (size_t i = 0; i < (size_t)W * H * 2; ++i) T.motion[i] = 0.0f; // both branches clear the motion plane
        }
        if (T.tri_id) for (size_t i = 0; i < (size_t)W * H; ++i) T.tri_id[i] = SHSB_TRI_ID_NONE;
        if (T.coverage) for (size_t i = 0; i < (size_t)W * H; ++i) T.coverage[i] = 0u;

        int shader_id = (fp->shading_model == SHSB_SHADING_BLINN_PHONG) ? SHSB_SHADER_BLINN_PHONG : SHSB_SHADER_PBR_MR;
        if (fp->debug_view == SHSB_DEBUG_ALBEDO) shader_id = SHSB_SHADER_DEBUG_ALBEDO;
        else if (fp->debug_view == SHSB_DEBUG_NORMAL) shader_id = SHSB_SHADER_DEBUG_NORMAL;
        else if (fp->debug_view == SHSB_DEBUG_DEPTH) shader_id = SHSB_SHADER_DEBUG_DEPTH;
        if (depth_only) shader_id = SHSB_SHADER_DEPTH_ONLY;

        uint32_t key_base = 0;
        for (uint32_t ii = 0; ii < s->n_items; ++ii)
        {
            const ShsbRenderItem& it = s->items[ii];
            if (!it.visible) continue;
            const ShsoMesh* mesh = find_mesh(assets, it.mesh);
            if (!mesh || mesh->n_positions == 0 || mesh->n_indices == 0) continue; // MeshData::empty(), mesh.hpp:32-35
            const uint32_t tri_count = mesh->n_indices / 3;

            DrawConst dc;
            model_from_transform(&it.tr, dc.model);
            std::memcpy(dc.viewproj, s->cam_viewproj, 64);
            normal_matrix(dc.model, dc.nrm);
            dc.has_varyings = !depth_only;
            // history: prev_model = last frame's model of this object if there was a frame, else the current one; prev_viewproj
            // likewise (pass_pbr_forward.hpp:149-161).  The depth pre-pass never writes motion (pass_adapters.hpp:521).
            set_motion(dc, !depth_only && fp->motion_vectors_enable != 0 && T.depth && T.motion,
                       prev_models16 ? prev_models16 + (size_t)ii * 16 : dc.model, prev_models16 ? s->cam_prev_viewproj : s->cam_viewproj);

            const ItemMaterial mat = resolve_material(it);
            ShsbUniforms u{};
            std::memcpy(u.light_dir_ws, s->sun_dir_ws, 12);
            std::memcpy(u.light_color, s->sun_color, 12);
            u.light_intensity = s->sun_intensity;
            std::memcpy(u.camera_pos, s->cam_pos, 12);
            u.base_color[0] = mat.base_color.x; u.base_color[1] = mat.base_color.y; u.base_color[2] = mat.base_color.z;
            u.metallic = mat.metallic; u.roughness = mat.roughness; u.ao = mat.ao;
            u.base_color_tex = mat.tex;
            const bool use_shadow = fp->shadow_enable && T.shadow && shadow_lvp; // pass_pbr_forward.hpp:185-194
            u.shadow_map = use_shadow ? 1u : 0u;
            if (use_shadow)
            {
                std::memcpy(u.light_viewproj, shadow_lvp, 64);
                u.shadow_bias_const = fp->shadow_bias_const;
                u.shadow_bias_slope = fp->shadow_bias_slope;
                u.shadow_pcf_radius = fp->shadow_pcf_radius;
                u.shadow_pcf_step = fp->shadow_pcf_step;
                u.shadow_strength = fp->shadow_strength;
            }
            FsEnv env;
            fill_env_from_uniforms(env, shader_id, &u, assets, &T);
            if (lights && counts && indices && (shader_id == SHSB_SHADER_PBR_MR || shader_id == SHSB_SHADER_BLINN_PHONG))
            {
                env.forward_plus = true;
                env.lights = (const uint8_t*)lights;
                env.n_lights = n_lights;
                env.tile_counts = counts;
                env.tile_indices = indices;
                env.tile_size = std::max(1u, fp->tile_size);
                env.max_per_tile = std::max(1u, fp->max_lights_per_tile);
                env.tiles_x = ((uint32_t)W + env.tile_size - 1u) / env.tile_size;
                env.tiles_y = ((uint32_t)H + env.tile_size - 1u) / env.tile_size;
            }
            draw_mesh(*mesh, dc, env, T, fp->cull_mode, fp->front_face_ccw != 0, key_base, st);
            key_base += tri_count * 8u;
        }
        if (out_stats) *out_stats = st;
        return SHSB_OK;
    }
}

extern "C" {

void shso_model_from_transform(const ShsbTransform* tr, float out_model[16]) { model_from_transform(tr, out_model); }

void shso_camera_viewproj(const float eye[3], const float target[3], const float up[3],
                          float fovy_radians, float aspect, float znear, float zfar, float out_viewproj[16])
{
    // glm::lookAtLH / perspectiveLH_NO (ext/matrix_transform.inl, ext/matrix_clip_space.inl), then proj*view
    const V3 E = load3(eye);
    const V3 f = normalize3(sub(load3(target), E));
    const V3 sv = normalize3(cross3(load3(up), f));
    const V3 uv = cross3(f, sv);
    float view[16] = {sv.x, uv.x, f.x, 0, sv.y, uv.y, f.y, 0, sv.z, uv.z, f.z, 0, -dot3(sv, E), -dot3(uv, E), -dot3(f, E), 1};
    const float th = std::tan(fovy_radians / 2.0f);
    float proj[16] = {0};
    proj[0] = 1.0f / (aspect * th);
    proj[5] = 1.0f / th;
    proj[10] = (zfar + znear) / (zfar - znear);
    proj[11] = 1.0f;
    proj[14] = -(2.0f * zfar * znear) / (zfar - znear);
    mat4_mul_mat4(proj, view, out_viewproj);
}

void shso_pack_point_light(const float pos[3], float range, const float color[3], float intensity,
                           uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff,
                           int32_t jolt_bounds, void* out_record160)
{
    // make_point_culling_light, lighting/light_types.hpp:327-351
    LightRec r{};
    const float rg = std::max(range, 0.0f);
    r.position_range[0] = pos[0]; r.position_range[1] = pos[1]; r.position_range[2] = pos[2]; r.position_range[3] = rg;
    for (int i = 0; i < 3; ++i) r.color_intensity[i] = fmax_glm(color[i], 0.0f);
    r.color_intensity[3] = std::max(intensity, 0.0f);
    r.direction_spot[1] = -1.0f; r.direction_spot[3] = 1.0f;
    r.axis_spot_outer[0] = 1.0f;
    r.up_shape_x[1] = 1.0f;
    r.shape_attenuation[0] = 0.0f;
    r.shape_attenuation[1] = std::max(atten_power, 0.001f);
    r.shape_attenuation[2] = std::max(atten_bias, 1e-5f);
    r.shape_attenuation[3] = std::max(atten_cutoff, 0.0f);
    r.type_shape_flags[0] = 1u; r.type_shape_flags[1] = 1u; r.type_shape_flags[2] = 7u; r.type_shape_flags[3] = atten_model;
    float sr = rg;
    if (jolt_bounds)
    {
        // scene_shape.hpp:56-81 on a Jolt sphere: box = c +- range; centre = 0.5*(min+max); radius = |0.5*(max-min)|
        float mn[3], mx[3], ce[3], ex[3];
        for (int i = 0; i < 3; ++i) { mn[i] = pos[i] - range; mx[i] = pos[i] + range; ce[i] = 0.5f * (mn[i] + mx[i]); ex[i] = 0.5f * (mx[i] - mn[i]); }
        sr = std::max(length3(V3{ex[0], ex[1], ex[2]}), 0.0f);
        for (int i = 0; i < 3; ++i) { r.cull_sphere[i] = ce[i]; r.cull_aabb_min[i] = mn[i]; r.cull_aabb_max[i] = mx[i]; }
    }
    else
    {
        for (int i = 0; i < 3; ++i) { r.cull_sphere[i] = pos[i]; r.cull_aabb_min[i] = pos[i] - rg; r.cull_aabb_max[i] = pos[i] + rg; }
    }
    r.cull_sphere[3] = sr;
    r.cull_aabb_min[3] = 1.0f; r.cull_aabb_max[3] = 1.0f;
    std::memcpy(out_record160, &r, sizeof(r));
}

void shso_pack_spot_light(const float pos[3], float range, const float color[3], float intensity,
                          const float dir[3], float inner_rad, float outer_rad,
                          uint32_t atten_model, float atten_power, float atten_bias, float atten_cutoff,
                          void* out_record160)
{
    // make_spot_culling_light, lighting/light_types.hpp:353-379
    LightRec r{};
    const float rg = std::max(range, 0.0f);
    V3 d = load3(dir);
    const float len2 = dot3(d, d);
    d = (len2 <= 1e-10f) ? V3{0.0f, -1.0f, 0.0f} : scale(d, 1.0f / std::sqrt(len2)); // normalize_or, volumes.hpp
    const float half_pi = 1.57079632679489661923f;
    const float inner = std_clampf(inner_rad, 0.01f, half_pi - 0.01f);
    const float outer = std_clampf(std::max(inner + 0.001f, outer_rad), inner + 0.001f, half_pi - 0.001f);
    r.position_range[0] = pos[0]; r.position_range[1] = pos[1]; r.position_range[2] = pos[2]; r.position_range[3] = rg;
    for (int i = 0; i < 3; ++i) r.color_intensity[i] = fmax_glm(color[i], 0.0f);
    r.color_intensity[3] = std::max(intensity, 0.0f);
    r.direction_spot[0] = d.x; r.direction_spot[1] = d.y; r.direction_spot[2] = d.z; r.direction_spot[3] = std::cos(inner);
    r.axis_spot_outer[0] = 1.0f; r.axis_spot_outer[3] = std::cos(outer);
    r.up_shape_x[1] = 1.0f;
    r.shape_attenuation[1] = std::max(atten_power, 0.001f);
    r.shape_attenuation[2] = std::max(atten_bias, 1e-5f);
    r.shape_attenuation[3] = std::max(atten_cutoff, 0.0f);
    r.type_shape_flags[0] = 2u; r.type_shape_flags[1] = 2u; r.type_shape_flags[2] = 7u; r.type_shape_flags[3] = atten_model;
    for (int i = 0; i < 3; ++i) { r.cull_sphere[i] = pos[i]; r.cull_aabb_min[i] = pos[i] - rg; r.cull_aabb_max[i] = pos[i] + rg; }
    r.cull_sphere[3] = rg;
    r.cull_aabb_min[3] = 1.0f; r.cull_aabb_max[3] = 1.0f;
    std::memcpy(out_record160, &r, sizeof(r));
}

int32_t shso_rasterize_mesh(const ShsoAssets* assets, shsb_mesh mesh_h, int32_t shader_id,
                            const ShsbUniforms* u, const ShsoTarget* tgt, const ShsbRasterCfg* cfg,
                            uint32_t key_base, ShsbStats* out_stats)
{
    const ShsoMesh* mesh = find_mesh(assets, mesh_h);
    if (!mesh || !u || !tgt || !tgt->hdr || !cfg) return SHSB_E_INVALID_ARGUMENT;
    if (shader_id < 0 || shader_id >= SHSB_SHADER_COUNT) return SHSB_E_UNSUPPORTED_SHADER;
    DrawConst dc;
    std::memcpy(dc.model, u->model, 64);
    std::memcpy(dc.viewproj, u->viewproj, 64);
    normal_matrix(dc.model, dc.nrm);
    dc.has_varyings = shader_id != SHSB_SHADER_DEPTH_ONLY;
    set_motion(dc, u->enable_motion_vectors != 0 && tgt->depth && tgt->motion, u->prev_model, u->prev_viewproj); // rasterizer.hpp:295
    FsEnv env;
    fill_env_from_uniforms(env, shader_id, u, assets, tgt);
    ShsbStats st{};
    draw_mesh(*mesh, dc, env, *tgt, cfg->cull_mode, cfg->front_face_ccw != 0, key_base, st);
    if (out_stats)
    {
        out_stats->tri_input += st.tri_input;
        out_stats->tri_after_clip += st.tri_after_clip;
        out_stats->tri_raster += st.tri_raster;
        out_stats->frag_covered += st.frag_covered;
    }
    return SHSB_OK;
}

int32_t shso_pass_pbr_forward(const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                              const ShsoTarget* tgt, const float* shadow_light_viewproj,
                              int32_t preserve_existing_depth, ShsbStats* out_stats)
{
    return forward_impl(assets, scene, fp, tgt, shadow_light_viewproj, preserve_existing_depth, false, nullptr, 0, nullptr, nullptr, out_stats);
}

int32_t shso_pass_pbr_forward_history(const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                                      const ShsoTarget* tgt, const float* shadow_light_viewproj, int32_t preserve_existing_depth,
                                      const float* prev_models16, ShsbStats* out_stats)
{
    return forward_impl(assets, scene, fp, tgt, shadow_light_viewproj, preserve_existing_depth, false, nullptr, 0, nullptr, nullptr, out_stats, prev_models16);
}

int32_t shso_pass_pbr_forward_plus(const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                                   const ShsoTarget* tgt, const float* shadow_light_viewproj,
                                   int32_t preserve_existing_depth,
                                   const void* records160, uint32_t n_lights,
                                   const uint32_t* counts, const uint32_t* indices, ShsbStats* out_stats)
{
    return forward_impl(assets, scene, fp, tgt, shadow_light_viewproj, preserve_existing_depth, false, records160, n_lights, counts, indices, out_stats);
}

int32_t shso_pass_depth_prepass(const ShsoAssets* assets, const ShsbScene* scene, const ShsbFrameParams* fp,
                                const ShsoTarget* tgt, ShsbStats* out_stats)
{
    return forward_impl(assets, scene, fp, tgt, nullptr, 0, true, nullptr, 0, nullptr, nullptr, out_stats);
}

int32_t shso_pass_shadow_map(const ShsoAssets* assets, const ShsbScene* s, const ShsbFrameParams* fp,
                             float* shadow, int32_t sw, int32_t sh, float out_light_viewproj[16])
{
    // PassShadowMap::execute, passes/pass_shadow_map.hpp:44-205
    if (!s || !fp || !shadow || sw <= 0 || sh <= 0) return SHSB_E_INVALID_ARGUMENT;
    if (!fp->shadow_enable) return SHSB_E_INVALID_ARGUMENT;
    for (size_t i = 0; i < (size_t)sw * sh; ++i) shadow[i] = 1.0f;

    // scene AABB over the 8 transformed corners of each caster's local bounds (:80-131)
    float amin[3] = {1e30f, 1e30f, 1e30f}, amax[3] = {-1e30f, -1e30f, -1e30f};
    auto expand = [&](float x, float y, float z) {
        amin[0] = fmin_glm(amin[0], x); amin[1] = fmin_glm(amin[1], y); amin[2] = fmin_glm(amin[2], z);
        amax[0] = fmax_glm(amax[0], x); amax[1] = fmax_glm(amax[1], y); amax[2] = fmax_glm(amax[2], z);
    };
    bool any = false;
    for (uint32_t ii = 0; ii < s->n_items; ++ii)
    {
        const ShsbRenderItem& it = s->items[ii];
        if (!it.visible || !it.casts_shadow) continue;
        const ShsoMesh* mesh = find_mesh(assets, it.mesh);
        if (mesh && mesh->n_positions)
        {
            float model[16];
            model_from_transform(&it.tr, model);
            float bmin[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
            float bmax[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
            for (uint32_t v = 0; v < mesh->n_positions; ++v)
                for (int a = 0; a < 3; ++a)
                {
                    bmin[a] = fmin_glm(bmin[a], mesh->positions[(size_t)v * 3 + a]);
                    bmax[a] = fmax_glm(bmax[a], mesh->positions[(size_t)v * 3 + a]);
                }
            for (int c = 0; c < 8; ++c)
            {
                const float x = (c & 1) ? bmax[0] : bmin[0], y = (c & 2) ? bmax[1] : bmin[1], z = (c & 4) ? bmax[2] : bmin[2];
                const V4 p = mat4_mul(model, x, y, z, 1.0f);
                expand(p.x, p.y, p.z);
            }
        }
        else expand(it.tr.pos[0], it.tr.pos[1], it.tr.pos[2]);
        any = true;
    }
    if (!any) { expand(-1, -1, -1); expand(1, 1, 1); }

    // build_dir_light_camera_aabb(sun, aabb, 10, shadow->w), camera/light_camera.hpp:33-99
    float lvp[16];
    {
        const float margin = 10.0f;
        const V3 dir = normalize3(load3(s->sun_dir_ws));
        const V3 up = (std::fabs(dir.y) > 0.95f) ? V3{0, 0, 1} : V3{0, 1, 0};
        const V3 mn = load3(amin), mx = load3(amax);
        const V3 c = scale(add(mn, mx), 0.5f);
        const float radius = length3(scale(sub(mx, mn), 0.5f)) + margin;
        const V3 pos = sub(c, scale(dir, radius * 2.0f));
        const V3 f = normalize3(sub(c, pos));
        const V3 sv = normalize3(cross3(up, f));
        const V3 uv = cross3(f, sv);
        const float view[16] = {sv.x, uv.x, f.x, 0, sv.y, uv.y, f.y, 0, sv.z, uv.z, f.z, 0, -dot3(sv, pos), -dot3(uv, pos), -dot3(f, pos), 1};
        float l = 1e30f, r = -1e30f, b = 1e30f, t = -1e30f, n = 1e30f, fa = -1e30f;
        for (int i = 0; i < 8; ++i)
        {
            const float x = (i & 1) ? mx.x : mn.x, y = (i & 2) ? mx.y : mn.y, z = (i & 4) ? mx.z : mn.z;
            const V4 p = mat4_mul(view, x, y, z, 1.0f);
            l = std::min(l, p.x); r = std::max(r, p.x);
            b = std::min(b, p.y); t = std::max(t, p.y);
            n = std::min(n, p.z); fa = std::max(fa, p.z);
        }
        l -= margin; r += margin; b -= margin; t += margin; n -= margin; fa += margin;
        const uint32_t res = (uint32_t)std::max(sw, 1);
        if (res > 0u)
        {
            const float span_x = std::max(r - l, 1e-5f), span_y = std::max(t - b, 1e-5f);
            const float inv_res = 1.0f / (float)res;
            const float texel_x = span_x * inv_res, texel_y = span_y * inv_res;
            float cx = 0.5f * (l + r), cy = 0.5f * (b + t);
            if (texel_x > 1e-6f) cx = std::floor(cx / texel_x + 0.5f) * texel_x;
            if (texel_y > 1e-6f) cy = std::floor(cy / texel_y + 0.5f) * texel_y;
            const float hx = 0.5f * span_x, hy = 0.5f * span_y;
            l = cx - hx; r = cx + hx; b = cy - hy; t = cy + hy;
        }
        float proj[16] = {0};
        proj[0] = 2.0f / (r - l);
        proj[5] = 2.0f / (t - b);
        proj[10] = 2.0f / (fa - n);
        proj[12] = -(r + l) / (r - l);
        proj[13] = -(t + b) / (t - b);
        proj[14] = -(fa + n) / (fa - n);
        proj[15] = 1.0f;
        mat4_mul_mat4(proj, view, lvp);
    }
    if (out_light_viewproj) std::memcpy(out_light_viewproj, lvp, 64);

    // depth-only raster, :144-204 (no clipping, no culling, affine NDC z, keep min)
    const float fw1 = (float)(sw - 1), fh1 = (float)(sh - 1);
    for (uint32_t ii = 0; ii < s->n_items; ++ii)
    {
        const ShsbRenderItem& it = s->items[ii];
        if (!it.visible || !it.casts_shadow) continue;
        const ShsoMesh* mesh = find_mesh(assets, it.mesh);
        if (!mesh || mesh->n_positions == 0) continue;
        float model[16];
        model_from_transform(&it.tr, model);
        const bool indexed = mesh->n_indices != 0;
        const size_t tri_count = indexed ? (mesh->n_indices / 3) : (mesh->n_positions / 3);
        for (size_t ti = 0; ti < tri_count; ++ti)
        {
            uint32_t id[3];
            for (int j = 0; j < 3; ++j) id[j] = indexed ? mesh->indices[ti * 3 + j] : (uint32_t)(ti * 3 + j);
            if (id[0] >= mesh->n_positions || id[1] >= mesh->n_positions || id[2] >= mesh->n_positions) continue;
            float nx[3], ny[3], nz[3];
            bool ok = true;
            for (int j = 0; j < 3; ++j)
            {
                const float* p = mesh->positions + (size_t)id[j] * 3;
                const V4 wp = mat4_mul(model, p[0], p[1], p[2], 1.0f);
                const V4 c = mat4_mul(lvp, wp.x, wp.y, wp.z, 1.0f);
                if (std::fabs(c.w) < 1e-8f) ok = false;
                nx[j] = c.x / c.w; ny[j] = c.y / c.w; nz[j] = c.z / c.w;
            }
            if (!ok) continue;
            if ((nx[0] < -1.0f && nx[1] < -1.0f && nx[2] < -1.0f) || (nx[0] > 1.0f && nx[1] > 1.0f && nx[2] > 1.0f)) continue;
            if ((ny[0] < -1.0f && ny[1] < -1.0f && ny[2] < -1.0f) || (ny[0] > 1.0f && ny[1] > 1.0f && ny[2] > 1.0f)) continue;
            if ((nz[0] < -1.0f && nz[1] < -1.0f && nz[2] < -1.0f) || (nz[0] > 1.0f && nz[1] > 1.0f && nz[2] > 1.0f)) continue;
            float sx[3], sy[3];
            for (int j = 0; j < 3; ++j) { sx[j] = (nx[j] * 0.5f + 0.5f) * fw1; sy[j] = (ny[j] * 0.5f + 0.5f) * fh1; }
            const int minx = std::max(0, (int)std::floor(std::min({sx[0], sx[1], sx[2]})));
            const int maxx = std::min(sw - 1, (int)std::ceil(std::max({sx[0], sx[1], sx[2]})));
            const int miny = std::max(0, (int)std::floor(std::min({sy[0], sy[1], sy[2]})));
            const int maxy = std::min(sh - 1, (int)std::ceil(std::max({sy[0], sy[1], sy[2]})));
            if (minx > maxx || miny > maxy) continue;
            const float v0x = sx[1] - sx[0], v0y = sy[1] - sy[0], v1x = sx[2] - sx[0], v1y = sy[2] - sy[0];
            const float den = v0x * v1y - v1x * v0y;
            if (std::fabs(den) < 1e-8f) continue;
            const float inv_den = 1.0f / den;
            for (int y = miny; y <= maxy; ++y)
                for (int x = minx; x <= maxx; ++x)
                {
                    const float v2x = ((float)x + 0.5f) - sx[0], v2y = ((float)y + 0.5f) - sy[0];
                    const float bv = (v2x * v1y - v1x * v2y) * inv_den;
                    const float bw = (v0x * v2y - v2x * v0y) * inv_den;
                    const float bu = 1.0f - bv - bw;
                    if (bu < 0.0f || bv < 0.0f || bw < 0.0f) continue;
                    const float z_ndc = bu * nz[0] + bv * nz[1] + bw * nz[2];
                    const float z01 = std_clampf(z_ndc * 0.5f + 0.5f, 0.0f, 1.0f);
                    float& zb = shadow[(size_t)y * sw + x];
                    if (z01 < zb) zb = z01;
                }
        }
    }
    return SHSB_OK;
}

int32_t shso_pass_tonemap(const float* hdr, int32_t w, int32_t h, float exposure_in, float gamma, uint8_t* out_ldr)
{
    // PassTonemap::execute, passes/pass_tonemap.hpp:49-81
    if (!hdr || !out_ldr || w <= 0 || h <= 0) return SHSB_E_INVALID_ARGUMENT;
    const float exposure = std::max(0.0001f, exposure_in);
    const float inv_gamma = 1.0f / std::max(0.001f, gamma);
    for (size_t i = 0; i < (size_t)w * h; ++i)
    {
        for (int c = 0; c < 3; ++c)
        {
            float v = std::max(0.0f, hdr[i * 4 + c] * exposure);
            v = v / (1.0f + v);
            v = std::pow(v, inv_gamma);
            out_ldr[i * 4 + c] = (uint8_t)std::clamp((int)std::lround(v * 255.0f), 0, 255);
        }
        out_ldr[i * 4 + 3] = 255;
    }
    return SHSB_OK;
}

int32_t shso_pass_motion_blur(const ShsbMotionBlurParams* p, const uint8_t* src, const float* motion, const float* depth,
                              int32_t w, int32_t h, uint8_t* out)
{
    // PassMotionBlur::execute, passes/pass_motion_blur.hpp:40-184
    if (!p || !src || !motion || !depth || !out || w <= 0 || h <= 0) return SHSB_E_INVALID_ARGUMENT;
    const size_t n = (size_t)w * h;
    if (!p->enable) { std::memcpy(out, src, n * 4); return SHSB_OK; } // :56-60
    const int samples = std::clamp(p->samples, 4, 32);                              // :79
    const float strength = std::max(0.0f, p->strength);                             // :80
    const float max_vel = std::max(1.0f, p->max_velocity_px);                       // :81
    const float min_vel = std::max(0.0f, p->min_velocity_px);                       // :82
    const float depth_eps = std::max(0.0f, p->depth_reject);                        // :83
    const float dt_scale = std::clamp(std::max(p->dt, 1e-4f) * 60.0f, 0.5f, 2.5f);  // :84
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const size_t i = (size_t)y * w + x;
            float vx = motion[i * 2 + 0] * strength * dt_scale; // :116-117
            float vy = motion[i * 2 + 1] * strength * dt_scale;
            const float len = std::sqrt(vx * vx + vy * vy);
            uint8_t* o = out + i * 4;
            if (len < min_vel) { std::memcpy(o, src + i * 4, 4); continue; } // :119-123
            if (len > max_vel && len > 1e-6f) // :124-129
            {
                const float s = max_vel / len;
                vx *= s;
                vy *= s;
            }
            const float zc = depth[i];
            float acc[3] = {0.0f, 0.0f, 0.0f}, kept = 0.0f;
            for (int k = 0; k < samples; ++k) // :136-149
            {
                const float t = ((float)k / (float)(samples - 1) - 0.5f);
                const int sx = std::clamp((int)std::lround((float)x + vx * t), 0, w - 1);
                const int sy = std::clamp((int)std::lround((float)y + vy * t), 0, h - 1);
                const size_t j = (size_t)sy * w + sx;
                if (std::abs(depth[j] - zc) > depth_eps) continue;
                acc[0] += (float)src[j * 4 + 0];
                acc[1] += (float)src[j * 4 + 1];
                acc[2] += (float)src[j * 4 + 2];
                kept += 1.0f;
            }
            if (kept < 1.0f) { std::memcpy(o, src + i * 4, 4); continue; } // :151-155
            for (int c = 0; c < 3; ++c) o[c] = (uint8_t)std::clamp((int)std::lround(acc[c] / kept), 0, 255); // :157-162
            o[3] = 255;
        }
    return SHSB_OK;
}

int32_t shso_pass_light_shafts(const ShsbLightShaftsParams* p, const uint8_t* src, const float* depth, int32_t w, int32_t h, uint8_t* out)
{
    // PassLightShafts::execute, passes/pass_light_shafts.hpp:43-214
    if (!p || !src || !out || w <= 0 || h <= 0) return SHSB_E_INVALID_ARGUMENT;
    const size_t n = (size_t)w * h;
    if (!p->enable) { std::memcpy(out, src, n * 4); return SHSB_OK; } // :53-67
    // projected sun, :77-93
    float sun_u = 0.5f, sun_v = 0.2f;
    bool sun_valid = false;
    {
        const V3 sp = add(load3(p->cam_pos), scale(neg(load3(p->sun_dir_ws)), 100.0f));
        const V4 clip = mat4_mul(p->cam_viewproj, sp.x, sp.y, sp.z, 1.0f);
        if (std::abs(clip.w) > 1e-6f)
        {
            const float nx = clip.x / clip.w, ny = clip.y / clip.w, nz = clip.z / clip.w;
            sun_u = nx * 0.5f + 0.5f;
            sun_v = ny * 0.5f + 0.5f;
            sun_valid = clip.w > 0.0f && nz >= -1.0f && nz <= 1.0f && sun_u >= 0.0f && sun_u <= 1.0f && sun_v >= 0.0f && sun_v <= 1.0f;
        }
    }
    if (!sun_valid) { std::memcpy(out, src, n * 4); return SHSB_OK; } // :96-108
    std::vector<float> luma(n); // :112-126
    for (size_t i = 0; i < n; ++i)
    {
        const float r = (float)src[i * 4 + 0] / 255.0f, g = (float)src[i * 4 + 1] / 255.0f, b = (float)src[i * 4 + 2] / 255.0f;
        luma[i] = 0.2126f * r + 0.7152f * g + 0.0722f * b;
    }
    const int steps = std::max(8, p->steps);                 // :135
    const float density = std::max(0.0f, p->density);        // :136
    const float weight = std::max(0.0f, p->weight);          // :137
    const float decay = std::clamp(p->decay, 0.0f, 1.0f);    // :138
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const float u = (float)x / (float)std::max(1, w - 1);
            const float v = (float)y / (float)std::max(1, h - 1);
            float fall = 1.0f, sum = 0.0f;
            for (int k = 0; k < steps; ++k) // :152-176
            {
                const float t = (float)k / (float)steps;
                const float su = u + (sun_u - u) * t * density;
                const float sv = v + (sun_v - v) * t * density;
                const int sx = std::clamp((int)std::lround(su * (float)(w - 1)), 0, w - 1);
                const int sy = std::clamp((int)std::lround(sv * (float)(h - 1)), 0, h - 1);
                float s = luma[(size_t)sy * w + sx];
                if (depth) s *= std::clamp(depth[(size_t)sy * w + sx], 0.0f, 1.0f); // :165-171
                sum += s * fall * weight;
                fall *= decay;
            }
            const uint8_t* b = src + ((size_t)y * w + x) * 4;
            uint8_t* o = out + ((size_t)y * w + x) * 4;
            const int boost = std::clamp((int)std::lround(sum * 80.0f), 0, 120); // :179
            o[0] = (uint8_t)std::clamp((int)b[0] + boost, 0, 255);
            o[1] = (uint8_t)std::clamp((int)b[1] + boost, 0, 255);
            o[2] = (uint8_t)std::clamp((int)b[2] + boost / 2, 0, 255);
            o[3] = 255;
        }
    return SHSB_OK;
}

int32_t shso_pass_taa(uint8_t* ldr, uint8_t* history, int32_t history_valid, size_t n_pixels)
{
    // PassTemporalAAAdapter::execute_resolved, pipeline/pass_adapters.hpp:1438-1491
    if (!ldr || !history) return SHSB_E_INVALID_ARGUMENT;
    if (!history_valid) { std::memcpy(history, ldr, n_pixels * 4); return SHSB_OK; } // :1461-1469
    const float blend = 0.12f, keep = 1.0f - blend;
    for (size_t i = 0; i < n_pixels; ++i)
    {
        for (int c = 0; c < 3; ++c)
        {
            const float v = keep * (float)ldr[i * 4 + c] + blend * (float)history[i * 4 + c]; // :1477-1481
            ldr[i * 4 + c] = (uint8_t)std::clamp((int)(v + 0.5f), 0, 255);
        }
        std::memcpy(history + i * 4, ldr + i * 4, 4); // alpha stays cur.a (:1486), history = out (:1488)
    }
    return SHSB_OK;
}

namespace
{
    // ndc_from_depth01_lh_no / ndc_from_view_depth_lh_no, lighting/jolt_light_culling.hpp:79-93
    inline float ndc_from_depth01(float depth01) { return std_clampf(depth01, 0.0f, 1.0f) * 2.0f - 1.0f; }
    inline float ndc_from_view_depth(float view_depth, float z_near, float z_far)
    {
        const float n = std::max(z_near, 1e-4f);
        const float f = std::max(z_far, n + 1e-3f);
        const float z = std_clampf(view_depth, n, f);
        const float denom = std::max(f - n, 1e-6f);
        return ((f + n) / denom) - ((2.0f * f * n) / (denom * z));
    }
}

// All four bin builders of lighting/jolt_light_culling.hpp share one loop: mode 0 = cull_lights_tiled (:135-187),
// 1 = cull_lights_tiled_depth01_range (:196-258), 2 = cull_lights_tiled_view_depth_range (:261-324),
// 3 = cull_lights_clustered (:341-412; bin = cz * tiles + ty * tiles_x + tx).  Pinned through oracle/ref_lightcull_harness.cpp.
static int32_t light_cull_bins(const void* records160, uint32_t n_lights, const float view_proj[16],
                               uint32_t vw, uint32_t vh, uint32_t ts, uint32_t max_per_tile,
                               int mode, uint32_t n_slices, const float* range_min, const float* range_max, float z_near, float z_far,
                               uint32_t* counts, uint32_t* indices)
{
    if (!view_proj || vw == 0 || vh == 0 || ts == 0 || max_per_tile == 0 || !counts || !indices) return SHSB_E_INVALID_ARGUMENT;
    if ((mode == 1 || mode == 2) && (!range_min || !range_max)) return SHSB_E_INVALID_ARGUMENT;
    if (mode != 3) n_slices = 1;
    if (n_slices == 0) return SHSB_E_INVALID_ARGUMENT;
    const uint32_t tiles_x = (vw + ts - 1) / ts, tiles_y = (vh + ts - 1) / ts;
    for (size_t t = 0; t < (size_t)tiles_x * tiles_y * n_slices; ++t) counts[t] = 0;
    if (n_lights == 0) return SHSB_OK;
    const uint8_t* recs = (const uint8_t*)records160;

    float inv_vp[16];
    mat4_inverse(view_proj, inv_vp);

    struct Pl { V3 n; float d; };
    auto sdist = [](const Pl& p, V3 x) { return dot3(p.n, x) + p.d; };
    // sphere/AABB vs plane set: geometry/jolt_culling.hpp:129-181,239-257 (eps 1e-5, :118-122)
    auto classify = [&](const Pl* planes, int np, const LightRec& lt) -> int { // 0 outside, 1 intersecting, 2 inside
        const V3 c = load3(lt.cull_sphere);
        const float r = std::max(lt.cull_sphere[3], 0.0f);
        bool inside = true;
        for (int i = 0; i < np; ++i)
        {
            const float dist = sdist(planes[i], c);
            if (dist < -(r + 1e-5f)) return 0;
            if (dist < (r + 1e-5f)) inside = false;
        }
        if (inside) return 2;
        inside = true;
        for (int i = 0; i < np; ++i)
        {
            const Pl& p = planes[i];
            const V3 pv{(p.n.x >= 0.0f) ? lt.cull_aabb_max[0] : lt.cull_aabb_min[0],
                        (p.n.y >= 0.0f) ? lt.cull_aabb_max[1] : lt.cull_aabb_min[1],
                        (p.n.z >= 0.0f) ? lt.cull_aabb_max[2] : lt.cull_aabb_min[2]};
            if (sdist(p, pv) < -1e-5f) return 0;
            const V3 nv{(p.n.x >= 0.0f) ? lt.cull_aabb_min[0] : lt.cull_aabb_max[0],
                        (p.n.y >= 0.0f) ? lt.cull_aabb_min[1] : lt.cull_aabb_max[1],
                        (p.n.z >= 0.0f) ? lt.cull_aabb_min[2] : lt.cull_aabb_max[2]};
            if (sdist(p, nv) < 1e-5f) inside = false;
        }
        return inside ? 2 : 1;
    };

    // extract_frustum_planes, geometry/frustum_culling.hpp:32-65
    Pl fr[6];
    {
        const float* m = view_proj;
        const float r0[4] = {m[0], m[4], m[8], m[12]}, r1[4] = {m[1], m[5], m[9], m[13]};
        const float r2[4] = {m[2], m[6], m[10], m[14]}, r3[4] = {m[3], m[7], m[11], m[15]};
        const float* rows[3] = {r0, r1, r2};
        for (int i = 0; i < 6; ++i)
        {
            float eq[4];
            const float* rr = rows[i / 2];
            for (int k = 0; k < 4; ++k) eq[k] = (i & 1) ? (r3[k] - rr[k]) : (r3[k] + rr[k]);
            const V3 n{eq[0], eq[1], eq[2]};
            const float len = length3(n);
            if (len <= 1e-8f) { fr[i].n = V3{0, 1, 0}; fr[i].d = eq[3]; }
            else { fr[i].n = divs(n, len); fr[i].d = eq[3] / len; }
        }
    }
    std::vector<uint8_t> visible(n_lights, 0);
    std::vector<LightRec> L(n_lights);
    for (uint32_t li = 0; li < n_lights; ++li)
    {
        std::memcpy(&L[li], recs + (size_t)li * SHSB_LIGHT_RECORD_BYTES, sizeof(LightRec));
        visible[li] = classify(fr, 6, L[li]) != 0;
    }

    auto unproject = [&](float x, float y, float z) {
        const V4 c = mat4_mul(inv_vp, x, y, z, 1.0f);
        return V3{c.x / c.w, c.y / c.w, c.z / c.w};
    };
    auto oriented = [&](V3 a, V3 b, V3 c, V3 inside) {
        Pl p;
        p.n = normalize3(cross3(sub(b, a), sub(c, a)));
        p.d = -dot3(p.n, a);
        if (dot3(p.n, inside) + p.d < 0.0f) { p.n = neg(p.n); p.d = -p.d; }
        return p;
    };

    const float log_ratio = (mode == 3) ? std::log(z_far / z_near) : 0.0f; // :373
    for (uint32_t cz = 0; cz < n_slices; ++cz)
    for (uint32_t ty = 0; ty < tiles_y; ++ty)
        for (uint32_t tx = 0; tx < tiles_x; ++tx)
        {
            float zn_ndc = -1.0f, zf_ndc = 1.0f;
            const uint32_t tile2d = ty * tiles_x + tx;
            if (mode == 1) { zn_ndc = ndc_from_depth01(range_min[tile2d]); zf_ndc = ndc_from_depth01(range_max[tile2d]); }                             // :232-235
            else if (mode == 2) { zn_ndc = ndc_from_view_depth(range_min[tile2d], z_near, z_far); zf_ndc = ndc_from_view_depth(range_max[tile2d], z_near, z_far); } // :300-303
            else if (mode == 3)
            {
                const float slice_near = z_near * std::exp(log_ratio * (float)cz / (float)n_slices);        // :377
                const float slice_far = z_near * std::exp(log_ratio * (float)(cz + 1) / (float)n_slices);   // :378
                zn_ndc = ndc_from_view_depth(slice_near, z_near, z_far);
                zf_ndc = ndc_from_view_depth(slice_far, z_near, z_far);
            }
            // make_screen_tile_cell, jolt_light_culling.hpp:95-133
            const float x0 = (float)(tx * ts) / (float)vw * 2.0f - 1.0f;
            const float x1 = (float)std::min((tx + 1) * ts, vw) / (float)vw * 2.0f - 1.0f;
            const float y_top = 1.0f - (float)(ty * ts) / (float)vh * 2.0f;
            const float y_bottom = 1.0f - (float)std::min((ty + 1) * ts, vh) / (float)vh * 2.0f;
            const V3 nbl = unproject(x0, y_bottom, zn_ndc), nbr = unproject(x1, y_bottom, zn_ndc);
            const V3 ntl = unproject(x0, y_top, zn_ndc), ntr = unproject(x1, y_top, zn_ndc);
            const V3 fbl = unproject(x0, y_bottom, zf_ndc), fbr = unproject(x1, y_bottom, zf_ndc);
            const V3 ftl = unproject(x0, y_top, zf_ndc), ftr = unproject(x1, y_top, zf_ndc);
            const V3 inside = scale(add(add(add(nbl, ntr), fbl), ftr), 0.25f);
            Pl cell[6];
            cell[0] = oriented(nbl, nbr, ntr, inside);
            cell[1] = oriented(fbr, fbl, ftl, inside);
            cell[2] = oriented(nbl, ntl, ftl, inside);
            cell[3] = oriented(nbr, fbr, ftr, inside);
            cell[4] = oriented(nbl, fbl, fbr, inside);
            cell[5] = oriented(ntl, ntr, ftr, inside);
            const size_t tile = (size_t)cz * tiles_x * tiles_y + tile2d;
            uint32_t cnt = 0;
            for (uint32_t li = 0; li < n_lights; ++li)
            {
                if (!visible[li]) continue;
                if (classify(cell, 6, L[li]) == 0) continue;
                if (cnt < max_per_tile) indices[tile * max_per_tile + cnt] = li;
                ++cnt;
            }
            counts[tile] = cnt;
        }
    return SHSB_OK;
}

int32_t shso_light_cull(const void* records160, uint32_t n_lights, const float view_proj[16],
                        uint32_t vw, uint32_t vh, uint32_t ts, uint32_t max_per_tile,
                        uint32_t* counts, uint32_t* indices)
{
    return light_cull_bins(records160, n_lights, view_proj, vw, vh, ts, max_per_tile, 0, 1, nullptr, nullptr, 0.1f, 1000.0f, counts, indices);
}

int32_t shso_light_cull_ex(const void* records160, uint32_t n_lights, const ShsbLightCullDesc* d,
                           const float* range_min, const float* range_max, uint32_t* counts, uint32_t* indices)
{
    if (!d || d->mode < 0 || d->mode > 3) return SHSB_E_INVALID_ARGUMENT;
    return light_cull_bins(records160, n_lights, d->view_proj, d->viewport_w, d->viewport_h, d->tile_size, d->max_per_bin, d->mode, d->depth_slices,
                           range_min, range_max, d->z_near, d->z_far, counts, indices);
}

int32_t shso_tile_depth_range(const float* depth, int32_t w, int32_t h, uint32_t ts, float zn, float zf, float* out_min, float* out_max)
{
    // Per-tile [min, max] linear view depth of a z-buffer written by rasterize_mesh (z01 = (view_z - zn) / (zf - zn),
    // sw_render/rasterizer.hpp:352-354, inverted as zn + z01 * (zf - zn)); cleared pixels (>= 1.0) are skipped, tiles
    // without geometry get [zn, zf] (light_culling_runtime.hpp:254-261).  Tile rows are top-anchored
    // (jolt_light_culling.hpp:105-107), framebuffer rows are y-up (gfx/rt_types.hpp:35-59).  This DEFINES the software
    // analogue of shaders/vulkan/fp_stress_depth_reduce.comp; the reference has no CPU depth reduce.
    if (!depth || !out_min || !out_max || w <= 0 || h <= 0 || ts == 0) return SHSB_E_INVALID_ARGUMENT;
    const uint32_t tiles_x = ((uint32_t)w + ts - 1) / ts, tiles_y = ((uint32_t)h + ts - 1) / ts;
    for (uint32_t ty = 0; ty < tiles_y; ++ty)
        for (uint32_t tx = 0; tx < tiles_x; ++tx)
        {
            bool any = false;
            float lo = zf, hi = zn;
            for (uint32_t r = ty * ts; r < std::min((ty + 1) * ts, (uint32_t)h); ++r)
                for (uint32_t x = tx * ts; x < std::min((tx + 1) * ts, (uint32_t)w); ++x)
                {
                    const float d = depth[(size_t)((uint32_t)h - 1 - r) * w + x];
                    if (d >= 1.0f) continue;
                    const float vz = zn + d * (zf - zn);
                    if (!any) { lo = hi = vz; any = true; }
                    else { lo = std::min(lo, vz); hi = std::max(hi, vz); }
                }
            out_min[ty * tiles_x + tx] = any ? lo : zn;
            out_max[ty * tiles_x + tx] = any ? hi : zf;
        }
    return SHSB_OK;
}

int32_t shso_tile_depth_range_ndc01(const float* depth, int32_t w, int32_t h, uint32_t ts, float z_near, float z_far, float* out_min, float* out_max)
{
    // shaders/vulkan/fp_stress_depth_reduce.comp restated: per-tile [min, max] view depth of a depth plane in the hardware's zero-to-one
    // encoding of the LH projection (the Vulkan path's depth attachment; the software rasteriser does not write it), inverted by
    // depth01_to_view_lh_no (:31-38).  Texels >= 1 are skipped (:64-65), a tile without any other texel gets (0, 0) (:73-77), min / max
    // start at 1e30 / 0 (:56-57).  Tile rows are top-anchored, buffer rows y-up.
    // PINNED against the shader's own text compiled as C++ (oracle/ref_glsl_a9_harness.cpp; tests/test_light_bins_cpu.py).
    if (!depth || !out_min || !out_max || w <= 0 || h <= 0 || ts == 0) return SHSB_E_INVALID_ARGUMENT;
    const uint32_t tiles_x = ((uint32_t)w + ts - 1) / ts, tiles_y = ((uint32_t)h + ts - 1) / ts;
    auto gmaxf = [](float a, float b) { return (a < b) ? b : a; };
    auto gminf = [](float a, float b) { return (b < a) ? b : a; };
    const float near_z = gmaxf(z_near, 0.001f), far_z = gmaxf(z_far, near_z + 0.01f);
    for (uint32_t ty = 0; ty < tiles_y; ++ty)
        for (uint32_t tx = 0; tx < tiles_x; ++tx)
        {
            bool any = false;
            float lo = 1e30f, hi = 0.0f;
            for (uint32_t r = ty * ts; r < std::min((ty + 1) * ts, (uint32_t)h); ++r)
                for (uint32_t x = tx * ts; x < std::min((tx + 1) * ts, (uint32_t)w); ++x)
                {
                    const float d01 = depth[(size_t)((uint32_t)h - 1 - r) * w + x];
                    if (d01 >= 1.0f) continue;
                    const float d = gminf(gmaxf(d01, 0.0f), 1.0f);
                    const float denom = gmaxf(far_z - d * (far_z - near_z), 1e-5f);
                    const float vz = (near_z * far_z) / denom;
                    lo = gminf(lo, vz); hi = gmaxf(hi, vz);
                    any = true;
                }
            out_min[ty * tiles_x + tx] = any ? lo : 0.0f;
            out_max[ty * tiles_x + tx] = any ? hi : 0.0f;
        }
    return SHSB_OK;
}

// ---- test hooks of row A9 (tests/test_a9_pinned_cpu.py): the restatement's per-light radiance, attenuation and list walk for
// ONE surface point, so that they can be held against the reference's GLSL compiled as C++ (oracle/ref_glsl_a9_harness.cpp)
void shso_eval_local_light(const void* records160, uint32_t idx, const float P[3], const float N[3], const float V[3], const float albedo[3],
                           float metallic, float roughness, uint32_t technique, float out3[3])
{
    FsEnv e{};
    e.lights = static_cast<const uint8_t*>(records160);
    const V3 r = eval_local_light(e, idx, load3(P), load3(N), load3(V), load3(albedo), metallic, roughness, technique == 1u);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

float shso_attenuation_quadratic(float distance, float range, uint32_t model, float power, float bias, float cutoff)
{
    return attenuation_quadratic(distance, range, model, power, bias, cutoff);
}

void shso_local_light_loop(const void* records160, uint32_t n_lights, const uint32_t* counts, const uint32_t* indices, uint32_t tiles_x, uint32_t tiles_y,
                           uint32_t max_per_tile, uint32_t tile_size, int32_t px, int32_t py, int32_t H, const float P[3], const float N[3], const float V[3],
                           const float albedo[3], float metallic, float roughness, uint32_t technique, float out3[3])
{
    FsEnv e{};
    e.lights = static_cast<const uint8_t*>(records160);
    e.n_lights = n_lights; e.tile_counts = counts; e.tile_indices = indices;
    e.tiles_x = tiles_x; e.tiles_y = tiles_y; e.max_per_tile = max_per_tile; e.tile_size = tile_size; e.H = H;
    Frag f{};
    f.world_pos = load3(P); f.px = px; f.py = py;
    const V3 r = forward_plus_lights(e, f, load3(N), load3(V), load3(albedo), metallic, roughness, technique == 1u);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

} // extern "C"
