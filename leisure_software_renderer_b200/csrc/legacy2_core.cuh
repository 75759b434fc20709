// legacy2_core.cuh -- the per-triangle and per-pixel arithmetic of the reference's legacy RENDER-TARGET demos (SURVEY.md section 8a
// rows L2 and L3) as host/device inline functions: legacy2.cu wraps them in kernels; tests/cpp/legacy2_emul_test.cpp compiles the
// very same functions with g++ (-ffp-contract=off == --fmad=false) and walks them pixel by pixel against the pinned oracle, so the
// order-independent reformulation below is checked on the CPU box as well.  That emulator is test code: the product launches kernels.
//
// Reference (cpp-folders/src/hello-render-target/):
//   hello_shadow_mapping_soft.cpp   vertex shaders :746-784, draw_triangle_tile_shadow :796-839, near-plane clip :869-905,
//                                   draw_triangle_tile_color_depth_softshadow :845-986, PCSS :252-445, fragment shader :991-1040
//   hello_pbr.cpp                   PBR:: :238-279, vertex shader :535-556, shadow_factor_pcf_2x2 :599-621, fragment_shader_pbr :627-727,
//                                   draw_triangle_tile_color_depth_motion :883-1045; shs/resources/ibl.hpp:215-287 (cube-map sampling)
//   hello-shs-renderer/shs_renderer.hpp:803-831 (barycentric_coordinate, clip_to_screen), :367-377 (sample_nearest)
//
// What differs from the L1 demo (legacy.cu): near-plane Sutherland-Hodgman clipping with a fan (up to two sub-triangles per source
// triangle), NO back-face cull (`|area| < 1e-8` only), depth = affinely interpolated VIEW-space z tested LESS, the z-buffer indexed
// with the flipped row, perspective-correct world position / uv but affine normal, and a depth write that precedes the
// `iw_sum <= 1e-8` reject of the varyings (:951-961): a pixel's colour is that of the LAST triangle in draw order that passed the
// depth test when it was drawn AND had a usable 1/w sum -- not necessarily the depth winner.  Per pixel the serial result is
// reproduced by visiting the candidate triangles in draw order and keeping (running minimum depth, last shadeable prefix minimum).
//
// Approximate operations (everything else is IEEE binary32, unfused, in the reference's order): powf (CUDA's, not glibc's; reaches
// the canvas through an 8-bit truncation -> the <= 1 LSB colour gate).  The PCSS kernel rotation's sinf / cosf are NOT approximated:
// the angle is hash01() * 6.2831853f, one of 2^24 values, and the product reads (sin, cos) from a table the host's libm filled
// (Draw::rot, api.cu) -- the double-precision sin / cos rounded to float would equal glibc's sinf / cosf on only 98.7 % of those
// angles (1 ULP off on the rest, counted exhaustively on the CPU), which moved a rotated tap across a texel boundary in a few pixels
// per frame.  L2_SINCOSF remains for builds without the table (the CPU emulator overrides it with libm's).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define L2_HD __host__ __device__ __forceinline__
#else
#define L2_HD inline
#endif

// hooks the CPU emulator overrides to switch between libm's and the device's flavour of the approximate operations
#ifndef L2_POWF
#define L2_POWF(x, y) powf((x), (y))
#endif
#ifndef L2_SINCOSF
#define L2_SINCOSF(a, s, c) do { double ds_, dc_; sincos((double)(a), &ds_, &dc_); (s) = (float)ds_; (c) = (float)dc_; } while (0)
#endif

namespace shsb
{
    namespace l2
    {
        enum Mode { MODE_SHADOW = 0, MODE_SOFTSHADOW = 1, MODE_PBR = 2 };
        constexpr int MAX_SPEC_MIPS = 16;

        struct V3 { float x, y, z; };
        L2_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
        L2_HD V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
        L2_HD V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
        L2_HD V3 kmul(float k, V3 a) { return v3(k * a.x, k * a.y, k * a.z); }  // float * vec3
        L2_HD V3 mulk(V3 a, float k) { return v3(a.x * k, a.y * k, a.z * k); }  // vec3 * float
        L2_HD V3 mul3(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
        L2_HD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
        L2_HD V3 normalize3(V3 v) { return mulk(v, 1.0f / sqrtf(dot3(v, v))); }  // glm::normalize: v * inversesqrt(dot)
        L2_HD float gmax(float a, float b) { return (a < b) ? b : a; }            // glm::max
        L2_HD float gmin(float a, float b) { return (b < a) ? b : a; }            // glm::min
        L2_HD V3 mix3(V3 a, V3 b, float t) { return add(mulk(a, 1.0f - t), mulk(b, t)); } // glm::mix
        L2_HD float clampf(float v, float lo, float hi) { if (v < lo) return lo; if (v > hi) return hi; return v; } // the demos' clampf / shs::Math::clampf
        L2_HD float std_clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }          // std::clamp
        L2_HD int clampi(int v, int lo, int hi) { if (v < lo) return lo; if (v > hi) return hi; return v; }
        L2_HD float saturate(float v) { return (v < 0.0f) ? 0.0f : (v > 1.0f ? 1.0f : v); }
        L2_HD bool finitef(float v) { return fabsf(v) <= FLT_MAX; } // false for inf and NaN

        // mat4 * vec4(p, 1): (m0*x + m1*y) + (m2*z + m3*w), per component (glm scalar path)
        L2_HD void mul_point(const float* m, V3 p, float out[4])
        {
            for (int r = 0; r < 4; ++r) out[r] = (m[r] * p.x + m[4 + r] * p.y) + (m[8 + r] * p.z + m[12 + r] * 1.0f);
        }

        struct Draw // one object of one pass of the demo's frame; the matrices are INPUTS of the path (struct Uniforms)
        {
            const float* positions;
            const float* normals;
            const float* uvs;          // may be null: uv = (0, 0)
            const uint32_t* indices;   // null: triangle soup
            uint32_t n_positions, n_normals, n_uvs, n_tris;
            int mode;
            int W, H;                  // the target of this pass: the shadow map (MODE_SHADOW) or the canvas
            int job_w, job_h;          // TILE_SIZE_X / TILE_SIZE_Y of the demo (160 x 160): see legacy.cu on why it shows
            float mvp[16], prev_mvp[16], model[16], mv[16], normal_mat[9], light_vp[16];
            float light_model[16];     // MODE_SHADOW: light_vp * model, the product taken first (:779), on the host
            float light_dir[3], camera_pos[3];
            unsigned char color[4];
            int use_texture;
            const unsigned char* tex;  // RGBA8, row-major
            int tex_w, tex_h;
            const float* shadow;       // shadow map depths (row = light-space screen y), null = no shadow term
            int sm_w, sm_h;
            const float* rot;          // PCSS kernel rotations: (sin, cos) of the 2^24 angles hash01() * 6.2831853f can take, tabulated by the
                                       // HOST's libm (api.cu: the sinf / cosf the reference itself calls); null = evaluate L2_SINCOSF
            // MODE_PBR
            float metallic, roughness, ao, ibl_diffuse, ibl_specular, ibl_reflection;
            const float* irradiance;   // 6 faces x irr_size^2 x RGB
            int irr_size;
            const float* prefiltered;  // n_mips cube maps, concatenated; null = no IBL term
            int n_mips;
            int spec_size[MAX_SPEC_MIPS];
            uint32_t spec_off[MAX_SPEC_MIPS]; // float offset of mip m
        };

        struct Vary { float pos[4], prev[4]; V3 world, normal; float u, v, view_z; };

        struct RasterRec // what the coverage / depth loop reads, one per slot (a slot = one fan triangle after clipping)
        {
            float ax, ay, v0x, v0y, v1x, v1y, d00, d01, d11, denom;
            float z0, z1, z2;          // MODE_SHADOW: NDC z; camera passes: view-space z
            float iw0, iw1, iw2;       // camera passes: 1 / w per corner (0 when |w| < 1e-6)
        };
        struct BoxRec { float minx, maxx, miny, maxy; }; // minx = NaN: the slot is empty
        struct ShadeRec { float pos[3][4], prev[3][4], world[3][3], normal[3][3], uv[3][2]; };

        L2_HD bool box_valid(const BoxRec& b) { return b.minx == b.minx; }

        // ---- vertex_shader_full (soft :746-766, pbr :535-556); prev is only meaningful in MODE_PBR
        L2_HD bool fetch_vertex(const Draw& d, uint32_t corner, uint32_t& vi)
        {
            vi = d.indices ? d.indices[corner] : corner;
            return vi < d.n_positions && (d.mode == MODE_SHADOW || vi < d.n_normals);
        }

        L2_HD void vertex_full(const Draw& d, uint32_t vi, Vary& o)
        {
            const V3 p = v3(d.positions[3 * vi], d.positions[3 * vi + 1], d.positions[3 * vi + 2]);
            const V3 n = v3(d.normals[3 * vi], d.normals[3 * vi + 1], d.normals[3 * vi + 2]);
            float wh[4], vp[4];
            mul_point(d.mvp, p, o.pos);
            if (d.mode == MODE_PBR) mul_point(d.prev_mvp, p, o.prev);
            else for (int k = 0; k < 4; ++k) o.prev[k] = 0.0f;
            mul_point(d.model, p, wh);
            o.world = v3(wh[0], wh[1], wh[2]);
            const float* nm = d.normal_mat; // mat3 * vec3 sums left to right
            o.normal = normalize3(v3(nm[0] * n.x + nm[3] * n.y + nm[6] * n.z, nm[1] * n.x + nm[4] * n.y + nm[7] * n.z, nm[2] * n.x + nm[5] * n.y + nm[8] * n.z));
            const bool has_uv = d.uvs && vi < d.n_uvs;
            o.u = has_uv ? d.uvs[2 * vi] : 0.0f;
            o.v = has_uv ? d.uvs[2 * vi + 1] : 0.0f;
            mul_point(d.mv, p, vp);
            o.view_z = vp[2];
        }

        L2_HD Vary lerp_vary(const Vary& a, const Vary& b, float t) // a + (b - a) * t, member by member
        {
            Vary o;
            for (int k = 0; k < 4; ++k) { o.pos[k] = a.pos[k] + (b.pos[k] - a.pos[k]) * t; o.prev[k] = a.prev[k] + (b.prev[k] - a.prev[k]) * t; }
            o.world = add(a.world, mulk(sub(b.world, a.world), t));
            o.normal = add(a.normal, mulk(sub(b.normal, a.normal), t));
            o.u = a.u + (b.u - a.u) * t;
            o.v = a.v + (b.v - a.v) * t;
            o.view_z = a.view_z + (b.view_z - a.view_z) * t;
            return o;
        }

        L2_HD bool clip_inside(const Vary& v) { return (v.pos[3] > 1e-6f) && (v.pos[2] >= 0.0f); }
        L2_HD Vary clip_intersect(const Vary& a, const Vary& b)
        {
            const float az = a.pos[2], bz = b.pos[2];
            const float denom = (bz - az);
            float t = (fabsf(denom) < 1e-8f) ? 0.0f : ((0.0f - az) / denom);
            t = clampf(t, 0.0f, 1.0f);
            return lerp_vary(a, b, t);
        }

        // clip_poly_near_z: Sutherland-Hodgman against (w > 1e-6 && z >= 0); at most 4 vertices come out of a triangle
        L2_HD int clip_near(const Vary vin[3], Vary poly[4])
        {
            int n = 0;
            for (int k = 0; k < 3; ++k)
            {
                const Vary& A = vin[k];
                const Vary& B = vin[(k + 1) % 3];
                const bool a_in = clip_inside(A), b_in = clip_inside(B);
                if (a_in && b_in) { if (n < 4) poly[n] = B; ++n; }
                else if (a_in && !b_in) { if (n < 4) poly[n] = clip_intersect(A, B); ++n; }
                else if (!a_in && b_in) { if (n < 4) poly[n] = clip_intersect(A, B); ++n; if (n < 4) poly[n] = B; ++n; }
            }
            return n < 4 ? n : 4;
        }

        // the P-independent half of Canvas::barycentric_coordinate + the float bounding box; false = the triangle draws nothing
        L2_HD bool finish_setup(const float sx[3], const float sy[3], RasterRec& r, BoxRec& b)
        {
            if (!(finitef(sx[0]) && finitef(sx[1]) && finitef(sx[2]) && finitef(sy[0]) && finitef(sy[1]) && finitef(sy[2]))) return false; // (int) of a non-finite float: UB in the reference
            const float area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sy[1] - sy[0]) * (sx[2] - sx[0]);
            if (fabsf(area) < 1e-8f) return false;
            const float v0x = sx[1] - sx[0], v0y = sy[1] - sy[0], v1x = sx[2] - sx[0], v1y = sy[2] - sy[0];
            const float d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
            const float denom = d00 * d11 - d01 * d01;
            if ((double)fabsf(denom) < 1e-5) return false; // compared against a DOUBLE literal; every pixel would get (-1, -1, -1)
            r.ax = sx[0]; r.ay = sy[0]; r.v0x = v0x; r.v0y = v0y; r.v1x = v1x; r.v1y = v1y;
            r.d00 = d00; r.d01 = d01; r.d11 = d11; r.denom = denom;
            b.minx = gmin(gmin(sx[0], sx[1]), sx[2]); b.maxx = gmax(gmax(sx[0], sx[1]), sx[2]);
            b.miny = gmin(gmin(sy[0], sy[1]), sy[2]); b.maxy = gmax(gmax(sy[0], sy[1]), sy[2]);
            return true;
        }

        L2_HD void invalidate(BoxRec& b) { b.minx = b.maxx = b.miny = b.maxy = nanf(""); }

        // ---- MODE_SHADOW set-up: shadow_vertex_shader x 3 (:776-784) + draw_triangle_tile_shadow's per-triangle part (:796-815)
        L2_HD void setup_shadow(const Draw& d, uint32_t t, RasterRec& r, BoxRec& b)
        {
            invalidate(b);
            float sx[3], sy[3], sz[3];
            for (int k = 0; k < 3; ++k)
            {
                uint32_t vi;
                if (!fetch_vertex(d, 3u * t + (uint32_t)k, vi)) return;
                float clip[4];
                mul_point(d.light_model, v3(d.positions[3 * vi], d.positions[3 * vi + 1], d.positions[3 * vi + 2]), clip);
                if (fabsf(clip[3]) < 1e-6f) return;
                const float ndx = clip[0] / clip[3], ndy = clip[1] / clip[3], ndz = clip[2] / clip[3];
                sx[k] = (ndx * 0.5f + 0.5f) * float(d.W - 1);
                sy[k] = (1.0f - (ndy * 0.5f + 0.5f)) * float(d.H - 1);
                sz[k] = ndz;
            }
            BoxRec bb;
            if (!finish_setup(sx, sy, r, bb)) return;
            r.z0 = sz[0]; r.z1 = sz[1]; r.z2 = sz[2];
            r.iw0 = r.iw1 = r.iw2 = 0.0f;
            b = bb;
        }

        // ---- camera set-up: 3 x vertex shader, clip, fan; writes slots 2t and 2t + 1
        L2_HD void setup_camera_slot(const Draw& d, const Vary& a, const Vary& b_, const Vary& c, RasterRec& r, BoxRec& b, ShadeRec& s)
        {
            invalidate(b);
            const Vary* tv[3] = {&a, &b_, &c};
            float sx[3], sy[3];
            for (int k = 0; k < 3; ++k)
            {
                if (tv[k]->pos[3] <= 1e-6f) return;
                const float ndx = tv[k]->pos[0] / tv[k]->pos[3], ndy = tv[k]->pos[1] / tv[k]->pos[3];
                sx[k] = (ndx + 1.0f) * 0.5f * float(d.W - 1); // Canvas::clip_to_screen
                sy[k] = (1.0f - ndy) * 0.5f * float(d.H - 1);
            }
            BoxRec bb;
            if (!finish_setup(sx, sy, r, bb)) return;
            r.z0 = a.view_z; r.z1 = b_.view_z; r.z2 = c.view_z;
            const float w0 = a.pos[3], w1 = b_.pos[3], w2 = c.pos[3];
            r.iw0 = (fabsf(w0) < 1e-6f) ? 0.0f : 1.0f / w0;
            r.iw1 = (fabsf(w1) < 1e-6f) ? 0.0f : 1.0f / w1;
            r.iw2 = (fabsf(w2) < 1e-6f) ? 0.0f : 1.0f / w2;
            for (int k = 0; k < 3; ++k)
            {
                for (int j = 0; j < 4; ++j) { s.pos[k][j] = tv[k]->pos[j]; s.prev[k][j] = tv[k]->prev[j]; }
                s.world[k][0] = tv[k]->world.x; s.world[k][1] = tv[k]->world.y; s.world[k][2] = tv[k]->world.z;
                s.normal[k][0] = tv[k]->normal.x; s.normal[k][1] = tv[k]->normal.y; s.normal[k][2] = tv[k]->normal.z;
                s.uv[k][0] = tv[k]->u; s.uv[k][1] = tv[k]->v;
            }
            b = bb;
        }

        L2_HD void setup_camera(const Draw& d, uint32_t t, RasterRec* r, BoxRec* b, ShadeRec* s) // r, b, s: the two slots of triangle t
        {
            invalidate(b[0]);
            invalidate(b[1]);
            Vary vin[3];
            for (int k = 0; k < 3; ++k)
            {
                uint32_t vi;
                if (!fetch_vertex(d, 3u * t + (uint32_t)k, vi)) return;
                vertex_full(d, vi, vin[k]);
            }
            Vary poly[4];
            const int n = clip_near(vin, poly);
            if (n < 3) return;
            setup_camera_slot(d, poly[0], poly[1], poly[2], r[0], b[0], s[0]);
            if (n == 4) setup_camera_slot(d, poly[0], poly[2], poly[3], r[1], b[1], s[1]);
        }

        // ---- per pixel
        // The pixel range a job tile [t0, t1] tests along one axis for a triangle spanning [lo, hi] (:806-814 / :917-925): the
        // demo's running clamp of bboxmin / bboxmax over the three corners collapses to these two expressions (only selections, no
        // rounding), both cast to int.
        L2_HD void job_range(float lo, float hi, int t0, int t1, int& i0, int& i1)
        {
            const float f0 = (float)t0, f1 = (float)t1;
            i0 = (int)gmax(f0, gmin(f1, lo));
            i1 = (int)gmin(f1, gmax(f0, hi));
        }

        L2_HD bool bary(const RasterRec& s, float Px, float Py, float& u, float& v, float& w) // shs_renderer.hpp:809-819
        {
            const float v2x = Px - s.ax, v2y = Py - s.ay;
            const float d20 = v2x * s.v0x + v2y * s.v0y;
            const float d21 = v2x * s.v1x + v2y * s.v1y;
            v = (s.d11 * d20 - s.d01 * d21) / s.denom;
            w = (s.d00 * d21 - s.d01 * d20) / s.denom;
            u = 1.0f - v - w;
            return !(u < 0.0f || v < 0.0f || w < 0.0f);
        }

        // MODE_SHADOW, one texel of one slot: the body of draw_triangle_tile_shadow's pixel loop (:816-836) up to the depth test.  The
        // shadow pass keeps a plain minimum per texel, so its result does not depend on the order of slots or texels.
        L2_HD bool shadow_texel_depth(const RasterRec& s, int px, int py, float& z)
        {
            float bu, bv, bw;
            if (!bary(s, (float)px + 0.5f, (float)py + 0.5f, bu, bv, bw)) return false;
            z = bu * s.z0 + bv * s.z1 + bw * s.z2;
            return !(z < 0.0f || z > 1.0f);
        }

        struct PixelState
        {
            float best_z;        // running minimum, starts as the buffer's content
            uint32_t shade_slot; // last prefix minimum with a usable 1/w sum; 0xFFFFFFFF = none
            bool wrote;          // the depth changed
        };

        // One candidate at one pixel, split in two: the order-free PROBE (does the job tile test the pixel for this slot, is the pixel
        // inside, its depth and whether the fragment's 1/w sum is usable) and the order-dependent UPDATE of the pixel's state, which must
        // see the probes that passed in draw order.  (jx0..jy1): the job tile of the pixel, inclusive.
        L2_HD bool pixel_probe(int mode, const RasterRec& s, const BoxRec& b, int px, int py, int jx0, int jx1, int jy0, int jy1, float& z, bool& usable)
        {
            // the job's pixel loops: px in [ix0, ix1] = [(int)max(jx0, min(jx1, minx)), (int)min(jx1, max(jx0, maxx))] (job_range), written as
            // compares -- with jx0 <= px <= jx1 and non-negative operands, ix0 <= px <=> minx < px + 1 or px == jx1, and ix1 >= px <=> maxx >= px
            // or px == jx0 (tests/test_legacy2_emul_cpu.py holds this form to the oracle's clamps and casts bit for bit)
            if (!((b.minx < (float)(px + 1) || px == jx1) && (b.maxx >= (float)px || px == jx0))) return false;
            if (!((b.miny < (float)(py + 1) || py == jy1) && (b.maxy >= (float)py || py == jy0))) return false;
            float bu, bv, bw;
            if (!bary(s, (float)px + 0.5f, (float)py + 0.5f, bu, bv, bw)) return false;
            z = bu * s.z0 + bv * s.z1 + bw * s.z2;
            usable = true;
            if (mode == MODE_SHADOW) return !(z < 0.0f || z > 1.0f);
            const float iw_sum = bu * s.iw0 + bv * s.iw1 + bw * s.iw2;
            usable = !(iw_sum <= 1e-8f);
            return true;
        }
        L2_HD void pixel_update(int mode, float z, bool usable, uint32_t slot, PixelState& st)
        {
            if (mode == MODE_SHADOW)
            {
                if (z < st.best_z) { st.best_z = z; st.wrote = true; }
                return;
            }
            if (!(z < st.best_z)) return;
            st.best_z = z;
            st.wrote = true;
            if (!usable) return; // the depth stays written (:951-961)
            st.shade_slot = slot;
        }
        L2_HD void pixel_visit(int mode, const RasterRec& s, const BoxRec& b, uint32_t slot, int px, int py, int jx0, int jx1, int jy0, int jy1, PixelState& st)
        {
            float z;
            bool usable;
            if (pixel_probe(mode, s, b, px, py, jx0, jx1, jy0, jy1, z, usable)) pixel_update(mode, z, usable, slot, st);
        }

        // ---- shading helpers
        L2_HD float shadow_sample_uv(const Draw& d, float u, float v) // :252-262 + ShadowMap::sample
        {
            if (u < 0.0f || u > 1.0f || v < 0.0f || v > 1.0f) return FLT_MAX;
            int x = (int)lroundf(u * float(d.sm_w - 1));
            int y = (int)lroundf(v * float(d.sm_h - 1));
            x = clampi(x, 0, d.sm_w - 1);
            y = clampi(y, 0, d.sm_h - 1);
            return d.shadow[(size_t)y * d.sm_w + x];
        }

        L2_HD uint32_t hash_u32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
        L2_HD float hash01(uint32_t x) { return float(hash_u32(x) & 0x00FFFFFFu) / float(0x01000000u); }

#define L2_POISSON_24                                                                                                                         \
    {-0.613392f, 0.617481f}, {0.170019f, -0.040254f}, {-0.299417f, 0.791925f}, {0.645680f, 0.493210f}, {-0.651784f, 0.717887f},                \
    {0.421003f, 0.027070f}, {-0.817194f, -0.271096f}, {-0.705374f, -0.668203f}, {0.977050f, -0.108615f}, {0.063326f, 0.142369f},               \
    {0.203528f, 0.214331f}, {-0.667531f, 0.326090f}, {-0.098422f, -0.295755f}, {-0.885922f, 0.215369f}, {0.566637f, 0.605213f},                \
    {0.039766f, -0.396100f}, {0.751946f, 0.453352f}, {0.078707f, -0.715323f}, {-0.075838f, -0.529344f}, {0.724479f, -0.580798f},               \
    {0.222999f, -0.215125f}, {-0.467574f, -0.405438f}, {-0.248268f, -0.814753f}, {0.354411f, -0.887570f}
        // POISSON_32 (:264-281); only the first 24 are ever read.  Constant bank on the device, a plain table on the host.
        static const float h_poisson[24][2] = {L2_POISSON_24};
#ifdef __CUDACC__
        static __constant__ float c_poisson[24][2] = {L2_POISSON_24};
#endif
        L2_HD void poisson(int i, float& x, float& y)
        {
#ifdef __CUDA_ARCH__
            x = c_poisson[i][0];
            y = c_poisson[i][1];
#else
            x = h_poisson[i][0];
            y = h_poisson[i][1];
#endif
        }

        // rotate2's std::cos / std::sin (:320-325) of ang = hash01(x) * 6.2831853f
        L2_HD void rotation(const Draw& d, uint32_t x, float ang, float& s, float& c)
        {
            if (d.rot)
            {
                const uint32_t k = hash_u32(x) & 0x00FFFFFFu; // hash01's numerator
                s = d.rot[2u * k];
                c = d.rot[2u * k + 1u];
                return;
            }
            L2_SINCOSF(ang, s, c);
        }

        L2_HD float pcss_shadow_factor(const Draw& d, float u, float v, float z_receiver, float bias, int px, int py) // :333-445
        {
            if (u < 0.0f || u > 1.0f || v < 0.0f || v > 1.0f) return 1.0f;
            const float center = shadow_sample_uv(d, u, v);
            if (center == FLT_MAX) return 1.0f;
            const float texelU = 1.0f / float(d.sm_w), texelV = 1.0f / float(d.sm_h);
            const float searchU = 18.0f * texelU, searchV = 18.0f * texelV;
            const uint32_t seed = (uint32_t)(px * 1973u ^ py * 9277u ^ 0x9e3779b9u);
            const float ang = hash01(seed) * 6.2831853f;
            float c, s;
            rotation(d, seed, ang, s, c);
            float blocker_sum = 0.0f;
            int blocker_cnt = 0;
            const float z_test = z_receiver - bias;
            for (int i = 0; i < 12; ++i)
            {
                float qx, qy;
                poisson(i, qx, qy);
                const float ox = c * qx - s * qy, oy = s * qx + c * qy;
                const float dd = shadow_sample_uv(d, u + ox * searchU, v + oy * searchV);
                if (dd == FLT_MAX) continue;
                if (dd < z_test) { blocker_sum += dd; blocker_cnt++; }
            }
            if (blocker_cnt <= 0) return 1.0f;
            const float avg = blocker_sum / float(blocker_cnt);
            const float zB = gmax(1e-5f, avg), zR = gmax(1e-5f, z_receiver);
            float ratio = (zR - zB) / zB;
            ratio = gmax(0.0f, ratio);
            float fU = 0.0035f * ratio, fV = 0.0035f * ratio;
            const float ftU = fU / texelU, ftV = fV / texelV;
            float ft = 0.5f * (ftU + ftV);
            ft = clampf(ft, 1.0f, 28.0f);
            fU = ft * texelU;
            fV = ft * texelV;
            float lit_sum = 0.0f;
            int lit_cnt = 0;
            const float ang2 = hash01(seed ^ 0xB5297A4Du) * 6.2831853f;
            rotation(d, seed ^ 0xB5297A4Du, ang2, s, c);
            for (int i = 0; i < 24; ++i)
            {
                float qx, qy;
                poisson(i, qx, qy);
                const float ox = c * qx - s * qy, oy = s * qx + c * qy;
                const float dd = shadow_sample_uv(d, u + ox * fU, v + oy * fV);
                if (dd == FLT_MAX) { lit_sum += 1.0f; lit_cnt++; continue; }
                lit_sum += (z_receiver <= dd + bias) ? 1.0f : 0.0f;
                lit_cnt++;
            }
            if (lit_cnt <= 0) return 1.0f;
            return lit_sum / float(lit_cnt);
        }

        L2_HD V3 albedo_srgb(const Draw& d, float fu, float fv) // shs::sample_nearest (shs_renderer.hpp:367-377) or the uniform colour, / 255
        {
            if (d.use_texture && d.tex && d.tex_w > 0 && d.tex_h > 0)
            {
                const float su = saturate(fu), sv = saturate(fv);
                int x = (int)lroundf(su * (float)(d.tex_w - 1));
                int y = (int)lroundf(sv * (float)(d.tex_h - 1));
                x = clampi(x, 0, d.tex_w - 1);
                y = clampi(y, 0, d.tex_h - 1);
                const unsigned char* t = d.tex + ((size_t)y * d.tex_w + x) * 4;
                return v3(float(t[0]) / 255.0f, float(t[1]) / 255.0f, float(t[2]) / 255.0f);
            }
            return v3(float(d.color[0]) / 255.0f, float(d.color[1]) / 255.0f, float(d.color[2]) / 255.0f);
        }

        struct Frag { V3 world, normal; float u, v; float cpos[4], ppos[4]; };

        // the varyings of the shaded triangle at this pixel (:951-985 / :990-1016): affine normal, perspective-correct world / uv
        L2_HD void interpolate(const RasterRec& r, const ShadeRec& s, float bu, float bv, float bw, Frag& f)
        {
            const float iw_sum = bu * r.iw0 + bv * r.iw1 + bw * r.iw2;
            for (int k = 0; k < 4; ++k)
            {
                f.cpos[k] = bu * s.pos[0][k] + bv * s.pos[1][k] + bw * s.pos[2][k];
                f.ppos[k] = bu * s.prev[0][k] + bv * s.prev[1][k] + bw * s.prev[2][k];
            }
            const V3 n0 = v3(s.normal[0][0], s.normal[0][1], s.normal[0][2]), n1 = v3(s.normal[1][0], s.normal[1][1], s.normal[1][2]), n2 = v3(s.normal[2][0], s.normal[2][1], s.normal[2][2]);
            const V3 w0 = v3(s.world[0][0], s.world[0][1], s.world[0][2]), w1 = v3(s.world[1][0], s.world[1][1], s.world[1][2]), w2 = v3(s.world[2][0], s.world[2][1], s.world[2][2]);
            f.normal = normalize3(add(add(kmul(bu, n0), kmul(bv, n1)), kmul(bw, n2)));
            const V3 wp = add(add(kmul(bu, mulk(w0, r.iw0)), kmul(bv, mulk(w1, r.iw1))), kmul(bw, mulk(w2, r.iw2)));
            f.world = v3(wp.x / iw_sum, wp.y / iw_sum, wp.z / iw_sum);
            f.u = (bu * (s.uv[0][0] * r.iw0) + bv * (s.uv[1][0] * r.iw1) + bw * (s.uv[2][0] * r.iw2)) / iw_sum;
            f.v = (bu * (s.uv[0][1] * r.iw0) + bv * (s.uv[1][1] * r.iw1) + bw * (s.uv[2][1] * r.iw2)) / iw_sum;
        }

        // ---- fragment_shader_softshadow (:991-1040) -> RGBA8 by truncation
        L2_HD void shade_softshadow(const Draw& d, const Frag& f, int px, int py, unsigned char out[4])
        {
            const V3 N = normalize3(f.normal);
            const V3 L = normalize3(v3(-d.light_dir[0], -d.light_dir[1], -d.light_dir[2]));
            const V3 Vd = normalize3(sub(v3(d.camera_pos[0], d.camera_pos[1], d.camera_pos[2]), f.world));
            const V3 base = albedo_srgb(d, f.u, f.v);
            const float ndl = dot3(N, L);
            const float diff = gmax(ndl, 0.0f);
            const V3 Hh = normalize3(add(L, Vd));
            const float spec = L2_POWF(gmax(dot3(N, Hh), 0.0f), 64.0f);
            const float specular = (0.45f * spec) * 1.0f;
            float shadow = 1.0f;
            if (d.shadow)
            {
                float clip[4];
                mul_point(d.light_vp, f.world, clip); // shadow_uvz_from_world :231-250
                if (!(fabsf(clip[3]) < 1e-6f))
                {
                    const float ndx = clip[0] / clip[3], ndy = clip[1] / clip[3], ndz = clip[2] / clip[3];
                    if (!(ndz < 0.0f || ndz > 1.0f))
                    {
                        const float suvx = ndx * 0.5f + 0.5f, suvy = 1.0f - (ndy * 0.5f + 0.5f);
                        const float slope = 1.0f - gmin(gmax(ndl, 0.0f), 1.0f);
                        const float bias = 0.0025f + 0.0100f * slope;
                        shadow = pcss_shadow_factor(d, suvx, suvy, ndz, bias, px, py);
                    }
                }
            }
            const float dcol = diff * 1.0f;
            const float bc3[3] = {base.x, base.y, base.z};
            for (int c = 0; c < 3; ++c)
            {
                const float amb = 0.22f * bc3[c];
                const float direct = shadow * (dcol * bc3[c] + specular);
                const float r = gmin(gmax(amb + direct, 0.0f), 1.0f);
                const float r2 = gmin(gmax(r, 0.0f), 1.0f) * 255.0f;
                out[c] = (unsigned char)r2;
            }
            out[3] = 255;
        }

        // ---- cube-map sampling, shs/resources/ibl.hpp:215-270
        L2_HD V3 cube_at(const float* data, int size, int f, int x, int y)
        {
            const float* p = data + (((size_t)f * size + (size_t)y) * size + (size_t)x) * 3;
            return v3(p[0], p[1], p[2]);
        }

        L2_HD V3 sample_face_bilinear(const float* data, int size, int face, float u, float v)
        {
            u = std_clampf(u, 0.0f, 1.0f);
            v = std_clampf(v, 0.0f, 1.0f);
            const float fx = u * float(size - 1), fy = v * float(size - 1);
            const int x0 = clampi((int)floorf(fx), 0, size - 1), y0 = clampi((int)floorf(fy), 0, size - 1);
            const int x1 = clampi(x0 + 1, 0, size - 1), y1 = clampi(y0 + 1, 0, size - 1);
            const float tx = fx - float(x0), ty = fy - float(y0);
            const V3 cx0 = mix3(cube_at(data, size, face, x0, y0), cube_at(data, size, face, x1, y0), tx);
            const V3 cx1 = mix3(cube_at(data, size, face, x0, y1), cube_at(data, size, face, x1, y1), tx);
            return mix3(cx0, cx1, ty);
        }

        L2_HD V3 sample_cubemap_linear(const float* data, int size, V3 dir)
        {
            if (size <= 0) return v3(0, 0, 0);
            V3 dd = dir;
            const float len = sqrtf(dot3(dd, dd));
            if (len < 1e-8f) return v3(0, 0, 0);
            dd = v3(dd.x / len, dd.y / len, dd.z / len);
            const float ax = fabsf(dd.x), ay = fabsf(dd.y), az = fabsf(dd.z);
            int face = 0;
            float u = 0.5f, v = 0.5f;
            if (ax >= ay && ax >= az)
            {
                if (dd.x > 0.0f) { face = 0; u = (-dd.z / ax); v = (dd.y / ax); }
                else { face = 1; u = (dd.z / ax); v = (dd.y / ax); }
            }
            else if (ay >= ax && ay >= az)
            {
                if (dd.y > 0.0f) { face = 2; u = (dd.x / ay); v = (-dd.z / ay); }
                else { face = 3; u = (dd.x / ay); v = (dd.z / ay); }
            }
            else
            {
                if (dd.z > 0.0f) { face = 4; u = (dd.x / az); v = (dd.y / az); }
                else { face = 5; u = (-dd.x / az); v = (dd.y / az); }
            }
            u = 0.5f * (u + 1.0f);
            v = 0.5f * (v + 1.0f);
            return sample_face_bilinear(data, size, face, u, v);
        }

        // ---- motion vector of draw_triangle_tile_color_depth_motion (:1018-1030)
        L2_HD void motion_vector(const Draw& d, const Frag& f, float& vx, float& vy)
        {
            const float csx = (f.cpos[0] / f.cpos[3] + 1.0f) * 0.5f * float(d.W - 1), csy = (1.0f - f.cpos[1] / f.cpos[3]) * 0.5f * float(d.H - 1);
            const float psx = (f.ppos[0] / f.ppos[3] + 1.0f) * 0.5f * float(d.W - 1), psy = (1.0f - f.ppos[1] / f.ppos[3]) * 0.5f * float(d.H - 1);
            vx = csx - psx;
            vy = -(csy - psy);
            const float vlen = sqrtf(vx * vx + vy * vy);
            if (vlen > 22.0f && vlen > 1e-6f) { const float k = 22.0f / vlen; vx *= k; vy *= k; }
        }

        // ---- fragment_shader_pbr (:627-727): Cook-Torrance sun + IBL, Reinhard + gamma in the shader -> RGBA8 by truncation
        L2_HD void shade_pbr(const Draw& d, const Frag& f, unsigned char out[4])
        {
            const float PI = 3.14159265358979323846f;
            const V3 N = normalize3(f.normal);
            const V3 Vd = normalize3(sub(v3(d.camera_pos[0], d.camera_pos[1], d.camera_pos[2]), f.world));
            const V3 L = normalize3(v3(-d.light_dir[0], -d.light_dir[1], -d.light_dir[2]));
            const V3 Hh = normalize3(add(Vd, L));
            const float NoV = gmax(0.0f, dot3(N, Vd)), NoL = gmax(0.0f, dot3(N, L)), NoH = gmax(0.0f, dot3(N, Hh));
            const V3 srgb = albedo_srgb(d, f.u, f.v);
            const V3 base = v3(L2_POWF(gmin(gmax(srgb.x, 0.0f), 1.0f), 2.2f), L2_POWF(gmin(gmax(srgb.y, 0.0f), 1.0f), 2.2f), L2_POWF(gmin(gmax(srgb.z, 0.0f), 1.0f), 2.2f));
            const float metallic = saturate(d.metallic);
            const float roughness = saturate(d.roughness) < 0.04f ? 0.04f : (d.roughness > 1.0f ? 1.0f : d.roughness); // clampf(r, 0.04, 1)
            const float ao = saturate(d.ao);
            const V3 F0 = mix3(v3(0.04f, 0.04f, 0.04f), base, metallic);
            const float fx = 1.0f - saturate(NoV); // PBR::fresnel_schlick
            const float fx2 = fx * fx;
            const float fx5 = fx2 * fx2 * fx;
            V3 F = add(F0, mulk(sub(v3(1.0f, 1.0f, 1.0f), F0), fx5));
            F = mul3(F, v3(1.0f, 0.96f, 0.90f));
            const V3 kd = mulk(sub(v3(1.0f, 1.0f, 1.0f), F), 1.0f - metallic);
            const float alpha = roughness * roughness;
            const float nh = saturate(NoH); // PBR::ndf_ggx
            const float a2 = alpha * alpha;
            const float dd = (nh * nh) * (a2 - 1.0f) + 1.0f;
            const float D = a2 / (PI * dd * dd);
            const float rr = (roughness < 0.04f ? 0.04f : (roughness > 1.0f ? 1.0f : roughness)) + 1.0f; // PBR::g_smith
            const float kk = (rr * rr) / 8.0f;
            const float nv = saturate(NoV), nl = saturate(NoL);
            const float gv = nv / (nv * (1.0f - kk) + kk), gl = nl / (nl * (1.0f - kk) + kk);
            const float G = gv * gl;
            const V3 direct_diffuse = mulk(mul3(kd, base), 1.0f / PI);
            const float spec_den = gmax(1e-6f, (4.0f * NoV * NoL));
            const V3 dgf = kmul(D * G, F);
            const V3 direct_specular = v3(dgf.x / spec_den, dgf.y / spec_den, dgf.z / spec_den);
            V3 direct = mulk(mul3(add(direct_diffuse, direct_specular), v3(3.0f, 3.0f, 3.0f)), NoL);
            float shadow = 1.0f;
            if (d.shadow)
            {
                float clip[4];
                mul_point(d.light_vp, f.world, clip);
                if (!(fabsf(clip[3]) < 1e-6f))
                {
                    const float ndx = clip[0] / clip[3], ndy = clip[1] / clip[3], ndz = clip[2] / clip[3];
                    if (!(ndz < 0.0f || ndz > 1.0f))
                    {
                        const float suvx = ndx * 0.5f + 0.5f, suvy = 1.0f - (ndy * 0.5f + 0.5f);
                        const float slope = 1.0f - gmin(gmax(dot3(N, L), 0.0f), 1.0f);
                        const float bias = 0.0025f + 0.0100f * slope;
                        // shadow_factor_pcf_2x2 (:599-621)
                        const float sfx = suvx * float(d.sm_w - 1), sfy = suvy * float(d.sm_h - 1);
                        const int x0 = clampi((int)floorf(sfx), 0, d.sm_w - 1);
                        const int y0 = clampi((int)floorf(sfy), 0, d.sm_h - 1);
                        const int x1 = clampi(x0 + 1, 0, d.sm_w - 1), y1 = clampi(y0 + 1, 0, d.sm_h - 1);
                        const float s00 = (ndz <= d.shadow[(size_t)y0 * d.sm_w + x0] + bias) ? 1.0f : 0.0f;
                        const float s10 = (ndz <= d.shadow[(size_t)y0 * d.sm_w + x1] + bias) ? 1.0f : 0.0f;
                        const float s01 = (ndz <= d.shadow[(size_t)y1 * d.sm_w + x0] + bias) ? 1.0f : 0.0f;
                        const float s11 = (ndz <= d.shadow[(size_t)y1 * d.sm_w + x1] + bias) ? 1.0f : 0.0f;
                        shadow = 0.25f * (s00 + s10 + s01 + s11);
                    }
                }
            }
            direct = mulk(direct, shadow);
            V3 ibl = v3(0.0f, 0.0f, 0.0f);
            if (d.prefiltered && d.irradiance && d.irr_size > 0 && d.n_mips > 0)
            {
                const V3 irradiance_c = sample_cubemap_linear(d.irradiance, d.irr_size, N);
                V3 diffuse_ibl = mul3(mul3(irradiance_c, base), kd);
                diffuse_ibl = mulk(diffuse_ibl, saturate(d.ibl_diffuse));
                const V3 I = v3(-Vd.x, -Vd.y, -Vd.z); // glm::reflect(-V, N) = I - N * dot(N, I) * 2
                const V3 R = sub(I, mulk(mulk(N, dot3(N, I)), 2.0f));
                float lod = roughness * float(d.n_mips - 1);
                const float mmax = float(d.n_mips - 1); // sample_prefiltered_spec_trilinear (ibl.hpp:272-287)
                lod = std_clampf(lod, 0.0f, mmax);
                const int m0 = (int)floorf(lod);
                const int m1 = (m0 + 1 < d.n_mips - 1) ? m0 + 1 : d.n_mips - 1;
                const float tl = lod - float(m0);
                const V3 prefiltered_c = mix3(sample_cubemap_linear(d.prefiltered + d.spec_off[m0], d.spec_size[m0], R),
                                              sample_cubemap_linear(d.prefiltered + d.spec_off[m1], d.spec_size[m1], R), tl);
                V3 spec_ibl = mul3(prefiltered_c, F);
                spec_ibl = mulk(spec_ibl, saturate(d.ibl_specular) * saturate(d.ibl_reflection));
                ibl = add(diffuse_ibl, spec_ibl);
            }
            ibl = mulk(ibl, ao);
            V3 color = add(direct, ibl);
            color = add(color, mulk(mulk(base, 0.03f), ao));
            color = mulk(color, 1.75f);
            color = v3(color.x / (1.0f + color.x), color.y / (1.0f + color.y), color.z / (1.0f + color.z)); // tonemap_reinhard
            const float inv_gamma = 1.0f / 2.2f;
            const float cs[3] = {L2_POWF(gmin(gmax(color.x, 0.0f), 1.0f), inv_gamma), L2_POWF(gmin(gmax(color.y, 0.0f), 1.0f), inv_gamma),
                                 L2_POWF(gmin(gmax(color.z, 0.0f), 1.0f), inv_gamma)};
            for (int c = 0; c < 3; ++c) out[c] = (unsigned char)(gmin(gmax(cs[c], 0.0f), 1.0f) * 255.0f);
            out[3] = 255;
        }

        // the whole of what a pixel does once its shaded slot is known; vel = null outside MODE_PBR
        L2_HD void shade_pixel(const Draw& d, const RasterRec& r, const ShadeRec& s, int px, int py, unsigned char out[4], float* vel)
        {
            float bu, bv, bw;
            bary(r, (float)px + 0.5f, (float)py + 0.5f, bu, bv, bw);
            Frag f;
            interpolate(r, s, bu, bv, bw, f);
            if (d.mode == MODE_PBR)
            {
                if (vel) motion_vector(d, f, vel[0], vel[1]);
                shade_pbr(d, f, out);
            }
            else shade_softshadow(d, f, px, py, out);
        }
    }
}
