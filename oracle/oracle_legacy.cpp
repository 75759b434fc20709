// oracle/oracle_legacy.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with the product).
//
// CPU restatement of the reference's LEGACY tile-job rasterizer (BASELINE configs[0] as shipped):
//   cpp-folders/src/hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp   Uniforms :35-41, blinn_phong_vertex_shader :48-58,
//       blinn_phong_fragment_shader :64-96, RendererSystem::draw_triangle_tile :189-242, ::process :244-313
//   cpp-folders/src/hello-shs-renderer/shs_renderer.hpp   Canvas::barycentric_coordinate :803-820, clip_to_screen :822-831,
//       ZBuffer::test_and_set_depth :659-669, Canvas::draw_pixel_screen_space :792-796, Camera3D::update :1223-1236
// Plain floats, explicit operation order (GLM's scalar path as stated in oracle/glm_shim).  PINNED: tests/test_legacy_cpu.py demands
// bit-equality with the reference's own sources compiled into oracle/_ref/libshs_legacy_ref.so (oracle/ref_legacy_harness.cpp).
// Part of liboracle.so (oracle/Makefile).
//
// Second half of the file: the legacy SOFT-SHADOW demo (config-3 flavour),
//   cpp-folders/src/hello-render-target/hello_shadow_mapping_soft.cpp   ShadowMap :191-229, shadow_uvz_from_world :231-250,
//       PCSS :252-445, vertex shaders :746-784, draw_triangle_tile_shadow :796-839, draw_triangle_tile_color_depth_softshadow :845-986,
//       fragment_shader_softshadow :991-1040
// PINNED the same way (oracle/ref_legacy2_harness.cpp -> oracle/_ref/libshs_legacy2_ref.so, tests/test_legacy2_cpu.py).  The CUDA
// path of this row (csrc/legacy2.cu) is checked against this restatement (tests/test_zz_gpu_legacy2.py).
//
// Third part: the legacy PBR / IBL demo (config-4 flavour, row L3), cpp-folders/src/hello-render-target/hello_pbr.cpp
//   PBR:: :238-279, shadow_factor_pcf_2x2 :599-621, fragment_shader_pbr :627-727, draw_triangle_tile_color_depth_motion :883-1045,
//   with shs-renderer-lib/include/shs/resources/ibl.hpp:215-287 (cube-map sampling).  Pinned by oracle/ref_legacy3_harness.cpp,
//   tests/test_legacy3_cpu.py.  CUDA path: csrc/legacy2.cu (MODE_PBR), same GPU test file.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace
{
    struct V3 { float x, y, z; };
    inline V3 add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
    inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
    inline V3 kmul(float k, V3 a) { return V3{k * a.x, k * a.y, k * a.z}; }      // float * vec3
    inline V3 mulk(V3 a, float k) { return V3{a.x * k, a.y * k, a.z * k}; }      // vec3 * float
    inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    inline V3 normalize3(V3 v) { return mulk(v, 1.0f / std::sqrt(dot3(v, v))); }
    inline V3 cross3(V3 x, V3 y) { return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
    inline float gmax(float a, float b) { return (a < b) ? b : a; }
    inline float gmin(float a, float b) { return (b < a) ? b : a; }

    // column-major 4x4 helpers in GLM's scalar order
    inline void mat_mul_point(const float* m, V3 p, float out[4])
    {
        for (int r = 0; r < 4; ++r) out[r] = (m[r] * p.x + m[4 + r] * p.y) + (m[8 + r] * p.z + m[12 + r] * 1.0f);
    }
    inline void mat_mul(const float* a, const float* b, float* out) // a * b: column i = a0*b[i][0] + a1*b[i][1] + a2*b[i][2] + a3*b[i][3]
    {
        float r[16];
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k)
                r[i * 4 + k] = a[k] * b[i * 4] + a[4 + k] * b[i * 4 + 1] + a[8 + k] * b[i * 4 + 2] + a[12 + k] * b[i * 4 + 3];
        std::memcpy(out, r, 64);
    }
    void mat_identity(float* m) { std::memset(m, 0, 64); m[0] = m[5] = m[10] = m[15] = 1.0f; }

    void mat_inverse(const float* src, float* out) // glm::inverse(mat4), cofactor expansion times 1 / det
    {
        float m[4][4];
        std::memcpy(m, src, 64);
        const float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
        const float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
        const float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
        const float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
        const float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
        const float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
        const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
        const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
        const float v0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, v1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
        const float v2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, v3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
        float inv[4][4];
        const float sa[4] = {+1, -1, +1, -1}, sb[4] = {-1, +1, -1, +1};
        for (int i = 0; i < 4; ++i)
        {
            const float i0 = v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i];
            const float i1 = v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i];
            const float i2 = v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i];
            const float i3 = v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i];
            inv[0][i] = i0 * sa[i]; inv[1][i] = i1 * sb[i]; inv[2][i] = i2 * sa[i]; inv[3][i] = i3 * sb[i];
        }
        const float row0[4] = {inv[0][0], inv[1][0], inv[2][0], inv[3][0]};
        const float d0 = m[0][0] * row0[0], d1 = m[0][1] * row0[1], d2 = m[0][2] * row0[2], d3 = m[0][3] * row0[3];
        const float det = (d0 + d1) + (d2 + d3);
        const float one_over = 1.0f / det;
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) out[c * 4 + r] = inv[c][r] * one_over;
    }
}

extern "C"
{
    // Camera3D::update (shs_renderer.hpp:1223-1236) under Viewer(position, speed, w, h) (:1322-1346): fov 60, aspect 4/3, z 0.1..1000
    void shso_legacy_camera(const float position[3], float horizontal_angle_deg, float vertical_angle_deg, float out_view[16], float out_proj[16])
    {
        const float deg = 0.01745329251994329576923690768489f; // glm::radians
        const float va = vertical_angle_deg * deg, ha = horizontal_angle_deg * deg;
        // unqualified cos / sin on floats = the C double functions; the products are narrowed by glm::vec3's converting constructor
        V3 dir{(float)(std::cos((double)va) * std::sin((double)ha)), (float)std::sin((double)va), (float)(std::cos((double)va) * std::cos((double)ha))};
        dir = normalize3(dir);
        const V3 world_up{0.0f, 1.0f, 0.0f};
        const V3 right = normalize3(cross3(world_up, dir));
        const V3 up = normalize3(cross3(dir, right));
        const V3 pos{position[0], position[1], position[2]};
        // glm::perspectiveLH_NO
        // (field_of_view is a run-time member in the reference: keep the compiler from folding tan() with its own, correctly
        // rounded, arithmetic -- libm's tanf is not always correctly rounded and is what the reference calls)
        volatile float field_of_view = 60.0f;
        const float fovy = field_of_view * deg, aspect = 4.0f / 3.0f, zn = 0.1f, zf = 1000.0f;
        const float thf = std::tan(fovy / 2.0f);
        std::memset(out_proj, 0, 64);
        out_proj[0] = 1.0f / (aspect * thf);
        out_proj[5] = 1.0f / thf;
        out_proj[10] = (zf + zn) / (zf - zn);
        out_proj[11] = 1.0f;
        out_proj[14] = -(2.0f * zf * zn) / (zf - zn);
        // glm::lookAtLH(eye, center, up)
        const V3 center = add(pos, dir);
        const V3 f = normalize3(sub(center, pos));
        const V3 s = normalize3(cross3(up, f));
        const V3 u = cross3(f, s);
        mat_identity(out_view);
        out_view[0] = s.x; out_view[4] = s.y; out_view[8] = s.z;
        out_view[1] = u.x; out_view[5] = u.y; out_view[9] = u.z;
        out_view[2] = f.x; out_view[6] = f.y; out_view[10] = f.z;
        out_view[12] = -dot3(s, pos); out_view[13] = -dot3(u, pos); out_view[14] = -dot3(f, pos);
    }

    // MonkeyObject::get_world_matrix (:122-128)
    void shso_legacy_world_matrix(const float position[3], const float scale[3], float rotation_angle_deg, float out_model[16])
    {
        float t[16], r[16], s[16], tr[16];
        mat_identity(t);
        // glm::translate(I, v): col3 = I0*v.x + I1*v.y + I2*v.z + I3
        for (int k = 0; k < 4; ++k) t[12 + k] = ((k == 0 ? 1.0f : 0.0f) * position[0] + (k == 1 ? 1.0f : 0.0f) * position[1]) + (k == 2 ? 1.0f : 0.0f) * position[2] + (k == 3 ? 1.0f : 0.0f);
        // glm::rotate(I, a, (0,1,0))
        const float a = rotation_angle_deg * 0.01745329251994329576923690768489f;
        const float c = std::cos(a), sn = std::sin(a);
        const V3 axis = normalize3(V3{0.0f, 1.0f, 0.0f});
        const V3 temp = mulk(axis, 1.0f - c);
        float R[3][3];
        R[0][0] = c + temp.x * axis.x; R[0][1] = temp.x * axis.y + sn * axis.z; R[0][2] = temp.x * axis.z - sn * axis.y;
        R[1][0] = temp.y * axis.x - sn * axis.z; R[1][1] = c + temp.y * axis.y; R[1][2] = temp.y * axis.z + sn * axis.x;
        R[2][0] = temp.z * axis.x + sn * axis.y; R[2][1] = temp.z * axis.y - sn * axis.x; R[2][2] = c + temp.z * axis.z;
        float I[16];
        mat_identity(I);
        mat_identity(r);
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 4; ++k) r[j * 4 + k] = I[k] * R[j][0] + I[4 + k] * R[j][1] + I[8 + k] * R[j][2];
        // glm::scale(I, v)
        mat_identity(s);
        for (int k = 0; k < 4; ++k) { s[k] = I[k] * scale[0]; s[4 + k] = I[4 + k] * scale[1]; s[8 + k] = I[8 + k] * scale[2]; }
        mat_mul(t, r, tr);
        mat_mul(tr, s, out_model);
    }

    void shso_legacy_mvp(const float proj[16], const float view[16], const float model[16], float out_mvp[16])
    {
        float pv[16];
        mat_mul(proj, view, pv);
        mat_mul(pv, model, out_mvp);
    }

    int32_t shso_legacy_draw(const float* positions, const float* normals, uint32_t n_vertices, const float mvp[16], const float model[16],
                             const float light_dir[3], const float camera_pos[3], const uint8_t color[4], int32_t W, int32_t H,
                             int32_t tile_w, int32_t tile_h, uint8_t* canvas_rgba, float* zbuffer)
    {
        if (!positions || !normals || !canvas_rgba || !zbuffer || W <= 0 || H <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        float inv[16], nm[9];
        mat_inverse(model, inv);
        for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row) nm[col * 3 + row] = inv[row * 4 + col]; // mat3(transpose(inverse(model)))
        const int cols = (W + tile_w - 1) / tile_w, rows = (H + tile_h - 1) / tile_h;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const int tminx = tx * tile_w, tminy = ty * tile_h;
                const int tmaxx = std::min((tx + 1) * tile_w, W) - 1, tmaxy = std::min((ty + 1) * tile_h, H) - 1;
                for (uint32_t i = 0; i + 2 < n_vertices; i += 3)
                {
                    // ---- vertex stage (:199-204)
                    float sx[3], sy[3], sz[3];
                    V3 vn[3], vw[3];
                    for (int k = 0; k < 3; ++k)
                    {
                        const V3 p{positions[3 * (i + k)], positions[3 * (i + k) + 1], positions[3 * (i + k) + 2]};
                        const V3 n{normals[3 * (i + k)], normals[3 * (i + k) + 1], normals[3 * (i + k) + 2]};
                        float clip[4], wp[4];
                        mat_mul_point(mvp, p, clip);
                        mat_mul_point(model, p, wp);
                        vw[k] = V3{wp[0], wp[1], wp[2]};
                        vn[k] = normalize3(V3{nm[0] * n.x + nm[3] * n.y + nm[6] * n.z, nm[1] * n.x + nm[4] * n.y + nm[7] * n.z, nm[2] * n.x + nm[5] * n.y + nm[8] * n.z});
                        const float ndx = clip[0] / clip[3], ndy = clip[1] / clip[3], ndz = clip[2] / clip[3];
                        sx[k] = (ndx + 1.0f) * 0.5f * float(W - 1);
                        sy[k] = (1.0f - ndy) * 0.5f * float(H - 1);
                        sz[k] = ndz;
                    }
                    // non-finite screen positions: the reference casts them to int (undefined behaviour); dropped (as in csrc/legacy.cu)
                    bool finite = true;
                    for (int k = 0; k < 3; ++k) finite = finite && std::isfinite(sx[k]) && std::isfinite(sy[k]);
                    if (!finite) continue;
                    // ---- bounding box clamped INTO the job tile (:206-214)
                    float bminx = (float)tmaxx, bminy = (float)tmaxy, bmaxx = (float)tminx, bmaxy = (float)tminy;
                    for (int k = 0; k < 3; ++k)
                    {
                        bminx = gmax((float)tminx, gmin(bminx, sx[k])); bminy = gmax((float)tminy, gmin(bminy, sy[k]));
                        bmaxx = gmin((float)tmaxx, gmax(bmaxx, sx[k])); bmaxy = gmin((float)tmaxy, gmax(bmaxy, sy[k]));
                    }
                    if (bminx > bmaxx || bminy > bmaxy) continue;
                    const float area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sy[1] - sy[0]) * (sx[2] - sx[0]);
                    if (area <= 0) continue;
                    // ---- fragment stage (:222-241), the reference's column-major pixel order
                    const float v0x = sx[1] - sx[0], v0y = sy[1] - sy[0], v1x = sx[2] - sx[0], v1y = sy[2] - sy[0];
                    for (int px = (int)bminx; px <= (int)bmaxx; ++px)
                        for (int py = (int)bminy; py <= (int)bmaxy; ++py)
                        {
                            const float Px = px + 0.5f, Py = py + 0.5f;
                            const float v2x = Px - sx[0], v2y = Py - sy[0];
                            const float d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
                            const float d20 = v2x * v0x + v2y * v0y, d21 = v2x * v1x + v2y * v1y;
                            const float denom = d00 * d11 - d01 * d01;
                            float bu, bv, bw;
                            if (std::abs(denom) < 1e-5) { bu = bv = bw = -1.0f; }
                            else
                            {
                                bv = (d11 * d20 - d01 * d21) / denom;
                                bw = (d00 * d21 - d01 * d20) / denom;
                                bu = 1.0f - bv - bw;
                            }
                            if (bu < 0 || bv < 0 || bw < 0) continue;
                            const float z = bu * sz[0] + bv * sz[1] + bw * sz[2];
                            if (px < 0 || px >= W || py < 0 || py >= H) continue;      // ZBuffer::test_and_set_depth bounds
                            float& d = zbuffer[(size_t)py * W + px];
                            if (!(z < d)) continue;
                            d = z;
                            const V3 in_normal = normalize3(add(add(kmul(bu, vn[0]), kmul(bv, vn[1])), kmul(bw, vn[2])));
                            const V3 world_pos = add(add(kmul(bu, vw[0]), kmul(bv, vw[1])), kmul(bw, vw[2]));
                            // blinn_phong_fragment_shader (:64-96)
                            const V3 norm = normalize3(in_normal);
                            const V3 ldir = normalize3(V3{-light_dir[0], -light_dir[1], -light_dir[2]});
                            const V3 vdir = normalize3(sub(V3{camera_pos[0], camera_pos[1], camera_pos[2]}, world_pos));
                            const float ambient = 0.15f * 1.0f;
                            const float diff = gmax(dot3(norm, ldir), 0.0f);
                            const float diffuse = diff * 1.0f;
                            const V3 halfway = normalize3(add(ldir, vdir));
                            const float spec = std::pow(gmax(dot3(norm, halfway), 0.0f), 64.0f);
                            const float specular = (0.5f * spec) * 1.0f;
                            const float lit = (ambient + diffuse) + specular;
                            const float oc[3] = {(float)color[0] / 255.0f, (float)color[1] / 255.0f, (float)color[2] / 255.0f};
                            uint8_t out[4];
                            for (int c = 0; c < 3; ++c)
                            {
                                const float r = gmin(gmax(lit * oc[c], 0.0f), 1.0f);
                                out[c] = (uint8_t)(r * 255);
                            }
                            out[3] = 255;
                            const int y_canvas = (H - 1) - py;                            // Canvas::draw_pixel_screen_space
                            std::memcpy(canvas_rgba + ((size_t)y_canvas * W + px) * 4, out, 4);
                        }
                }
            }
        return 0;
    }
}

// =====================================================================================================================
// L2: the legacy soft-shadow demo (hello_shadow_mapping_soft.cpp)
// =====================================================================================================================
namespace
{
    const float FLT_MAXV = std::numeric_limits<float>::max();
    inline float clampf_l2(float v, float lo, float hi) { if (v < lo) return lo; if (v > hi) return hi; return v; }   // the demo's clampf :127-132
    inline int clampi_l2(int v, int lo, int hi) { if (v < lo) return lo; if (v > hi) return hi; return v; }
    inline float saturate_hdr(float v) { return (v < 0.0f) ? 0.0f : (v > 1.0f ? 1.0f : v); }                            // shs::Math::saturate

    // Canvas::barycentric_coordinate, shs_renderer.hpp:803-820
    inline void bary_legacy(float Px, float Py, const float sx[3], const float sy[3], float& bu, float& bv, float& bw)
    {
        const float v0x = sx[1] - sx[0], v0y = sy[1] - sy[0], v1x = sx[2] - sx[0], v1y = sy[2] - sy[0];
        const float v2x = Px - sx[0], v2y = Py - sy[0];
        const float d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
        const float d20 = v2x * v0x + v2y * v0y, d21 = v2x * v1x + v2y * v1y;
        const float denom = d00 * d11 - d01 * d01;
        if (std::abs(denom) < 1e-5) { bu = bv = bw = -1.0f; return; }
        bv = (d11 * d20 - d01 * d21) / denom;
        bw = (d00 * d21 - d01 * d20) / denom;
        bu = 1.0f - bv - bw;
    }

    // the job tile's clamped bounding box (:806-814 / :917-925); false = empty
    inline bool job_bbox(const float sx[3], const float sy[3], int tminx, int tminy, int tmaxx, int tmaxy, float& bminx, float& bminy, float& bmaxx, float& bmaxy)
    {
        bminx = (float)tmaxx; bminy = (float)tmaxy; bmaxx = (float)tminx; bmaxy = (float)tminy;
        for (int k = 0; k < 3; ++k)
        {
            bminx = gmax((float)tminx, gmin(bminx, sx[k])); bminy = gmax((float)tminy, gmin(bminy, sy[k]));
            bmaxx = gmin((float)tmaxx, gmax(bmaxx, sx[k])); bmaxy = gmin((float)tmaxy, gmax(bmaxy, sy[k]));
        }
        return !(bminx > bmaxx || bminy > bmaxy);
    }

    struct ShadowView { const float* depth; int w, h; };

    inline float shadow_sample_depth_uv(const ShadowView& sm, float u, float v) // :252-262 + ShadowMap::sample :223-228
    {
        if (u < 0.0f || u > 1.0f || v < 0.0f || v > 1.0f) return FLT_MAXV;
        int x = (int)std::lround(u * float(sm.w - 1));
        int y = (int)std::lround(v * float(sm.h - 1));
        x = clampi_l2(x, 0, sm.w - 1);
        y = clampi_l2(y, 0, sm.h - 1);
        return sm.depth[(size_t)y * sm.w + x];
    }

    const float POISSON_32[32][2] = {
        {-0.613392f, 0.617481f}, {0.170019f, -0.040254f}, {-0.299417f, 0.791925f}, {0.645680f, 0.493210f}, {-0.651784f, 0.717887f},
        {0.421003f, 0.027070f}, {-0.817194f, -0.271096f}, {-0.705374f, -0.668203f}, {0.977050f, -0.108615f}, {0.063326f, 0.142369f},
        {0.203528f, 0.214331f}, {-0.667531f, 0.326090f}, {-0.098422f, -0.295755f}, {-0.885922f, 0.215369f}, {0.566637f, 0.605213f},
        {0.039766f, -0.396100f}, {0.751946f, 0.453352f}, {0.078707f, -0.715323f}, {-0.075838f, -0.529344f}, {0.724479f, -0.580798f},
        {0.222999f, -0.215125f}, {-0.467574f, -0.405438f}, {-0.248268f, -0.814753f}, {0.354411f, -0.887570f}, {0.175817f, 0.382366f},
        {0.487472f, -0.063082f}, {-0.084078f, 0.898312f}, {0.488876f, -0.783441f}, {0.470016f, 0.217933f}, {-0.696890f, -0.549791f},
        {-0.149693f, 0.605762f}, {0.034211f, 0.979980f}};

    inline uint32_t hash_u32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
    inline float hash01(uint32_t x) { return float(hash_u32(x) & 0x00FFFFFFu) / float(0x01000000u); }
    inline void rotate2(float px, float py, float a, float& ox, float& oy)
    {
        const float c = std::cos(a), s = std::sin(a); // float overloads = libm cosf / sinf
        ox = c * px - s * py;
        oy = s * px + c * py;
    }

    float pcss_shadow_factor(const ShadowView& sm, float u, float v, float z_receiver, float bias, int px, int py) // :333-445
    {
        if (u < 0.0f || u > 1.0f || v < 0.0f || v > 1.0f) return 1.0f;
        const float center = shadow_sample_depth_uv(sm, u, v);
        if (center == FLT_MAXV) return 1.0f;
        const float texelU = 1.0f / float(sm.w), texelV = 1.0f / float(sm.h);
        const float searchU = 18.0f * texelU, searchV = 18.0f * texelV;
        const uint32_t seed = (uint32_t)(px * 1973u ^ py * 9277u ^ 0x9e3779b9u);
        const float ang = hash01(seed) * 6.2831853f;
        float blocker_sum = 0.0f;
        int blocker_cnt = 0;
        const float z_test = z_receiver - bias;
        for (int i = 0; i < 12; ++i)
        {
            float ox, oy;
            rotate2(POISSON_32[i & 31][0], POISSON_32[i & 31][1], ang, ox, oy);
            const float d = shadow_sample_depth_uv(sm, u + ox * searchU, v + oy * searchV);
            if (d == FLT_MAXV) continue;
            if (d < z_test) { blocker_sum += d; blocker_cnt++; }
        }
        if (blocker_cnt <= 0) return 1.0f;
        const float avg = blocker_sum / float(blocker_cnt);
        const float zB = gmax(1e-5f, avg), zR = gmax(1e-5f, z_receiver);
        float ratio = (zR - zB) / zB;
        ratio = gmax(0.0f, ratio);
        float fU = 0.0035f * ratio, fV = 0.0035f * ratio;
        const float ftU = fU / texelU, ftV = fV / texelV;
        float ft = 0.5f * (ftU + ftV);
        ft = clampf_l2(ft, 1.0f, 28.0f);
        fU = ft * texelU;
        fV = ft * texelV;
        float lit_sum = 0.0f;
        int lit_cnt = 0;
        const float ang2 = hash01(seed ^ 0xB5297A4Du) * 6.2831853f;
        for (int i = 0; i < 24; ++i)
        {
            float ox, oy;
            rotate2(POISSON_32[i & 31][0], POISSON_32[i & 31][1], ang2, ox, oy);
            const float d = shadow_sample_depth_uv(sm, u + ox * fU, v + oy * fV);
            if (d == FLT_MAXV) { lit_sum += 1.0f; lit_cnt++; continue; }
            lit_sum += (z_receiver <= d + bias) ? 1.0f : 0.0f;
            lit_cnt++;
        }
        if (lit_cnt <= 0) return 1.0f;
        return lit_sum / float(lit_cnt);
    }

    struct L2Vary { float pos[4]; V3 world, normal; float u, v, view_z; };

    inline L2Vary lerp_vary(const L2Vary& a, const L2Vary& b, float t) // :859-867: a + (b - a) * t
    {
        L2Vary o;
        for (int k = 0; k < 4; ++k) o.pos[k] = a.pos[k] + (b.pos[k] - a.pos[k]) * t;
        o.world = add(a.world, mulk(sub(b.world, a.world), t));
        o.normal = add(a.normal, mulk(sub(b.normal, a.normal), t));
        o.u = a.u + (b.u - a.u) * t;
        o.v = a.v + (b.v - a.v) * t;
        o.view_z = a.view_z + (b.view_z - a.view_z) * t;
        return o;
    }
}

extern "C"
{
    struct ShsoL2Uniforms // struct Uniforms, hello_shadow_mapping_soft.cpp:714-732, as plain data (same layout as in ref_legacy2_harness.cpp)
    {
        float mvp[16], model[16], mv[16], normal_mat[9], light_vp[16];
        float light_dir_world[3], camera_pos[3];
        uint8_t base_color[4];
        int32_t use_texture;
    };

    int32_t shso_l2_shadow_draw(const float* positions, uint32_t n_vertices, const float model[16], const float light_vp[16],
                                int32_t sm_w, int32_t sm_h, int32_t tile_w, int32_t tile_h, float* shadow_depth)
    {
        if (!positions || !shadow_depth || sm_w <= 0 || sm_h <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        float lm[16];
        mat_mul(light_vp, model, lm); // u.light_vp * u.model * vec4: the matrix product first (:779)
        const int cols = (sm_w + tile_w - 1) / tile_w, rows = (sm_h + tile_h - 1) / tile_h;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const int tminx = tx * tile_w, tminy = ty * tile_h;
                const int tmaxx = std::min((tx + 1) * tile_w, sm_w) - 1, tmaxy = std::min((ty + 1) * tile_h, sm_h) - 1;
                for (uint32_t i = 0; i + 2 < n_vertices; i += 3)
                {
                    float sx[3], sy[3], sz[3];
                    bool ok = true;
                    for (int k = 0; k < 3 && ok; ++k)
                    {
                        const V3 p{positions[3 * (i + k)], positions[3 * (i + k) + 1], positions[3 * (i + k) + 2]};
                        float clip[4];
                        mat_mul_point(lm, p, clip);
                        if (std::abs(clip[3]) < 1e-6f) { ok = false; break; }
                        const float ndx = clip[0] / clip[3], ndy = clip[1] / clip[3], ndz = clip[2] / clip[3];
                        sx[k] = (ndx * 0.5f + 0.5f) * float(sm_w - 1);
                        sy[k] = (1.0f - (ndy * 0.5f + 0.5f)) * float(sm_h - 1);
                        sz[k] = ndz;
                    }
                    if (!ok) continue;
                    bool finite = true;
                    for (int k = 0; k < 3; ++k) finite = finite && std::isfinite(sx[k]) && std::isfinite(sy[k]);
                    if (!finite) continue; // the reference casts them to int: undefined behaviour
                    float bminx, bminy, bmaxx, bmaxy;
                    if (!job_bbox(sx, sy, tminx, tminy, tmaxx, tmaxy, bminx, bminy, bmaxx, bmaxy)) continue;
                    const float area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sy[1] - sy[0]) * (sx[2] - sx[0]);
                    if (std::abs(area) < 1e-8f) continue;
                    for (int px = (int)bminx; px <= (int)bmaxx; ++px)
                        for (int py = (int)bminy; py <= (int)bmaxy; ++py)
                        {
                            float bu, bv, bw;
                            bary_legacy(px + 0.5f, py + 0.5f, sx, sy, bu, bv, bw);
                            if (bu < 0 || bv < 0 || bw < 0) continue;
                            const float z = bu * sz[0] + bv * sz[1] + bw * sz[2];
                            if (z < 0.0f || z > 1.0f) continue;
                            if (px < 0 || px >= sm_w || py < 0 || py >= sm_h) continue;
                            float& d = shadow_depth[(size_t)py * sm_w + px];
                            if (z < d) d = z;
                        }
                }
            }
        return 0;
    }

    int32_t shso_l2_camera_draw(const float* positions, const float* normals, const float* uvs, uint32_t n_vertices, const ShsoL2Uniforms* un,
                                const uint8_t* texture_rgba, int32_t tex_w, int32_t tex_h, const float* shadow_depth, int32_t sm_w, int32_t sm_h,
                                int32_t W, int32_t H, int32_t tile_w, int32_t tile_h, uint8_t* canvas_rgba, float* zbuffer)
    {
        if (!positions || !normals || !uvs || !un || !canvas_rgba || !zbuffer || W <= 0 || H <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        const bool has_tex = texture_rgba && tex_w > 0 && tex_h > 0;
        const bool has_shadow = shadow_depth && sm_w > 0 && sm_h > 0;
        const ShadowView sm{shadow_depth, sm_w, sm_h};
        const float* nm = un->normal_mat;
        const int cols = (W + tile_w - 1) / tile_w, rows = (H + tile_h - 1) / tile_h;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const int tminx = tx * tile_w, tminy = ty * tile_h;
                const int tmaxx = std::min((tx + 1) * tile_w, W) - 1, tmaxy = std::min((ty + 1) * tile_h, H) - 1;
                for (uint32_t i = 0; i + 2 < n_vertices; i += 3)
                {
                    // ---- vertex_shader_full x 3 (:746-766)
                    L2Vary vin[3];
                    for (int k = 0; k < 3; ++k)
                    {
                        const V3 p{positions[3 * (i + k)], positions[3 * (i + k) + 1], positions[3 * (i + k) + 2]};
                        const V3 n{normals[3 * (i + k)], normals[3 * (i + k) + 1], normals[3 * (i + k) + 2]};
                        float wh[4], vp[4];
                        mat_mul_point(un->mvp, p, vin[k].pos);
                        mat_mul_point(un->model, p, wh);
                        vin[k].world = V3{wh[0], wh[1], wh[2]};
                        vin[k].normal = normalize3(V3{nm[0] * n.x + nm[3] * n.y + nm[6] * n.z, nm[1] * n.x + nm[4] * n.y + nm[7] * n.z, nm[2] * n.x + nm[5] * n.y + nm[8] * n.z});
                        vin[k].u = uvs[2 * (i + k)];
                        vin[k].v = uvs[2 * (i + k) + 1];
                        mat_mul_point(un->mv, p, vp);
                        vin[k].view_z = vp[2];
                    }
                    // ---- clip_poly_near_z (:869-905): Sutherland-Hodgman against z >= 0 (and w > 1e-6)
                    std::vector<L2Vary> poly;
                    poly.reserve(6);
                    auto inside = [](const L2Vary& v) { return (v.pos[3] > 1e-6f) && (v.pos[2] >= 0.0f); };
                    auto intersect = [](const L2Vary& a, const L2Vary& b) {
                        const float az = a.pos[2], bz = b.pos[2];
                        const float denom = (bz - az);
                        float t = (std::abs(denom) < 1e-8f) ? 0.0f : ((0.0f - az) / denom);
                        t = clampf_l2(t, 0.0f, 1.0f);
                        return lerp_vary(a, b, t);
                    };
                    for (int k = 0; k < 3; ++k)
                    {
                        const L2Vary& A = vin[k];
                        const L2Vary& B = vin[(k + 1) % 3];
                        const bool a_in = inside(A), b_in = inside(B);
                        if (a_in && b_in) poly.push_back(B);
                        else if (a_in && !b_in) poly.push_back(intersect(A, B));
                        else if (!a_in && b_in) { poly.push_back(intersect(A, B)); poly.push_back(B); }
                    }
                    if (poly.size() < 3) continue;
                    for (int ti = 1; ti + 1 < (int)poly.size(); ++ti)
                    {
                        const L2Vary tv[3] = {poly[0], poly[(size_t)ti], poly[(size_t)ti + 1]};
                        bool tri_ok = true;
                        float sx[3], sy[3];
                        for (int k = 0; k < 3; ++k)
                        {
                            if (tv[k].pos[3] <= 1e-6f) { tri_ok = false; break; }
                            const float ndx = tv[k].pos[0] / tv[k].pos[3], ndy = tv[k].pos[1] / tv[k].pos[3];
                            sx[k] = (ndx + 1.0f) * 0.5f * float(W - 1);       // Canvas::clip_to_screen
                            sy[k] = (1.0f - ndy) * 0.5f * float(H - 1);
                        }
                        if (!tri_ok) continue;
                        bool finite = true;
                        for (int k = 0; k < 3; ++k) finite = finite && std::isfinite(sx[k]) && std::isfinite(sy[k]);
                        if (!finite) continue; // undefined behaviour in the reference ((int) of a non-finite float)
                        float bminx, bminy, bmaxx, bmaxy;
                        if (!job_bbox(sx, sy, tminx, tminy, tmaxx, tmaxy, bminx, bminy, bmaxx, bmaxy)) continue;
                        const float area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sy[1] - sy[0]) * (sx[2] - sx[0]);
                        if (std::abs(area) < 1e-8f) continue;
                        for (int px = (int)bminx; px <= (int)bmaxx; ++px)
                            for (int py = (int)bminy; py <= (int)bmaxy; ++py)
                            {
                                float bu, bv, bw;
                                bary_legacy(px + 0.5f, py + 0.5f, sx, sy, bu, bv, bw);
                                if (bu < 0 || bv < 0 || bw < 0) continue;
                                const float vz = bu * tv[0].view_z + bv * tv[1].view_z + bw * tv[2].view_z;
                                // ZBuffer::test_and_set_depth_screen_space: the z-buffer is indexed with the FLIPPED row here
                                const int y_canvas = (H - 1) - py;
                                if (px < 0 || px >= W || y_canvas < 0 || y_canvas >= H) continue;
                                float& d = zbuffer[(size_t)y_canvas * W + px];
                                if (!(vz < d)) continue;
                                d = vz;
                                const float w0 = tv[0].pos[3], w1 = tv[1].pos[3], w2 = tv[2].pos[3];
                                const float iw0 = (std::abs(w0) < 1e-6f) ? 0.0f : 1.0f / w0;
                                const float iw1 = (std::abs(w1) < 1e-6f) ? 0.0f : 1.0f / w1;
                                const float iw2 = (std::abs(w2) < 1e-6f) ? 0.0f : 1.0f / w2;
                                const float iw_sum = bu * iw0 + bv * iw1 + bw * iw2;
                                if (iw_sum <= 1e-8f) continue; // depth already written (:951-961)
                                const V3 in_normal = normalize3(add(add(kmul(bu, tv[0].normal), kmul(bv, tv[1].normal)), kmul(bw, tv[2].normal)));
                                const V3 wp_over_w = add(add(kmul(bu, mulk(tv[0].world, iw0)), kmul(bv, mulk(tv[1].world, iw1))), kmul(bw, mulk(tv[2].world, iw2)));
                                const V3 world_pos{wp_over_w.x / iw_sum, wp_over_w.y / iw_sum, wp_over_w.z / iw_sum};
                                const float uw = bu * (tv[0].u * iw0) + bv * (tv[1].u * iw1) + bw * (tv[2].u * iw2);
                                const float vw = bu * (tv[0].v * iw0) + bv * (tv[1].v * iw1) + bw * (tv[2].v * iw2);
                                const float fu = uw / iw_sum, fv = vw / iw_sum;

                                // ---- fragment_shader_softshadow (:991-1040)
                                const V3 N = normalize3(in_normal);
                                const V3 L = normalize3(V3{-un->light_dir_world[0], -un->light_dir_world[1], -un->light_dir_world[2]});
                                const V3 Vd = normalize3(sub(V3{un->camera_pos[0], un->camera_pos[1], un->camera_pos[2]}, world_pos));
                                V3 base;
                                if (un->use_texture && has_tex)
                                {
                                    // shs::sample_nearest (shs_renderer.hpp:367-377)
                                    const float su = saturate_hdr(fu), sv = saturate_hdr(fv);
                                    int x = (int)std::lround(su * (float)(tex_w - 1));
                                    int y = (int)std::lround(sv * (float)(tex_h - 1));
                                    x = clampi_l2(x, 0, tex_w - 1);
                                    y = clampi_l2(y, 0, tex_h - 1);
                                    const uint8_t* t = texture_rgba + ((size_t)y * tex_w + x) * 4;
                                    base = V3{float(t[0]) / 255.0f, float(t[1]) / 255.0f, float(t[2]) / 255.0f};
                                }
                                else base = V3{float(un->base_color[0]) / 255.0f, float(un->base_color[1]) / 255.0f, float(un->base_color[2]) / 255.0f};
                                const float ndl = dot3(N, L);
                                const float diff = gmax(ndl, 0.0f);
                                const V3 Hh = normalize3(add(L, Vd));
                                const float spec = std::pow(gmax(dot3(N, Hh), 0.0f), 64.0f);
                                const float specular = (0.45f * spec) * 1.0f;
                                float shadow = 1.0f;
                                if (has_shadow)
                                {
                                    float clip[4];
                                    mat_mul_point(un->light_vp, world_pos, clip);          // shadow_uvz_from_world :231-250
                                    if (!(std::abs(clip[3]) < 1e-6f))
                                    {
                                        const float ndx = clip[0] / clip[3], ndy = clip[1] / clip[3], ndz = clip[2] / clip[3];
                                        if (!(ndz < 0.0f || ndz > 1.0f))
                                        {
                                            const float suvx = ndx * 0.5f + 0.5f, suvy = 1.0f - (ndy * 0.5f + 0.5f);
                                            const float slope = 1.0f - gmin(gmax(ndl, 0.0f), 1.0f);
                                            const float bias = 0.0025f + 0.0100f * slope;
                                            shadow = pcss_shadow_factor(sm, suvx, suvy, ndz, bias, px, py);
                                        }
                                    }
                                }
                                const float dcol = diff * 1.0f;
                                uint8_t out[4];
                                const float bc3[3] = {base.x, base.y, base.z};
                                for (int c = 0; c < 3; ++c)
                                {
                                    const float amb = 0.22f * bc3[c];
                                    const float direct = shadow * (dcol * bc3[c] + specular);
                                    const float r = gmin(gmax(amb + direct, 0.0f), 1.0f);   // glm::clamp(amb + direct, 0, 1)
                                    const float r2 = gmin(gmax(r, 0.0f), 1.0f) * 255.0f;    // rgb01_to_color clamps again, then scales
                                    out[c] = (uint8_t)r2;
                                }
                                out[3] = 255;
                                std::memcpy(canvas_rgba + ((size_t)y_canvas * W + px) * 4, out, 4);
                            }
                    }
                }
            }
        return 0;
    }
}

// =====================================================================================================================
// L3: the legacy PBR / IBL demo (hello_pbr.cpp)
// =====================================================================================================================
namespace
{
    inline float std_clamp(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); } // std::clamp
    inline int std_clampi(int v, int lo, int hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
    inline V3 mul3(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
    inline V3 mix3(V3 a, V3 b, float t) { return add(mulk(a, 1.0f - t), mulk(b, t)); }                   // glm::mix

    struct CubeView { const float* data; int size; }; // 6 faces x size^2 x RGB

    inline V3 cube_at(const CubeView& c, int f, int x, int y)
    {
        const float* p = c.data + (((size_t)f * c.size + (size_t)y) * c.size + (size_t)x) * 3;
        return V3{p[0], p[1], p[2]};
    }

    V3 sample_face_bilinear(const CubeView& cm, int face, float u, float v) // ibl.hpp:215-235
    {
        u = std_clamp(u, 0.0f, 1.0f);
        v = std_clamp(v, 0.0f, 1.0f);
        const float fx = u * float(cm.size - 1), fy = v * float(cm.size - 1);
        const int x0 = std_clampi((int)std::floor(fx), 0, cm.size - 1), y0 = std_clampi((int)std::floor(fy), 0, cm.size - 1);
        const int x1 = std_clampi(x0 + 1, 0, cm.size - 1), y1 = std_clampi(y0 + 1, 0, cm.size - 1);
        const float tx = fx - float(x0), ty = fy - float(y0);
        const V3 cx0 = mix3(cube_at(cm, face, x0, y0), cube_at(cm, face, x1, y0), tx);
        const V3 cx1 = mix3(cube_at(cm, face, x0, y1), cube_at(cm, face, x1, y1), tx);
        return mix3(cx0, cx1, ty);
    }

    V3 sample_cubemap_linear(const CubeView& cm, V3 dir) // ibl.hpp:237-270
    {
        if (cm.size <= 0) return V3{0, 0, 0};
        V3 d = dir;
        const float len = std::sqrt(dot3(d, d));
        if (len < 1e-8f) return V3{0, 0, 0};
        d = V3{d.x / len, d.y / len, d.z / len};
        const float ax = std::abs(d.x), ay = std::abs(d.y), az = std::abs(d.z);
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
        return sample_face_bilinear(cm, face, u, v);
    }

    struct L3Vary { float pos[4], prev[4]; V3 world, normal; float u, v, view_z; };

    inline L3Vary lerp_vary3(const L3Vary& a, const L3Vary& b, float t)
    {
        L3Vary o;
        for (int k = 0; k < 4; ++k) { o.pos[k] = a.pos[k] + (b.pos[k] - a.pos[k]) * t; o.prev[k] = a.prev[k] + (b.prev[k] - a.prev[k]) * t; }
        o.world = add(a.world, mulk(sub(b.world, a.world), t));
        o.normal = add(a.normal, mulk(sub(b.normal, a.normal), t));
        o.u = a.u + (b.u - a.u) * t;
        o.v = a.v + (b.v - a.v) * t;
        o.view_z = a.view_z + (b.view_z - a.view_z) * t;
        return o;
    }
}

extern "C"
{
    struct ShsoL3Uniforms // same layout as in ref_legacy3_harness.cpp
    {
        float mvp[16], prev_mvp[16], model[16], mv[16], normal_mat[9], light_vp[16];
        float light_dir_world[3], camera_pos[3];
        uint8_t base_color_srgb[4];
        float metallic, roughness, ao;
        int32_t use_texture;
        float ibl_diffuse_intensity, ibl_specular_intensity, ibl_reflection_strength;
    };

    // draw_triangle_tile_shadow of hello_pbr.cpp:827-875 is the same function as the soft-shadow demo's
    int32_t shso_l3_shadow_draw(const float* positions, uint32_t n_vertices, const float model[16], const float light_vp[16],
                                int32_t sm_w, int32_t sm_h, int32_t tile_w, int32_t tile_h, float* shadow_depth)
    {
        return shso_l2_shadow_draw(positions, n_vertices, model, light_vp, sm_w, sm_h, tile_w, tile_h, shadow_depth);
    }

    int32_t shso_l3_camera_draw(const float* positions, const float* normals, const float* uvs, uint32_t n_vertices, const ShsoL3Uniforms* un,
                                const uint8_t* texture_rgba, int32_t tex_w, int32_t tex_h, const float* shadow_depth, int32_t sm_w, int32_t sm_h,
                                const float* irradiance, int32_t irr_size, const float* prefiltered, const int32_t* spec_sizes, int32_t n_mips,
                                int32_t W, int32_t H, int32_t tile_w, int32_t tile_h, uint8_t* canvas_rgba, float* zbuffer, float* velocity)
    {
        if (!positions || !normals || !uvs || !un || !canvas_rgba || !zbuffer || !velocity || W <= 0 || H <= 0 || tile_w <= 0 || tile_h <= 0) return 1;
        const bool has_tex = texture_rgba && tex_w > 0 && tex_h > 0;
        const bool has_shadow = shadow_depth && sm_w > 0 && sm_h > 0;
        const bool has_ibl = irradiance && irr_size > 0 && prefiltered && spec_sizes && n_mips > 0;
        const CubeView irr{irradiance, irr_size};
        std::vector<CubeView> mips;
        if (has_ibl)
        {
            size_t off = 0;
            for (int m = 0; m < n_mips; ++m) { mips.push_back(CubeView{prefiltered + off, spec_sizes[m]}); off += (size_t)6 * spec_sizes[m] * spec_sizes[m] * 3; }
        }
        const float PI = 3.14159265358979323846f;
        const float* nm = un->normal_mat;
        const int cols = (W + tile_w - 1) / tile_w, rows = (H + tile_h - 1) / tile_h;
        for (int ty = 0; ty < rows; ++ty)
            for (int tx = 0; tx < cols; ++tx)
            {
                const int tminx = tx * tile_w, tminy = ty * tile_h;
                const int tmaxx = std::min((tx + 1) * tile_w, W) - 1, tmaxy = std::min((ty + 1) * tile_h, H) - 1;
                for (uint32_t i = 0; i + 2 < n_vertices; i += 3)
                {
                    L3Vary vin[3];
                    for (int k = 0; k < 3; ++k) // vertex_shader_full :535-556
                    {
                        const V3 p{positions[3 * (i + k)], positions[3 * (i + k) + 1], positions[3 * (i + k) + 2]};
                        const V3 n{normals[3 * (i + k)], normals[3 * (i + k) + 1], normals[3 * (i + k) + 2]};
                        float wh[4], vp[4];
                        mat_mul_point(un->mvp, p, vin[k].pos);
                        mat_mul_point(un->prev_mvp, p, vin[k].prev);
                        mat_mul_point(un->model, p, wh);
                        vin[k].world = V3{wh[0], wh[1], wh[2]};
                        vin[k].normal = normalize3(V3{nm[0] * n.x + nm[3] * n.y + nm[6] * n.z, nm[1] * n.x + nm[4] * n.y + nm[7] * n.z, nm[2] * n.x + nm[5] * n.y + nm[8] * n.z});
                        vin[k].u = uvs[2 * (i + k)];
                        vin[k].v = uvs[2 * (i + k) + 1];
                        mat_mul_point(un->mv, p, vp);
                        vin[k].view_z = vp[2];
                    }
                    std::vector<L3Vary> poly;
                    poly.reserve(6);
                    auto inside = [](const L3Vary& v) { return (v.pos[3] > 1e-6f) && (v.pos[2] >= 0.0f); };
                    auto intersect = [](const L3Vary& a, const L3Vary& b) {
                        const float az = a.pos[2], bz = b.pos[2];
                        const float denom = (bz - az);
                        float t = (std::abs(denom) < 1e-8f) ? 0.0f : ((0.0f - az) / denom);
                        t = saturate_hdr(t); // shs::Math::clampf(t, 0, 1)
                        return lerp_vary3(a, b, t);
                    };
                    for (int k = 0; k < 3; ++k)
                    {
                        const L3Vary& A = vin[k];
                        const L3Vary& B = vin[(k + 1) % 3];
                        const bool a_in = inside(A), b_in = inside(B);
                        if (a_in && b_in) poly.push_back(B);
                        else if (a_in && !b_in) poly.push_back(intersect(A, B));
                        else if (!a_in && b_in) { poly.push_back(intersect(A, B)); poly.push_back(B); }
                    }
                    if (poly.size() < 3) continue;
                    for (int ti = 1; ti + 1 < (int)poly.size(); ++ti)
                    {
                        const L3Vary tv[3] = {poly[0], poly[(size_t)ti], poly[(size_t)ti + 1]};
                        bool tri_ok = true;
                        float sx[3], sy[3];
                        for (int k = 0; k < 3; ++k)
                        {
                            if (tv[k].pos[3] <= 1e-6f) { tri_ok = false; break; }
                            sx[k] = (tv[k].pos[0] / tv[k].pos[3] + 1.0f) * 0.5f * float(W - 1);
                            sy[k] = (1.0f - tv[k].pos[1] / tv[k].pos[3]) * 0.5f * float(H - 1);
                        }
                        if (!tri_ok) continue;
                        bool finite = true;
                        for (int k = 0; k < 3; ++k) finite = finite && std::isfinite(sx[k]) && std::isfinite(sy[k]);
                        if (!finite) continue;
                        float bminx, bminy, bmaxx, bmaxy;
                        if (!job_bbox(sx, sy, tminx, tminy, tmaxx, tmaxy, bminx, bminy, bmaxx, bmaxy)) continue;
                        const float area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sy[1] - sy[0]) * (sx[2] - sx[0]);
                        if (std::abs(area) < 1e-8f) continue;
                        for (int px = (int)bminx; px <= (int)bmaxx; ++px)
                            for (int py = (int)bminy; py <= (int)bmaxy; ++py)
                            {
                                float bu, bv, bw;
                                bary_legacy(px + 0.5f, py + 0.5f, sx, sy, bu, bv, bw);
                                if (bu < 0 || bv < 0 || bw < 0) continue;
                                const float vz = bu * tv[0].view_z + bv * tv[1].view_z + bw * tv[2].view_z;
                                const int y_canvas = (H - 1) - py;
                                if (px < 0 || px >= W || y_canvas < 0 || y_canvas >= H) continue;
                                float& d = zbuffer[(size_t)y_canvas * W + px];
                                if (!(vz < d)) continue;
                                d = vz;
                                const float w0 = tv[0].pos[3], w1 = tv[1].pos[3], w2 = tv[2].pos[3];
                                const float iw0 = (std::abs(w0) < 1e-6f) ? 0.0f : 1.0f / w0;
                                const float iw1 = (std::abs(w1) < 1e-6f) ? 0.0f : 1.0f / w1;
                                const float iw2 = (std::abs(w2) < 1e-6f) ? 0.0f : 1.0f / w2;
                                const float iw_sum = bu * iw0 + bv * iw1 + bw * iw2;
                                if (iw_sum <= 1e-8f) continue;
                                float cpos[4], ppos[4];
                                for (int k = 0; k < 4; ++k)
                                {
                                    cpos[k] = bu * tv[0].pos[k] + bv * tv[1].pos[k] + bw * tv[2].pos[k];
                                    ppos[k] = bu * tv[0].prev[k] + bv * tv[1].prev[k] + bw * tv[2].prev[k];
                                }
                                const V3 in_normal = normalize3(add(add(kmul(bu, tv[0].normal), kmul(bv, tv[1].normal)), kmul(bw, tv[2].normal)));
                                const V3 wp_over_w = add(add(kmul(bu, mulk(tv[0].world, iw0)), kmul(bv, mulk(tv[1].world, iw1))), kmul(bw, mulk(tv[2].world, iw2)));
                                const V3 world_pos{wp_over_w.x / iw_sum, wp_over_w.y / iw_sum, wp_over_w.z / iw_sum};
                                const float fu = (bu * (tv[0].u * iw0) + bv * (tv[1].u * iw1) + bw * (tv[2].u * iw2)) / iw_sum;
                                const float fv = (bu * (tv[0].v * iw0) + bv * (tv[1].v * iw1) + bw * (tv[2].v * iw2)) / iw_sum;

                                // ---- motion vector (:1018-1030): screen-space difference of the affinely interpolated clip positions
                                const float csx = (cpos[0] / cpos[3] + 1.0f) * 0.5f * float(W - 1), csy = (1.0f - cpos[1] / cpos[3]) * 0.5f * float(H - 1);
                                const float psx = (ppos[0] / ppos[3] + 1.0f) * 0.5f * float(W - 1), psy = (1.0f - ppos[1] / ppos[3]) * 0.5f * float(H - 1);
                                float vx = csx - psx, vy = -(csy - psy);
                                const float vlen = std::sqrt(vx * vx + vy * vy);
                                if (vlen > 22.0f && vlen > 1e-6f) { const float k = 22.0f / vlen; vx *= k; vy *= k; }
                                velocity[((size_t)y_canvas * W + px) * 2] = vx;
                                velocity[((size_t)y_canvas * W + px) * 2 + 1] = vy;

                                // ---- fragment_shader_pbr (:627-727)
                                const V3 N = normalize3(in_normal);
                                const V3 Vd = normalize3(sub(V3{un->camera_pos[0], un->camera_pos[1], un->camera_pos[2]}, world_pos));
                                const V3 L = normalize3(V3{-un->light_dir_world[0], -un->light_dir_world[1], -un->light_dir_world[2]});
                                const V3 Hh = normalize3(add(Vd, L));
                                const float NoV = gmax(0.0f, dot3(N, Vd)), NoL = gmax(0.0f, dot3(N, L)), NoH = gmax(0.0f, dot3(N, Hh));
                                V3 srgb;
                                if (un->use_texture && has_tex)
                                {
                                    const float su = saturate_hdr(fu), sv = saturate_hdr(fv);
                                    int x = (int)std::lround(su * (float)(tex_w - 1));
                                    int y = (int)std::lround(sv * (float)(tex_h - 1));
                                    x = clampi_l2(x, 0, tex_w - 1);
                                    y = clampi_l2(y, 0, tex_h - 1);
                                    const uint8_t* t = texture_rgba + ((size_t)y * tex_w + x) * 4;
                                    srgb = V3{float(t[0]) / 255.0f, float(t[1]) / 255.0f, float(t[2]) / 255.0f};
                                }
                                else srgb = V3{float(un->base_color_srgb[0]) / 255.0f, float(un->base_color_srgb[1]) / 255.0f, float(un->base_color_srgb[2]) / 255.0f};
                                // srgb_to_linear: pow(clamp(x, 0, 1), 2.2)
                                const V3 base{std::pow(gmin(gmax(srgb.x, 0.0f), 1.0f), 2.2f), std::pow(gmin(gmax(srgb.y, 0.0f), 1.0f), 2.2f), std::pow(gmin(gmax(srgb.z, 0.0f), 1.0f), 2.2f)};
                                const float metallic = saturate_hdr(un->metallic);
                                const float roughness = saturate_hdr(un->roughness) < 0.04f ? 0.04f : (un->roughness > 1.0f ? 1.0f : un->roughness); // clampf(r, 0.04, 1)
                                const float ao = saturate_hdr(un->ao);
                                const V3 F0 = mix3(V3{0.04f, 0.04f, 0.04f}, base, metallic);
                                // PBR::fresnel_schlick
                                const float fx = 1.0f - saturate_hdr(NoV);
                                const float fx2 = fx * fx;
                                const float fx5 = fx2 * fx2 * fx;
                                V3 F = add(F0, mulk(sub(V3{1.0f, 1.0f, 1.0f}, F0), fx5));
                                F = mul3(F, V3{1.0f, 0.96f, 0.90f});
                                const V3 kd = mulk(sub(V3{1.0f, 1.0f, 1.0f}, F), 1.0f - metallic);
                                const float alpha = roughness * roughness;
                                // PBR::ndf_ggx
                                const float nh = saturate_hdr(NoH);
                                const float a2 = alpha * alpha;
                                const float dd = (nh * nh) * (a2 - 1.0f) + 1.0f;
                                const float D = a2 / (PI * dd * dd);
                                // PBR::g_smith
                                const float rr = (roughness < 0.04f ? 0.04f : (roughness > 1.0f ? 1.0f : roughness)) + 1.0f;
                                const float kk = (rr * rr) / 8.0f;
                                const float nv = saturate_hdr(NoV), nl = saturate_hdr(NoL);
                                const float gv = nv / (nv * (1.0f - kk) + kk), gl = nl / (nl * (1.0f - kk) + kk);
                                const float G = gv * gl;
                                const V3 direct_diffuse = mulk(mul3(kd, base), 1.0f / PI);
                                const float spec_den = gmax(1e-6f, (4.0f * NoV * NoL));
                                const V3 dgf = kmul(D * G, F);
                                const V3 direct_specular{dgf.x / spec_den, dgf.y / spec_den, dgf.z / spec_den};
                                V3 direct = mulk(mul3(add(direct_diffuse, direct_specular), V3{3.0f, 3.0f, 3.0f}), NoL);
                                float shadow = 1.0f;
                                if (has_shadow)
                                {
                                    float clip[4];
                                    mat_mul_point(un->light_vp, world_pos, clip);
                                    if (!(std::abs(clip[3]) < 1e-6f))
                                    {
                                        const float ndx = clip[0] / clip[3], ndy = clip[1] / clip[3], ndz = clip[2] / clip[3];
                                        if (!(ndz < 0.0f || ndz > 1.0f))
                                        {
                                            const float suvx = ndx * 0.5f + 0.5f, suvy = 1.0f - (ndy * 0.5f + 0.5f);
                                            const float slope = 1.0f - gmin(gmax(dot3(N, L), 0.0f), 1.0f);
                                            const float bias = 0.0025f + 0.0100f * slope;
                                            // shadow_factor_pcf_2x2 (:599-621); shs::ShadowMap::sample returns FLT_MAX out of bounds
                                            const float sfx = suvx * float(sm_w - 1), sfy = suvy * float(sm_h - 1);
                                            const int x0 = std_clampi((int)std::floor(sfx), 0, sm_w - 1);
                                            const int y0 = std_clampi((int)std::floor(sfy), 0, sm_h - 1);
                                            const int x1 = std_clampi(x0 + 1, 0, sm_w - 1), y1 = std_clampi(y0 + 1, 0, sm_h - 1);
                                            auto smp = [&](int x, int y) { return shadow_depth[(size_t)y * sm_w + x]; };
                                            const float s00 = (ndz <= smp(x0, y0) + bias) ? 1.0f : 0.0f;
                                            const float s10 = (ndz <= smp(x1, y0) + bias) ? 1.0f : 0.0f;
                                            const float s01 = (ndz <= smp(x0, y1) + bias) ? 1.0f : 0.0f;
                                            const float s11 = (ndz <= smp(x1, y1) + bias) ? 1.0f : 0.0f;
                                            shadow = 0.25f * (s00 + s10 + s01 + s11);
                                        }
                                    }
                                }
                                direct = mulk(direct, shadow);
                                V3 ibl{0.0f, 0.0f, 0.0f};
                                if (has_ibl)
                                {
                                    const V3 irradiance_c = sample_cubemap_linear(irr, N);
                                    V3 diffuse_ibl = mul3(mul3(irradiance_c, base), kd);
                                    diffuse_ibl = mulk(diffuse_ibl, saturate_hdr(un->ibl_diffuse_intensity));
                                    // glm::reflect(-V, N) = I - N * dot(N, I) * 2
                                    const V3 I{-Vd.x, -Vd.y, -Vd.z};
                                    const V3 R = sub(I, mulk(mulk(N, dot3(N, I)), 2.0f));
                                    float lod = roughness * float(n_mips - 1);
                                    // sample_prefiltered_spec_trilinear (ibl.hpp:272-287)
                                    const float mmax = float(n_mips - 1);
                                    lod = std_clamp(lod, 0.0f, mmax);
                                    const int m0 = (int)std::floor(lod);
                                    const int m1 = std::min(m0 + 1, n_mips - 1);
                                    const float tl = lod - float(m0);
                                    const V3 prefiltered_c = mix3(sample_cubemap_linear(mips[(size_t)m0], R), sample_cubemap_linear(mips[(size_t)m1], R), tl);
                                    V3 spec_ibl = mul3(prefiltered_c, F);
                                    spec_ibl = mulk(spec_ibl, saturate_hdr(un->ibl_specular_intensity) * saturate_hdr(un->ibl_reflection_strength));
                                    ibl = add(diffuse_ibl, spec_ibl);
                                }
                                ibl = mulk(ibl, ao);
                                V3 color = add(direct, ibl);
                                color = add(color, mulk(mulk(base, 0.03f), ao));
                                color = mulk(color, 1.75f);
                                color = V3{color.x / (1.0f + color.x), color.y / (1.0f + color.y), color.z / (1.0f + color.z)}; // tonemap_reinhard
                                const float inv_gamma = 1.0f / 2.2f;
                                const float cs[3] = {std::pow(gmin(gmax(color.x, 0.0f), 1.0f), inv_gamma), std::pow(gmin(gmax(color.y, 0.0f), 1.0f), inv_gamma),
                                                     std::pow(gmin(gmax(color.z, 0.0f), 1.0f), inv_gamma)};
                                uint8_t out[4];
                                for (int c = 0; c < 3; ++c) out[c] = (uint8_t)(gmin(gmax(cs[c], 0.0f), 1.0f) * 255.0f);
                                out[3] = 255;
                                std::memcpy(canvas_rgba + ((size_t)y_canvas * W + px) * 4, out, 4);
                            }
                    }
                }
            }
        return 0;
    }
}
