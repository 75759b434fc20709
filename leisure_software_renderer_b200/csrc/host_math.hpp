// host_math.hpp -- host-side float32 restatements the passes need before they can launch kernels.
//
// These run on the CPU because they are per-draw / per-frame scalar work whose bits must equal the
// reference's (the reference computes them with GLM on the host as well):
//   model matrix        T * Rx * Ry * Rz * S via glm::translate/rotate/scale   passes/pass_pbr_forward.hpp:136-141
//   normal matrix       transpose(inverse(mat3(model))) if |det| > 1e-8          shader/builtin_shaders.hpp:92-95
//   camera viewproj     perspectiveLH_NO * lookAtLH                              camera/convention.hpp:19-27
//   inverse(view_proj)  for tile cells                                           lighting/jolt_light_culling.hpp:150
//   frustum planes      normalised rows r3 +- r{0,1,2}                           geometry/frustum_culling.hpp:32-65
//   light camera        texel-snapped ortho camera over the scene AABB           camera/light_camera.hpp:33-99
// Operation order follows GLM's scalar code path (see DESIGN.md "Arithmetic contract"); the translation
// unit is compiled with -ffp-contract=off.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>

namespace shsb_host
{
    struct vec3f { float x, y, z; };
    struct vec4f { float x, y, z, w; };
    struct mat4f { vec4f col[4]; };

    inline vec3f operator+(vec3f a, vec3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
    inline vec3f operator-(vec3f a, vec3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
    inline vec3f operator*(vec3f a, float k) { return {a.x * k, a.y * k, a.z * k}; }
    inline vec4f operator+(vec4f a, vec4f b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
    inline vec4f operator*(vec4f a, float k) { return {a.x * k, a.y * k, a.z * k, a.w * k}; }
    inline float dot(vec3f a, vec3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    inline vec3f cross(vec3f a, vec3f b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
    inline float length(vec3f a) { return std::sqrt(dot(a, a)); }
    inline vec3f normalize(vec3f a) { return a * (1.0f / std::sqrt(dot(a, a))); }
    inline float gmin(float a, float b) { return (b < a) ? b : a; }
    inline float gmax(float a, float b) { return (a < b) ? b : a; }

    inline mat4f load(const float* m) { mat4f r; std::memcpy(&r, m, 64); return r; }
    inline void store(const mat4f& m, float* out) { std::memcpy(out, &m, 64); }
    inline mat4f identity() { return {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}}; }

    inline vec4f mul_v(const mat4f& m, vec4f v)
    {
        const vec4f a = m.col[0] * v.x + m.col[1] * v.y;
        const vec4f b = m.col[2] * v.z + m.col[3] * v.w;
        return a + b;
    }

    inline mat4f mul(const mat4f& a, const mat4f& b)
    {
        mat4f r;
        for (int i = 0; i < 4; ++i)
        {
            const vec4f bi = b.col[i];
            r.col[i] = a.col[0] * bi.x + a.col[1] * bi.y + a.col[2] * bi.z + a.col[3] * bi.w;
        }
        return r;
    }

    inline mat4f translate(const mat4f& m, vec3f v)
    {
        mat4f r = m;
        r.col[3] = m.col[0] * v.x + m.col[1] * v.y + m.col[2] * v.z + m.col[3];
        return r;
    }

    inline mat4f rotate(const mat4f& m, float angle, vec3f v)
    {
        const float c = std::cos(angle), s = std::sin(angle);
        const vec3f axis = normalize(v);
        const vec3f t = axis * (1.0f - c);
        const float ax[3] = {axis.x, axis.y, axis.z}, tp[3] = {t.x, t.y, t.z};
        float R[3][3];
        R[0][0] = c + tp[0] * ax[0];
        R[0][1] = tp[0] * ax[1] + s * ax[2];
        R[0][2] = tp[0] * ax[2] - s * ax[1];
        R[1][0] = tp[1] * ax[0] - s * ax[2];
        R[1][1] = c + tp[1] * ax[1];
        R[1][2] = tp[1] * ax[2] + s * ax[0];
        R[2][0] = tp[2] * ax[0] + s * ax[1];
        R[2][1] = tp[2] * ax[1] - s * ax[0];
        R[2][2] = c + tp[2] * ax[2];
        mat4f r;
        for (int j = 0; j < 3; ++j) r.col[j] = m.col[0] * R[j][0] + m.col[1] * R[j][1] + m.col[2] * R[j][2];
        r.col[3] = m.col[3];
        return r;
    }

    inline mat4f scale(const mat4f& m, vec3f v)
    {
        mat4f r;
        r.col[0] = m.col[0] * v.x;
        r.col[1] = m.col[1] * v.y;
        r.col[2] = m.col[2] * v.z;
        r.col[3] = m.col[3];
        return r;
    }

    inline mat4f model_from_transform(const float pos[3], const float rot[3], const float scl[3])
    {
        mat4f m = identity();
        m = translate(m, {pos[0], pos[1], pos[2]});
        m = rotate(m, rot[0], {1.0f, 0.0f, 0.0f});
        m = rotate(m, rot[1], {0.0f, 1.0f, 0.0f});
        m = rotate(m, rot[2], {0.0f, 0.0f, 1.0f});
        m = scale(m, {scl[0], scl[1], scl[2]});
        return m;
    }

    // glm::determinant(mat4) (func_matrix.inl compute_determinant<4,4>)
    inline float determinant(const mat4f& mm)
    {
        const float* M = &mm.col[0].x;
        auto m = [&](int c, int r) { return M[c * 4 + r]; };
        const float s00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3);
        const float s01 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3);
        const float s02 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2);
        const float s03 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3);
        const float s04 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2);
        const float s05 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1);
        const float d0 = +(m(1, 1) * s00 - m(1, 2) * s01 + m(1, 3) * s02);
        const float d1 = -(m(1, 0) * s00 - m(1, 2) * s03 + m(1, 3) * s04);
        const float d2 = +(m(1, 0) * s01 - m(1, 1) * s03 + m(1, 3) * s05);
        const float d3 = -(m(1, 0) * s02 - m(1, 1) * s04 + m(1, 2) * s05);
        return m(0, 0) * d0 + m(0, 1) * d1 + m(0, 2) * d2 + m(0, 3) * d3;
    }

    inline void normal_matrix(const mat4f& model, float n9[9])
    {
        const float* M = &model.col[0].x;
        auto m = [&](int c, int r) { return M[c * 4 + r]; };
        const float det = +m(0, 0) * (m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2))
                          - m(1, 0) * (m(0, 1) * m(2, 2) - m(2, 1) * m(0, 2))
                          + m(2, 0) * (m(0, 1) * m(1, 2) - m(1, 1) * m(0, 2));
        if (std::fabs(det) > 1e-8f)
        {
            const float k = 1.0f / det;
            float inv[3][3]; // inv[c][r]
            inv[0][0] = +(m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2)) * k;
            inv[1][0] = -(m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2)) * k;
            inv[2][0] = +(m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1)) * k;
            inv[0][1] = -(m(0, 1) * m(2, 2) - m(2, 1) * m(0, 2)) * k;
            inv[1][1] = +(m(0, 0) * m(2, 2) - m(2, 0) * m(0, 2)) * k;
            inv[2][1] = -(m(0, 0) * m(2, 1) - m(2, 0) * m(0, 1)) * k;
            inv[0][2] = +(m(0, 1) * m(1, 2) - m(1, 1) * m(0, 2)) * k;
            inv[1][2] = -(m(0, 0) * m(1, 2) - m(1, 0) * m(0, 2)) * k;
            inv[2][2] = +(m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1)) * k;
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) n9[c * 3 + r] = inv[r][c];
        }
        else
        {
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) n9[c * 3 + r] = m(c, r);
        }
    }

    inline mat4f look_at_lh(vec3f eye, vec3f center, vec3f up)
    {
        const vec3f f = normalize(center - eye);
        const vec3f s = normalize(cross(up, f));
        const vec3f u = cross(f, s);
        mat4f r = identity();
        r.col[0] = {s.x, u.x, f.x, 0.0f};
        r.col[1] = {s.y, u.y, f.y, 0.0f};
        r.col[2] = {s.z, u.z, f.z, 0.0f};
        r.col[3] = {-dot(s, eye), -dot(u, eye), -dot(f, eye), 1.0f};
        return r;
    }

    inline mat4f perspective_lh_no(float fovy, float aspect, float zn, float zf)
    {
        const float t = std::tan(fovy / 2.0f);
        mat4f r{};
        r.col[0].x = 1.0f / (aspect * t);
        r.col[1].y = 1.0f / t;
        r.col[2].z = (zf + zn) / (zf - zn);
        r.col[2].w = 1.0f;
        r.col[3].z = -(2.0f * zf * zn) / (zf - zn);
        return r;
    }

    inline mat4f ortho_lh_no(float l, float r, float b, float t, float n, float f)
    {
        mat4f m = identity();
        m.col[0].x = 2.0f / (r - l);
        m.col[1].y = 2.0f / (t - b);
        m.col[2].z = 2.0f / (f - n);
        m.col[3].x = -(r + l) / (r - l);
        m.col[3].y = -(t + b) / (t - b);
        m.col[3].z = -(f + n) / (f - n);
        return m;
    }

    inline mat4f inverse(const mat4f& mm)
    {
        const float* M = &mm.col[0].x;
        auto m = [&](int c, int r) { return M[c * 4 + r]; };
        const float c00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), c02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), c03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
        const float c04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), c06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), c07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
        const float c08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), c10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), c11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
        const float c12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), c14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), c15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
        const float c16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), c18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), c19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
        const float c20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1), c22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), c23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
        const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
        const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
        const float v0[4] = {m(1, 0), m(0, 0), m(0, 0), m(0, 0)}, v1[4] = {m(1, 1), m(0, 1), m(0, 1), m(0, 1)};
        const float v2[4] = {m(1, 2), m(0, 2), m(0, 2), m(0, 2)}, v3[4] = {m(1, 3), m(0, 3), m(0, 3), m(0, 3)};
        const float sa[4] = {1.0f, -1.0f, 1.0f, -1.0f}, sb[4] = {-1.0f, 1.0f, -1.0f, 1.0f};
        float inv[16];
        for (int i = 0; i < 4; ++i)
        {
            inv[0 + i] = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa[i];
            inv[4 + i] = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb[i];
            inv[8 + i] = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa[i];
            inv[12 + i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb[i];
        }
        const float d0 = m(0, 0) * inv[0], d1 = m(0, 1) * inv[4], d2 = m(0, 2) * inv[8], d3 = m(0, 3) * inv[12];
        const float k = 1.0f / ((d0 + d1) + (d2 + d3));
        mat4f r;
        float* o = &r.col[0].x;
        for (int i = 0; i < 16; ++i) o[i] = inv[i] * k;
        return r;
    }

    // 6 planes (nx, ny, nz, d): Left Right Bottom Top Near Far
    inline void frustum_planes(const mat4f& vp, float out24[24])
    {
        const float* M = &vp.col[0].x;
        const float rows[4][4] = {{M[0], M[4], M[8], M[12]}, {M[1], M[5], M[9], M[13]}, {M[2], M[6], M[10], M[14]}, {M[3], M[7], M[11], M[15]}};
        for (int i = 0; i < 6; ++i)
        {
            float eq[4];
            for (int k = 0; k < 4; ++k) eq[k] = (i & 1) ? (rows[3][k] - rows[i / 2][k]) : (rows[3][k] + rows[i / 2][k]);
            const vec3f n{eq[0], eq[1], eq[2]};
            const float len = length(n);
            if (len <= 1e-8f) { out24[i * 4 + 0] = 0.0f; out24[i * 4 + 1] = 1.0f; out24[i * 4 + 2] = 0.0f; out24[i * 4 + 3] = eq[3]; }
            else { out24[i * 4 + 0] = n.x / len; out24[i * 4 + 1] = n.y / len; out24[i * 4 + 2] = n.z / len; out24[i * 4 + 3] = eq[3] / len; }
        }
    }

    struct Aabb
    {
        vec3f mn{1e30f, 1e30f, 1e30f}, mx{-1e30f, -1e30f, -1e30f};
        void expand(vec3f p)
        {
            mn = {gmin(mn.x, p.x), gmin(mn.y, p.y), gmin(mn.z, p.z)};
            mx = {gmax(mx.x, p.x), gmax(mx.y, p.y), gmax(mx.z, p.z)};
        }
    };

    // build_dir_light_camera_aabb(sun_dir, aabb, 10, resolution).viewproj
    inline mat4f light_camera_viewproj(vec3f sun_dir, const Aabb& box, float margin, unsigned resolution)
    {
        const vec3f dir = normalize(sun_dir);
        const vec3f up = (std::fabs(dir.y) > 0.95f) ? vec3f{0, 0, 1} : vec3f{0, 1, 0};
        const vec3f c = (box.mn + box.mx) * 0.5f;
        const float radius = length((box.mx - box.mn) * 0.5f) + margin;
        const vec3f pos = c - dir * (radius * 2.0f);
        const mat4f view = look_at_lh(pos, c, up);
        float l = 1e30f, r = -1e30f, b = 1e30f, t = -1e30f, n = 1e30f, f = -1e30f;
        for (int i = 0; i < 8; ++i)
        {
            const vec4f p = mul_v(view, {(i & 1) ? box.mx.x : box.mn.x, (i & 2) ? box.mx.y : box.mn.y, (i & 4) ? box.mx.z : box.mn.z, 1.0f});
            l = std::min(l, p.x); r = std::max(r, p.x);
            b = std::min(b, p.y); t = std::max(t, p.y);
            n = std::min(n, p.z); f = std::max(f, p.z);
        }
        l -= margin; r += margin; b -= margin; t += margin; n -= margin; f += margin;
        if (resolution > 0u)
        {
            const float span_x = std::max(r - l, 1e-5f), span_y = std::max(t - b, 1e-5f);
            const float inv_res = 1.0f / static_cast<float>(resolution);
            const float texel_x = span_x * inv_res, texel_y = span_y * inv_res;
            float cx = 0.5f * (l + r), cy = 0.5f * (b + t);
            if (texel_x > 1e-6f) cx = std::floor(cx / texel_x + 0.5f) * texel_x;
            if (texel_y > 1e-6f) cy = std::floor(cy / texel_y + 0.5f) * texel_y;
            const float hx = 0.5f * span_x, hy = 0.5f * span_y;
            l = cx - hx; r = cx + hx; b = cy - hy; t = cy + hy;
        }
        return mul(ortho_lh_no(l, r, b, t, n, f), view);
    }

    // ndc_from_view_depth_lh_no, lighting/jolt_light_culling.hpp:85-93
    inline float ndc_from_view_depth_lh_no(float view_depth, float z_near, float z_far)
    {
        const float n = std::max(z_near, 1e-4f);
        const float f = std::max(z_far, n + 1e-3f);
        const float z = std::clamp(view_depth, n, f);
        const float denom = std::max(f - n, 1e-6f);
        return ((f + n) / denom) - ((2.0f * f * n) / (denom * z));
    }

    // ---------------------------------------------------------------- sort-first partitions (ShsbFrameParams::own_row_*)
    // Tile row ty (0 = top of the frame, `tile` pixels high) is owned iff ty >= first and (ty - first) % stride < count; count <= 0 owns all.
    struct RowOwnership
    {
        int H, tile, first, count, stride;
        bool owns(int ty) const { return count <= 0 || (ty >= first && ((ty - first) % stride) < count); }
        // any owned tile row among the rows that framebuffer (bottom-origin) pixel rows [ymin, ymax] fall into?
        bool any_owned(float ymin, float ymax) const
        {
            if (ymax < 0.0f || ymin > (float)(H - 1)) return false; // entirely above or below the frame: no pixels at all
            const int py0 = (int)std::floor(std::max(ymin, 0.0f)), py1 = (int)std::ceil(std::min(ymax, (float)(H - 1)));
            const int ty0 = (H - 1 - py1) / tile, ty1 = (H - 1 - py0) / tile; // tile rows count from the top
            for (int ty = ty0; ty <= ty1; ++ty) if (owns(ty)) return true;
            return false;
        }
    };

    // Can any triangle of a draw reach a tile row this submission owns?  Conservative: the 8 corners of the mesh's local bounds are
    // projected like vertices (sw_render/rasterizer.hpp:260-269); if all lie in front of the camera, every triangle's clamped pixel
    // bbox lies within the corners' screen-y range (+- two pixels of rounding slack, then whole tile rows).  Any corner at w <= 0 or
    // non-finite means "cannot bound": keep the draw.  A dropped draw contributes no pixels to the owned rows.
    inline bool bounds_touch_owned_rows(const RowOwnership& own, const mat4f& viewproj, const mat4f& model, vec3f bmin, vec3f bmax)
    {
        float ymin = 3.0e38f, ymax = -3.0e38f;
        for (int c = 0; c < 8; ++c)
        {
            const vec4f wp = mul_v(model, {(c & 1) ? bmax.x : bmin.x, (c & 2) ? bmax.y : bmin.y, (c & 4) ? bmax.z : bmin.z, 1.0f});
            const vec4f clip = mul_v(viewproj, {wp.x, wp.y, wp.z, 1.0f});
            if (!(clip.w > 1e-4f) || !std::isfinite(clip.y) || !std::isfinite(clip.w)) return true;
            const float sy = (clip.y / clip.w * 0.5f + 0.5f) * (float)(own.H - 1);
            if (!std::isfinite(sy)) return true;
            ymin = std::min(ymin, sy);
            ymax = std::max(ymax, sy);
        }
        return own.any_owned(ymin - 2.0f, ymax + 2.0f);
    }

    // rows of a view-projection matrix that produce clip.y and clip.w, with the norms of their xyz parts
    struct VpRows { float y[4], w[4], gy, gw; };
    inline VpRows vp_rows_of(const float* m)
    {
        VpRows v{};
        v.y[0] = m[1]; v.y[1] = m[5]; v.y[2] = m[9]; v.y[3] = m[13];
        v.w[0] = m[3]; v.w[1] = m[7]; v.w[2] = m[11]; v.w[3] = m[15];
        v.gy = std::sqrt(m[1] * m[1] + m[5] * m[5] + m[9] * m[9]);
        v.gw = std::sqrt(m[3] * m[3] + m[7] * m[7] + m[11] * m[11]);
        return v;
    }

    // The same question answered without building the model matrix: model = T * R * S maps the mesh's local bounds into the sphere
    // (pos, max|scl| * max distance of a bounds corner from the local origin), whatever the rotation.  Over that sphere clip.y and
    // clip.w vary by at most radius * |gradient|, so ndc.y lies between the extreme ratios of the two intervals (all w > 0, else
    // "cannot bound").  Conservative, never exact: a draw it keeps is tested again with its real matrix.
    inline bool sphere_may_touch_owned_rows(const RowOwnership& own, const VpRows& v, const float pos[3], const float scl[3], vec3f bmin, vec3f bmax)
    {
        const float ex = std::max(std::fabs(bmin.x), std::fabs(bmax.x)), ey = std::max(std::fabs(bmin.y), std::fabs(bmax.y)),
                    ez = std::max(std::fabs(bmin.z), std::fabs(bmax.z));
        const float smax = std::max(std::fabs(scl[0]), std::max(std::fabs(scl[1]), std::fabs(scl[2])));
        const float radius = smax * std::sqrt(ex * ex + ey * ey + ez * ez) * 1.001f + 1e-4f;
        const float yc = v.y[0] * pos[0] + v.y[1] * pos[1] + v.y[2] * pos[2] + v.y[3];
        const float wc = v.w[0] * pos[0] + v.w[1] * pos[1] + v.w[2] * pos[2] + v.w[3];
        const float dy = radius * v.gy, dw = radius * v.gw;
        const float w0 = wc - dw, w1 = wc + dw;
        if (!(w0 > 1e-3f) || !std::isfinite(yc) || !std::isfinite(w1) || !std::isfinite(dy)) return true;
        const float y0 = yc - dy, y1 = yc + dy;
        const float nlo = std::min(std::min(y0 / w0, y0 / w1), std::min(y1 / w0, y1 / w1));
        const float nhi = std::max(std::max(y0 / w0, y0 / w1), std::max(y1 / w0, y1 / w1));
        const float ymin = (nlo * 0.5f + 0.5f) * (float)(own.H - 1) - 2.0f, ymax = (nhi * 0.5f + 0.5f) * (float)(own.H - 1) + 2.0f;
        if (!std::isfinite(ymin) || !std::isfinite(ymax)) return true;
        return own.any_owned(ymin, ymax);
    }
}
