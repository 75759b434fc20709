// oracle/oracle_scene_cull.cpp -- TEST INFRASTRUCTURE ONLY: CPU restatement of the scene-level steps upstream of draw submission
// (SURVEY.md section 8f row 1), each function citing the reference lines it follows
// (/root/reference/cpp-folders/src/shs-renderer-lib/include/shs/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// leg may call it; the product path never does.
// PINNED against the reference's own headers compiled with the JoltPhysics declaration shim (oracle/ref_lightcull_harness.cpp:
// shsref_cull_objects, shsref_collect_object_lights; tests/test_scene_cull_cpu.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace
{
    struct P4 { float nx, ny, nz, d; };
    inline float sdist(const P4& p, float x, float y, float z) { return (p.nx * x + p.ny * y + p.nz * z) + p.d; } // Plane::signed_distance, geometry/volumes.hpp:47-50

    // extract_frustum_planes + make_plane_from_vec4, geometry/frustum_culling.hpp:32-65 (Left, Right, Bottom, Top, Near, Far)
    void frustum_planes(const float* m, P4 out[6])
    {
        const float r0[4] = {m[0], m[4], m[8], m[12]}, r1[4] = {m[1], m[5], m[9], m[13]}, r2[4] = {m[2], m[6], m[10], m[14]}, r3[4] = {m[3], m[7], m[11], m[15]};
        const float* rows[3] = {r0, r1, r2};
        for (int i = 0; i < 6; ++i)
        {
            float eq[4];
            for (int k = 0; k < 4; ++k) eq[k] = (i & 1) ? (r3[k] - rows[i / 2][k]) : (r3[k] + rows[i / 2][k]);
            const float len = std::sqrt(eq[0] * eq[0] + eq[1] * eq[1] + eq[2] * eq[2]);
            if (len <= 1e-8f) out[i] = P4{0.0f, 1.0f, 0.0f, eq[3]};
            else out[i] = P4{eq[0] / len, eq[1] / len, eq[2] / len, eq[3] / len};
        }
    }
}

extern "C"
{
    // cull_vs_frustum<SceneShape> (geometry/jolt_culling.hpp:279-306) with classify_vs_frustum (:258-275), classify_sphere_vs_frustum
    // (:183-197) and classify_aabb_vs_frustum (:201-225).  bounds10: per object sphere centre xyz, radius, aabb min xyz, aabb max xyz
    // (SceneShape::bounding_sphere / world_aabb).  classes: CullClass (0 outside, 1 intersecting, 2 inside); visible: ascending object
    // indices of the non-outside objects; counts5: tested, outside, intersecting, inside, visible.
    int32_t shso_cull_objects(const float* bounds10, uint32_t n, const float view_proj[16], uint8_t* classes, uint32_t* visible, uint32_t counts5[5])
    {
        if ((n && !bounds10) || !view_proj || !classes || !visible || !counts5) return 1;
        P4 fr[6];
        frustum_planes(view_proj, fr);
        uint32_t cnt[3] = {0, 0, 0}, nv = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            const float* b = bounds10 + (size_t)i * 10;
            const float r = std::max(b[3], 0.0f);
            int cls = 2;
            bool decided = false;
            for (int k = 0; k < 6; ++k)
            {
                const float dist = sdist(fr[k], b[0], b[1], b[2]);
                if (dist < -(r + 1e-5f)) { cls = 0; decided = true; break; }
                if (dist < (r + 1e-5f)) cls = 1;
            }
            if (!decided && cls == 1) // the sphere intersects: the world AABB decides
            {
                bool inside = true;
                for (int k = 0; k < 6 && cls != 0; ++k)
                {
                    const P4& p = fr[k];
                    if (sdist(p, (p.nx >= 0.0f) ? b[7] : b[4], (p.ny >= 0.0f) ? b[8] : b[5], (p.nz >= 0.0f) ? b[9] : b[6]) < -1e-5f) { cls = 0; break; }
                    if (sdist(p, (p.nx >= 0.0f) ? b[4] : b[7], (p.ny >= 0.0f) ? b[5] : b[8], (p.nz >= 0.0f) ? b[6] : b[9]) < 1e-5f) inside = false;
                }
                if (cls != 0) cls = inside ? 2 : 1;
            }
            classes[i] = (uint8_t)cls;
            ++cnt[cls];
            if (cls != 0) visible[nv++] = i;
        }
        counts5[0] = n; counts5[1] = cnt[0]; counts5[2] = cnt[1]; counts5[3] = cnt[2]; counts5[4] = nv;
        return 0;
    }

    // collect_object_lights (lighting/light_runtime.hpp:592-616) for n_objects objects: light_affects_object (:570-590),
    // intersect_sphere_aabb / intersect_aabb_aabb (:239-252), add_light_candidate (:263-289).  records160: CullingLightGPU records
    // (LightInstance::packed; position_range.xyz stands for LightProperties::position_ws, which the packers copy there,
    // light_types.hpp:332).  visible: light indices in visit order.  Outputs per object: count, 8 indices, 8 squared distances
    // (unused slots zero, like the value-initialised LightSelection).
    int32_t shso_collect_object_lights(const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const void* records160, uint32_t n_lights,
                                       int32_t cull_mode, uint32_t* out_counts, uint32_t* out_indices8, float* out_dist2_8)
    {
        if ((n_objects && !object_aabbs6) || (n_visible && !visible) || (n_lights && !records160) || !out_counts || !out_indices8 || !out_dist2_8) return 1;
        if (cull_mode < 0 || cull_mode > 2) return 1;
        const uint8_t* recs = (const uint8_t*)records160;
        for (uint32_t o = 0; o < n_objects; ++o)
        {
            const float* mn = object_aabbs6 + (size_t)o * 6;
            const float* mx = mn + 3;
            const float c[3] = {0.5f * (mn[0] + mx[0]), 0.5f * (mn[1] + mx[1]), 0.5f * (mn[2] + mx[2])}; // AABB::center, geometry/aabb.hpp:26
            uint32_t* idx = out_indices8 + (size_t)o * 8;
            float* d2s = out_dist2_8 + (size_t)o * 8;
            for (int k = 0; k < 8; ++k) { idx[k] = 0; d2s[k] = 0.0f; }
            uint32_t count = 0;
            for (uint32_t v = 0; v < n_visible; ++v)
            {
                const uint32_t li = visible[v];
                if (li >= n_lights) continue;
                float rec[40];
                std::memcpy(rec, recs + (size_t)li * 160, 160);
                const float* pos = rec;            // position_range
                const float* sph = rec + 28;       // cull_sphere
                const float* amin = rec + 32;      // cull_aabb_min
                const float* amax = rec + 36;      // cull_aabb_max
                bool affects = true;
                if (cull_mode == 1)
                {
                    const float radius = std::max(sph[3], 0.0f);
                    float d[3];
                    for (int k = 0; k < 3; ++k)
                    {
                        const float lo = (sph[k] < mn[k]) ? mn[k] : sph[k];   // glm::clamp = min(max(x, lo), hi)
                        const float cl = (mx[k] < lo) ? mx[k] : lo;
                        d[k] = sph[k] - cl;
                    }
                    affects = (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) <= radius * radius;
                }
                else if (cull_mode == 2)
                {
                    for (int k = 0; k < 3 && affects; ++k)
                        if (amax[k] < mn[k] || amin[k] > mx[k]) affects = false;
                }
                if (!affects) continue;
                const float dx = pos[0] - c[0], dy = pos[1] - c[1], dz = pos[2] - c[2];
                const float dist2 = dx * dx + dy * dy + dz * dz;
                if (count < 8) { idx[count] = li; d2s[count] = dist2; ++count; continue; }
                uint32_t farthest = 0;
                float far_d2 = d2s[0];
                for (uint32_t i = 1; i < 8; ++i)
                    if (d2s[i] > far_d2) { farthest = i; far_d2 = d2s[i]; }
                if (dist2 < far_d2) { idx[farthest] = li; d2s[farthest] = dist2; }
            }
            out_counts[o] = count;
        }
        return 0;
    }

    // build_tile_view_depth_range_from_scene (lighting/light_culling_runtime.hpp:188-264) with project_aabb_bounds (:92-153),
    // aabb_corners (:78-90), ndc_x_to_bin / ndc_y_to_bin_top_origin (:155-167).  Objects enter as world AABBs (SceneShape::world_aabb);
    // visible: scene indices in visit order (entries >= n_objects are skipped).  out_min / out_max: one float per tile, tile (0,0) top-left.
    int32_t shso_tile_depth_range_from_scene(const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const float view[16],
                                             const float view_proj[16], uint32_t viewport_w, uint32_t viewport_h, uint32_t tile_size, float z_near, float z_far,
                                             float* out_min, float* out_max)
    {
        if (!view || !view_proj || !out_min || !out_max || viewport_w == 0 || viewport_h == 0 || tile_size == 0) return 1;
        const uint32_t tiles_x = (viewport_w + tile_size - 1u) / tile_size, tiles_y = (viewport_h + tile_size - 1u) / tile_size, total = tiles_x * tiles_y;
        std::vector<uint8_t> has(total, 0);
        for (uint32_t t = 0; t < total; ++t) { out_min[t] = z_far; out_max[t] = z_near; }
        auto mul = [](const float* m, int row, float x, float y, float z) { return (m[row] * x + m[4 + row] * y) + (m[8 + row] * z + m[12 + row] * 1.0f); }; // mat4 * vec4(p, 1), GLM scalar order
        auto xbin = [](float ndc_x, uint32_t bins) { const float u = std::clamp(ndc_x * 0.5f + 0.5f, 0.0f, 0.999999f); return std::min((uint32_t)(u * (float)bins), bins - 1u); };
        auto ybin = [](float ndc_y, uint32_t bins) { const float v = std::clamp(1.0f - (ndc_y * 0.5f + 0.5f), 0.0f, 0.999999f); return std::min((uint32_t)(v * (float)bins), bins - 1u); };
        for (uint32_t vi = 0; vi < n_visible; ++vi)
        {
            const uint32_t o = visible[vi];
            if (o >= n_objects) continue;
            const float* b = object_aabbs6 + (size_t)o * 6;
            bool any = false;
            float min_x = 1.0f, max_x = -1.0f, min_y = 1.0f, max_y = -1.0f, min_d = z_far, max_d = z_near;
            for (int c = 0; c < 8; ++c)
            {
                const float x = (c & 1) ? b[3] : b[0], y = (c & 2) ? b[4] : b[1], z = (c & 4) ? b[5] : b[2];
                const float cw = mul(view_proj, 3, x, y, z);
                if (cw <= 1e-5f) continue;
                const float nx = mul(view_proj, 0, x, y, z) / cw, ny = mul(view_proj, 1, x, y, z) / cw;
                min_x = std::min(min_x, nx); max_x = std::max(max_x, nx);
                min_y = std::min(min_y, ny); max_y = std::max(max_y, ny);
                const float vd = mul(view, 2, x, y, z);
                if (vd > 1e-5f) { min_d = std::min(min_d, vd); max_d = std::max(max_d, vd); }
                any = true;
            }
            if (!any) continue;
            min_x = std::clamp(min_x, -1.0f, 1.0f); max_x = std::clamp(max_x, -1.0f, 1.0f);
            min_y = std::clamp(min_y, -1.0f, 1.0f); max_y = std::clamp(max_y, -1.0f, 1.0f);
            if (min_x > max_x) std::swap(min_x, max_x);
            if (min_y > max_y) std::swap(min_y, max_y);
            min_d = std::clamp(min_d, z_near, z_far);
            max_d = std::clamp(max_d, z_near, z_far);
            if (min_d > max_d) { min_d = z_near; max_d = z_far; }
            const uint32_t tx0 = xbin(min_x, tiles_x), tx1 = xbin(max_x, tiles_x), ty0 = ybin(max_y, tiles_y), ty1 = ybin(min_y, tiles_y);
            for (uint32_t ty = ty0; ty <= ty1; ++ty)
                for (uint32_t tx = tx0; tx <= tx1; ++tx)
                {
                    const uint32_t t = ty * tiles_x + tx;
                    if (t >= total) continue;
                    out_min[t] = std::min(out_min[t], min_d);
                    out_max[t] = std::max(out_max[t], max_d);
                    has[t] = 1;
                }
        }
        for (uint32_t t = 0; t < total; ++t)
            if (!has[t] || out_min[t] > out_max[t]) { out_min[t] = z_near; out_max[t] = z_far; }
        return 0;
    }

    // gather_light_scene_candidates_for_aabb (lighting/light_culling_runtime.hpp:373-449) followed by collect_object_lights
    // (lighting/light_runtime.hpp:592-616), per object, as exp-plumbing/hello_light_types_culling_sw.cpp:968-996 chains them.
    // The bins are LightBinCullingData::bin_local_light_lists in the capped device layout (bin_counts[bins], bin_indices[bins * max_per_bin],
    // bin = tz * bins_x * bins_y + ty * bins_x + tx); the visible-light list is the identity, so local index == scene index == light index
    // and the fallback (no bins / nothing of the box in front of the camera) is every light in order.  z_near / z_far are
    // LightBinCullingData's (max(cfg.z_near, 1e-4), max(cfg.z_far, z_near + 1e-3), :278-279).  out_candidates: the gathered list's length.
    int32_t shso_select_object_lights_from_bins(const float* object_aabbs6, uint32_t n_objects, const float view[16], const float view_proj[16], uint32_t bins_x, uint32_t bins_y,
                                                uint32_t bins_z, int32_t clustered, float z_near, float z_far, uint32_t max_per_bin, const uint32_t* bin_counts,
                                                const uint32_t* bin_indices, const void* records160, uint32_t n_lights, int32_t cull_mode, uint32_t* out_counts,
                                                uint32_t* out_indices8, float* out_dist2_8, uint32_t* out_candidates)
    {
        if (!view || !view_proj || !out_counts || !out_indices8 || !out_dist2_8 || !out_candidates || cull_mode < 0 || cull_mode > 2) return 1;
        auto mul = [](const float* m, int row, float x, float y, float z) { return (m[row] * x + m[4 + row] * y) + (m[8 + row] * z + m[12 + row] * 1.0f); };
        auto xbin = [](float ndc_x, uint32_t bins) { const float u = std::clamp(ndc_x * 0.5f + 0.5f, 0.0f, 0.999999f); return std::min((uint32_t)(u * (float)bins), bins - 1u); };
        auto ybin = [](float ndc_y, uint32_t bins) { const float v = std::clamp(1.0f - (ndc_y * 0.5f + 0.5f), 0.0f, 0.999999f); return std::min((uint32_t)(v * (float)bins), bins - 1u); };
        auto slice = [](float view_depth, float zn_, float zf_, uint32_t slices) -> uint32_t { // view_depth_to_cluster_slice :170-186
            if (slices <= 1u) return 0u;
            const float zn = std::max(zn_, 1e-4f), zf = std::max(zf_, zn + 1e-3f);
            const float d = std::clamp(view_depth, zn, zf);
            const float log_ratio = std::log(zf / zn);
            if (log_ratio <= 1e-6f) return 0u;
            const float t = std::clamp(std::log(d / zn) / log_ratio, 0.0f, 0.999999f);
            return std::min((uint32_t)(t * (float)slices), slices - 1u);
        };
        const bool has_bins = bins_x > 0 && bins_y > 0 && bins_z > 0 && bin_counts && bin_indices;
        std::vector<uint32_t> cand, identity(n_lights);
        for (uint32_t i = 0; i < n_lights; ++i) identity[i] = i;
        for (uint32_t o = 0; o < n_objects; ++o)
        {
            const float* b = object_aabbs6 + (size_t)o * 6;
            const std::vector<uint32_t>* list = &identity;
            bool any = false;
            float min_x = 1.0f, max_x = -1.0f, min_y = 1.0f, max_y = -1.0f, min_d = z_far, max_d = z_near;
            if (has_bins)
            {
                for (int c = 0; c < 8; ++c) // project_aabb_bounds :92-153
                {
                    const float x = (c & 1) ? b[3] : b[0], y = (c & 2) ? b[4] : b[1], z = (c & 4) ? b[5] : b[2];
                    const float cw = mul(view_proj, 3, x, y, z);
                    if (cw <= 1e-5f) continue;
                    const float nx = mul(view_proj, 0, x, y, z) / cw, ny = mul(view_proj, 1, x, y, z) / cw;
                    min_x = std::min(min_x, nx); max_x = std::max(max_x, nx);
                    min_y = std::min(min_y, ny); max_y = std::max(max_y, ny);
                    const float vd = mul(view, 2, x, y, z);
                    if (vd > 1e-5f) { min_d = std::min(min_d, vd); max_d = std::max(max_d, vd); }
                    any = true;
                }
            }
            if (any)
            {
                min_x = std::clamp(min_x, -1.0f, 1.0f); max_x = std::clamp(max_x, -1.0f, 1.0f);
                min_y = std::clamp(min_y, -1.0f, 1.0f); max_y = std::clamp(max_y, -1.0f, 1.0f);
                if (min_x > max_x) std::swap(min_x, max_x);
                if (min_y > max_y) std::swap(min_y, max_y);
                min_d = std::clamp(min_d, z_near, z_far);
                max_d = std::clamp(max_d, z_near, z_far);
                if (min_d > max_d) { min_d = z_near; max_d = z_far; }
                const uint32_t tx0 = xbin(min_x, bins_x), tx1 = xbin(max_x, bins_x), ty0 = ybin(max_y, bins_y), ty1 = ybin(min_y, bins_y);
                uint32_t tz0 = 0u, tz1 = std::max(bins_z, 1u) - 1u;
                if (clustered && bins_z > 1u)
                {
                    tz0 = slice(min_d, z_near, z_far, bins_z);
                    tz1 = slice(max_d, z_near, z_far, bins_z);
                    if (tz0 > tz1) std::swap(tz0, tz1);
                }
                cand.clear();
                for (uint32_t tz = tz0; tz <= tz1; ++tz)
                    for (uint32_t ty = ty0; ty <= ty1; ++ty)
                        for (uint32_t tx = tx0; tx <= tx1; ++tx)
                        {
                            const uint32_t bin = tz * (bins_x * bins_y) + ty * bins_x + tx;
                            const uint32_t n = std::min(bin_counts[bin], max_per_bin);
                            for (uint32_t k = 0; k < n; ++k)
                            {
                                const uint32_t li = bin_indices[(size_t)bin * max_per_bin + k];
                                if (li >= n_lights) continue;
                                if (std::find(cand.begin(), cand.end(), li) == cand.end()) cand.push_back(li);
                            }
                        }
                list = &cand;
            }
            out_candidates[o] = (uint32_t)list->size();
            if (int32_t rc = shso_collect_object_lights(b, 1, list->data(), (uint32_t)list->size(), records160, n_lights, cull_mode, out_counts + o, out_indices8 + (size_t)o * 8,
                                                        out_dist2_8 + (size_t)o * 8))
                return rc;
        }
        return 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Software occlusion, geometry/culling_software.hpp:41-333, as scene/scene_culling.hpp:186-222 drives it.  PINNED against the
// reference's own header compiled with the JoltPhysics declaration shim (oracle/ref_occlusion_harness.cpp; tests/test_occlusion_cpu.py).
namespace
{
    inline void mul4(const float* m, float x, float y, float z, float w, float out[4]) // glm's scalar mat4 * vec4
    {
        for (int r = 0; r < 4; ++r) out[r] = (m[r] * x + m[4 + r] * y) + (m[8 + r] * z + m[12 + r] * w);
    }
    inline float edge_fn(const float* a, const float* b, const float* p) { return (p[0] - a[0]) * (b[1] - a[1]) - (p[1] - a[1]) * (b[0] - a[0]); } // :41-44

    bool world_to_screen(const float* world, const float* vp, int w, int h, float xy[2], float& depth01) // :46-63
    {
        float clip[4];
        mul4(vp, world[0], world[1], world[2], 1.0f, clip);
        if (clip[3] <= 0.001f) return false;
        const float ndc[3] = {clip[0] / clip[3], clip[1] / clip[3], clip[2] / clip[3]};
        if (ndc[2] < -1.0f || ndc[2] > 1.0f) return false;
        xy[0] = (ndc[0] + 1.0f) * 0.5f * (float)w;
        xy[1] = (ndc[1] + 1.0f) * 0.5f * (float)h;
        depth01 = ndc[2] * 0.5f + 0.5f;
        return true;
    }

    void depth_triangle(float* depth, int w, int h, const float* p0, float z0, const float* p1, float z1, const float* p2, float z2) // :65-115
    {
        const float area = edge_fn(p0, p1, p2);
        if (std::abs(area) <= 1e-6f) return;
        const int min_x = std::max(0, (int)std::floor(std::min(p0[0], std::min(p1[0], p2[0]))));
        const int min_y = std::max(0, (int)std::floor(std::min(p0[1], std::min(p1[1], p2[1]))));
        const int max_x = std::min(w - 1, (int)std::ceil(std::max(p0[0], std::max(p1[0], p2[0]))));
        const int max_y = std::min(h - 1, (int)std::ceil(std::max(p0[1], std::max(p1[1], p2[1]))));
        if (min_x > max_x || min_y > max_y) return;
        const bool ccw = area > 0.0f;
        for (int y = min_y; y <= max_y; ++y)
            for (int x = min_x; x <= max_x; ++x)
            {
                const float p[2] = {(float)x + 0.5f, (float)y + 0.5f};
                const float w0 = edge_fn(p1, p2, p), w1 = edge_fn(p2, p0, p), w2 = edge_fn(p0, p1, p);
                if (!(ccw ? (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) : (w0 <= 0.0f && w1 <= 0.0f && w2 <= 0.0f))) continue;
                const float d = (w0 / area) * z0 + (w1 / area) * z1 + (w2 / area) * z2;
                if (d < 0.0f || d > 1.0f) continue;
                float& dst = depth[(size_t)y * w + x];
                if (d < dst) dst = d;
            }
    }
}

extern "C" int32_t shso_software_occlusion(const float* object_aabbs6, uint32_t n_objects, const uint32_t* frustum_visible, uint32_t n_visible, const uint32_t* object_mesh,
                                           const float* object_models16, const uint32_t* mesh_table3, uint32_t n_meshes, const float* vertices, uint32_t n_vertices,
                                           const uint32_t* indices, uint32_t n_indices, const float view[16], const float view_proj[16], int32_t occ_w, int32_t occ_h,
                                           float depth_epsilon, int32_t enable_occlusion, uint8_t* out_occluded, uint32_t* out_visible, uint32_t out_counts4[4], float* out_depth)
{
    (void)n_vertices; (void)n_indices;
    std::vector<float> depth((size_t)occ_w * occ_h, 1.0f);
    for (uint32_t i = 0; i < n_objects; ++i) out_occluded[i] = 0;
    uint32_t nv = 0;
    if (!enable_occlusion) // :270-284
    {
        for (uint32_t k = 0; k < n_visible; ++k) if (frustum_visible[k] < n_objects) out_visible[nv++] = frustum_visible[k];
    }
    else
    {
        auto view_depth = [&](uint32_t i) { // view_depth_of_aabb_center, :224-231
            const float* b = object_aabbs6 + (size_t)i * 6;
            float v[4];
            mul4(view, 0.5f * (b[0] + b[3]), 0.5f * (b[1] + b[4]), 0.5f * (b[2] + b[5]), 1.0f, v);
            return v[2];
        };
        std::vector<uint32_t> sorted(frustum_visible, frustum_visible + n_visible);
        std::sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) { // :288-297
            if (a >= n_objects) return false;
            if (b >= n_objects) return true;
            return view_depth(a) < view_depth(b);
        });
        for (const uint32_t idx : sorted)
        {
            if (idx >= n_objects) continue;
            // project_aabb_to_screen_rect, :145-199
            const float* b = object_aabbs6 + (size_t)idx * 6;
            float min_x = (float)occ_w, min_y = (float)occ_h, max_x = -1.0f, max_y = -1.0f, near_depth = 1.0f;
            bool any = false;
            for (int c = 0; c < 8; ++c)
            {
                float clip[4];
                mul4(view_proj, b[(c & 1) ? 3 : 0], b[(c & 2) ? 4 : 1], b[(c & 4) ? 5 : 2], 1.0f, clip);
                if (clip[3] <= 0.001f) continue;
                const float ndc[3] = {clip[0] / clip[3], clip[1] / clip[3], clip[2] / clip[3]};
                const float z01 = ndc[2] * 0.5f + 0.5f;
                if (z01 < 0.0f || z01 > 1.0f) continue;
                const float sx = (ndc[0] + 1.0f) * 0.5f * (float)occ_w, sy = (ndc[1] + 1.0f) * 0.5f * (float)occ_h;
                min_x = std::min(min_x, sx); min_y = std::min(min_y, sy); max_x = std::max(max_x, sx); max_y = std::max(max_y, sy);
                near_depth = std::min(near_depth, z01);
                any = true;
            }
            bool occluded = false;
            if (any)
            {
                const int x0 = std::max(0, (int)std::floor(min_x)), y0 = std::max(0, (int)std::floor(min_y));
                const int x1 = std::min(occ_w - 1, (int)std::ceil(max_x)), y1 = std::min(occ_h - 1, (int)std::ceil(max_y));
                const float z_near = std::clamp(near_depth, 0.0f, 1.0f);
                if (x0 <= x1 && y0 <= y1) // is_rect_occluded, :201-222
                {
                    occluded = true;
                    for (int y = y0; y <= y1 && occluded; ++y)
                        for (int x = x0; x <= x1; ++x)
                            if (z_near <= depth[(size_t)y * occ_w + x] + depth_epsilon) { occluded = false; break; }
                }
            }
            out_occluded[idx] = occluded ? 1 : 0;
            if (occluded) continue;
            out_visible[nv++] = idx;
            const uint32_t m = object_mesh[idx];
            if (m >= n_meshes) continue;
            const uint32_t first = mesh_table3[3 * m], count = mesh_table3[3 * m + 1], base_v = mesh_table3[3 * m + 2];
            const float* model = object_models16 + (size_t)idx * 16;
            for (uint32_t i = 0; i + 2 < count; i += 3) // rasterize_mesh_depth_transformed, :116-143
            {
                float s[3][2], z[3];
                bool ok = true;
                for (int v = 0; v < 3 && ok; ++v)
                {
                    const float* lp = vertices + (size_t)(base_v + indices[first + i + v]) * 3;
                    float wp[4];
                    mul4(model, lp[0], lp[1], lp[2], 1.0f, wp);
                    ok = world_to_screen(wp, view_proj, occ_w, occ_h, s[v], z[v]);
                }
                if (ok) depth_triangle(depth.data(), occ_w, occ_h, s[0], z[0], s[1], z[1], s[2], z[2]);
            }
        }
    }
    out_counts4[0] = n_objects;
    out_counts4[1] = std::max(n_visible, nv);
    out_counts4[2] = nv;
    out_counts4[3] = out_counts4[1] - nv;
    if (out_depth) std::memcpy(out_depth, depth.data(), depth.size() * sizeof(float));
    return 0;
}
