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
