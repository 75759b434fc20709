// tests/cpp/scene_cull_emul.cpp -- TEST INFRASTRUCTURE ONLY.  The device functions of leisure_software_renderer_b200/csrc/scene_cull_core.cuh
// compiled by g++ (-ffp-contract=off == nvcc --fmad=false) and driven like scene_cull.cu drives them (one object per thread; an
// ordered compaction of the non-outside objects), with the C signatures of oracle/oracle_scene_cull.cpp under the prefix shsemu_.
// Nothing in the product links or loads it.
#include <cmath>
#include <cstring>
#define SC_LOGF(x) logf(x) /* the host libm flavour (the device narrows a double log); see scene_cull_core.cuh */
#include <algorithm>
#include <vector>
#include "scene_cull_core.cuh"

using namespace shsb::sc;

extern "C"
{
    int32_t shsemu_cull_objects(const float* bounds10, uint32_t n, const float view_proj[16], uint8_t* classes, uint32_t* visible, uint32_t counts5[5])
    {
        // the planes are computed on the host by the product too (hm::frustum_planes in api.cu); restated here in the same order
        const float* m = view_proj;
        const float r0[4] = {m[0], m[4], m[8], m[12]}, r1[4] = {m[1], m[5], m[9], m[13]}, r2[4] = {m[2], m[6], m[10], m[14]}, r3[4] = {m[3], m[7], m[11], m[15]};
        const float* rows[3] = {r0, r1, r2};
        float planes[24];
        for (int i = 0; i < 6; ++i)
        {
            float eq[4];
            for (int k = 0; k < 4; ++k) eq[k] = (i & 1) ? (r3[k] - rows[i / 2][k]) : (r3[k] + rows[i / 2][k]);
            const float len = std::sqrt(eq[0] * eq[0] + eq[1] * eq[1] + eq[2] * eq[2]);
            if (len <= 1e-8f) { planes[4 * i] = 0; planes[4 * i + 1] = 1; planes[4 * i + 2] = 0; planes[4 * i + 3] = eq[3]; }
            else { planes[4 * i] = eq[0] / len; planes[4 * i + 1] = eq[1] / len; planes[4 * i + 2] = eq[2] / len; planes[4 * i + 3] = eq[3] / len; }
        }
        uint32_t cnt[3] = {0, 0, 0}, nv = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            const int c = classify_object(bounds10 + (size_t)i * 10, planes);
            classes[i] = (uint8_t)c;
            ++cnt[c];
            if (c != OUTSIDE) visible[nv++] = i;
        }
        counts5[0] = n; counts5[1] = cnt[0]; counts5[2] = cnt[1]; counts5[3] = cnt[2]; counts5[4] = nv;
        return 0;
    }

    int32_t shsemu_collect_object_lights(const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const void* records160, uint32_t n_lights,
                                         int32_t cull_mode, uint32_t* out_counts, uint32_t* out_indices8, float* out_dist2_8)
    {
        for (uint32_t o = 0; o < n_objects; ++o)
            out_counts[o] = collect_lights(object_aabbs6 + (size_t)o * 6, visible, n_visible, (const float*)records160, n_lights, cull_mode, out_indices8 + (size_t)o * 8,
                                           out_dist2_8 + (size_t)o * 8);
        return 0;
    }
}

extern "C" int32_t shsemu_tile_depth_range_from_scene(const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const float view[16],
                                                       const float view_proj[16], uint32_t viewport_w, uint32_t viewport_h, uint32_t tile_size, float z_near, float z_far,
                                                       float* out_min, float* out_max)
{
    // like scene_cull.cu: keys initialised, every visible object folds its range into its rectangle (atomicMin / atomicMax on ordered
    // keys -- visited here in REVERSE order to show that the order does not matter), then the closing pass
    const uint32_t tiles_x = (viewport_w + tile_size - 1u) / tile_size, tiles_y = (viewport_h + tile_size - 1u) / tile_size, tiles = tiles_x * tiles_y;
    std::vector<uint32_t> kmin(tiles, depth_key(z_far)), kmax(tiles, depth_key(z_near)), has(tiles, 0u);
    for (uint32_t v = n_visible; v-- > 0;)
    {
        const uint32_t o = visible[v];
        if (o >= n_objects) continue;
        TileRect r;
        if (!project_object(object_aabbs6 + (size_t)o * 6, view, view_proj, z_near, z_far, tiles_x, tiles_y, r)) continue;
        const uint32_t lo = depth_key(r.min_depth), hi = depth_key(r.max_depth);
        for (uint32_t ty = r.ty0; ty <= r.ty1; ++ty)
            for (uint32_t tx = r.tx0; tx <= r.tx1; ++tx)
            {
                const uint32_t t = ty * tiles_x + tx;
                if (lo < kmin[t]) kmin[t] = lo;
                if (hi > kmax[t]) kmax[t] = hi;
                has[t] = 1u;
            }
    }
    for (uint32_t t = 0; t < tiles; ++t) finish_tile(has[t], key_depth(kmin[t]), key_depth(kmax[t]), z_near, z_far, out_min[t], out_max[t]);
    return 0;
}

extern "C" int32_t shsemu_select_object_lights_from_bins(const float* object_aabbs6, uint32_t n_objects, const float view[16], const float view_proj[16], uint32_t bins_x, uint32_t bins_y,
                                                          uint32_t bins_z, int32_t clustered, float z_near, float z_far, uint32_t max_per_bin, const uint32_t* bin_counts,
                                                          const uint32_t* bin_indices, const void* records160, uint32_t n_lights, int32_t cull_mode, uint32_t* out_counts,
                                                          uint32_t* out_indices8, float* out_dist2_8, uint32_t* out_candidates)
{
    BinGrid g{bins_x, bins_y, bins_z, clustered, z_near, z_far, max_per_bin};
    std::vector<uint32_t> seen((n_lights + 31) / 32 + 1);
    for (uint32_t o = 0; o < n_objects; ++o)
    {
        std::fill(seen.begin(), seen.end(), 0u);
        Selection sel;
        out_candidates[o] = select_from_bins(object_aabbs6 + (size_t)o * 6, view, view_proj, g, bin_counts, bin_indices, (const float*)records160, n_lights, cull_mode, seen.data(), sel);
        out_counts[o] = sel.count;
        for (uint32_t k = 0; k < 8; ++k) { out_indices8[8 * o + k] = sel.idx[k]; out_dist2_8[8 * o + k] = sel.d2[k]; }
    }
    return 0;
}

// The software-occlusion kernel's walk (scene_cull.cu: software_occlusion_kernel) with the device functions: objects in the host's
// sorted order; per object the rectangle test over its texels, then every triangle's bbox texels through occ_texel_depth with a
// minimum on the depth's BIT PATTERN (what the device's atomicMin does) -- triangles and texels visited in REVERSE order to show
// that the minimum does not care.
extern "C" int32_t shsemu_software_occlusion(const float* object_aabbs6, uint32_t n_objects, const uint32_t* frustum_visible, uint32_t n_visible, const uint32_t* object_mesh,
                                             const float* object_models16, const uint32_t* mesh_table3, uint32_t n_meshes, const float* vertices, uint32_t n_vertices,
                                             const uint32_t* indices, uint32_t n_indices, const float view[16], const float view_proj[16], int32_t occ_w, int32_t occ_h,
                                             float depth_epsilon, int32_t enable_occlusion, uint8_t* out_occluded, uint32_t* out_visible, uint32_t out_counts4[4], float* out_depth)
{
    (void)n_indices;
    std::vector<uint32_t> bits((size_t)occ_w * occ_h, 0x3F800000u);
    auto as_float = [](uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; };
    auto as_bits = [](float f) { uint32_t b; std::memcpy(&b, &f, 4); return b; };
    for (uint32_t i = 0; i < n_objects; ++i) out_occluded[i] = 0;
    uint32_t nv = 0;
    if (!enable_occlusion)
    {
        for (uint32_t k = 0; k < n_visible; ++k) if (frustum_visible[k] < n_objects) out_visible[nv++] = frustum_visible[k];
    }
    else
    {
        std::vector<float> key(n_objects);
        for (uint32_t i = 0; i < n_objects; ++i) key[i] = occ_view_depth(object_aabbs6 + (size_t)i * 6, view);
        std::vector<uint32_t> sorted(frustum_visible, frustum_visible + n_visible);
        std::sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) { if (a >= n_objects) return false; if (b >= n_objects) return true; return key[a] < key[b]; });
        for (const uint32_t idx : sorted)
        {
            if (idx >= n_objects) continue;
            const OccRect r = occ_project_rect(object_aabbs6 + (size_t)idx * 6, view_proj, occ_w, occ_h);
            bool shows = false;
            if (r.valid)
                for (int y = r.y_max; y >= r.y_min; --y)
                    for (int x = r.x_max; x >= r.x_min; --x) shows = shows || occ_texel_shows(r.z_near, as_float(bits[(size_t)y * occ_w + x]), depth_epsilon);
            const bool occluded = r.valid && !shows;
            out_occluded[idx] = occluded ? 1 : 0;
            if (occluded) continue;
            out_visible[nv++] = idx;
            const uint32_t m = object_mesh[idx];
            if (m >= n_meshes) continue;
            const uint32_t first = mesh_table3[3 * m], count = mesh_table3[3 * m + 1], base_v = mesh_table3[3 * m + 2];
            const float* model = object_models16 + (size_t)idx * 16;
            const uint32_t n_tri = count / 3;
            for (uint32_t ti = n_tri; ti-- > 0;)
            {
                const uint32_t t = ti * 3;
                if (!(t + 2 < count)) continue;
                float xy[3][2], z[3];
                bool ok = true;
                for (int v = 0; v < 3; ++v)
                {
                    const uint32_t vi = base_v + indices[first + t + v];
                    ok = ok && vi < n_vertices && occ_project_vertex(model, vertices + (size_t)vi * 3, view_proj, occ_w, occ_h, xy[v], z[v]);
                }
                if (!ok) continue;
                const OccTri tri = occ_setup_triangle(xy[0], z[0], xy[1], z[1], xy[2], z[2], occ_w, occ_h);
                if (!tri.valid) continue;
                for (int y = tri.max_y; y >= tri.min_y; --y)
                    for (int x = tri.max_x; x >= tri.min_x; --x)
                    {
                        float d;
                        if (!occ_texel_depth(tri, x, y, d)) continue;
                        const uint32_t b = as_bits(d == 0.0f ? 0.0f : d);
                        uint32_t& dst = bits[(size_t)y * occ_w + x];
                        if (b < dst) dst = b;
                    }
            }
        }
    }
    out_counts4[0] = n_objects; out_counts4[1] = std::max(n_visible, nv); out_counts4[2] = nv; out_counts4[3] = out_counts4[1] - nv;
    if (out_depth) for (size_t i = 0; i < bits.size(); ++i) out_depth[i] = as_float(bits[i]);
    return 0;
}
