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
