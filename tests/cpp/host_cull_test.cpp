// CPU property test of the sort-first draw culling in leisure_software_renderer_b200/csrc/host_math.hpp (no CUDA, no reference tree
// needed): whenever bounds_touch_owned_rows / sphere_may_touch_owned_rows say a draw CANNOT reach the owned tile rows, no vertex of a
// mesh inside its bounds may project (with the rasterizer's own screen map, sw_render/rasterizer.hpp:260-269, bbox :280-289) into a
// pixel row of an owned tile row.  Random cameras, transforms (incl. mirrored / flattened scales), bounds and partitions.
#include <cstdint>
#include <cstdio>
#include <random>

#include "host_math.hpp"

namespace hm = shsb_host;

int main()
{
    std::mt19937 rng(12345);
    auto U = [&](float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); };
    long dropped_exact = 0, dropped_sphere = 0, kept = 0, violations = 0, checked = 0;
    for (int iter = 0; iter < 60000; ++iter)
    {
        const int H = 16 + (int)(rng() % 1200), W = 64;
        const int tile = 16, tile_rows = (H + tile - 1) / tile;
        hm::RowOwnership own{H, tile, (int)(rng() % (unsigned)tile_rows), 1 + (int)(rng() % 4), 1 + (int)(rng() % 9)};
        if (own.stride < own.count) own.stride = own.count;
        if (iter % 7 == 0) own.stride = 1 << 24; // a contiguous band
        const float eye[3] = {U(-5, 5), U(0.5f, 8), U(-25, -5)}, target[3] = {U(-2, 2), U(-1, 2), U(-2, 2)}, up[3] = {0, 1, 0};
        const hm::mat4f vp = hm::mul(hm::perspective_lh_no(U(0.5f, 1.3f), (float)W / (float)H * U(0.8f, 2.0f), 0.1f, 200.0f),
                                     hm::look_at_lh({eye[0], eye[1], eye[2]}, {target[0], target[1], target[2]}, {up[0], up[1], up[2]}));
        float vpf[16];
        hm::store(vp, vpf);
        const hm::VpRows rows = hm::vp_rows_of(vpf);
        const float pos[3] = {U(-20, 20), U(-3, 6), U(-20, 30)}, rot[3] = {U(-3, 3), U(-3, 3), U(-3, 3)};
        float scl[3] = {U(0.2f, 3), U(0.2f, 3), U(0.2f, 3)};
        if (iter % 11 == 0) scl[1] = 0.0f;
        if (iter % 13 == 0) scl[0] = -scl[0];
        const hm::vec3f bmin{U(-2, 0.5f), U(-2, 0.5f), U(-2, 0.5f)}, bmax{bmin.x + U(0, 3), bmin.y + U(0, 3), bmin.z + U(0, 3)};
        const hm::mat4f model = hm::model_from_transform(pos, rot, scl);
        const bool a = hm::bounds_touch_owned_rows(own, vp, model, bmin, bmax);
        const bool b = hm::sphere_may_touch_owned_rows(own, rows, pos, scl, bmin, bmax);
        if (a && b) { ++kept; continue; }
        if (!a) ++dropped_exact;
        if (!b) ++dropped_sphere;
        if (!b && a) { /* the sphere test dropped a draw the exact test keeps: must still be a true negative, checked below */ }
        // brute force: points of the bounds (corners + random interior points) through the rasterizer's screen map
        for (int k = 0; k < 64; ++k)
        {
            const float fx = k < 8 ? (float)(k & 1) : U(0, 1), fy = k < 8 ? (float)((k >> 1) & 1) : U(0, 1), fz = k < 8 ? (float)((k >> 2) & 1) : U(0, 1);
            const hm::vec4f wp = hm::mul_v(model, {bmin.x + (bmax.x - bmin.x) * fx, bmin.y + (bmax.y - bmin.y) * fy, bmin.z + (bmax.z - bmin.z) * fz, 1.0f});
            const hm::vec4f clip = hm::mul_v(vp, {wp.x, wp.y, wp.z, 1.0f});
            if (!(clip.w > 0.0f)) { ++violations; continue; } // a dropped draw must lie entirely in front of the camera
            const float sy = (clip.y / clip.w * 0.5f + 0.5f) * (float)(H - 1);
            // a triangle with this vertex touches pixel rows floor(sy) .. ceil(sy) at least (rasterizer.hpp:285-289)
            for (int py : {(int)std::floor(sy), (int)std::ceil(sy)})
            {
                if (py < 0 || py > H - 1) continue;
                ++checked;
                if (own.owns((H - 1 - py) / tile)) ++violations;
            }
        }
    }
    std::printf("kept %ld  dropped by bounds %ld  dropped by sphere %ld  points checked %ld  violations %ld\n", kept, dropped_exact, dropped_sphere, checked, violations);
    return (violations == 0 && dropped_exact > 1000 && dropped_sphere > 1000) ? 0 : 1;
}
