// scene_cull_core.cuh -- scene-level steps immediately upstream of draw submission as host/device
// inline functions; scene_cull.cu wraps them in kernels, tests/cpp/scene_cull_emul.cpp compiles the same functions with g++ and checks
// them against the pinned oracle on the CPU box (test code: the product launches kernels).
//   classify_object      classify_vs_frustum for a FastCullable + HasWorldAABB object, geometry/jolt_culling.hpp:129-181, 239-275
//                        (sphere first, AABB p-vertex / n-vertex test when the sphere intersects; tolerances 1e-5, :118-122)
//   collect_lights       collect_object_lights + light_affects_object + add_light_candidate, lighting/light_runtime.hpp:239-252,
//                        263-289, 570-616: up to 8 nearest affecting lights per object; the SLOT a light lands in depends on visit
//                        order (replace the first farthest slot when strictly nearer), so each object walks its candidates serially
// IEEE binary32, unfused, GLM's scalar order (compiled --fmad=false / -ffp-contract=off).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define SC_HD __host__ __device__ __forceinline__
#else
#define SC_HD inline
#endif

namespace shsb
{
    namespace sc
    {
        constexpr uint32_t LIGHT_SELECTION_CAPACITY = 8u;      // kLightSelectionCapacity, light_runtime.hpp:22
        constexpr uint32_t LIGHT_RECORD_FLOATS = 40u;          // CullingLightGPU, 160 bytes
        constexpr uint32_t REC_POSITION = 0u, REC_CULL_SPHERE = 28u, REC_CULL_AABB_MIN = 32u, REC_CULL_AABB_MAX = 36u; // float offsets
        enum CullClass { OUTSIDE = 0, INTERSECTING = 1, INSIDE = 2 };
        enum LightObjectCullMode { MODE_NONE = 0, MODE_SPHERE_AABB = 1, MODE_VOLUME_AABB = 2 };

        SC_HD float gmax(float a, float b) { return (a < b) ? b : a; }
        SC_HD float gmin(float a, float b) { return (b < a) ? b : a; }
        SC_HD float sdist(const float* pl, float x, float y, float z) { return (pl[0] * x + pl[1] * y + pl[2] * z) + pl[3]; } // dot(normal, p) + d

        // bounds10: sphere centre xyz, radius, aabb min xyz, aabb max xyz; planes24: 6 x (nx, ny, nz, d)
        SC_HD int classify_object(const float* b, const float* planes24)
        {
            const float r = gmax(b[3], 0.0f);
            bool inside = true;
            for (int i = 0; i < 6; ++i)
            {
                const float dist = sdist(planes24 + 4 * i, b[0], b[1], b[2]);
                if (dist < -(r + 1e-5f)) return OUTSIDE;
                if (dist < (r + 1e-5f)) inside = false;
            }
            if (inside) return INSIDE;
            inside = true;
            for (int i = 0; i < 6; ++i)
            {
                const float* p = planes24 + 4 * i;
                const float px = (p[0] >= 0.0f) ? b[7] : b[4], py = (p[1] >= 0.0f) ? b[8] : b[5], pz = (p[2] >= 0.0f) ? b[9] : b[6];
                if (sdist(p, px, py, pz) < -1e-5f) return OUTSIDE;
                const float nx = (p[0] >= 0.0f) ? b[4] : b[7], ny = (p[1] >= 0.0f) ? b[5] : b[8], nz = (p[2] >= 0.0f) ? b[6] : b[9];
                if (sdist(p, nx, ny, nz) < 1e-5f) inside = false;
            }
            return inside ? INSIDE : INTERSECTING;
        }

        SC_HD bool light_affects_object(const float* rec, const float* box6, int mode)
        {
            if (mode == MODE_NONE) return true;
            if (mode == MODE_SPHERE_AABB)
            {
                const float* s = rec + REC_CULL_SPHERE;
                const float radius = gmax(s[3], 0.0f);
                float d2 = 0.0f; // glm::dot(d, d) with d = centre - clamp(centre, min, max)
                float dd[3];
                for (int k = 0; k < 3; ++k) dd[k] = s[k] - gmin(gmax(s[k], box6[k]), box6[3 + k]);
                d2 = dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2];
                return d2 <= radius * radius;
            }
            const float* mn = rec + REC_CULL_AABB_MIN;
            const float* mx = rec + REC_CULL_AABB_MAX;
            for (int k = 0; k < 3; ++k)
                if (mx[k] < box6[k] || mn[k] > box6[3 + k]) return false;
            return true;
        }

        // visible: indices into the record array, visited in order; entries >= n_lights are skipped (:604-606)
        SC_HD uint32_t collect_lights(const float* box6, const uint32_t* visible, uint32_t n_visible, const float* records, uint32_t n_lights, int mode,
                                      uint32_t* out_idx, float* out_d2)
        {
            uint32_t count = 0;
            const float cx = 0.5f * (box6[0] + box6[3]), cy = 0.5f * (box6[1] + box6[4]), cz = 0.5f * (box6[2] + box6[5]);
            for (uint32_t k = 0; k < LIGHT_SELECTION_CAPACITY; ++k) { out_idx[k] = 0u; out_d2[k] = 0.0f; }
            for (uint32_t v = 0; v < n_visible; ++v)
            {
                const uint32_t li = visible[v];
                if (li >= n_lights) continue;
                const float* rec = records + (size_t)li * LIGHT_RECORD_FLOATS;
                if (!light_affects_object(rec, box6, mode)) continue;
                const float dx = rec[REC_POSITION] - cx, dy = rec[REC_POSITION + 1] - cy, dz = rec[REC_POSITION + 2] - cz;
                const float d2 = dx * dx + dy * dy + dz * dz;
                if (count < LIGHT_SELECTION_CAPACITY) { out_idx[count] = li; out_d2[count] = d2; ++count; continue; }
                uint32_t farthest = 0;
                float far_d2 = out_d2[0];
                for (uint32_t i = 1; i < LIGHT_SELECTION_CAPACITY; ++i)
                    if (out_d2[i] > far_d2) { farthest = i; far_d2 = out_d2[i]; }
                if (d2 < far_d2) { out_idx[farthest] = li; out_d2[farthest] = d2; }
            }
            return count;
        }
    }
}
