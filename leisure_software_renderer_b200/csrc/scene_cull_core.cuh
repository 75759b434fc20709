// scene_cull_core.cuh -- scene-level steps immediately upstream of draw submission as host/device
// inline functions; scene_cull.cu wraps them in kernels, tests/cpp/scene_cull_emul.cpp compiles the same functions with g++ and checks
// them against the pinned oracle on the CPU box (test code: the product launches kernels).
//   classify_object      classify_vs_frustum for a FastCullable + HasWorldAABB object, geometry/jolt_culling.hpp:129-181, 239-275
//                        (sphere first, AABB p-vertex / n-vertex test when the sphere intersects; tolerances 1e-5, :118-122)
//   collect_lights       collect_object_lights + light_affects_object + add_light_candidate, lighting/light_runtime.hpp:239-252,
//                        263-289, 570-616: up to 8 nearest affecting lights per object; the SLOT a light lands in depends on visit
//                        order (replace the first farthest slot when strictly nearer), so each object walks its candidates serially
//   project_object       project_aabb_bounds + the tile rectangle of build_tile_view_depth_range_from_scene,
//                        lighting/light_culling_runtime.hpp:92-168, 188-264: an object's screen-tile rectangle and view-depth range;
//                        the per-tile min / max over objects is order-independent and is taken with atomics on ordered keys
//   select_from_bins     gather_light_scene_candidates_for_aabb (lighting/light_culling_runtime.hpp:373-449) fused with
//                        collect_object_lights: the bins an object's projected AABB touches are walked in (z, y, x) order, each light
//                        is considered at its FIRST occurrence (a per-object bit set replaces the reference's std::find) and fed to
//                        the selection in that order.  The cluster slice uses logf: CUDA's double log rounded to float on the device.
// IEEE binary32, unfused, GLM's scalar order (compiled --fmad=false / -ffp-contract=off).
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

#ifndef SC_LOGF
#define SC_LOGF(x) ((float)log((double)(x)))
#endif

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

        SC_HD float std_clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); } // std::clamp
        SC_HD float std_minf(float a, float b) { return (b < a) ? b : a; }                                  // std::min
        SC_HD float std_maxf(float a, float b) { return (a < b) ? b : a; }                                  // std::max

        SC_HD uint32_t ndc_x_to_bin(float ndc_x, uint32_t bins) // :155-160
        {
            const float u = std_clampf(ndc_x * 0.5f + 0.5f, 0.0f, 0.999999f);
            const uint32_t b = (uint32_t)(u * (float)bins);
            return b < bins - 1u ? b : bins - 1u;
        }
        SC_HD uint32_t ndc_y_to_bin_top(float ndc_y, uint32_t bins) // :162-167
        {
            const float v = std_clampf(1.0f - (ndc_y * 0.5f + 0.5f), 0.0f, 0.999999f);
            const uint32_t b = (uint32_t)(v * (float)bins);
            return b < bins - 1u ? b : bins - 1u;
        }

        struct TileRect { uint32_t tx0, tx1, ty0, ty1; float min_depth, max_depth; };

        // false: no corner in front of the camera (the object contributes nothing)
        SC_HD bool project_object(const float* box6, const float* view, const float* view_proj, float z_near, float z_far, uint32_t tiles_x, uint32_t tiles_y, TileRect& r)
        {
            bool any = false;
            float min_x = 1.0f, max_x = -1.0f, min_y = 1.0f, max_y = -1.0f, min_d = z_far, max_d = z_near;
            for (int c = 0; c < 8; ++c) // aabb_corners order (:80-90): x fastest, then y, then z
            {
                const float x = box6[(c & 1) ? 3 : 0], y = box6[(c & 2) ? 4 : 1], z = box6[(c & 4) ? 5 : 2];
                const float* m = view_proj;
                const float cw = (m[3] * x + m[7] * y) + (m[11] * z + m[15] * 1.0f);
                if (cw <= 1e-5f) continue;
                const float cx = (m[0] * x + m[4] * y) + (m[8] * z + m[12] * 1.0f), cy = (m[1] * x + m[5] * y) + (m[9] * z + m[13] * 1.0f);
                const float nx = cx / cw, ny = cy / cw;
                min_x = std_minf(min_x, nx); max_x = std_maxf(max_x, nx);
                min_y = std_minf(min_y, ny); max_y = std_maxf(max_y, ny);
                const float vd = (view[2] * x + view[6] * y) + (view[10] * z + view[14] * 1.0f);
                if (vd > 1e-5f) { min_d = std_minf(min_d, vd); max_d = std_maxf(max_d, vd); }
                any = true;
            }
            if (!any) return false;
            min_x = std_clampf(min_x, -1.0f, 1.0f); max_x = std_clampf(max_x, -1.0f, 1.0f);
            min_y = std_clampf(min_y, -1.0f, 1.0f); max_y = std_clampf(max_y, -1.0f, 1.0f);
            if (min_x > max_x) { const float t = min_x; min_x = max_x; max_x = t; }
            if (min_y > max_y) { const float t = min_y; min_y = max_y; max_y = t; }
            min_d = std_clampf(min_d, z_near, z_far);
            max_d = std_clampf(max_d, z_near, z_far);
            if (min_d > max_d) { min_d = z_near; max_d = z_far; }
            r.tx0 = ndc_x_to_bin(min_x, tiles_x); r.tx1 = ndc_x_to_bin(max_x, tiles_x);
            r.ty0 = ndc_y_to_bin_top(max_y, tiles_y); r.ty1 = ndc_y_to_bin_top(min_y, tiles_y);
            r.min_depth = min_d; r.max_depth = max_d;
            return true;
        }

        // order-preserving float <-> uint32 keys (so that integer atomicMin / atomicMax give the float min / max)
        SC_HD uint32_t depth_key(float f) { uint32_t b; memcpy(&b, &f, 4); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
        SC_HD float key_depth(uint32_t k) { const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k; float f; memcpy(&f, &b, 4); return f; }

        // the closing loop of build_tile_view_depth_range_from_scene (:254-261)
        SC_HD void finish_tile(uint32_t has, float mn, float mx, float z_near, float z_far, float& out_min, float& out_max)
        {
            if (has == 0u || mn > mx) { out_min = z_near; out_max = z_far; }
            else { out_min = mn; out_max = mx; }
        }

        // ---- light selection state shared by collect_lights and select_from_bins
        struct Selection { uint32_t idx[LIGHT_SELECTION_CAPACITY]; float d2[LIGHT_SELECTION_CAPACITY]; uint32_t count; };
        SC_HD void selection_clear(Selection& s) { s.count = 0; for (uint32_t k = 0; k < LIGHT_SELECTION_CAPACITY; ++k) { s.idx[k] = 0u; s.d2[k] = 0.0f; } }
        // add_light_candidate, light_runtime.hpp:263-289: the order-dependent part (the kernels feed it in the reference's visit order)
        SC_HD void selection_insert(Selection& s, uint32_t li, float d2)
        {
            if (s.count < LIGHT_SELECTION_CAPACITY) { s.idx[s.count] = li; s.d2[s.count] = d2; ++s.count; return; }
            uint32_t farthest = 0;
            float far_d2 = s.d2[0];
            for (uint32_t i = 1; i < LIGHT_SELECTION_CAPACITY; ++i)
                if (s.d2[i] > far_d2) { farthest = i; far_d2 = s.d2[i]; }
            if (d2 < far_d2) { s.idx[farthest] = li; s.d2[farthest] = d2; }
        }
        // the order-free part: does the light reach the object, and how far is it from the box centre
        SC_HD bool light_candidate(const float* rec, const float* box6, float cx, float cy, float cz, int mode, float& d2)
        {
            if (!light_affects_object(rec, box6, mode)) return false;
            const float dx = rec[REC_POSITION] - cx, dy = rec[REC_POSITION + 1] - cy, dz = rec[REC_POSITION + 2] - cz;
            d2 = dx * dx + dy * dy + dz * dz;
            return true;
        }
        SC_HD void consider_light(Selection& s, uint32_t li, const float* rec, const float* box6, float cx, float cy, float cz, int mode)
        {
            float d2;
            if (light_candidate(rec, box6, cx, cy, cz, mode, d2)) selection_insert(s, li, d2);
        }

        SC_HD uint32_t view_depth_to_cluster_slice(float view_depth, float z_near, float z_far, uint32_t slices) // :170-186
        {
            if (slices <= 1u) return 0u;
            const float zn = std_maxf(z_near, 1e-4f);
            const float zf = std_maxf(z_far, zn + 1e-3f);
            const float d = std_clampf(view_depth, zn, zf);
            const float log_ratio = SC_LOGF(zf / zn);
            if (log_ratio <= 1e-6f) return 0u;
            const float t = std_clampf(SC_LOGF(d / zn) / log_ratio, 0.0f, 0.999999f);
            const uint32_t b = (uint32_t)(t * (float)slices);
            return b < slices - 1u ? b : slices - 1u;
        }

        struct BinGrid
        {
            uint32_t bins_x, bins_y, bins_z; // bins_z = 1 for the tiled builders
            int clustered;                   // LightCullingMode::Clustered
            float z_near, z_far;             // LightBinCullingData::z_near / z_far (already max(.., 1e-4) / max(.., zn + 1e-3), :278-279)
            uint32_t max_per_bin;            // row stride of bin_indices; a bin holds min(count, max_per_bin) entries
        };

        // The bins an object's projected AABB touches (:389-420): false = no bins / nothing in front of the camera -> every light is a candidate
        SC_HD bool object_bin_range(const float* box6, const float* view, const float* view_proj, const BinGrid& g, TileRect& r, uint32_t& tz0, uint32_t& tz1)
        {
            const bool has_bins = g.bins_x > 0u && g.bins_y > 0u && g.bins_z > 0u;
            if (!has_bins || !project_object(box6, view, view_proj, g.z_near, g.z_far, g.bins_x, g.bins_y, r)) return false;
            tz0 = 0u; tz1 = (g.bins_z > 1u ? g.bins_z : 1u) - 1u;
            if (g.clustered && g.bins_z > 1u)
            {
                tz0 = view_depth_to_cluster_slice(r.min_depth, g.z_near, g.z_far, g.bins_z);
                tz1 = view_depth_to_cluster_slice(r.max_depth, g.z_near, g.z_far, g.bins_z);
                if (tz0 > tz1) { const uint32_t t = tz0; tz0 = tz1; tz1 = t; }
            }
            return true;
        }

        // seen: ceil(n_lights / 32) words of scratch owned by this object, zeroed on entry.  Returns the number of candidates
        // (gather's list length); the selection is what collect_object_lights makes of that list.
        SC_HD uint32_t select_from_bins(const float* box6, const float* view, const float* view_proj, const BinGrid& g, const uint32_t* bin_counts, const uint32_t* bin_indices,
                                        const float* records, uint32_t n_lights, int mode, uint32_t* seen, Selection& sel)
        {
            selection_clear(sel);
            const float cx = 0.5f * (box6[0] + box6[3]), cy = 0.5f * (box6[1] + box6[4]), cz = 0.5f * (box6[2] + box6[5]);
            TileRect r;
            uint32_t tz0, tz1;
            if (!object_bin_range(box6, view, view_proj, g, r, tz0, tz1))
            {
                // fallback_light_scene_candidates: every visible light, in order
                for (uint32_t li = 0; li < n_lights; ++li) consider_light(sel, li, records + (size_t)li * LIGHT_RECORD_FLOATS, box6, cx, cy, cz, mode);
                return n_lights;
            }
            uint32_t n_candidates = 0;
            for (uint32_t tz = tz0; tz <= tz1; ++tz)
                for (uint32_t ty = r.ty0; ty <= r.ty1; ++ty)
                    for (uint32_t tx = r.tx0; tx <= r.tx1; ++tx)
                    {
                        const uint32_t bin = tz * (g.bins_x * g.bins_y) + ty * g.bins_x + tx;
                        const uint32_t n = bin_counts[bin] < g.max_per_bin ? bin_counts[bin] : g.max_per_bin;
                        for (uint32_t k = 0; k < n; ++k)
                        {
                            const uint32_t li = bin_indices[(size_t)bin * g.max_per_bin + k];
                            if (li >= n_lights) continue;
                            const uint32_t bit = 1u << (li & 31u);
                            if (seen[li >> 5] & bit) continue;
                            seen[li >> 5] |= bit;
                            ++n_candidates;
                            consider_light(sel, li, records + (size_t)li * LIGHT_RECORD_FLOATS, box6, cx, cy, cz, mode);
                        }
                    }
            return n_candidates;
        }

        // ---------------------------------------------------------------- software occlusion (geometry/culling_software.hpp:41-222)
        // The pass of run_software_occlusion_pass (:253-333) walks the frustum-visible objects front to back; each object's screen
        // rectangle is tested against the occlusion depth buffer the objects before it rasterised their occluder meshes into.  The
        // ORDER of objects is serial; the work of one object -- an AND over the rectangle's texels, a minimum-write per triangle
        // texel -- is order-free and is what the device spreads over a CTA.
        SC_HD void mul4(const float* m, float x, float y, float z, float w, float out[4]) // glm's scalar mat4 * vec4
        {
            for (int r = 0; r < 4; ++r) out[r] = (m[r] * x + m[4 + r] * y) + (m[8 + r] * z + m[12 + r] * w);
        }

        struct OccRect { int x_min, y_min, x_max, y_max; float z_near; bool valid; };

        // project_aabb_to_screen_rect, :145-199
        SC_HD OccRect occ_project_rect(const float* box6, const float* view_proj, int width, int height)
        {
            OccRect out{0, 0, -1, -1, 1.0f, false};
            if (width <= 0 || height <= 0) return out;
            float min_x = (float)width, min_y = (float)height, max_x = -1.0f, max_y = -1.0f, near_depth = 1.0f;
            bool any = false;
            for (int c = 0; c < 8; ++c)
            {
                float clip[4];
                mul4(view_proj, box6[(c & 1) ? 3 : 0], box6[(c & 2) ? 4 : 1], box6[(c & 4) ? 5 : 2], 1.0f, clip);
                if (clip[3] <= 0.001f) continue;
                const float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
                const float z01 = nz * 0.5f + 0.5f;
                if (z01 < 0.0f || z01 > 1.0f) continue;
                const float sx = (nx + 1.0f) * 0.5f * (float)width;
                const float sy = (ny + 1.0f) * 0.5f * (float)height;
                min_x = std_minf(min_x, sx); min_y = std_minf(min_y, sy);
                max_x = std_maxf(max_x, sx); max_y = std_maxf(max_y, sy);
                near_depth = std_minf(near_depth, z01);
                any = true;
            }
            if (!any) return out;
            const int fx = (int)floorf(min_x), fy = (int)floorf(min_y), cx = (int)ceilf(max_x), cy = (int)ceilf(max_y);
            out.x_min = fx > 0 ? fx : 0; out.y_min = fy > 0 ? fy : 0;
            out.x_max = cx < width - 1 ? cx : width - 1; out.y_max = cy < height - 1 ? cy : height - 1;
            out.z_near = std_clampf(near_depth, 0.0f, 1.0f);
            out.valid = out.x_min <= out.x_max && out.y_min <= out.y_max;
            return out;
        }

        // one texel of is_rect_occluded, :201-222: true = the rectangle is NOT hidden at this texel
        SC_HD bool occ_texel_shows(float z_near, float depth, float epsilon) { return z_near <= depth + epsilon; }

        // project_world_to_screen, :46-63, after vec3(model * vec4(local, 1)) of rasterize_mesh_depth_transformed :131-133
        SC_HD bool occ_project_vertex(const float* model, const float* local3, const float* view_proj, int width, int height, float xy[2], float& depth01)
        {
            float wp[4], clip[4];
            mul4(model, local3[0], local3[1], local3[2], 1.0f, wp);
            mul4(view_proj, wp[0], wp[1], wp[2], 1.0f, clip);
            if (clip[3] <= 0.001f) return false;
            const float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
            if (nz < -1.0f || nz > 1.0f) return false;
            xy[0] = (nx + 1.0f) * 0.5f * (float)width;
            xy[1] = (ny + 1.0f) * 0.5f * (float)height;
            depth01 = nz * 0.5f + 0.5f;
            return true;
        }

        SC_HD float occ_edge(const float* a, const float* b, float px, float py) { return (px - a[0]) * (b[1] - a[1]) - (py - a[1]) * (b[0] - a[0]); } // :41-44

        struct OccTri { float p0[2], p1[2], p2[2], z0, z1, z2, area; int min_x, min_y, max_x, max_y; bool valid; };

        // the per-triangle part of rasterize_depth_triangle, :65-90
        SC_HD OccTri occ_setup_triangle(const float* p0, float z0, const float* p1, float z1, const float* p2, float z2, int width, int height)
        {
            OccTri t;
            t.p0[0] = p0[0]; t.p0[1] = p0[1]; t.p1[0] = p1[0]; t.p1[1] = p1[1]; t.p2[0] = p2[0]; t.p2[1] = p2[1];
            t.z0 = z0; t.z1 = z1; t.z2 = z2;
            t.area = occ_edge(p0, p1, p2[0], p2[1]);
            t.valid = false;
            t.min_x = t.min_y = 0; t.max_x = t.max_y = -1;
            if (fabsf(t.area) <= 1e-6f) return t;
            const float min_xf = std_minf(p0[0], std_minf(p1[0], p2[0])), min_yf = std_minf(p0[1], std_minf(p1[1], p2[1]));
            const float max_xf = std_maxf(p0[0], std_maxf(p1[0], p2[0])), max_yf = std_maxf(p0[1], std_maxf(p1[1], p2[1]));
            const int fx = (int)floorf(min_xf), fy = (int)floorf(min_yf), cx = (int)ceilf(max_xf), cy = (int)ceilf(max_yf);
            t.min_x = fx > 0 ? fx : 0; t.min_y = fy > 0 ? fy : 0;
            t.max_x = cx < width - 1 ? cx : width - 1; t.max_y = cy < height - 1 ? cy : height - 1;
            t.valid = t.min_x <= t.max_x && t.min_y <= t.max_y;
            return t;
        }

        // the per-texel part, :92-114: true + depth when the texel takes part in the minimum
        SC_HD bool occ_texel_depth(const OccTri& t, int x, int y, float& depth)
        {
            const float px = (float)x + 0.5f, py = (float)y + 0.5f;
            const float w0 = occ_edge(t.p1, t.p2, px, py), w1 = occ_edge(t.p2, t.p0, px, py), w2 = occ_edge(t.p0, t.p1, px, py);
            const bool inside = (t.area > 0.0f) ? (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) : (w0 <= 0.0f && w1 <= 0.0f && w2 <= 0.0f);
            if (!inside) return false;
            const float iw0 = w0 / t.area, iw1 = w1 / t.area, iw2 = w2 / t.area;
            depth = iw0 * t.z0 + iw1 * t.z1 + iw2 * t.z2;
            return !(depth < 0.0f || depth > 1.0f);
        }

        // view_depth_of_aabb_center, :224-231 (the sort key)
        SC_HD float occ_view_depth(const float* box6, const float* view)
        {
            float v[4];
            mul4(view, 0.5f * (box6[0] + box6[3]), 0.5f * (box6[1] + box6[4]), 0.5f * (box6[2] + box6[5]), 1.0f, v);
            return v[2];
        }
    }
}
