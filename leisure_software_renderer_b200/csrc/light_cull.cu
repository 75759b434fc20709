// light_cull.cu -- K4: Forward+ tile light-list builder.
//
// Replaces cull_lights_tiled (lighting/jolt_light_culling.hpp:135-187): per 2-D screen tile, build the
// 6-plane cell from 8 unprojected NDC corners (make_screen_tile_cell, :95-133) and keep, in ASCENDING
// light index, every frustum-visible light whose bounding sphere is not Outside the cell and -- if the
// sphere only intersects -- whose AABB p-vertex is not Outside either (classify_vs_cell,
// geometry/jolt_culling.hpp:239-257; sphere :129-147; AABB :152-181; eps 1e-5 :118-122).  Bounds come from
// the CullingLightGPU record's cull_sphere / cull_aabb_min / cull_aabb_max (lighting/light_types.hpp:141-167).
//
// Two levels: K4a filters the lights once per 12x12-tile macro cell (exact camera-frustum pre-filter + a
// conservative macro-cell test), K4b runs the exact per-tile test over the macro cell's candidates only
// (about 6x fewer sphere/AABB-vs-planes tests than tiles x lights).  Both compact with warp ballots + a
// per-warp prefix, so list order is the ascending light order of the serial reference loop.
// Output: counts[T] (uncapped) and indices[T * max_per_tile] (first max_per_tile survivors).
//
// Every expression decides list membership => exact helpers only (and the TU is built with --fmad=false).
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        constexpr int CULL_THREADS = 128;   // per-tile kernel: 4 tiles per CTA, one warp each
        constexpr int MACRO_THREADS = 256;  // per-macro-cell kernel walks every light

        struct Planes6 { float4 p[6]; };

        __device__ __forceinline__ float plane_dist(const float4& pl, float x, float y, float z)
        {
            return xadd(xadd(xadd(xmul(pl.x, x), xmul(pl.y, y)), xmul(pl.z, z)), pl.w); // dot(n, p) + d
        }

        // 0 = Outside, 1 = Intersecting, 2 = Inside
        __device__ __forceinline__ int classify(const float4* __restrict__ planes, const float4 sphere, const float4 bmin, const float4 bmax)
        {
            // The result is a pure function of the six per-plane distances (Outside if any fails, Inside if all clear), so
            // the planes may be visited in any order: the four side planes (2..5) reject far more lights than near / far.
            const float r = fmaxf(sphere.w, 0.0f);
            const float r_eps = xadd(r, 1e-5f);
            bool inside = true;
#pragma unroll
            for (int k = 0; k < 6; ++k)
            {
                const int i = (k + 2) % 6;
                const float d = plane_dist(planes[i], sphere.x, sphere.y, sphere.z);
                if (d < -r_eps) return 0;
                if (d < r_eps) inside = false;
            }
            if (inside) return 2;
            inside = true;
#pragma unroll
            for (int k = 0; k < 6; ++k)
            {
                const int i = (k + 2) % 6;
                const float4 pl = planes[i];
                const float d = plane_dist(pl, (pl.x >= 0.0f) ? bmax.x : bmin.x, (pl.y >= 0.0f) ? bmax.y : bmin.y, (pl.z >= 0.0f) ? bmax.z : bmin.z);
                if (d < -1e-5f) return 0;
                const float dn = plane_dist(pl, (pl.x >= 0.0f) ? bmin.x : bmax.x, (pl.y >= 0.0f) ? bmin.y : bmax.y, (pl.z >= 0.0f) ? bmin.z : bmax.z);
                if (dn < 1e-5f) inside = false;
            }
            return inside ? 2 : 1;
        }

        struct CullParams
        {
            float inv_vp[16];
            uint32_t vw, vh, ts, max_per_tile, tiles_x, tiles_y, n_lights;
            uint32_t macro_x, macro_y; // macro cells (MACRO x MACRO tiles) per row / column
            int own_first, own_count, own_stride; // sort-first partition: tile rows whose lists are needed (count 0 = all)
        };

        __device__ __forceinline__ bool cull_row_owned(const CullParams& cp, uint32_t ty)
        {
            return cp.own_count <= 0 || ((int)ty >= cp.own_first && (((int)ty - cp.own_first) % cp.own_stride) < cp.own_count);
        }

        // Tiles per macro-cell edge.  Measured on the C2 bench (profiles/r2_tile_kernel_specialisation.md): every macro cell walks ALL lights,
        // and that work shares the SMs with the previous frame's tile kernel -- 4: 8973 frames/s, 6: 9814, 8: 10 200, 12: 10 365, 16: 10 347.
#ifndef SHSB_MACRO
#define SHSB_MACRO 12
#endif
        constexpr uint32_t MACRO = SHSB_MACRO; // tiles per macro-cell edge

        // make_screen_tile_cell, jolt_light_culling.hpp:95-133, evaluated by threads 0..7 (corners) and 0..5 (planes)
        // of the calling CTA; both stages are followed by a __syncthreads() in the caller.
        __device__ __forceinline__ void cell_corner(const CullParams& cp, uint32_t tx, uint32_t ty, int c, float out[3], float near_ndc = -1.0f, float far_ndc = 1.0f)
        {
            // corner order nbl nbr ntl ntr fbl fbr ftl ftr (:103-117)
            const float fw = (float)cp.vw, fh = (float)cp.vh;
            const float x0 = xsub(xmul(xdiv((float)(tx * cp.ts), fw), 2.0f), 1.0f);
            const float x1 = xsub(xmul(xdiv((float)min((tx + 1u) * cp.ts, cp.vw), fw), 2.0f), 1.0f);
            const float y_top = xsub(1.0f, xmul(xdiv((float)(ty * cp.ts), fh), 2.0f));
            const float y_bottom = xsub(1.0f, xmul(xdiv((float)min((ty + 1u) * cp.ts, cp.vh), fh), 2.0f));
            const float x = (c & 1) ? x1 : x0;
            const float y = (c & 2) ? y_top : y_bottom;
            const float z = (c & 4) ? far_ndc : near_ndc; // tile_near_ndc / tile_far_ndc, :101-102
            const float4 q = xmat4_mul(cp.inv_vp, x, y, z, 1.0f);
            out[0] = xdiv(q.x, q.w);
            out[1] = xdiv(q.y, q.w);
            out[2] = xdiv(q.z, q.w);
        }

        __device__ __forceinline__ float4 cell_plane(const float (*corner)[3], int i)
        {
            enum { NBL = 0, NBR = 1, NTL = 2, NTR = 3, FBL = 4, FBR = 5, FTL = 6, FTR = 7 };
            // plane vertex triples, jolt_light_culling.hpp:125-130: near far left right bottom top
            const int tri[6][3] = {{NBL, NBR, NTR}, {FBR, FBL, FTL}, {NBL, NTL, FTL}, {NBR, FBR, FTR}, {NBL, FBL, FBR}, {NTL, NTR, FTR}};
            const float* A = corner[tri[i][0]];
            const float* B = corner[tri[i][1]];
            const float* Cc = corner[tri[i][2]];
            // inside = (nbl + ntr + fbl + ftr) * 0.25
            F3 in;
            in.x = xmul(xadd(xadd(xadd(corner[NBL][0], corner[NTR][0]), corner[FBL][0]), corner[FTR][0]), 0.25f);
            in.y = xmul(xadd(xadd(xadd(corner[NBL][1], corner[NTR][1]), corner[FBL][1]), corner[FTR][1]), 0.25f);
            in.z = xmul(xadd(xadd(xadd(corner[NBL][2], corner[NTR][2]), corner[FBL][2]), corner[FTR][2]), 0.25f);
            // make_oriented_plane_from_points, :53-68
            const F3 e1{xsub(B[0], A[0]), xsub(B[1], A[1]), xsub(B[2], A[2])};
            const F3 e2{xsub(Cc[0], A[0]), xsub(Cc[1], A[1]), xsub(Cc[2], A[2])};
            F3 nrm{xsub(xmul(e1.y, e2.z), xmul(e2.y, e1.z)), xsub(xmul(e1.z, e2.x), xmul(e2.z, e1.x)), xsub(xmul(e1.x, e2.y), xmul(e2.x, e1.y))};
            nrm = xnormalize3(nrm);
            float d = -xdot3(nrm, F3{A[0], A[1], A[2]});
            if (xadd(xdot3(nrm, in), d) < 0.0f) { nrm.x = -nrm.x; nrm.y = -nrm.y; nrm.z = -nrm.z; d = -d; }
            return make_float4(nrm.x, nrm.y, nrm.z, d);
        }

        // K4a -- one CTA per macro cell (8x8 tiles): the exact camera-frustum pre-filter of
        // jolt_light_culling.hpp:152-161 plus a CONSERVATIVE macro-cell test, producing an ascending candidate list
        // for the tiles of the cell.  Conservative means: every light the exact per-tile test keeps is a candidate.
        // Each tile plane family (left / right / bottom / top) is a pencil of planes through the eye; over the
        // few degrees a macro cell spans, a sphere's signed distance to the planes of one family is monotone or has
        // a positive interior maximum, so a light can pass a tile's plane only if it passes the family's first or
        // last plane inside the macro cell.  Both are built exactly like tile planes (first and last tile of the
        // cell) and tested with a slack of 1 cm + 0.1 % of the radius + a NOISE term.
        //
        // The noise term: the reference builds every tile's planes from the tile's OWN unprojected corners, and the
        // near corners of a tile are sub-millimetre apart at coordinates of tens to hundreds of metres, so a tile's
        // plane is tilted about its corner ray by (corner rounding error) / (near-edge length): 1e-3 rad on the 1080p
        // frame, several 1e-2 rad on the 8K frame with its far camera -- the exact per-tile test reproduces that tilt
        // bit for bit (same operations), the pre-filter must allow for it.  It is MEASURED per macro cell: the near
        // (far) planes of all tiles of a cell are one geometric plane, the left / right planes of a tile column and
        // the bottom / top planes of a tile row likewise, so the spread of their computed normals is the tilt.  A tilt
        // dn moves a plane by dn * (distance of the point from the tilt axis); the slack adds 4 x that bound.
        __global__ void __launch_bounds__(MACRO_THREADS) macro_cull_kernel(const DevLightRec* __restrict__ lights, const CullParams cp, const Planes6 frustum,
                                                                          uint32_t* __restrict__ macro_counts, uint32_t* __restrict__ macro_lists,
                                                                          float4* __restrict__ tile_planes)
        {
            __shared__ float s_corner[2][8][3];
            __shared__ float4 s_plane[2][6];
            __shared__ float4 s_tile_n[MACRO * MACRO][6];
            __shared__ unsigned s_dev[6];
            __shared__ uint32_t s_warp_count[MACRO_THREADS / 32];
            const uint32_t mc = blockIdx.x;
            const uint32_t mx = mc % cp.macro_x, my = mc / cp.macro_x;
            const uint32_t tx0 = mx * MACRO, ty0 = my * MACRO;
            const uint32_t tx1 = min(tx0 + MACRO, cp.tiles_x) - 1u, ty1 = min(ty0 + MACRO, cp.tiles_y) - 1u;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            {
                bool any_owned = false; // CTA-uniform
                for (uint32_t ty = ty0; ty <= ty1; ++ty) any_owned = any_owned || cull_row_owned(cp, ty);
                if (!any_owned) { if (threadIdx.x == 0) macro_counts[mc] = 0; return; }
            }
            if (threadIdx.x < 16)
            {
                const int which = threadIdx.x >> 3;
                cell_corner(cp, which ? tx1 : tx0, which ? ty1 : ty0, threadIdx.x & 7, s_corner[which][threadIdx.x & 7]);
            }
            if (threadIdx.x < 6) s_dev[threadIdx.x] = 0u;
            // every tile of the cell builds its own six planes, exactly as the per-tile kernel will
            const uint32_t t_col = threadIdx.x % MACRO, t_row = threadIdx.x / MACRO;
            const bool t_valid = threadIdx.x < MACRO * MACRO && tx0 + t_col <= tx1 && ty0 + t_row <= ty1;
            if (t_valid)
            {
                float c8[8][3];
#pragma unroll
                for (int c = 0; c < 8; ++c) cell_corner(cp, tx0 + t_col, ty0 + t_row, c, c8[c]);
                // ... and the per-tile kernel reads them back instead of building them a second time (same functions, same bits)
                float4* out = tile_planes + (size_t)((ty0 + t_row) * cp.tiles_x + (tx0 + t_col)) * 6;
#pragma unroll
                for (int i = 0; i < 6; ++i) { const float4 pl = cell_plane(c8, i); s_tile_n[threadIdx.x][i] = pl; out[i] = pl; }
            }
            __syncthreads();
            if (threadIdx.x < 12)
            {
                const int which = threadIdx.x / 6;
                s_plane[which][threadIdx.x % 6] = cell_plane(s_corner[which], threadIdx.x % 6);
            }
            if (t_valid)
            {
                // reference copy of the same geometric plane: near / far -> tile (0, 0); left / right -> same column, row 0;
                // bottom / top -> same row, column 0
#pragma unroll
                for (int i = 0; i < 6; ++i)
                {
                    const uint32_t ref = (i < 2) ? 0u : ((i < 4) ? t_col : t_row * MACRO);
                    const float4 a = s_tile_n[threadIdx.x][i], b = s_tile_n[ref][i];
                    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
                    const float dn = sqrtf(dx * dx + dy * dy + dz * dz);
                    if (dn == dn) atomicMax(&s_dev[i], __float_as_uint(dn)); // non-negative floats order like their bit patterns
                }
            }
            __syncthreads();
            float4 pa[6], pb[6];
            float dev[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { pa[i] = s_plane[0][i]; pb[i] = s_plane[1][i]; }
            {
                // the tilt of the near plane (its three corners are the closest together) bounds the side planes' where a
                // cell has a single row / column of tiles and nothing to compare them with
                const float dn_near = __uint_as_float(s_dev[0]);
#pragma unroll
                for (int i = 0; i < 6; ++i) dev[i] = fmaxf(__uint_as_float(s_dev[i]), (i >= 2) ? dn_near : 0.0f) + 1e-7f;
            }
            // tilt axes: the corner rays of the first and of the last tile of the cell; points are measured from the near corner
            const float ax0 = s_corner[0][0][0], ay0 = s_corner[0][0][1], az0 = s_corner[0][0][2];
            float ux = s_corner[0][4][0] - ax0, uy = s_corner[0][4][1] - ay0, uz = s_corner[0][4][2] - az0;
            float vx = s_corner[1][7][0] - s_corner[1][3][0], vy = s_corner[1][7][1] - s_corner[1][3][1], vz = s_corner[1][7][2] - s_corner[1][3][2];
            {
                const float ul = rsqrtf(ux * ux + uy * uy + uz * uz), vl = rsqrtf(vx * vx + vy * vy + vz * vz);
                ux *= ul; uy *= ul; uz *= ul; vx *= vl; vy *= vl; vz *= vl;
            }
            const float cxx = uy * vz - uz * vy, cxy = uz * vx - ux * vz, cxz = ux * vy - uy * vx;
            const float sin_span = fminf(1.0f, sqrtf(cxx * cxx + cxy * cxy + cxz * cxz) * 1.25f + 1e-4f); // angle the cell's corner rays span
            const float fx0 = s_corner[0][4][0], fy0 = s_corner[0][4][1], fz0 = s_corner[0][4][2];            // a far corner (far plane's anchor)

            uint32_t total = 0;
            uint32_t* list = macro_lists + (size_t)mc * cp.n_lights;
            for (uint32_t base = 0; base < cp.n_lights; base += MACRO_THREADS)
            {
                const uint32_t li = base + threadIdx.x;
                bool keep = false;
                if (li < cp.n_lights)
                {
                    const float4 sp = __ldg(reinterpret_cast<const float4*>(lights[li].cull_sphere));
                    const float4 mn = __ldg(reinterpret_cast<const float4*>(lights[li].cull_aabb_min));
                    const float4 mx4 = __ldg(reinterpret_cast<const float4*>(lights[li].cull_aabb_max));
                    keep = classify(frustum.p, sp, mn, mx4) != 0; // exact frustum_visible[li]
                    if (keep)
                    {
                        const float r = fmaxf(sp.w, 0.0f);
                        const float wx = sp.x - ax0, wy = sp.y - ay0, wz = sp.z - az0;
                        const float wl = sqrtf(wx * wx + wy * wy + wz * wz);
                        const float along = wx * ux + wy * uy + wz * uz;
                        // distance from the tilt axis of ANY tile of the cell: from the first corner ray, plus what the rays diverge by
                        const float rho = sqrtf(fmaxf(wl * wl - along * along, 0.0f)) + wl * sin_span + 0.5f;
                        const float gx = sp.x - fx0, gy = sp.y - fy0, gz = sp.z - fz0;
                        const float far_l = sqrtf(gx * gx + gy * gy + gz * gz) + 0.5f;
                        const float base_slack = r * 1.001f + 1e-2f;
#pragma unroll
                        for (int i = 0; i < 6; ++i)
                        {
                            const float arm = (i == 0) ? (wl + 0.5f) : ((i == 1) ? far_l : rho);
                            const float slack = -(base_slack + 4.0f * dev[i] * arm);
                            const float da = plane_dist(pa[i], sp.x, sp.y, sp.z), db = plane_dist(pb[i], sp.x, sp.y, sp.z);
                            if (da < slack && db < slack) keep = false;
                        }
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                __syncthreads();
                if (lane == 0) s_warp_count[warp] = (uint32_t)__popc(m);
                __syncthreads();
                uint32_t before = 0, chunk_total = 0;
#pragma unroll
                for (int w = 0; w < MACRO_THREADS / 32; ++w)
                {
                    const uint32_t c = s_warp_count[w];
                    if (w < warp) before += c;
                    chunk_total += c;
                }
                if (keep) list[total + before + (uint32_t)__popc(m & ((1u << lane) - 1u))] = li;
                total += chunk_total;
            }
            if (threadIdx.x == 0) macro_counts[mc] = total;
        }

        // K4b -- one WARP per tile: the exact test of cull_lights_tiled over the macro cell's candidates.  The tile's planes come
        // from K4a (which builds every tile's planes anyway, to measure their tilt); the warp walks the candidates 32 at a time and
        // compacts with one ballot -- no CTA barrier, so a tile holds 32 threads' worth of registers for its lifetime
        // (this kernel shares the SMs with the previous frame's tile kernel).
        __global__ void __launch_bounds__(CULL_THREADS) tile_cull_kernel(const DevLightRec* __restrict__ lights, const CullParams cp,
                                                                         const uint32_t* __restrict__ macro_counts, const uint32_t* __restrict__ macro_lists,
                                                                         const float4* __restrict__ tile_planes, uint32_t* __restrict__ counts, uint32_t* __restrict__ indices)
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const uint32_t tile = blockIdx.x * (CULL_THREADS / 32) + (uint32_t)warp;
            if (tile >= cp.tiles_x * cp.tiles_y) return; // warp-uniform
            const uint32_t tx = tile % cp.tiles_x, ty = tile / cp.tiles_x;
            if (!cull_row_owned(cp, ty)) { if (lane == 0) counts[tile] = 0; return; } // another rank's row: nobody reads this list
            // the tile's six planes (make_screen_tile_cell) as the macro-cell kernel built and kept them
            float4 planes[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) planes[i] = __ldg(tile_planes + (size_t)tile * 6 + i);

            const uint32_t mc = (ty / MACRO) * cp.macro_x + (tx / MACRO);
            const uint32_t n_cand = macro_counts[mc];
            const uint32_t* cand = macro_lists + (size_t)mc * cp.n_lights;
            uint32_t total = 0;
            for (uint32_t base = 0; base < n_cand; base += 32u)
            {
                const uint32_t ci = base + (uint32_t)lane;
                bool keep = false;
                uint32_t li = 0;
                if (ci < n_cand)
                {
                    li = cand[ci];
                    const float4 sp = __ldg(reinterpret_cast<const float4*>(lights[li].cull_sphere));
                    const float4 mn = __ldg(reinterpret_cast<const float4*>(lights[li].cull_aabb_min));
                    const float4 mx = __ldg(reinterpret_cast<const float4*>(lights[li].cull_aabb_max));
                    keep = classify(planes, sp, mn, mx) != 0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep)
                {
                    const uint32_t pos = total + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    if (pos < cp.max_per_tile) indices[(size_t)tile * cp.max_per_tile + pos] = li;
                }
                total += (uint32_t)__popc(m);
            }
            if (lane == 0) counts[tile] = total;
        }
    }

    namespace
    {
        // ---------------------------------------------------------------- depth-range and clustered modes
        // cull_lights_tiled_depth01_range / _view_depth_range / cull_lights_clustered (jolt_light_culling.hpp:196-412) share
        // make_screen_tile_cell with a per-cell near / far NDC.  A thin depth range makes the side planes of a cell
        // numerically unrelated to the tile's geometric planes (and a zero range makes them NaN, which the reference's
        // comparisons then treat as "not outside"), so no conservative pre-filter is valid here: every cell runs the
        // exact test over the camera-frustum-visible lights.

        // ndc_from_depth01_lh_no, :79-83
        __device__ __forceinline__ float ndc_from_depth01(float depth01) { return xsub(xmul(sclamp(depth01, 0.0f, 1.0f), 2.0f), 1.0f); }

        // ndc_from_view_depth_lh_no, :85-93
        __device__ __forceinline__ float ndc_from_view_depth(float view_depth, float z_near, float z_far)
        {
            const float n = gmax(z_near, 1e-4f);
            const float f = gmax(z_far, xadd(n, 1e-3f));
            const float z = sclamp(view_depth, n, f);
            const float denom = gmax(xsub(f, n), 1e-6f);
            return xsub(xdiv(xadd(f, n), denom), xdiv(xmul(xmul(2.0f, f), n), xmul(denom, z)));
        }

        // frustum_visible[] of :150-161 as an ascending list: vis[0] = count, vis[1..] = light indices.  One CTA.
        __global__ void __launch_bounds__(MACRO_THREADS) frustum_list_kernel(const DevLightRec* __restrict__ lights, uint32_t n_lights, const Planes6 frustum,
                                                                            uint32_t* __restrict__ vis)
        {
            __shared__ uint32_t s_warp_count[MACRO_THREADS / 32];
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            uint32_t total = 0;
            for (uint32_t base = 0; base < n_lights; base += MACRO_THREADS)
            {
                const uint32_t li = base + threadIdx.x;
                bool keep = false;
                if (li < n_lights)
                {
                    const float4 sp = __ldg(reinterpret_cast<const float4*>(lights[li].cull_sphere));
                    const float4 mn = __ldg(reinterpret_cast<const float4*>(lights[li].cull_aabb_min));
                    const float4 mx4 = __ldg(reinterpret_cast<const float4*>(lights[li].cull_aabb_max));
                    keep = classify(frustum.p, sp, mn, mx4) != 0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                __syncthreads();
                if (lane == 0) s_warp_count[warp] = (uint32_t)__popc(m);
                __syncthreads();
                uint32_t before = 0, chunk_total = 0;
#pragma unroll
                for (int w = 0; w < MACRO_THREADS / 32; ++w)
                {
                    const uint32_t c = s_warp_count[w];
                    if (w < warp) before += c;
                    chunk_total += c;
                }
                if (keep) vis[1 + total + before + (uint32_t)__popc(m & ((1u << lane) - 1u))] = li;
                total += chunk_total;
            }
            if (threadIdx.x == 0) vis[0] = total;
        }

        // One WARP per cell (tile, or tile x depth slice): exact classify_vs_cell over the frustum-visible lights,
        // ascending, ballot-compacted.  Cell index = cz * tiles + ty * tiles_x + tx (:394-396).
        __global__ void __launch_bounds__(CULL_THREADS) cell_cull_kernel(const DevLightRec* __restrict__ lights, const CullParams cp, int mode, uint32_t n_slices,
                                                                         const float* __restrict__ range_min, const float* __restrict__ range_max,
                                                                         const float2* __restrict__ slice_ndc, float z_near, float z_far,
                                                                         const uint32_t* __restrict__ vis, uint32_t* __restrict__ counts, uint32_t* __restrict__ indices)
        {
            __shared__ float s_corner[CULL_THREADS / 32][8][3];
            __shared__ float4 s_plane[CULL_THREADS / 32][6];
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const uint32_t tiles = cp.tiles_x * cp.tiles_y;
            const uint32_t cell = blockIdx.x * (CULL_THREADS / 32) + (uint32_t)warp;
            if (cell >= tiles * n_slices) return; // warp-uniform
            const uint32_t tile = cell % tiles, cz = cell / tiles;
            const uint32_t tx = tile % cp.tiles_x, ty = tile / cp.tiles_x;
            float near_ndc = -1.0f, far_ndc = 1.0f;
            if (mode == 1) { near_ndc = ndc_from_depth01(__ldg(range_min + tile)); far_ndc = ndc_from_depth01(__ldg(range_max + tile)); }                                  // :232-235
            else if (mode == 2) { near_ndc = ndc_from_view_depth(__ldg(range_min + tile), z_near, z_far); far_ndc = ndc_from_view_depth(__ldg(range_max + tile), z_near, z_far); } // :300-303
            else if (mode == 3) { const float2 sl = __ldg(slice_ndc + cz); near_ndc = sl.x; far_ndc = sl.y; }                                                                 // :375-379
            if (lane < 8) cell_corner(cp, tx, ty, lane, s_corner[warp][lane], near_ndc, far_ndc);
            __syncwarp();
            if (lane < 6) s_plane[warp][lane] = cell_plane(s_corner[warp], lane);
            __syncwarp();
            float4 planes[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) planes[i] = s_plane[warp][i];
            const uint32_t n_cand = vis[0];
            uint32_t total = 0;
            for (uint32_t base = 0; base < n_cand; base += 32u)
            {
                const uint32_t ci = base + (uint32_t)lane;
                bool keep = false;
                uint32_t li = 0;
                if (ci < n_cand)
                {
                    li = vis[1 + ci];
                    const float4 sp = __ldg(reinterpret_cast<const float4*>(lights[li].cull_sphere));
                    const float4 mn = __ldg(reinterpret_cast<const float4*>(lights[li].cull_aabb_min));
                    const float4 mx = __ldg(reinterpret_cast<const float4*>(lights[li].cull_aabb_max));
                    keep = classify(planes, sp, mn, mx) != 0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep)
                {
                    const uint32_t pos = total + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    if (pos < cp.max_per_tile) indices[(size_t)cell * cp.max_per_tile + pos] = li;
                }
                total += (uint32_t)__popc(m);
            }
            if (lane == 0) counts[cell] = total;
        }

        // Per-tile depth range of the z-buffer, as linear view depth (the inverse of rasterizer.hpp:352-354:
        // view_z = zn + z01 * (zf - zn)); pixels still at the clear value (>= 1.0) are skipped and a tile without any
        // geometry gets [zn, zf] like build_tile_view_depth_range_from_scene (light_culling_runtime.hpp:254-261).  The
        // software analogue of shaders/vulkan/fp_stress_depth_reduce.comp.  Light tiles are top-anchored
        // (jolt_light_culling.hpp:105-107) while the framebuffer is y-up: tile row ty covers rows H-1-py in [ty*ts, (ty+1)*ts).
        // ndc01 != 0: the buffer holds the hardware's zero-to-one projective depth (uploaded by the caller) and the result is the compute shader's own:
        // view_z = near * far / max(far - clamp(d) * (far - near), 1e-5) with near = max(zn, 0.001), far = max(zf, near + 0.01)
        // (fp_stress_depth_reduce.comp:31-38), min / max starting at 1e30 / 0, (0, 0) for a tile without geometry (:56-57, :73-77).  Every
        // step of that mapping is monotone in d under IEEE rounding, so it commutes with the reduction here as well.
        __global__ void __launch_bounds__(CULL_THREADS) tile_depth_range_kernel(const float* __restrict__ depth, int W, int H, uint32_t ts, uint32_t tiles_x, uint32_t tiles_y,
                                                                                float zn, float zf, int ndc01, float* __restrict__ out_min, float* __restrict__ out_max)
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const uint32_t tile = blockIdx.x * (CULL_THREADS / 32) + (uint32_t)warp;
            if (tile >= tiles_x * tiles_y) return;
            const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
            const int x0 = (int)(tx * ts), x1 = min((int)((tx + 1u) * ts), W);
            const int r0 = (int)(ty * ts), r1 = min((int)((ty + 1u) * ts), H); // rows counted from the top
            const int tw = x1 - x0, n = tw * (r1 - r0);
            float lo = INFINITY, hi = -INFINITY; // any value below 1 takes part, negative ones included
            for (int i = lane; i < n; i += 32)
            {
                const int px = x0 + i % tw, py = H - 1 - (r0 + i / tw);
                const float d = __ldg(depth + (size_t)py * W + px);
                if (d >= 1.0f) continue;
                lo = fminf(lo, d);
                hi = fmaxf(hi, d);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if (lane == 0)
            {
                const bool any = hi > -INFINITY;
                if (ndc01)
                {
                    const float near_z = gmax(zn, 0.001f), far_z = gmax(zf, xadd(near_z, 0.01f));
                    const float span = xsub(far_z, near_z), nf = xmul(near_z, far_z);
                    const float vlo = xdiv(nf, gmax(xsub(far_z, xmul(gclamp(lo, 0.0f, 1.0f), span)), 1e-5f));
                    const float vhi = xdiv(nf, gmax(xsub(far_z, xmul(gclamp(hi, 0.0f, 1.0f), span)), 1e-5f));
                    out_min[tile] = any ? gmin(1e30f, vlo) : 0.0f;
                    out_max[tile] = any ? gmax(0.0f, vhi) : 0.0f;
                    return;
                }
                const float k = xsub(zf, zn);
                out_min[tile] = any ? xadd(zn, xmul(lo, k)) : zn; // monotone in d, so min / max commute with the mapping
                out_max[tile] = any ? xadd(zn, xmul(hi, k)) : zf;
            }
        }
    }

    size_t light_cull_scratch_words(uint32_t n_lights, uint32_t vw, uint32_t vh, uint32_t ts)
    {
        const uint32_t tiles_x = (vw + ts - 1) / ts, tiles_y = (vh + ts - 1) / ts;
        const size_t n_macro = (size_t)((tiles_x + MACRO - 1) / MACRO) * ((tiles_y + MACRO - 1) / MACRO);
        return n_macro * ((size_t)n_lights + 1) + 4 + (size_t)tiles_x * tiles_y * 24; // + the tiles' planes (6 x float4 each, 16-byte aligned)
    }

    void launch_light_cull(const DevLightRec* lights, uint32_t n_lights, const float* frustum_planes24, const float* inv_view_proj,
                           uint32_t vw, uint32_t vh, uint32_t ts, uint32_t max_per_tile,
                           uint32_t* scratch, uint32_t* counts, uint32_t* indices, cudaStream_t s, uint64_t* launches,
                           int own_first, int own_count, int own_stride)
    {
        CullParams cp;
        for (int i = 0; i < 16; ++i) cp.inv_vp[i] = inv_view_proj[i];
        cp.vw = vw; cp.vh = vh; cp.ts = ts; cp.max_per_tile = max_per_tile;
        cp.tiles_x = (vw + ts - 1) / ts;
        cp.tiles_y = (vh + ts - 1) / ts;
        cp.n_lights = n_lights;
        cp.own_first = own_first; cp.own_count = own_count; cp.own_stride = own_stride > 0 ? own_stride : 1;
        cp.macro_x = (cp.tiles_x + MACRO - 1) / MACRO;
        cp.macro_y = (cp.tiles_y + MACRO - 1) / MACRO;
        Planes6 fr;
        for (int i = 0; i < 6; ++i) fr.p[i] = make_float4(frustum_planes24[i * 4], frustum_planes24[i * 4 + 1], frustum_planes24[i * 4 + 2], frustum_planes24[i * 4 + 3]);
        const uint32_t n_macro = cp.macro_x * cp.macro_y;
        uint32_t* macro_counts = scratch;
        uint32_t* macro_lists = scratch + n_macro;
        const size_t planes_at = ((size_t)n_macro * ((size_t)n_lights + 1) + 3) & ~(size_t)3; // in words, from a 256-byte aligned allocation
        float4* tile_planes = reinterpret_cast<float4*>(scratch + planes_at);
        macro_cull_kernel<<<n_macro, MACRO_THREADS, 0, s>>>(lights, cp, fr, macro_counts, macro_lists, tile_planes);
        tile_cull_kernel<<<(cp.tiles_x * cp.tiles_y + CULL_THREADS / 32 - 1) / (CULL_THREADS / 32), CULL_THREADS, 0, s>>>(lights, cp, macro_counts, macro_lists, tile_planes, counts,
                                                                                                                          indices);
        *launches += 2;
    }

    void launch_light_cull_cells(const DevLightRec* lights, uint32_t n_lights, const float* frustum_planes24, const float* inv_view_proj,
                                 uint32_t vw, uint32_t vh, uint32_t ts, uint32_t max_per_bin, int mode, uint32_t n_slices,
                                 const float* range_min, const float* range_max, const float2* slice_ndc, float z_near, float z_far,
                                 uint32_t* vis_scratch, uint32_t* counts, uint32_t* indices, cudaStream_t s, uint64_t* launches)
    {
        CullParams cp;
        for (int i = 0; i < 16; ++i) cp.inv_vp[i] = inv_view_proj[i];
        cp.vw = vw; cp.vh = vh; cp.ts = ts; cp.max_per_tile = max_per_bin;
        cp.tiles_x = (vw + ts - 1) / ts;
        cp.tiles_y = (vh + ts - 1) / ts;
        cp.n_lights = n_lights;
        cp.own_first = 0; cp.own_count = 0; cp.own_stride = 1;
        cp.macro_x = cp.macro_y = 0;
        Planes6 fr;
        for (int i = 0; i < 6; ++i) fr.p[i] = make_float4(frustum_planes24[i * 4], frustum_planes24[i * 4 + 1], frustum_planes24[i * 4 + 2], frustum_planes24[i * 4 + 3]);
        frustum_list_kernel<<<1, MACRO_THREADS, 0, s>>>(lights, n_lights, fr, vis_scratch);
        const uint32_t cells = cp.tiles_x * cp.tiles_y * n_slices;
        cell_cull_kernel<<<(cells + CULL_THREADS / 32 - 1) / (CULL_THREADS / 32), CULL_THREADS, 0, s>>>(lights, cp, mode, n_slices, range_min, range_max, slice_ndc, z_near, z_far,
                                                                                                      vis_scratch, counts, indices);
        *launches += 2;
    }

    void launch_tile_depth_range(const float* depth, int W, int H, uint32_t ts, float zn, float zf, int ndc01, float* out_min, float* out_max, cudaStream_t s, uint64_t* launches)
    {
        const uint32_t tiles_x = ((uint32_t)W + ts - 1) / ts, tiles_y = ((uint32_t)H + ts - 1) / ts;
        const uint32_t tiles = tiles_x * tiles_y;
        if (!tiles) return;
        tile_depth_range_kernel<<<(tiles + CULL_THREADS / 32 - 1) / (CULL_THREADS / 32), CULL_THREADS, 0, s>>>(depth, W, H, ts, tiles_x, tiles_y, zn, zf, ndc01, out_min, out_max);
        *launches += 1;
    }
}
