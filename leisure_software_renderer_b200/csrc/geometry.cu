// geometry.cu -- K1: vertex transform, frustum clip, triangle setup.
//
// Replaces the serial per-triangle front half of rasterize_mesh (sw_render/rasterizer.hpp:206-328):
// VS x3 (builtin make_default_vertex_out, shader/builtin_shaders.hpp:87-103), trivially-inside test
// (:232-249), Sutherland-Hodgman clip against the 6 homogeneous planes (:69-164), fan triangulation
// (:253-258), NDC -> screen map (:260-269), signed-area cull (:271-278), clamped bbox (:280-290) and the
// 1/w pre-multiplication of attributes (:292-328) -- one CUDA thread per source triangle, all draws of a
// frame in one launch.  Set-up triangles are appended (warp-aggregated atomic) to the RasterRec/ShadeRec
// arrays; their position in memory is arbitrary because every record carries its draw-order key and the
// tile kernel resolves visibility by (depth, key), which reproduces "earliest draw wins a depth tie"
// (rasterizer.hpp:359) without ordered lists.
//
// Compiled with --fmad=false AND written with the exact helpers: every expression here decides bits of
// coverage, depth or clip topology.
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        constexpr int GEOM_THREADS = 128;
        constexpr int NATTR = 8;      // WorldPos.xyz, NormalWS.xyz, UV0.xy (the varyings the builtin programs read)
        constexpr int MAX_POLY = 10;  // a triangle clipped by 6 planes has at most 9 corners

        struct Corner
        {
            float clip[4];
            float a[NATTR];
        };

        __device__ __forceinline__ Corner run_vs(const DevItem& it, const DevMesh& mesh, uint32_t idx, const FrameConst& fc, bool varyings)
        {
            const float px = mesh.positions[(size_t)idx * 3 + 0];
            const float py = mesh.positions[(size_t)idx * 3 + 1];
            const float pz = mesh.positions[(size_t)idx * 3 + 2];
            Corner o;
            const float4 wp = xmat4_mul(it.model, px, py, pz, 1.0f);
            const float4 cl = xmat4_mul(fc.viewproj, wp.x, wp.y, wp.z, wp.w);
            o.clip[0] = cl.x; o.clip[1] = cl.y; o.clip[2] = cl.z; o.clip[3] = cl.w;
            if (varyings)
            {
                float nx = 0.0f, ny = 1.0f, nz = 0.0f, tu = 0.0f, tv = 0.0f; // read_v defaults, rasterizer.hpp:196-202
                if (idx < mesh.n_normals) { nx = mesh.normals[(size_t)idx * 3 + 0]; ny = mesh.normals[(size_t)idx * 3 + 1]; nz = mesh.normals[(size_t)idx * 3 + 2]; }
                if (idx < mesh.n_uvs) { tu = mesh.uvs[(size_t)idx * 2 + 0]; tv = mesh.uvs[(size_t)idx * 2 + 1]; }
                const float* n = it.nrm;
                F3 nn;
                nn.x = xadd(xadd(xmul(n[0], nx), xmul(n[3], ny)), xmul(n[6], nz));
                nn.y = xadd(xadd(xmul(n[1], nx), xmul(n[4], ny)), xmul(n[7], nz));
                nn.z = xadd(xadd(xmul(n[2], nx), xmul(n[5], ny)), xmul(n[8], nz));
                nn = xnormalize3(nn);
                o.a[0] = wp.x; o.a[1] = wp.y; o.a[2] = wp.z;
                o.a[3] = nn.x; o.a[4] = nn.y; o.a[5] = nn.z;
                o.a[6] = tu; o.a[7] = tv;
            }
            else
            {
#pragma unroll
                for (int i = 0; i < NATTR; ++i) o.a[i] = 0.0f;
            }
            return o;
        }

        __device__ __forceinline__ bool corner_inside(const Corner& c)
        {
            const float x = c.clip[0], y = c.clip[1], z = c.clip[2], w = c.clip[3];
            if (!(w > 0.0f)) return false;
            return (x >= -w && x <= w) && (y >= -w && y <= w) && (z >= -w && z <= w);
        }

        __device__ __forceinline__ float plane_dist(const Corner& c, int plane)
        {
            switch (plane)
            {
            case 0: return xadd(c.clip[0], c.clip[3]);
            case 1: return xsub(c.clip[3], c.clip[0]);
            case 2: return xadd(c.clip[1], c.clip[3]);
            case 3: return xsub(c.clip[3], c.clip[1]);
            case 4: return xadd(c.clip[2], c.clip[3]);
            default: return xsub(c.clip[3], c.clip[2]);
            }
        }

        __device__ __forceinline__ float xmix(float a, float b, float t, float one_minus_t) { return xadd(xmul(a, one_minus_t), xmul(b, t)); }

        __device__ Corner lerp_corner(const Corner& a, const Corner& b, float t)
        {
            Corner o;
            const float omt = xsub(1.0f, t);
#pragma unroll
            for (int i = 0; i < 4; ++i) o.clip[i] = xmix(a.clip[i], b.clip[i], t, omt);
#pragma unroll
            for (int i = 0; i < NATTR; ++i) o.a[i] = xmix(a.a[i], b.a[i], t, omt);
            return o;
        }

        // Projects one fan triangle, applies the cull / bbox rules and, if it survives, fills the two records.
        // Returns: bit0 = counted in tri_raster, bit1 = record must be emitted.
        __device__ __forceinline__ int setup_triangle(const Corner& c0, const Corner& c1, const Corner& c2, const FrameConst& fc,
                                                      uint32_t key, uint32_t item_index, RasterRec& rr, ShadeRec& sr)
        {
            const Corner* c[3] = {&c0, &c1, &c2};
            float sx[3], sy[3];
            const float fw1 = (float)(fc.W - 1), fh1 = (float)(fc.H - 1);
#pragma unroll
            for (int j = 0; j < 3; ++j)
            {
                const float w = c[j]->clip[3];
                const float nx = xdiv(c[j]->clip[0], w), ny = xdiv(c[j]->clip[1], w), nz = xdiv(c[j]->clip[2], w);
                if (!isfinite(nx) || !isfinite(ny) || !isfinite(nz)) return 0;
                sx[j] = xmul(xadd(xmul(nx, 0.5f), 0.5f), fw1);
                sy[j] = xmul(xadd(xmul(ny, 0.5f), 0.5f), fh1);
            }
            const float v0x = xsub(sx[1], sx[0]), v0y = xsub(sy[1], sy[0]);
            const float v1x = xsub(sx[2], sx[0]), v1y = xsub(sy[2], sy[0]);
            const float area2 = xsub(xmul(v0x, v1y), xmul(v0y, v1x)); // == barycentric den (v1x*v0y commutes)
            if (fabsf(area2) < 1e-10f) return 0;
            const bool ccw = area2 > 0.0f;
            const bool is_front = (ccw == (fc.front_face_ccw != 0));
            if (fc.cull_mode == 1 && !is_front) return 0;
            if (fc.cull_mode == 2 && is_front) return 0;

            const float minxf = fminf(fminf(sx[0], sx[1]), sx[2]), maxxf = fmaxf(fmaxf(sx[0], sx[1]), sx[2]);
            const float minyf = fminf(fminf(sy[0], sy[1]), sy[2]), maxyf = fmaxf(fmaxf(sy[0], sy[1]), sy[2]);
            const int minx = max(0, (int)floorf(minxf));
            const int maxx = min(fc.W - 1, (int)ceilf(maxxf));
            const int miny = max(0, (int)floorf(minyf));
            const int maxy = min(fc.H - 1, (int)ceilf(maxyf));
            if (minx > maxx || miny > maxy) return 0;
            if (fabsf(area2) < 1e-8f) return 1; // barycentric_2d returns (-1,-1,-1): counted, covers nothing (rasterizer.hpp:173)

            float iw[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) iw[j] = xrcp(c[j]->clip[3]);
            rr.ax = sx[0]; rr.ay = sy[0];
            rr.v0x = v0x; rr.v0y = v0y; rr.v1x = v1x; rr.v1y = v1y;
            rr.inv_den = xrcp(area2);
            rr.iw0 = iw[0]; rr.iw1 = iw[1]; rr.iw2 = iw[2];
            rr.zw0 = xmul(c0.clip[2], iw[0]); rr.zw1 = xmul(c1.clip[2], iw[1]); rr.zw2 = xmul(c2.clip[2], iw[2]);
            rr.bbox_x = (uint32_t)minx | ((uint32_t)maxx << 16);
            rr.bbox_y = (uint32_t)miny | ((uint32_t)maxy << 16);
            rr.key = key + 1u;
#pragma unroll
            for (int j = 0; j < 3; ++j)
            {
                sr.wp[j][0] = xmul(c[j]->a[0], iw[j]); sr.wp[j][1] = xmul(c[j]->a[1], iw[j]); sr.wp[j][2] = xmul(c[j]->a[2], iw[j]);
                sr.n[j][0] = xmul(c[j]->a[3], iw[j]); sr.n[j][1] = xmul(c[j]->a[4], iw[j]); sr.n[j][2] = xmul(c[j]->a[5], iw[j]);
                sr.uv[j][0] = xmul(c[j]->a[6], iw[j]); sr.uv[j][1] = xmul(c[j]->a[7], iw[j]);
            }
            sr.item = item_index;
            return 3;
        }

        __device__ __forceinline__ void store_records(const Geometry& g, uint32_t slot, const RasterRec& rr, const ShadeRec& sr)
        {
            float4* d = reinterpret_cast<float4*>(g.rrecs + slot);
            const float4* s = reinterpret_cast<const float4*>(&rr);
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = s[i];
            float4* d2 = reinterpret_cast<float4*>(g.srecs + slot);
            const float4* s2 = reinterpret_cast<const float4*>(&sr);
#pragma unroll
            for (int i = 0; i < 7; ++i) d2[i] = s2[i]; // 25 words used; the pad tail is never read
        }

        // Per-tile list sizes are counted where the record is born (one pass less than a separate count kernel).
        // Tile rows are TOP-anchored (y is up in the render target): tile_y = (H-1-y) / TILE.
        struct TileRange { int tx0, tx1, ty0, ty1; };
        __device__ __forceinline__ TileRange tile_range(uint32_t bbox_x, uint32_t bbox_y, int H)
        {
            TileRange t;
            t.tx0 = (int)(bbox_x & 0xffffu) / TILE;
            t.tx1 = (int)(bbox_x >> 16) / TILE;
            t.ty0 = (H - 1 - (int)(bbox_y >> 16)) / TILE_H;
            t.ty1 = (H - 1 - (int)(bbox_y & 0xffffu)) / TILE_H;
            return t;
        }

        // all 32 lanes of the warp must call this; `emit` lanes contribute their record's bbox
        __device__ __forceinline__ void warp_count_tiles(const FrameConst& fc, const Geometry& g, bool emit, uint32_t bbox_x, uint32_t bbox_y)
        {
            TileRange tr{0, -1, 0, -1};
            int n_tiles = 0;
            if (emit)
            {
                tr = tile_range(bbox_x, bbox_y, fc.H);
                n_tiles = (tr.tx1 - tr.tx0 + 1) * (tr.ty1 - tr.ty0 + 1);
            }
            if (n_tiles > 0 && n_tiles <= 8)
            {
                for (int ty = tr.ty0; ty <= tr.ty1; ++ty)
                {
                    if (!owned_row(fc, ty)) continue; // sort-first: another rank's tile row
                    for (int tx = tr.tx0; tx <= tr.tx1; ++tx) atomicAdd(&g.tile_count[(uint32_t)ty * (uint32_t)fc.tiles_x + (uint32_t)tx], 1u);
                }
            }
            unsigned big = __ballot_sync(0xffffffffu, n_tiles > 8);
            const int lane = threadIdx.x & 31;
            while (big)
            {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                const int tx0 = __shfl_sync(0xffffffffu, tr.tx0, src), tx1 = __shfl_sync(0xffffffffu, tr.tx1, src);
                const int ty0 = __shfl_sync(0xffffffffu, tr.ty0, src), ty1 = __shfl_sync(0xffffffffu, tr.ty1, src);
                const int wx = tx1 - tx0 + 1, total = wx * (ty1 - ty0 + 1);
                for (int k = lane; k < total; k += 32)
                    if (owned_row(fc, ty0 + k / wx)) atomicAdd(&g.tile_count[(uint32_t)(ty0 + k / wx) * (uint32_t)fc.tiles_x + (uint32_t)(tx0 + k % wx)], 1u);
            }
        }

        __device__ __forceinline__ void thread_count_tiles(const FrameConst& fc, const Geometry& g, uint32_t bbox_x, uint32_t bbox_y)
        {
            const TileRange tr = tile_range(bbox_x, bbox_y, fc.H);
            for (int ty = tr.ty0; ty <= tr.ty1; ++ty)
            {
                if (!owned_row(fc, ty)) continue;
                for (int tx = tr.tx0; tx <= tr.tx1; ++tx) atomicAdd(&g.tile_count[(uint32_t)ty * (uint32_t)fc.tiles_x + (uint32_t)tx], 1u);
            }
        }

        __device__ __forceinline__ DevStats* stats_shard(const Geometry& g) { return g.stats + (blockIdx.x & (STAT_SHARDS - 1)); }

        // CTA-wide statistics: warp reduce -> shared -> ONE atomic per counter and CTA, into the CTA's shard.
        // Contains a __syncthreads(): every thread of the CTA must call it.
        __device__ __forceinline__ void block_add_stats(const Geometry& g, unsigned tri_input, unsigned after_clip, unsigned raster)
        {
            __shared__ unsigned s_stat[GEOM_THREADS / 32][3];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                tri_input += __shfl_down_sync(0xffffffffu, tri_input, o);
                after_clip += __shfl_down_sync(0xffffffffu, after_clip, o);
                raster += __shfl_down_sync(0xffffffffu, raster, o);
            }
            if ((threadIdx.x & 31) == 0) { s_stat[threadIdx.x >> 5][0] = tri_input; s_stat[threadIdx.x >> 5][1] = after_clip; s_stat[threadIdx.x >> 5][2] = raster; }
            __syncthreads();
            if (threadIdx.x == 0)
            {
                unsigned a = 0, b = 0, c = 0;
#pragma unroll
                for (int w = 0; w < GEOM_THREADS / 32; ++w) { a += s_stat[w][0]; b += s_stat[w][1]; c += s_stat[w][2]; }
                DevStats* st = stats_shard(g);
                if (a) atomicAdd(&st->tri_input, (unsigned long long)a);
                if (b) atomicAdd(&st->tri_after_clip, (unsigned long long)b);
                if (c) atomicAdd(&st->tri_raster, (unsigned long long)c);
            }
        }

        // CTA-wide record allocation: ONE atomicAdd on the global record counter per CTA (a counter hit once per warp
        // serialises ~10^5 same-address atomics on a multi-million-triangle frame).  Returns the calling thread's slot
        // (meaningful where `emit`).  Contains __syncthreads(): every thread of the CTA must call it.
        __device__ __forceinline__ uint32_t block_alloc_records(const Geometry& g, bool emit)
        {
            __shared__ uint32_t s_wcount[GEOM_THREADS / 32];
            __shared__ uint32_t s_base;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const unsigned mask = __ballot_sync(0xffffffffu, emit);
            if (lane == 0) s_wcount[warp] = (uint32_t)__popc(mask);
            __syncthreads();
            if (threadIdx.x == 0)
            {
                uint32_t total = 0;
#pragma unroll
                for (int w = 0; w < GEOM_THREADS / 32; ++w) total += s_wcount[w];
                s_base = total ? atomicAdd(g.rec_count, total) : 0u;
            }
            __syncthreads();
            uint32_t before = 0;
#pragma unroll
            for (int w = 0; w < GEOM_THREADS / 32; ++w) if (w < warp) before += s_wcount[w];
            return s_base + before + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        }

#ifndef SHSB_GEOM_CTAS
#define SHSB_GEOM_CTAS 1
#endif
        __global__ void __launch_bounds__(GEOM_THREADS, SHSB_GEOM_CTAS) geometry_kernel(const FrameConst fc, const Geometry g)
        {
            const uint2 blk = g.block_table[blockIdx.x];
            const DevItem& it = g.items[blk.x];
            const DevMesh mesh = g.meshes[it.mesh];
            const uint32_t ti = blk.y + threadIdx.x;
            const bool active = ti < it.tri_count;
            const bool varyings = fc.shader_id != 5; // SHSB_SHADER_DEPTH_ONLY: pass_adapters.hpp:335-353 sets no varyings

            unsigned n_input = 0, n_after = 0, n_raster = 0;
            bool emit = false;
            RasterRec rr;
            ShadeRec sr;
            if (active)
            {
                n_input = 1;
                uint32_t i0, i1, i2;
                if (mesh.n_indices) { i0 = mesh.indices[(size_t)ti * 3]; i1 = mesh.indices[(size_t)ti * 3 + 1]; i2 = mesh.indices[(size_t)ti * 3 + 2]; }
                else { i0 = ti * 3; i1 = i0 + 1; i2 = i0 + 2; }
                if (i0 < mesh.n_positions && i1 < mesh.n_positions && i2 < mesh.n_positions)
                {
                    const Corner c0 = run_vs(it, mesh, i0, fc, varyings);
                    const Corner c1 = run_vs(it, mesh, i1, fc, varyings);
                    const Corner c2 = run_vs(it, mesh, i2, fc, varyings);
                    if (corner_inside(c0) && corner_inside(c1) && corner_inside(c2))
                    {
                        n_after = 1;
                        const int r = setup_triangle(c0, c1, c2, fc, (it.tri_offset + ti) * 8u, blk.x, rr, sr);
                        n_raster = r & 1;
                        emit = (r & 2) != 0;
                    }
                    else
                    {
                        // A triangle whose three corners are beyond the SAME clip plane leaves the Sutherland-Hodgman loop
                        // (rasterizer.hpp:111-164) empty: every vertex an earlier plane inserts is a convex combination of
                        // corners, so it is beyond that plane too.  Such triangles -- most of what a camera inside a large scene
                        // does not see -- are dropped here instead of being queued; the margin (1e-4 of the operands) keeps
                        // the shortcut away from every case where rounding in the interpolation could decide differently.
                        bool beyond = false;
#pragma unroll
                        for (int p = 0; p < 6; ++p)
                        {
                            const int a = p >> 1;
                            const float m0 = 1e-4f * (fabsf(c0.clip[3]) + fabsf(c0.clip[a])), m1 = 1e-4f * (fabsf(c1.clip[3]) + fabsf(c1.clip[a])),
                                        m2 = 1e-4f * (fabsf(c2.clip[3]) + fabsf(c2.clip[a]));
                            if (plane_dist(c0, p) < -m0 && plane_dist(c1, p) < -m1 && plane_dist(c2, p) < -m2) beyond = true;
                        }
                        if (!beyond)
                        {
                            const uint32_t q = atomicAdd(g.clipq_count, 1u);
                            if (q < g.clipq_capacity) g.clip_queue[q] = make_uint2(blk.x, ti);
                            else { atomicAdd(&stats_shard(g)->overflow_clipq, 1u); *g.overflow_flag = 1u; }
                        }
                    }
                }
            }
            // CTA-aggregated append
            const uint32_t slot = block_alloc_records(g, emit);
            if (emit)
            {
                if (slot < g.rec_capacity) store_records(g, slot, rr, sr);
                else { atomicAdd(&stats_shard(g)->overflow_recs, 1u); *g.overflow_flag = 1u; emit = false; }
            }
            if (__ballot_sync(0xffffffffu, emit)) warp_count_tiles(fc, g, emit, rr.bbox_x, rr.bbox_y);
            block_add_stats(g, n_input, n_after, n_raster);
        }

        // Rare path: triangles with at least one corner outside the clip volume.
        __global__ void __launch_bounds__(GEOM_THREADS) clip_kernel(const FrameConst fc, const Geometry g)
        {
            const uint32_t n = min(*g.clipq_count, g.clipq_capacity);
            const bool varyings = fc.shader_id != 5;
            unsigned n_after = 0, n_raster = 0;
            for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
            {
                const uint2 e = g.clip_queue[q];
                const DevItem& it = g.items[e.x];
                const DevMesh mesh = g.meshes[it.mesh];
                const uint32_t ti = e.y;
                uint32_t i0, i1, i2;
                if (mesh.n_indices) { i0 = mesh.indices[(size_t)ti * 3]; i1 = mesh.indices[(size_t)ti * 3 + 1]; i2 = mesh.indices[(size_t)ti * 3 + 2]; }
                else { i0 = ti * 3; i1 = i0 + 1; i2 = i0 + 2; }

                Corner bufA[MAX_POLY], bufB[MAX_POLY];
                Corner* src = bufA;
                Corner* dst = bufB;
                src[0] = run_vs(it, mesh, i0, fc, varyings);
                src[1] = run_vs(it, mesh, i1, fc, varyings);
                src[2] = run_vs(it, mesh, i2, fc, varyings);
                int np = 3;
                for (int plane = 0; plane < 6 && np > 0; ++plane)
                {
                    int m = 0;
                    for (int i = 0; i < np; ++i)
                    {
                        const Corner& cur = src[i];
                        const Corner& nxt = src[(i + 1 == np) ? 0 : i + 1];
                        const float da = plane_dist(cur, plane);
                        const float db = plane_dist(nxt, plane);
                        const bool cur_in = da >= 0.0f, nxt_in = db >= 0.0f;
                        if (m + 2 > MAX_POLY) break;
                        if (cur_in && nxt_in) dst[m++] = nxt;
                        else if (cur_in != nxt_in)
                        {
                            const float denom = xsub(da, db);
                            if (fabsf(denom) > 1e-8f) dst[m++] = lerp_corner(cur, nxt, xdiv(da, denom));
                            if (nxt_in) dst[m++] = nxt;
                        }
                    }
                    Corner* t = src; src = dst; dst = t;
                    np = m;
                }
                if (np < 3) continue;
                for (int k = 1; k + 1 < np; ++k)
                {
                    ++n_after;
                    RasterRec rr;
                    ShadeRec sr;
                    const int r = setup_triangle(src[0], src[k], src[k + 1], fc, (it.tri_offset + ti) * 8u + (uint32_t)(k - 1), e.x, rr, sr);
                    n_raster += r & 1;
                    if (r & 2)
                    {
                        const uint32_t slot = atomicAdd(g.rec_count, 1u);
                        if (slot < g.rec_capacity) { store_records(g, slot, rr, sr); thread_count_tiles(fc, g, rr.bbox_x, rr.bbox_y); }
                        else { atomicAdd(&stats_shard(g)->overflow_recs, 1u); *g.overflow_flag = 1u; }
                    }
                }
            }
            block_add_stats(g, 0, n_after, n_raster);
        }

        // PassShadowMap set-up rules (passes/pass_shadow_map.hpp:155-190): world = vec3(model * p), clip = light_vp * (world, 1),
        // reject |w| < 1e-8, reject only if all three corners are beyond the same NDC bound, NO clipping, NO culling.
        // One texel of PassShadowMap's inline raster (pass_shadow_map.hpp:185-201) for one set-up triangle: the same operations, in the
        // same order, as the tile kernel's shadow mode.  The pass keeps the MINIMUM depth per texel, which does not depend on the
        // order triangles arrive in, so an atomic minimum on the float's bit pattern (depths are in [0, 1]: bit order == value
        // order) reproduces the serial loop bit for bit.  NaN depths fail `z01 < 1` exactly as they fail the reference's `z01 < zbuf`.
        __device__ __forceinline__ void shadow_texel(const RasterRec& r, int x, int y, uint32_t* __restrict__ depth_bits, int W)
        {
            const float pxf = xadd((float)x, 0.5f), pyf = xadd((float)y, 0.5f);
            const float v2x = xsub(pxf, r.ax), v2y = xsub(pyf, r.ay);
            const float bv = xmul(xsub(xmul(v2x, r.v1y), xmul(r.v1x, v2y)), r.inv_den);
            const float bw = xmul(xsub(xmul(r.v0x, v2y), xmul(v2x, r.v0y)), r.inv_den);
            const float bu = xsub(xsub(1.0f, bv), bw);
            if (bu < 0.0f || bv < 0.0f || bw < 0.0f) return;
            const float z_ndc = xadd(xadd(xmul(bu, r.zw0), xmul(bv, r.zw1)), xmul(bw, r.zw2));
            const float z01 = sclamp(xadd(xmul(z_ndc, 0.5f), 0.5f), 0.0f, 1.0f);
            if (z01 < 1.0f) atomicMin(depth_bits + (size_t)y * (size_t)W + (size_t)x, __float_as_uint(z01));
        }

        constexpr int SHADOW_SMALL_AREA = 64;     // clamped bbox of at most this many texels: walked by the triangle's own lane
        constexpr int SHADOW_MEDIUM_AREA = 4096;  // up to this many: walked by the whole warp; larger triangles go through the tile bins

        // PassShadowMap::execute, pass_shadow_map.hpp:144-203: one thread per caster triangle.  Vertex transform, the all-outside
        // reject, the screen map and the bbox are the reference's; what happens to a surviving triangle depends on its size.  A
        // shadow map sees the scene from far away, so nearly all of its triangles cover a handful of texels: binning them into
        // 16x16 tiles and testing 32 texels per (block, triangle) pair spends 250 instructions on a triangle that covers 8 texels.
        // With `fc.direct_depth` the kernel rasterises them itself into the pre-cleared depth plane -- small ones in the lane that
        // set them up, medium ones by the whole warp -- and only what is larger than 64x64 texels is emitted as a record for the
        // binned tile kernel (which then LOADS the plane, load_depth = 1).
        __global__ void __launch_bounds__(GEOM_THREADS) shadow_geometry_kernel(const FrameConst fc, const Geometry g)
        {
            const uint2 blk = g.block_table[blockIdx.x];
            const DevItem& it = g.items[blk.x];
            const DevMesh mesh = g.meshes[it.mesh];
            const uint32_t ti = blk.y + threadIdx.x;
            bool emit = false;
            RasterRec rr;
            int minx = 0, maxx = -1, miny = 0, maxy = -1;
            if (ti < it.tri_count)
            {
                uint32_t id[3];
                if (mesh.n_indices) { id[0] = mesh.indices[(size_t)ti * 3]; id[1] = mesh.indices[(size_t)ti * 3 + 1]; id[2] = mesh.indices[(size_t)ti * 3 + 2]; }
                else { id[0] = ti * 3; id[1] = id[0] + 1; id[2] = id[0] + 2; }
                if (id[0] < mesh.n_positions && id[1] < mesh.n_positions && id[2] < mesh.n_positions)
                {
                    float nx[3], ny[3], nz[3];
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                    {
                        const float* p = mesh.positions + (size_t)id[j] * 3;
                        const float4 wp = xmat4_mul(it.model, p[0], p[1], p[2], 1.0f);
                        const float4 c = xmat4_mul(fc.viewproj, wp.x, wp.y, wp.z, 1.0f);
                        if (fabsf(c.w) < 1e-8f) ok = false;
                        nx[j] = xdiv(c.x, c.w); ny[j] = xdiv(c.y, c.w); nz[j] = xdiv(c.z, c.w);
                    }
                    if (ok &&
                        !((nx[0] < -1.0f && nx[1] < -1.0f && nx[2] < -1.0f) || (nx[0] > 1.0f && nx[1] > 1.0f && nx[2] > 1.0f)) &&
                        !((ny[0] < -1.0f && ny[1] < -1.0f && ny[2] < -1.0f) || (ny[0] > 1.0f && ny[1] > 1.0f && ny[2] > 1.0f)) &&
                        !((nz[0] < -1.0f && nz[1] < -1.0f && nz[2] < -1.0f) || (nz[0] > 1.0f && nz[1] > 1.0f && nz[2] > 1.0f)))
                    {
                        const float fw1 = (float)(fc.W - 1), fh1 = (float)(fc.H - 1);
                        float sx[3], sy[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                        {
                            sx[j] = xmul(xadd(xmul(nx[j], 0.5f), 0.5f), fw1);
                            sy[j] = xmul(xadd(xmul(ny[j], 0.5f), 0.5f), fh1);
                        }
                        // floor/ceil of +-inf/NaN are UB on the CPU side as well; a sane light camera never produces them.
                        const float minxf = fminf(fminf(sx[0], sx[1]), sx[2]), maxxf = fmaxf(fmaxf(sx[0], sx[1]), sx[2]);
                        const float minyf = fminf(fminf(sy[0], sy[1]), sy[2]), maxyf = fmaxf(fmaxf(sy[0], sy[1]), sy[2]);
                        minx = max(0, (int)floorf(minxf));
                        maxx = min(fc.W - 1, (int)ceilf(maxxf));
                        miny = max(0, (int)floorf(minyf));
                        maxy = min(fc.H - 1, (int)ceilf(maxyf));
                        const float v0x = xsub(sx[1], sx[0]), v0y = xsub(sy[1], sy[0]);
                        const float v1x = xsub(sx[2], sx[0]), v1y = xsub(sy[2], sy[0]);
                        const float den = xsub(xmul(v0x, v1y), xmul(v1x, v0y));
                        if (minx <= maxx && miny <= maxy && !(fabsf(den) < 1e-8f))
                        {
                            rr.ax = sx[0]; rr.ay = sy[0];
                            rr.v0x = v0x; rr.v0y = v0y; rr.v1x = v1x; rr.v1y = v1y;
                            rr.inv_den = xrcp(den);
                            rr.iw0 = rr.iw1 = rr.iw2 = 1.0f;
                            rr.zw0 = nz[0]; rr.zw1 = nz[1]; rr.zw2 = nz[2];
                            rr.bbox_x = (uint32_t)minx | ((uint32_t)maxx << 16);
                            rr.bbox_y = (uint32_t)miny | ((uint32_t)maxy << 16);
                            rr.key = (it.tri_offset + ti) * 8u + 1u;
                            emit = true;
                        }
                    }
                }
            }
            if (fc.direct_depth)
            {
                uint32_t* depth_bits = reinterpret_cast<uint32_t*>(fc.direct_depth);
                const int bw = maxx - minx + 1, bh = maxy - miny + 1;
                const long long area = emit ? (long long)bw * (long long)bh : 0;
                if (emit && area <= SHADOW_SMALL_AREA)
                {
                    for (int y = miny; y <= maxy; ++y)
                        for (int x = minx; x <= maxx; ++x) shadow_texel(rr, x, y, depth_bits, fc.W);
                    emit = false;
                }
                // medium triangles: the warp walks each one's bbox together (a serial walk of a 40x40 bbox by one lane would stall 31 others)
                unsigned med = __ballot_sync(0xffffffffu, emit && area <= SHADOW_MEDIUM_AREA);
                const int lane = threadIdx.x & 31;
                if (emit && area <= SHADOW_MEDIUM_AREA) emit = false;
                while (med)
                {
                    const int src = __ffs(med) - 1;
                    med &= med - 1;
                    RasterRec q;
                    q.ax = __shfl_sync(0xffffffffu, rr.ax, src); q.ay = __shfl_sync(0xffffffffu, rr.ay, src);
                    q.v0x = __shfl_sync(0xffffffffu, rr.v0x, src); q.v0y = __shfl_sync(0xffffffffu, rr.v0y, src);
                    q.v1x = __shfl_sync(0xffffffffu, rr.v1x, src); q.v1y = __shfl_sync(0xffffffffu, rr.v1y, src);
                    q.inv_den = __shfl_sync(0xffffffffu, rr.inv_den, src);
                    q.zw0 = __shfl_sync(0xffffffffu, rr.zw0, src); q.zw1 = __shfl_sync(0xffffffffu, rr.zw1, src); q.zw2 = __shfl_sync(0xffffffffu, rr.zw2, src);
                    const int x0 = __shfl_sync(0xffffffffu, minx, src), y0 = __shfl_sync(0xffffffffu, miny, src);
                    const int w = __shfl_sync(0xffffffffu, bw, src), total = w * __shfl_sync(0xffffffffu, bh, src);
                    for (int k = lane; k < total; k += 32) shadow_texel(q, x0 + k % w, y0 + k / w, depth_bits, fc.W);
                }
            }
            uint32_t slot;
            if (fc.direct_depth)
            {
                // what is left are the few triangles larger than 64x64 texels: one counter atomic per warp that has any, no CTA
                // barrier (the lanes of a CTA finish their texel walks at very different times)
                const unsigned mask = __ballot_sync(0xffffffffu, emit);
                if (!mask) return;
                const int lane = threadIdx.x & 31;
                uint32_t base = 0;
                if (lane == __ffs(mask) - 1) base = atomicAdd(g.rec_count, (uint32_t)__popc(mask));
                slot = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1) + (uint32_t)__popc(mask & ((1u << lane) - 1u));
            }
            else slot = block_alloc_records(g, emit);
            if (emit)
            {
                if (slot < g.rec_capacity)
                {
                    float4* d = reinterpret_cast<float4*>(g.rrecs + slot);
                    const float4* s = reinterpret_cast<const float4*>(&rr);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] = s[i];
                }
                else { atomicAdd(&stats_shard(g)->overflow_recs, 1u); *g.overflow_flag = 1u; emit = false; }
            }
            if (fc.direct_depth) return; // large triangles are walked by shadow_large_kernel, not binned
            if (__ballot_sync(0xffffffffu, emit)) warp_count_tiles(fc, g, emit, rr.bbox_x, rr.bbox_y);
        }

        // The shadow pass's large triangles (bbox above 64x64 texels; a ground plane under the casters is the typical one): their
        // records are cut into 64x64-texel chunks and the chunks dealt round-robin over a fixed grid, so that one triangle spanning
        // the whole map is walked by every SM.  Every CTA scans the (short) record list and keeps a running chunk count; a chunk is
        // walked by the 256 threads of its CTA, 16 texels each, with the same exact per-texel test and atomic minimum as above.
        constexpr int SHADOW_CHUNK = 64;
        __global__ void __launch_bounds__(256) shadow_large_kernel(const FrameConst fc, const Geometry g)
        {
            const uint32_t n = min(*g.rec_count, g.rec_capacity);
            uint32_t* depth_bits = reinterpret_cast<uint32_t*>(fc.direct_depth);
            uint32_t seen = 0; // chunks of the records before this one
            for (uint32_t ri = 0; ri < n; ++ri)
            {
                const RasterRec* rp = g.rrecs + ri;
                const uint32_t bx = rp->bbox_x, by = rp->bbox_y;
                const int minx = (int)(bx & 0xffffu), maxx = (int)(bx >> 16), miny = (int)(by & 0xffffu), maxy = (int)(by >> 16);
                const uint32_t cx = (uint32_t)(maxx - minx) / SHADOW_CHUNK + 1u, cy = (uint32_t)(maxy - miny) / SHADOW_CHUNK + 1u;
                const uint32_t chunks = cx * cy;
                // first chunk of this record that falls to this CTA: global chunk index == blockIdx.x (mod gridDim.x)
                uint32_t c = (blockIdx.x + gridDim.x - seen % gridDim.x) % gridDim.x;
                if (c < chunks)
                {
                    const RasterRec r = *rp;
                    for (; c < chunks; c += gridDim.x)
                    {
                        const int x0 = minx + (int)(c % cx) * SHADOW_CHUNK, y0 = miny + (int)(c / cx) * SHADOW_CHUNK;
#pragma unroll 4
                        for (int k = threadIdx.x; k < SHADOW_CHUNK * SHADOW_CHUNK; k += 256)
                        {
                            const int x = x0 + (k % SHADOW_CHUNK), y = y0 + (k / SHADOW_CHUNK);
                            if (x <= maxx && y <= maxy) shadow_texel(r, x, y, depth_bits, fc.W);
                        }
                    }
                }
                seen += chunks;
            }
        }
    }

    void launch_geometry(const FrameConst& fc, const Geometry& g, cudaStream_t s, uint64_t* launches)
    {
        if (g.n_blocks == 0) return;
        if (fc.shadow_mode)
        {
            shadow_geometry_kernel<<<g.n_blocks, GEOM_THREADS, 0, s>>>(fc, g);
            *launches += 1;
            if (fc.direct_depth)
            {
                shadow_large_kernel<<<148 * 2, 256, 0, s>>>(fc, g);
                *launches += 1;
            }
            return;
        }
        geometry_kernel<<<g.n_blocks, GEOM_THREADS, 0, s>>>(fc, g);
        // The clip queue length lives on the device; a fixed persistent-style grid (4 CTAs per SM) strides over it.
        clip_kernel<<<148 * 4, GEOM_THREADS, 0, s>>>(fc, g);
        *launches += 2;
    }
}
