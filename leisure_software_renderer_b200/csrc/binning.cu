// binning.cu -- K2: triangle -> screen-tile lists (CSR).
//
// The reference has no binning: rasterize_mesh scans each triangle's clamped bbox serially
// (sw_render/rasterizer.hpp:280-289, 330-338) and the legacy path brute-forces tiles x triangles
// (hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp:260-310).  Here every set-up triangle is
// appended to the list of each 16x16 tile its clamped bbox touches.  Binning by the SAME integer bbox the
// reference iterates is trivially conservative: a pixel outside the bbox is never tested by the reference,
// and the tile kernel repeats that bbox test per pixel.
//
// Per-tile list sizes are counted by the geometry kernels when a record is emitted.  Two launches remain:
//   alloc : one thread per tile reserves its list segment with one atomicAdd on a global cursor (segment
//           order in memory is arbitrary -- no serial scan) and files the tile into a scheduling class so
//           that the tile kernel starts the most expensive tiles first;
//   fill  : every record appends its index to the segment of each tile its bbox touches.  Small triangles
//           (<= 8 tiles) are handled by their own lane; larger ones are walked by the whole warp so that a
//           full-screen triangle does not serialise thousands of atomics in one thread.
// List order is arbitrary; visibility is resolved by (depth, draw-order key) in the tile kernel.
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        constexpr int BIN_THREADS = 256;
        constexpr int SMALL_TILES = 8;

        struct TileRange { int tx0, tx1, ty0, ty1; };

        __device__ __forceinline__ TileRange tile_range(const RasterRec& r, int H)
        {
            const int minx = (int)(r.bbox_x & 0xffffu), maxx = (int)(r.bbox_x >> 16);
            const int miny = (int)(r.bbox_y & 0xffffu), maxy = (int)(r.bbox_y >> 16);
            TileRange t;
            t.tx0 = minx / TILE;
            t.tx1 = maxx / TILE;
            t.ty0 = (H - 1 - maxy) / TILE_H; // tile rows are anchored at the TOP of the frame (y is up in the RT)
            t.ty1 = (H - 1 - miny) / TILE_H;
            return t;
        }

        __global__ void __launch_bounds__(BIN_THREADS) bin_kernel(const FrameConst fc, const Geometry g)
        {
            const uint32_t n_recs = min(*g.rec_count, g.rec_capacity);
            const int lane = threadIdx.x & 31;
            // grid-stride over warps' worth of records so that every lane of a warp stays in the loop together
            for (uint32_t base = (blockIdx.x * BIN_THREADS + (threadIdx.x & ~31u)); base < n_recs; base += gridDim.x * BIN_THREADS)
            {
                const uint32_t i = base + lane;
                TileRange tr{0, -1, 0, -1};
                int n_tiles = 0;
                if (i < n_recs)
                {
                    RasterRec tmp;
                    tmp.bbox_x = g.rrecs[i].bbox_x; // offsets 52 / 56: two 4-byte loads (not 8-byte aligned)
                    tmp.bbox_y = g.rrecs[i].bbox_y;
                    tr = tile_range(tmp, fc.H);
                    n_tiles = (tr.tx1 - tr.tx0 + 1) * (tr.ty1 - tr.ty0 + 1);
                }
                if (n_tiles > 0 && n_tiles <= SMALL_TILES)
                {
                    for (int ty = tr.ty0; ty <= tr.ty1; ++ty)
                    {
                        if (!owned_row(fc, ty)) continue; // must mirror the counting in geometry.cu exactly
                        for (int tx = tr.tx0; tx <= tr.tx1; ++tx)
                        {
                            const uint32_t t = (uint32_t)ty * (uint32_t)fc.tiles_x + (uint32_t)tx;
                            const uint32_t pos = g.tile_offset[t] + atomicAdd(&g.tile_fill[t], 1u);
                            if (pos < g.list_capacity) g.tile_list[pos] = i;
                        }
                    }
                }
                unsigned big = __ballot_sync(0xffffffffu, n_tiles > SMALL_TILES);
                while (big)
                {
                    const int src = __ffs(big) - 1;
                    big &= big - 1;
                    const int tx0 = __shfl_sync(0xffffffffu, tr.tx0, src), tx1 = __shfl_sync(0xffffffffu, tr.tx1, src);
                    const int ty0 = __shfl_sync(0xffffffffu, tr.ty0, src), ty1 = __shfl_sync(0xffffffffu, tr.ty1, src);
                    const uint32_t rec = base + (uint32_t)src;
                    const int wx = tx1 - tx0 + 1;
                    const int total = wx * (ty1 - ty0 + 1);
                    for (int k = lane; k < total; k += 32)
                    {
                        const int ty = ty0 + k / wx, tx = tx0 + k % wx;
                        if (!owned_row(fc, ty)) continue;
                        const uint32_t t = (uint32_t)ty * (uint32_t)fc.tiles_x + (uint32_t)tx;
                        const uint32_t pos = g.tile_offset[t] + atomicAdd(&g.tile_fill[t], 1u);
                        if (pos < g.list_capacity) g.tile_list[pos] = rec;
                    }
                }
            }
        }

        // One thread per tile: reserve the list segment, reset the write cursor, choose the scheduling class.
        // Atomics are warp-aggregated: one atomicAdd per warp on the list cursor and one per (warp, class).
        __global__ void __launch_bounds__(BIN_THREADS) alloc_kernel(const FrameConst fc, const Geometry g, uint32_t n_tiles)
        {
            const uint32_t t = blockIdx.x * BIN_THREADS + threadIdx.x;
            const int lane = threadIdx.x & 31;
            const bool in_frame = t < n_tiles;
            const bool live = in_frame && owned_row(fc, (int)(t / (uint32_t)fc.tiles_x)); // rows of other ranks are not scheduled at all
            const uint32_t c = live ? g.tile_count[t] : 0u;
            // warp exclusive scan of the counts -> one cursor atomic per warp
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t base = 0;
            if (lane == 31 && warp_total) base = atomicAdd(g.list_cursor, warp_total);
            base = __shfl_sync(0xffffffffu, base, 31);
            const uint32_t off = base + incl - c;
            if (lane == 31 && warp_total && base + warp_total > g.list_capacity) { atomicAdd(&g.stats[blockIdx.x & (STAT_SHARDS - 1)].overflow_lists, 1u); *g.overflow_flag = 1u; }
            // class 0: geometry + saturated light list (walks all lights), 1: geometry + long light list,
            // 2: geometry, 3: background only
            uint32_t cls = c ? 2u : 3u;
            if (live && c && fc.forward_plus && fc.light_tile_size == (uint32_t)TILE)
            {
                // the light tile (16 x 16) this raster tile lies in
                const uint32_t lc = fc.tile_counts[min((t / (uint32_t)fc.tiles_x) * (uint32_t)TILE_H / (uint32_t)TILE, fc.light_tiles_y - 1u) * fc.light_tiles_x +
                                                   min(t % (uint32_t)fc.tiles_x, fc.light_tiles_x - 1u)];
                if (lc >= fc.max_per_tile) cls = 0u;
                else if (lc >= 48u) cls = 1u;
            }
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
            {
                const unsigned m = __ballot_sync(0xffffffffu, live && cls == k);
                if (!m) continue;
                uint32_t cbase = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) cbase = atomicAdd(&g.class_count[k], (uint32_t)__popc(m));
                cbase = __shfl_sync(0xffffffffu, cbase, leader);
                if (live && cls == k) g.tile_order[(size_t)k * n_tiles + cbase + (uint32_t)__popc(m & ((1u << lane) - 1u))] = (t % (uint32_t)fc.tiles_x) | ((t / (uint32_t)fc.tiles_x) << 16);
            }
            if (in_frame)
            {
                g.tile_offset[t] = off;
                g.tile_fill[t] = 0;
            }
        }
    }

    void launch_binning(const FrameConst& fc, const Geometry& g, cudaStream_t s, uint64_t* launches)
    {
        const uint32_t n_tiles = (uint32_t)fc.tiles_x * (uint32_t)fc.tiles_y;
        alloc_kernel<<<(n_tiles + BIN_THREADS - 1) / BIN_THREADS, BIN_THREADS, 0, s>>>(fc, g, n_tiles);
        bin_kernel<<<148 * 8, BIN_THREADS, 0, s>>>(fc, g); // persistent-style grid: 148 SMs x 8 CTAs, grid-stride
        *launches += 2;
    }
}
