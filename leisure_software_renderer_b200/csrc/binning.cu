// binning.cu -- K2: triangle -> screen-tile lists (CSR).
//
// The reference has no binning: rasterize_mesh scans each triangle's clamped bbox serially
// (sw_render/rasterizer.hpp:280-289, 330-338) and the legacy path brute-forces tiles x triangles
// (hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp:260-310).  Here every set-up triangle is
// appended to the list of each 16x16 tile its clamped bbox touches.  Binning by the SAME integer bbox the
// reference iterates is trivially conservative: a pixel outside the bbox is never tested by the reference,
// and the tile kernel repeats that bbox test per pixel.
//
// Three launches: count (atomicAdd per tile), exclusive scan over tiles, fill.  Small triangles (<= 8
// tiles) are handled by their own lane; larger ones are walked by the whole warp so that a full-screen
// triangle does not serialise thousands of atomics in one thread.  List order is arbitrary; visibility
// is resolved by (depth, draw-order key) in the tile kernel.
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
            t.ty0 = (H - 1 - maxy) / TILE; // tile rows are anchored at the TOP of the frame (y is up in the RT)
            t.ty1 = (H - 1 - miny) / TILE;
            return t;
        }

        template <bool FILL>
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
                        for (int tx = tr.tx0; tx <= tr.tx1; ++tx)
                        {
                            const uint32_t t = (uint32_t)ty * (uint32_t)fc.tiles_x + (uint32_t)tx;
                            if (FILL)
                            {
                                const uint32_t pos = g.tile_offset[t] + atomicAdd(&g.tile_fill[t], 1u);
                                if (pos < g.list_capacity) g.tile_list[pos] = i;
                            }
                            else atomicAdd(&g.tile_count[t], 1u);
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
                        const uint32_t t = (uint32_t)ty * (uint32_t)fc.tiles_x + (uint32_t)tx;
                        if (FILL)
                        {
                            const uint32_t pos = g.tile_offset[t] + atomicAdd(&g.tile_fill[t], 1u);
                            if (pos < g.list_capacity) g.tile_list[pos] = rec;
                        }
                        else atomicAdd(&g.tile_count[t], 1u);
                    }
                }
            }
        }

        // Exclusive scan of tile_count -> tile_offset (n+1 entries) by one 1024-thread CTA; also zeroes tile_fill.
        __global__ void __launch_bounds__(1024) scan_kernel(const Geometry g, uint32_t n_tiles)
        {
            __shared__ uint32_t warp_sums[32];
            __shared__ uint32_t carry;
            const uint32_t per = (n_tiles + 1023u) / 1024u;
            const uint32_t begin = min(threadIdx.x * per, n_tiles), end = min(begin + per, n_tiles);
            uint32_t sum = 0;
            for (uint32_t t = begin; t < end; ++t) sum += g.tile_count[t];
            // block-wide exclusive scan of the per-thread sums
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) warp_sums[warp] = incl;
            __syncthreads();
            if (warp == 0)
            {
                uint32_t w = warp_sums[lane];
                uint32_t wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += v;
                }
                warp_sums[lane] = wi - w; // exclusive
                if (lane == 31) carry = wi;
            }
            __syncthreads();
            uint32_t run = warp_sums[warp] + (incl - sum);
            for (uint32_t t = begin; t < end; ++t)
            {
                const uint32_t c = g.tile_count[t];
                g.tile_offset[t] = run;
                g.tile_fill[t] = 0;
                run += c;
            }
            if (threadIdx.x == 0)
            {
                g.tile_offset[n_tiles] = carry;
                if (carry > g.list_capacity) atomicAdd(&g.stats->overflow_lists, 1u);
            }
        }
    }

    void launch_binning(const FrameConst& fc, const Geometry& g, cudaStream_t s, uint64_t* launches)
    {
        const uint32_t n_tiles = (uint32_t)fc.tiles_x * (uint32_t)fc.tiles_y;
        // tile_count was zeroed by the frame's arena reset.  Persistent-style grid: 148 SMs x 8 CTAs.
        const int grid = 148 * 8;
        bin_kernel<false><<<grid, BIN_THREADS, 0, s>>>(fc, g);
        scan_kernel<<<1, 1024, 0, s>>>(g, n_tiles);
        bin_kernel<true><<<grid, BIN_THREADS, 0, s>>>(fc, g);
        *launches += 3;
    }
}
