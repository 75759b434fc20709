// flat_draw.cu -- the reference's flat-shaded software mesh draws (sw_render/debug_draw.hpp:64-203; the multi-light variant of
// exp-plumbing/hello_light_types_culling_sw.cpp:366-422 that consumes the per-object LightSelection) for a whole batch of draws.
//
// The reference walks draws, triangles and texels serially: a fragment replaces the texel when its depth is strictly below the
// stored one, so among equal depths the EARLIEST triangle of the earliest draw stays.  That is the minimum of (depth, order) with
// order = the running triangle number of the batch, and a minimum is order-free:
//   flat_init_kernel     zkey[texel] = depth bits << 32 | 0        (what is already in the depth buffer wins every tie)
//   flat_setup_kernel    one thread per (draw, triangle): world positions, projection, face normal, the flat colour (all lights of the
//                        draw's selection), triangle setup; the triangle's bounding box is cut into 64 x 64 texel chunks and one work
//                        item per chunk is appended (a floor that covers the canvas becomes ~100 items, a far capsule facet one)
//   flat_raster_kernel   one warp per work item: lanes walk the chunk's texels, atomicMin(zkey, depth bits << 32 | 1 + order)
//   flat_resolve_kernel  texels whose key carries an order take that triangle's colour and the new depth
// Depths are in [0, 1] (draw_filled_triangle rejects the rest), so their bit patterns order like the values.
// Compiled --fmad=false: every expression is the reference's, unfused (flat_draw_core.cuh).
#include "shsb_dev.cuh"
#include "flat_draw_core.cuh"

namespace shsb
{
    namespace
    {
        constexpr int CHUNK_SHIFT = 6; // 64 x 64 texels per work item

        struct Batch
        {
            float view_proj[16];
            float camera[3];
            float L[3];
            int W, H, mode;
            uint32_t n_draws, n_tris, n_lights, item_cap;
        };

        __global__ void flat_init_kernel(const float* __restrict__ depth, unsigned long long* __restrict__ zkey, uint32_t n)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i < n) zkey[i] = (unsigned long long)__float_as_uint(depth[i]) << 32;
        }

        __global__ void __launch_bounds__(128) flat_setup_kernel(Batch b, const fd::DrawRec* __restrict__ draws, const fd::LightProps* __restrict__ lights,
                                                                 sc::OccTri* __restrict__ tris, uint32_t* __restrict__ colours, uint2* __restrict__ items,
                                                                 uint32_t* __restrict__ item_total)
        {
            const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
            if (g >= b.n_tris) return;
            // the draw this triangle belongs to: last draw whose tri_base <= g
            uint32_t lo = 0, hi = b.n_draws - 1;
            while (lo < hi)
            {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (draws[mid].tri_base <= g) lo = mid; else hi = mid - 1;
            }
            const fd::DrawRec& d = draws[lo];
            const uint32_t t = g - d.tri_base;
            sc::OccTri rec;
            rec.valid = false;
            const uint32_t i0 = d.indices[3 * t], i1 = d.indices[3 * t + 1], i2 = d.indices[3 * t + 2];
            if (i0 < d.n_positions && i1 < d.n_positions && i2 < d.n_positions)
            {
                float w0[4], w1[4], w2[4], s0[2], s1[2], s2[2], z0, z1, z2;
                sc::mul4(d.model, d.positions[3 * i0], d.positions[3 * i0 + 1], d.positions[3 * i0 + 2], 1.0f, w0);
                sc::mul4(d.model, d.positions[3 * i1], d.positions[3 * i1 + 1], d.positions[3 * i1 + 2], 1.0f, w1);
                sc::mul4(d.model, d.positions[3 * i2], d.positions[3 * i2 + 1], d.positions[3 * i2 + 2], 1.0f, w2);
                fd::V3 n;
                if (fd::project_world(w0, b.view_proj, b.W, b.H, s0, z0) && fd::project_world(w1, b.view_proj, b.W, b.H, s1, z1) &&
                    fd::project_world(w2, b.view_proj, b.W, b.H, s2, z2) && fd::face_normal(fd::v3(w0), fd::v3(w1), fd::v3(w2), n))
                {
                    rec = sc::occ_setup_triangle(s0, z0, s1, z1, s2, z2, b.W, b.H);
                    if (rec.valid)
                        colours[g] = (b.mode == fd::MODE_BLINN_PHONG)
                                         ? fd::blinn_phong_colour(fd::v3(w0), fd::v3(w1), fd::v3(w2), n, fd::v3(b.camera), fd::v3(b.L), fd::v3(d.base))
                                         : fd::multi_light_colour(fd::v3(w0), fd::v3(w1), fd::v3(w2), n, fd::v3(b.camera), fd::v3(d.base), lights, b.n_lights,
                                                                  d.selection, d.selection_count);
                }
            }
            tris[g] = rec;
            if (!rec.valid) return;
            const int cx0 = rec.min_x >> CHUNK_SHIFT, cx1 = rec.max_x >> CHUNK_SHIFT, cy0 = rec.min_y >> CHUNK_SHIFT, cy1 = rec.max_y >> CHUNK_SHIFT;
            const uint32_t n_items = (uint32_t)((cx1 - cx0 + 1) * (cy1 - cy0 + 1));
            const uint32_t at = atomicAdd(item_total, n_items);
            if (at + n_items > b.item_cap) return; // the host sees the total, grows the list and runs the batch again
            uint32_t k = at;
            for (int cy = cy0; cy <= cy1; ++cy)
                for (int cx = cx0; cx <= cx1; ++cx) items[k++] = make_uint2(g, ((uint32_t)cy << 16) | (uint32_t)cx);
        }

        __global__ void __launch_bounds__(256) flat_raster_kernel(const sc::OccTri* __restrict__ tris, const uint2* __restrict__ items, uint32_t n_items,
                                                                  unsigned long long* __restrict__ zkey, int W)
        {
            const uint32_t lane = threadIdx.x & 31u;
            const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
            for (uint32_t it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += warps)
            {
                const uint2 item = items[it];
                const sc::OccTri t = tris[item.x];
                const int cx = (int)(item.y & 0xFFFFu), cy = (int)(item.y >> 16);
                const int x0 = max(t.min_x, cx << CHUNK_SHIFT), x1 = min(t.max_x, (cx << CHUNK_SHIFT) + 63);
                const int y0 = max(t.min_y, cy << CHUNK_SHIFT), y1 = min(t.max_y, (cy << CHUNK_SHIFT) + 63);
                const int w = x1 - x0 + 1, n = w * (y1 - y0 + 1);
                const unsigned long long order = (unsigned long long)item.x + 1ull;
                for (int i = (int)lane; i < n; i += 32)
                {
                    const int y = y0 + i / w, x = x0 + i % w;
                    float depth;
                    if (!sc::occ_texel_depth(t, x, y, depth)) continue;
                    depth = depth + 0.0f; // -0 compares equal to +0 in the reference's test: one bit pattern for both
                    atomicMin(&zkey[(size_t)y * W + x], ((unsigned long long)__float_as_uint(depth) << 32) | order);
                }
            }
        }

        __global__ void flat_resolve_kernel(const unsigned long long* __restrict__ zkey, const uint32_t* __restrict__ colours, float* __restrict__ depth,
                                            uint32_t* __restrict__ canvas, uint32_t n)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n) return;
            const unsigned long long k = zkey[i];
            const uint32_t order = (uint32_t)k;
            if (order == 0u) return;
            depth[i] = __uint_as_float((uint32_t)(k >> 32));
            canvas[i] = colours[order - 1u];
        }
    }

    void launch_flat_draw_batch(const fd::BatchDesc& bd, const fd::DrawRec* draws, const fd::LightProps* lights, void* tris, uint32_t* colours, uint2* items, uint32_t item_cap,
                                uint32_t* item_total, unsigned long long* zkey, float* depth, bool init_keys, cudaStream_t s, uint64_t* launches)
    {
        Batch b;
        for (int i = 0; i < 16; ++i) b.view_proj[i] = bd.view_proj[i];
        for (int i = 0; i < 3; ++i) { b.camera[i] = bd.camera[i]; b.L[i] = bd.L[i]; }
        b.W = bd.W; b.H = bd.H; b.mode = bd.mode; b.n_draws = bd.n_draws; b.n_tris = bd.n_tris; b.n_lights = bd.n_lights; b.item_cap = item_cap;
        const uint32_t n_px = (uint32_t)bd.W * (uint32_t)bd.H;
        uint64_t n = 0;
        if (init_keys) { flat_init_kernel<<<(n_px + 255) / 256, 256, 0, s>>>(depth, zkey, n_px); ++n; }
        cudaMemsetAsync(item_total, 0, sizeof(uint32_t), s);
        if (bd.n_tris) { flat_setup_kernel<<<(bd.n_tris + 127) / 128, 128, 0, s>>>(b, draws, lights, (sc::OccTri*)tris, colours, items, item_total); ++n; }
        if (launches) *launches += n;
    }

    void launch_flat_raster_resolve(const fd::BatchDesc& bd, const void* tris, const uint32_t* colours, const uint2* items, uint32_t n_items,
                                    unsigned long long* zkey, float* depth, uchar4* canvas, cudaStream_t s, uint64_t* launches)
    {
        const uint32_t n_px = (uint32_t)bd.W * (uint32_t)bd.H;
        uint64_t n = 1;
        if (n_items)
        {
            const uint32_t blocks = (n_items + 7u) / 8u; // one warp per work item, 8 warps per block, at most 16 resident blocks per SM's worth of grid
            flat_raster_kernel<<<blocks < 148u * 16u ? blocks : 148u * 16u, 256, 0, s>>>((const sc::OccTri*)tris, items, n_items, zkey, bd.W);
            ++n;
        }
        flat_resolve_kernel<<<(n_px + 255) / 256, 256, 0, s>>>(zkey, colours, depth, (uint32_t*)canvas, n_px);
        if (launches) *launches += n;
    }

    size_t flat_tri_record_bytes() { return sizeof(sc::OccTri); }
}
