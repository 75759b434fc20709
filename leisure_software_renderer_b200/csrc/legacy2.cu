// legacy2.cu -- the reference's legacy render-target demos as a device path:
// hello-render-target/hello_shadow_mapping_soft.cpp (shadow pass + PCSS soft-shadow lit pass) and hello_pbr.cpp (shadow pass +
// Cook-Torrance / IBL lit pass with motion vectors).  The arithmetic lives in legacy2_core.cuh (shared with the CPU emulator test);
// this file is the two kernels around it.  Compiled --fmad=false like legacy.cu.
//
// set-up kernel : one thread per source triangle -> 1 slot (shadow pass) or 2 slots (camera passes: near-plane clip + fan), split in
//                 three arrays so that the staging filter below reads 16 bytes per slot and the shading record only for the winner.
// raster kernel : one CTA per 16x16 pixel tile, one pixel per thread; slots are staged 256 at a time through shared memory with an
//                 ORDER-PRESERVING ballot compaction (a pixel must see its candidates in draw order, see legacy2_core.cuh) after a
//                 bounding-box filter that reproduces the job tiles' clamped integer boxes; then every pixel walks the staged
//                 records, keeps (running minimum depth, last shadeable prefix minimum), and shades once at the end.
// Like legacy.cu the staging is O(tiles x slots / 256): demo-sized scenes; the production path with binned lists is tile_raster.cu.
// shadow kernel : the shadow pass is a plain per-texel minimum, hence order-free: slots are rasterised directly, one thread per
//                 (slot, job tile, 32x32 block), large ranges by the whole warp, atomic minimum on the depth's bit pattern (round 2; the
//                 tiled kernel needed 16 384 CTAs for the 2048^2 map, each staging every slot).
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        using l2::Draw; using l2::BoxRec; using l2::PixelState; using l2::MODE_SHADOW; using l2::MODE_PBR;
        using LRaster = l2::RasterRec; using LShade = l2::ShadeRec;
        constexpr int L2_TILE = 16;
        constexpr int L2_CHUNK = 256;

        __global__ void __launch_bounds__(128) legacy2_setup_kernel(const Draw d, LRaster* __restrict__ rr, BoxRec* __restrict__ bb, LShade* __restrict__ ss)
        {
            const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
            if (t >= d.n_tris) return;
            if (d.mode == MODE_SHADOW) l2::setup_shadow(d, t, rr[t], bb[t]);
            else l2::setup_camera(d, t, rr + 2 * (size_t)t, bb + 2 * (size_t)t, ss + 2 * (size_t)t);
        }

        // Does any job tile J overlapping the pixel rectangle T = [tx0, tx1] x [ty0, ty1] test one of T's pixels for a triangle whose float box
        // is [lox, hix] x [loy, hiy]?  J tests the columns [i0, i1] = [(int)max(j0, min(j1, lox)), (int)min(j1, max(j0, hix))] (job_range,
        // legacy2_core.cuh), a non-empty interval inside J; with [a0, a1] = T's columns inside J and all operands non-negative,
        //     i0 <= a1  <=>  max(j0, min(j1, lox)) < a1 + 1  <=>  lox < a1 + 1  or  a1 == j1      (j0 <= a1 always)
        //     i1 >= a0  <=>  min(j1, max(j0, hix)) >= a0     <=>  hix >= a0     or  a0 == j0      (j1 >= a0 always)
        // -- the same decisions as evaluating job_range and intersecting, in four float compares per axis instead of two clamps and casts.
        __device__ __forceinline__ bool tile_tests_box(const Draw& d, int tx0, int tx1, int ty0, int ty1, float lox, float hix, float loy, float hiy)
        {
            for (int jy = (ty0 / d.job_h) * d.job_h; jy <= ty1; jy += d.job_h)
            {
                const int jy1 = min(jy + d.job_h, d.H) - 1, a0 = max(ty0, jy), a1 = min(ty1, jy1);
                if (!((loy < (float)(a1 + 1) || a1 == jy1) && (hiy >= (float)a0 || a0 == jy))) continue;
                for (int jx = (tx0 / d.job_w) * d.job_w; jx <= tx1; jx += d.job_w)
                {
                    const int jx1 = min(jx + d.job_w, d.W) - 1, c0 = max(tx0, jx), c1 = min(tx1, jx1);
                    if ((lox < (float)(c1 + 1) || c1 == jx1) && (hix >= (float)c0 || c0 == jx)) return true;
                }
            }
            return false;
        }

        struct Staged { LRaster r; BoxRec b; uint32_t slot; };
#ifdef SHSB_PHASE_CLOCKS
        // Debug build only (tools/l2_clocks.py): per CTA of the last raster launch, cycles in the staging / visit loop and in the final shading
        __device__ unsigned long long g_l2_clk[2][8192];
#endif

        __global__ void __launch_bounds__(L2_TILE * L2_TILE) legacy2_raster_kernel(const Draw d, const uint32_t n_slots, const LRaster* __restrict__ rr,
                                                                                    const BoxRec* __restrict__ bb, const LShade* __restrict__ ss,
                                                                                    uchar4* __restrict__ canvas, float* __restrict__ zbuf, float2* __restrict__ velocity)
        {
            // two staging buffers: a chunk needs two barriers (counts published; records published) and none at its end, because the next
            // chunk fills the OTHER buffer and the one after that is behind two more barriers
            __shared__ Staged s_tri[2][L2_CHUNK];
            __shared__ uint32_t s_warp_cnt[2][L2_TILE * L2_TILE / 32];
            const int tx0 = blockIdx.x * L2_TILE, ty0 = blockIdx.y * L2_TILE;
            const int tx1 = min(tx0 + L2_TILE, d.W) - 1, ty1 = min(ty0 + L2_TILE, d.H) - 1;
            const int px = tx0 + (int)(threadIdx.x % L2_TILE), py = ty0 + (int)(threadIdx.x / L2_TILE); // screen space, y down
            const bool inside = px < d.W && py < d.H;
            const int jx0 = (px / d.job_w) * d.job_w, jx1 = min(jx0 + d.job_w, d.W) - 1;
            const int jy0 = (py / d.job_h) * d.job_h, jy1 = min(jy0 + d.job_h, d.H) - 1;
            // the shadow map is indexed with the screen row, the camera passes' z-buffer with the FLIPPED row
            // (ZBuffer::test_and_set_depth_screen_space) like the canvas and the velocity buffer
            const size_t row = (d.mode == MODE_SHADOW) ? (size_t)py : (size_t)((d.H - 1) - py);
            const size_t at = row * (size_t)d.W + (size_t)px;
            PixelState st;
            st.best_z = inside ? zbuf[at] : 0.0f;
            st.shade_slot = 0xFFFFFFFFu;
            st.wrote = false;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const bool corner = (px == jx0 || px == jx1) && (py == jy0 || py == jy1); // a corner texel of its job tile
#ifdef SHSB_PHASE_CLOCKS
            const long long clk0 = clock64();
#endif

            int buf = 0;
            for (uint32_t base = 0; base < n_slots; base += L2_CHUNK, buf ^= 1)
            {
                const uint32_t t = base + threadIdx.x;
                bool keep = false;
                BoxRec b;
                if (t < n_slots)
                {
                    b = bb[t];
                    // does any job tile overlapping this CTA's 16x16 tile test one of its pixels for this triangle?
                    if (l2::box_valid(b)) keep = tile_tests_box(d, tx0, tx1, ty0, ty1, b.minx, b.maxx, b.miny, b.maxy);
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) s_warp_cnt[buf][warp] = __popc(ballot);
                __syncthreads();
                uint32_t before = 0, n = 0;
#pragma unroll
                for (int i = 0; i < L2_TILE * L2_TILE / 32; ++i) { const uint32_t c = s_warp_cnt[buf][i]; if (i < warp) before += c; n += c; }
                if (n == 0) continue; // CTA-uniform: nothing of this chunk reaches the tile
                if (keep)
                {
                    Staged& sd = s_tri[buf][before + __popc(ballot & ((1u << lane) - 1u))]; // order-preserving
                    sd.r = rr[t];
                    sd.b = b;
                    sd.slot = t;
                }
                __syncthreads();
                if (inside && !corner)
                    for (uint32_t i = 0; i < n; ++i) l2::pixel_visit(d.mode, s_tri[buf][i].r, s_tri[buf][i].b, s_tri[buf][i].slot, px, py, jx0, jx1, jy0, jy1, st);
                // The corner pixels of a job tile are tested for EVERY slot that lies outside the job tile in both directions (job_range clamps
                // such a box to the corner texel): one lane walking ~all slots of the draw was the kernel's critical path (224 us of a 245 us
                // launch, tools/l2_clocks.py).  Their warp probes 32 slots at a time for them and applies the few hits in slot order.
                unsigned hot = __ballot_sync(0xffffffffu, inside && corner);
                while (hot)
                {
                    const int src = __ffs(hot) - 1;
                    hot &= hot - 1u;
                    const int cpx = __shfl_sync(0xffffffffu, px, src), cpy = __shfl_sync(0xffffffffu, py, src);
                    const int cjx0 = __shfl_sync(0xffffffffu, jx0, src), cjx1 = __shfl_sync(0xffffffffu, jx1, src);
                    const int cjy0 = __shfl_sync(0xffffffffu, jy0, src), cjy1 = __shfl_sync(0xffffffffu, jy1, src);
                    PixelState cs;
                    cs.best_z = __shfl_sync(0xffffffffu, st.best_z, src);
                    cs.shade_slot = __shfl_sync(0xffffffffu, st.shade_slot, src);
                    cs.wrote = __shfl_sync(0xffffffffu, (int)st.wrote, src) != 0;
                    for (uint32_t i0 = 0; i0 < n; i0 += 32u)
                    {
                        const uint32_t i = i0 + (uint32_t)lane;
                        float z = 0.0f;
                        bool usable = false, cand = false;
                        if (i < n) cand = l2::pixel_probe(d.mode, s_tri[buf][i].r, s_tri[buf][i].b, cpx, cpy, cjx0, cjx1, cjy0, cjy1, z, usable);
                        unsigned m = __ballot_sync(0xffffffffu, cand);
                        while (m)
                        {
                            const int l = __ffs(m) - 1;
                            m &= m - 1u;
                            l2::pixel_update(d.mode, __shfl_sync(0xffffffffu, z, l), __shfl_sync(0xffffffffu, (int)usable, l) != 0, s_tri[buf][i0 + (uint32_t)l].slot, cs);
                        }
                    }
                    if (lane == src) st = cs;
                }
            }
#ifdef SHSB_PHASE_CLOCKS
            const long long clk1 = clock64();
            const uint32_t cta = blockIdx.y * gridDim.x + blockIdx.x;
            if (threadIdx.x == 0 && cta < 8192u) { g_l2_clk[0][cta] = (unsigned long long)(clk1 - clk0); g_l2_clk[1][cta] = 0ull; }
            __syncthreads();
#endif
            if (!inside || !st.wrote) return;
            zbuf[at] = st.best_z;
            if (d.mode == MODE_SHADOW || st.shade_slot == 0xFFFFFFFFu) return;
            unsigned char out[4];
            float vel[2] = {0.0f, 0.0f};
            l2::shade_pixel(d, rr[st.shade_slot], ss[st.shade_slot], px, py, out, vel);
            canvas[at] = make_uchar4(out[0], out[1], out[2], out[3]);
            if (d.mode == MODE_PBR && velocity) velocity[at] = make_float2(vel[0], vel[1]);
#ifdef SHSB_PHASE_CLOCKS
            if (cta < 8192u) atomicMax(&g_l2_clk[1][cta], (unsigned long long)(clock64() - clk1));
#endif
        }
    }

    namespace
    {
        // ---- MODE_SHADOW without tiles: the shadow pass keeps min(depth) per texel (pixel_visit's MODE_SHADOW branch), which is order-free,
        // so the slots are rasterised DIRECTLY: one thread per (slot, job tile, 32x32 block of it) computes the texels of the block that the
        // job tile tests for the slot (job_range: the slot's clamped box, or -- the demo's quirk -- one border texel of the job tile when
        // the box lies outside it); up to 4 texels are finished by that thread, larger ranges by its whole warp, one after the other.
        // Depths that pass are in [0, 1], so a signed integer minimum on their bit patterns is the float minimum (and never beats a
        // negative content).
        __device__ __forceinline__ void shadow_texel(const LRaster& r, int px, int py, int W, float* __restrict__ zbuf)
        {
            float z;
            if (!l2::shadow_texel_depth(r, px, py, z)) return;
            atomicMin(reinterpret_cast<int*>(zbuf) + ((size_t)py * (size_t)W + (size_t)px), __float_as_int(z + 0.0f)); // -0 -> +0
        }

        constexpr int SUB = 32; // a job tile is cut into SUB x SUB texel blocks: the unit of work of one thread / one warp round

        __global__ void __launch_bounds__(256) legacy2_shadow_direct_kernel(const Draw d, const uint32_t n_slots, const int jobs_x, const int jobs_y, const int subs_x, const int subs_y,
                                                                            const LRaster* __restrict__ rr, const BoxRec* __restrict__ bb, float* __restrict__ zbuf)
        {
            // item = (job tile, block of the job tile, slot), slot fastest: the 32 lanes of a warp hold 32 consecutive slots against ONE block,
            // of which few cover it -- the large ranges of a big triangle are spread over many warps instead of queueing up in one
            const uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
            const uint32_t slot = (uint32_t)(item % n_slots);
            const uint32_t cell = (uint32_t)(item / n_slots);
            const uint32_t subs = (uint32_t)(subs_x * subs_y), job = cell / subs, sub = cell % subs;
            int ix0 = 0, ix1 = -1, iy0 = 0, iy1 = -1;
            if (job < (uint32_t)(jobs_x * jobs_y))
            {
                const BoxRec b = bb[slot];
                const int jx = (int)(job % (uint32_t)jobs_x) * d.job_w, jy = (int)(job / (uint32_t)jobs_x) * d.job_h;
                const int jx1 = min(jx + d.job_w, d.W) - 1, jy1 = min(jy + d.job_h, d.H) - 1;
                const int bx = jx + (int)(sub % (uint32_t)subs_x) * SUB, by = jy + (int)(sub / (uint32_t)subs_x) * SUB;
                const int bx1 = min(bx + SUB - 1, jx1), by1 = min(by + SUB - 1, jy1);
                // nearly every (slot, job tile, block) is empty: the compare form of "job_range meets the block" (tile_tests_box) decides that
                // before the clamps and casts are evaluated
                if (l2::box_valid(b) && bx <= jx1 && by <= jy1 && (b.minx < (float)(bx1 + 1) || bx1 == jx1) && (b.maxx >= (float)bx || bx == jx) &&
                    (b.miny < (float)(by1 + 1) || by1 == jy1) && (b.maxy >= (float)by || by == jy))
                {
                    l2::job_range(b.minx, b.maxx, jx, jx1, ix0, ix1);
                    l2::job_range(b.miny, b.maxy, jy, jy1, iy0, iy1);
                    ix0 = max(ix0, bx); ix1 = min(ix1, bx1);
                    iy0 = max(iy0, by); iy1 = min(iy1, by1);
                }
            }
            const int w = ix1 - ix0 + 1, h = iy1 - iy0 + 1;
            const int area = (w > 0 && h > 0) ? w * h : 0;
            if (area > 0 && area <= 4)
            {
                const LRaster r = rr[slot];
                for (int i = 0; i < area; ++i) shadow_texel(r, ix0 + i % w, iy0 + i / w, d.W, zbuf);
            }
            unsigned big = __ballot_sync(0xffffffffu, area > 4);
            const int lane = threadIdx.x & 31;
            while (big)
            {
                const int src = __ffs(big) - 1;
                big &= big - 1u;
                const uint32_t s_slot = __shfl_sync(0xffffffffu, slot, src);
                const int x0 = __shfl_sync(0xffffffffu, ix0, src), y0 = __shfl_sync(0xffffffffu, iy0, src);
                const int ww = __shfl_sync(0xffffffffu, w, src), n = __shfl_sync(0xffffffffu, area, src);
                const LRaster r = rr[s_slot];
                for (int i = lane; i < n; i += 32) shadow_texel(r, x0 + i % ww, y0 + i / ww, d.W, zbuf);
            }
        }
    }

#ifdef SHSB_PHASE_CLOCKS
    extern "C" __attribute__((visibility("default"))) int shsb_debug_l2_clocks(unsigned long long* out16384)
    {
        cudaDeviceSynchronize();
        return cudaMemcpyFromSymbol(out16384, g_l2_clk, sizeof(unsigned long long) * 2 * 8192) == cudaSuccess ? 0 : 1;
    }
#endif

    uint32_t legacy2_slots(const l2::Draw& d) { return d.mode == l2::MODE_SHADOW ? d.n_tris : 2u * d.n_tris; }

    void launch_legacy2_draw(const l2::Draw& d, l2::RasterRec* rr, l2::BoxRec* bb, l2::ShadeRec* ss, uchar4* canvas, float* zbuf, float2* velocity,
                             cudaStream_t s, uint64_t* launches)
    {
        if (d.n_tris == 0 || d.W <= 0 || d.H <= 0) return;
        legacy2_setup_kernel<<<(d.n_tris + 127) / 128, 128, 0, s>>>(d, rr, bb, ss);
        if (d.mode == l2::MODE_SHADOW)
        {
            const int jobs_x = (d.W + d.job_w - 1) / d.job_w, jobs_y = (d.H + d.job_h - 1) / d.job_h;
            const int subs_x = (d.job_w + SUB - 1) / SUB, subs_y = (d.job_h + SUB - 1) / SUB;
            const uint64_t items = (uint64_t)d.n_tris * (uint64_t)jobs_x * (uint64_t)jobs_y * (uint64_t)subs_x * (uint64_t)subs_y;
            if (items <= (1ull << 28)) // demo-sized passes; beyond that the tiled kernel below
            {
                legacy2_shadow_direct_kernel<<<(unsigned)((items + 255) / 256), 256, 0, s>>>(d, d.n_tris, jobs_x, jobs_y, subs_x, subs_y, rr, bb, zbuf);
                if (launches) *launches += 2;
                return;
            }
        }
        const dim3 grid((unsigned)((d.W + L2_TILE - 1) / L2_TILE), (unsigned)((d.H + L2_TILE - 1) / L2_TILE));
        legacy2_raster_kernel<<<grid, L2_TILE * L2_TILE, 0, s>>>(d, legacy2_slots(d), rr, bb, ss, canvas, zbuf, velocity);
        if (launches) *launches += 2;
    }
}
