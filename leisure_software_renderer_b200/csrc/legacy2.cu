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

        struct Staged { LRaster r; BoxRec b; uint32_t slot; };

        __global__ void __launch_bounds__(L2_TILE * L2_TILE) legacy2_raster_kernel(const Draw d, const uint32_t n_slots, const LRaster* __restrict__ rr,
                                                                                    const BoxRec* __restrict__ bb, const LShade* __restrict__ ss,
                                                                                    uchar4* __restrict__ canvas, float* __restrict__ zbuf, float2* __restrict__ velocity)
        {
            __shared__ Staged s_tri[L2_CHUNK];
            __shared__ uint32_t s_warp_base[L2_TILE * L2_TILE / 32];
            __shared__ uint32_t s_count;
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

            for (uint32_t base = 0; base < n_slots; base += L2_CHUNK)
            {
                const uint32_t t = base + threadIdx.x;
                bool keep = false;
                BoxRec b;
                if (t < n_slots)
                {
                    b = bb[t];
                    if (l2::box_valid(b))
                    {
                        // does any job tile overlapping this CTA's 16x16 tile test one of its pixels for this triangle?
                        for (int jy = (ty0 / d.job_h) * d.job_h; jy <= ty1 && !keep; jy += d.job_h)
                            for (int jx = (tx0 / d.job_w) * d.job_w; jx <= tx1 && !keep; jx += d.job_w)
                            {
                                int ix0, ix1, iy0, iy1;
                                l2::job_range(b.minx, b.maxx, jx, min(jx + d.job_w, d.W) - 1, ix0, ix1);
                                l2::job_range(b.miny, b.maxy, jy, min(jy + d.job_h, d.H) - 1, iy0, iy1);
                                keep = max(ix0, tx0) <= min(ix1, tx1) && max(iy0, ty0) <= min(iy1, ty1);
                            }
                    }
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) s_warp_base[warp] = __popc(ballot);
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    uint32_t acc = 0;
                    for (int i = 0; i < L2_TILE * L2_TILE / 32; ++i) { const uint32_t c = s_warp_base[i]; s_warp_base[i] = acc; acc += c; }
                    s_count = acc;
                }
                __syncthreads();
                if (keep)
                {
                    Staged& s = s_tri[s_warp_base[warp] + __popc(ballot & ((1u << lane) - 1u))];
                    s.r = rr[t];
                    s.b = b;
                    s.slot = t;
                }
                __syncthreads();
                const uint32_t n = s_count;
                if (inside)
                    for (uint32_t i = 0; i < n; ++i) l2::pixel_visit(d.mode, s_tri[i].r, s_tri[i].b, s_tri[i].slot, px, py, jx0, jx1, jy0, jy1, st);
                __syncthreads(); // s_tri / s_warp_base are rewritten by the next chunk
            }
            if (!inside || !st.wrote) return;
            zbuf[at] = st.best_z;
            if (d.mode == MODE_SHADOW || st.shade_slot == 0xFFFFFFFFu) return;
            unsigned char out[4];
            float vel[2] = {0.0f, 0.0f};
            l2::shade_pixel(d, rr[st.shade_slot], ss[st.shade_slot], px, py, out, vel);
            canvas[at] = make_uchar4(out[0], out[1], out[2], out[3]);
            if (d.mode == MODE_PBR && velocity) velocity[at] = make_float2(vel[0], vel[1]);
        }
    }

    uint32_t legacy2_slots(const l2::Draw& d) { return d.mode == l2::MODE_SHADOW ? d.n_tris : 2u * d.n_tris; }

    void launch_legacy2_draw(const l2::Draw& d, l2::RasterRec* rr, l2::BoxRec* bb, l2::ShadeRec* ss, uchar4* canvas, float* zbuf, float2* velocity,
                             cudaStream_t s, uint64_t* launches)
    {
        if (d.n_tris == 0 || d.W <= 0 || d.H <= 0) return;
        legacy2_setup_kernel<<<(d.n_tris + 127) / 128, 128, 0, s>>>(d, rr, bb, ss);
        const dim3 grid((unsigned)((d.W + L2_TILE - 1) / L2_TILE), (unsigned)((d.H + L2_TILE - 1) / L2_TILE));
        legacy2_raster_kernel<<<grid, L2_TILE * L2_TILE, 0, s>>>(d, legacy2_slots(d), rr, bb, ss, canvas, zbuf, velocity);
        if (launches) *launches += 2;
    }
}
