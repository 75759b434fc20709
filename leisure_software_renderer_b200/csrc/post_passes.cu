// post_passes.cu -- the full-frame passes that run right after the raster path:
//   PassMotionBlur::execute   passes/pass_motion_blur.hpp:40-184
//   PassLightShafts::execute  passes/pass_light_shafts.hpp:43-214
//   PassTemporalAAAdapter     pipeline/pass_adapters.hpp:1402-1496
//
// All three produce RGBA8, so parity is bit-exact: this file is compiled with --fmad=false and every expression
// keeps the reference's operation order (x86-64 -O3 without FMA).  They are gather / streaming kernels over
// planes that fit the 126 MB L2 at every BASELINE size up to 4K, so the bound is HBM for the first touch of each
// plane (algorithmic bytes in DESIGN.md section 4.3) and L2 gather throughput for the taps.
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        // std::lround(float): round half away from zero.  v - trunc(v) is exact in binary32, so this is exact.
        __device__ __forceinline__ int lround_f(float v)
        {
            float r = truncf(v);
            const float d = v - r;
            if (fabsf(d) >= 0.5f) r += copysignf(1.0f, v);
            return (int)r;
        }
        __device__ __forceinline__ int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

        // clamp((int)std::lround(v), 0, hi) for 0 <= hi < 2^22, without the conversion unit.  lround is monotone and
        // maps 0 -> 0 and hi -> hi, so clamping FIRST (in float) gives the same integer; the clamped value is >= 0 and
        // small, so adding 1.5 * 2^23 rounds it to an integer (ties to even) whose value sits in the low mantissa bits,
        // and the one case where ties-to-even differs from half-away-from-zero (remainder exactly +0.5) is fixed up.
        // FRND / F2I run at a quarter of the FP32 rate and four of them per tap made the march conversion-bound.
        // (|v| beyond the int range is undefined behaviour in the reference's (int) cast; here it clamps like any other value.)
        __device__ __forceinline__ int lround_clamp0(float v, float hi_f)
        {
            const float MAGIC = 12582912.0f; // 1.5 * 2^23
            const float vc = fminf(fmaxf(v, 0.0f), hi_f);
            const float t = vc + MAGIC;
            int r = __float_as_int(t) - 0x4B400000;
            if (vc - (t - MAGIC) == 0.5f) r += 1;
            return r;
        }

        // ------------------------------------------------------------------ PassMotionBlur
        // pass_motion_blur.hpp:109-161: per pixel, `samples` taps along the (scaled, clamped) velocity, each
        // rejected when its depth differs from the centre's by more than depth_reject; mean of the kept taps.
        __global__ void __launch_bounds__(256) motion_blur_kernel(PostMotionBlur a)
        {
            // t_k = k / (samples - 1) - 0.5 is the same for every pixel: one exact division per tap per CTA, not per pixel
            __shared__ float t_tab[32];
            if (threadIdx.x < a.samples) t_tab[threadIdx.x] = (float)threadIdx.x / (float)(a.samples - 1) - 0.5f;
            __syncthreads();
            const int x = blockIdx.x * 32 + (threadIdx.x & 31);
            const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
            if (x >= a.w || y >= a.h) return;
            const uchar4 centre = a.src[(size_t)y * a.src_w + x];
            const float2 mv = a.motion[(size_t)y * a.mot_w + x];
            float vx = mv.x * a.strength * a.dt_scale; // :116-117
            float vy = mv.y * a.strength * a.dt_scale;
            const float len = sqrtf(vx * vx + vy * vy); // :118 (sqrtf is IEEE-exact, no FMA in this TU)
            uchar4 out = centre;
            if (!(len < a.min_vel)) // :119-123
            {
                if (len > a.max_vel && len > 1e-6f) // :124-129
                {
                    const float s = a.max_vel / len;
                    vx *= s;
                    vy *= s;
                }
                const float centre_depth = a.depth[(size_t)y * a.mot_w + x];
                float ar = 0.0f, ag = 0.0f, ab = 0.0f, aw = 0.0f;
                const float fx = (float)x, fy = (float)y, wm1 = (float)(a.w - 1), hm1 = (float)(a.h - 1);
                for (int i = 0; i < a.samples; ++i) // :136-149
                {
                    const float t = t_tab[i];
                    const int sx = lround_clamp0(fx + vx * t, wm1);
                    const int sy = lround_clamp0(fy + vy * t, hm1);
                    const float sd = __ldg(a.depth + (size_t)sy * a.mot_w + sx);
                    if (fabsf(sd - centre_depth) > a.depth_eps) continue;
                    const uchar4 sc = __ldg(a.src + (size_t)sy * a.src_w + sx);
                    ar += (float)sc.x;
                    ag += (float)sc.y;
                    ab += (float)sc.z;
                    aw += 1.0f;
                }
                if (!(aw < 1.0f)) // :151-163
                {
                    out.x = (unsigned char)iclamp(lround_f(ar / aw), 0, 255);
                    out.y = (unsigned char)iclamp(lround_f(ag / aw), 0, 255);
                    out.z = (unsigned char)iclamp(lround_f(ab / aw), 0, 255);
                    out.w = 255;
                }
            }
            a.dst[(size_t)y * a.dst_w + x] = out;
        }

        // ------------------------------------------------------------------ PassLightShafts
        // pass_light_shafts.hpp:112-126 (luma) and :165-171 (the per-tap depth factor).  The reference multiplies
        // every tap's luma by clamp(depth,0,1) of the SAME clamped pixel, so the product is a per-pixel plane:
        // computing it once is the identical single multiply, and the march then gathers one float per tap.
        __global__ void __launch_bounds__(256) shafts_luma_kernel(const uchar4* __restrict__ src, int src_w, const float* __restrict__ depth, int depth_w,
                                                                  float* __restrict__ lumad, int w, int h)
        {
            __shared__ float unorm[256]; // (float)c / 255.0f, one exact division per value per CTA instead of three per pixel
            unorm[threadIdx.x] = (float)threadIdx.x / 255.0f;
            __syncthreads();
            const int x = blockIdx.x * 32 + (threadIdx.x & 31);
            const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
            if (x >= w || y >= h) return;
            const uchar4 c = src[(size_t)y * src_w + x];
            const float r = unorm[c.x], g = unorm[c.y], b = unorm[c.z];
            float s = 0.2126f * r + 0.7152f * g + 0.0722f * b;
            if (depth) s *= sclamp(depth[(size_t)y * depth_w + x], 0.0f, 1.0f);
            lumad[(size_t)y * w + x] = s;
        }

        // pass_light_shafts.hpp:145-192: march from the pixel towards the sun's screen position.
        constexpr int SHAFT_TABLE = 1024; // steps beyond this (reference default: 48) divide per pixel
        __global__ void __launch_bounds__(256) shafts_march_kernel(PostLightShafts a)
        {
            // t_i = i / steps is the same for every pixel: one exact division per step per CTA
            __shared__ float t_tab[SHAFT_TABLE];
            for (int i = threadIdx.x; i < min(a.steps, SHAFT_TABLE); i += 256) t_tab[i] = (float)i / (float)a.steps;
            __syncthreads();
            const int x = blockIdx.x * 32 + (threadIdx.x & 31);
            const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
            if (x >= a.w || y >= a.h) return;
            const float u = (float)x / (float)max(1, a.w - 1);
            const float v = (float)y / (float)max(1, a.h - 1);
            const float du = a.sun_u - u, dv = a.sun_v - v;
            const float wm1 = (float)(a.w - 1), hm1 = (float)(a.h - 1), fsteps = (float)a.steps;
            float illum_decay = 1.0f, accum = 0.0f;
            auto tap = [&](float t) {
                const float su = u + du * t * a.density;
                const float sv = v + dv * t * a.density;
                const int sx = lround_clamp0(su * wm1, wm1);
                const int sy = lround_clamp0(sv * hm1, hm1);
                const float s = __ldg(a.lumad + (sy * a.w + sx));
                accum += s * illum_decay * a.weight;
                illum_decay *= a.decay;
            };
            const int n_tab = min(a.steps, SHAFT_TABLE);
            for (int i = 0; i < n_tab; ++i) tap(t_tab[i]);
            for (int i = n_tab; i < a.steps; ++i) tap((float)i / fsteps); // beyond the table: divide per pixel
            const uchar4 base = a.src[(size_t)y * a.src_w + x];
            const int boost = iclamp(lround_f(accum * 80.0f), 0, 120);
            uchar4 out;
            out.x = (unsigned char)iclamp((int)base.x + boost, 0, 255);
            out.y = (unsigned char)iclamp((int)base.y + boost, 0, 255);
            out.z = (unsigned char)iclamp((int)base.z + boost / 2, 0, 255);
            out.w = 255;
            a.dst[(size_t)y * a.dst_w + x] = out;
        }

        // ------------------------------------------------------------------ PassTemporalAAAdapter
        // pass_adapters.hpp:1471-1489: out = (int)(0.88*cur + 0.12*prev + 0.5) per channel, alpha from cur; the
        // result is both the frame and the next history.  4 pixels per thread (128-bit loads / stores).
        __device__ __forceinline__ uint32_t taa_px(uint32_t cur, uint32_t prev, float keep, float blend)
        {
            uint32_t out = cur & 0xFF000000u;
#pragma unroll
            for (int s = 0; s < 24; s += 8)
            {
                const float vv = keep * (float)((cur >> s) & 0xFFu) + blend * (float)((prev >> s) & 0xFFu);
                out |= (uint32_t)iclamp((int)(vv + 0.5f), 0, 255) << s;
            }
            return out;
        }

        __global__ void __launch_bounds__(256) taa_kernel(uint4* __restrict__ ldr, uint4* __restrict__ hist, size_t n4, float keep, float blend)
        {
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
            {
                const uint4 c = ldr[i], p = hist[i];
                uint4 o;
                o.x = taa_px(c.x, p.x, keep, blend);
                o.y = taa_px(c.y, p.y, keep, blend);
                o.z = taa_px(c.z, p.z, keep, blend);
                o.w = taa_px(c.w, p.w, keep, blend);
                ldr[i] = o;
                hist[i] = o;
            }
        }

        __global__ void taa_tail_kernel(uint32_t* __restrict__ ldr, uint32_t* __restrict__ hist, size_t begin, size_t n, float keep, float blend)
        {
            const size_t i = begin + threadIdx.x;
            if (i < n)
            {
                const uint32_t o = taa_px(ldr[i], hist[i], keep, blend);
                ldr[i] = o;
                hist[i] = o;
            }
        }

        // copy_ldr (pass_motion_blur.hpp:187-199): cropped row copy between targets of different widths
        __global__ void __launch_bounds__(256) copy_ldr_kernel(const uchar4* __restrict__ src, int src_w, uchar4* __restrict__ dst, int dst_w, int w, int h)
        {
            const int x = blockIdx.x * 32 + (threadIdx.x & 31);
            const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
            if (x < w && y < h) dst[(size_t)y * dst_w + x] = src[(size_t)y * src_w + x];
        }

        inline dim3 grid2d(int w, int h) { return dim3((unsigned)((w + 31) / 32), (unsigned)((h + 7) / 8)); }
    }

    void launch_motion_blur(const PostMotionBlur& a, cudaStream_t s, uint64_t* launches)
    {
        if (a.w <= 0 || a.h <= 0) return;
        motion_blur_kernel<<<grid2d(a.w, a.h), 256, 0, s>>>(a);
        *launches += 1;
    }

    void launch_light_shafts(const PostLightShafts& a, const float* depth, int depth_w, float* lumad, cudaStream_t s, uint64_t* launches)
    {
        if (a.w <= 0 || a.h <= 0) return;
        shafts_luma_kernel<<<grid2d(a.w, a.h), 256, 0, s>>>(a.src, a.src_w, depth, depth_w, lumad, a.w, a.h);
        PostLightShafts b = a;
        b.lumad = lumad;
        shafts_march_kernel<<<grid2d(a.w, a.h), 256, 0, s>>>(b);
        *launches += 2;
    }

    void launch_taa(uchar4* ldr, uchar4* hist, size_t n_pixels, float keep, float blend, cudaStream_t s, uint64_t* launches)
    {
        const size_t n4 = n_pixels / 4;
        if (n4)
        {
            const int grid = (int)min((n4 + 255) / 256, (size_t)148 * 16);
            taa_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<uint4*>(ldr), reinterpret_cast<uint4*>(hist), n4, keep, blend);
            *launches += 1;
        }
        if (n4 * 4 < n_pixels)
        {
            taa_tail_kernel<<<1, 4, 0, s>>>(reinterpret_cast<uint32_t*>(ldr), reinterpret_cast<uint32_t*>(hist), n4 * 4, n_pixels, keep, blend);
            *launches += 1;
        }
    }

    void launch_copy_ldr(const uchar4* src, int src_w, uchar4* dst, int dst_w, int w, int h, cudaStream_t s, uint64_t* launches)
    {
        if (w <= 0 || h <= 0) return;
        copy_ldr_kernel<<<grid2d(w, h), 256, 0, s>>>(src, src_w, dst, dst_w, w, h);
        *launches += 1;
    }
}
