// legacy.cu -- the reference's LEGACY tile-job rasterizer (BASELINE configs[0] as shipped) as a device
// path of its own: hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp:48-96 (Blinn-Phong vertex / fragment shaders),
// :189-242 (RendererSystem::draw_triangle_tile), hello-shs-renderer/shs_renderer.hpp:802-831 (barycentric_coordinate,
// clip_to_screen), :652-669 (ZBuffer::test_and_set_depth).  It differs from the library path (tile_raster.cu) in every decision
// that selects a pixel: y-FLIPPED screen map (1 - ndc.y), no clipping at all (triangles behind the camera project through), cull on
// `area <= 0`, dot-product barycentrics with a double-precision 1e-5 degeneracy test, depth = affine NDC z tested LESS against a
// FLT_MAX-cleared buffer, AFFINE (not perspective-correct) normal / world position, fragment colour truncated to RGBA8.
//
// Compiled --fmad=false: every expression below is the reference's, IEEE binary32, unfused, in the reference's order (GLM's scalar
// path as stated in oracle/glm_shim).  The only approximate operation is powf in the specular term (CUDA's powf is not glibc's); it
// reaches the output only through an 8-bit truncation, which is what the <= 1 LSB colour gate is for.
//
// Order independence: the demo runs one job per 80x80 screen tile and, inside a tile, objects and triangles in order with a strict
// `z < zbuf` test; the fragment shader is pure.  The serial result therefore is: per pixel, the covering triangle with the smallest z,
// the earliest one on ties, provided it beats the depth already in the buffer.  Each pixel keeps min (z, triangle index) and shades
// the winner once.  The demo's job-tile size (80x80) does show in the result in one corner case and is therefore a parameter: a job
// clamps the triangle's bounding box INTO its tile before casting it to int (:206-212), so a tile that the box does not reach still
// tests its border column / row nearest to the triangle, and an ill-conditioned sliver can pass the dot-product barycentric test
// there.  Every pixel applies exactly that clamped integer box of its own job tile before the coverage test.
//
// Two kernels: set-up (one thread per source triangle: 3 x vertex shader, screen map, cull, per-triangle barycentric constants) and
// raster (one CTA per 16x16 pixel tile, triangles staged 256 at a time through shared memory with a bounding-box filter).  The
// raster kernel is O(tiles x triangles / 256) in staging work -- adequate for the demo-sized scenes this variant exists for; the
// production path with binned tile lists is tile_raster.cu.
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        struct V3 { float x, y, z; };
        __device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
        __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
        __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
        __device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
        __device__ __forceinline__ V3 operator*(float k, V3 a) { return v3(k * a.x, k * a.y, k * a.z); }
        __device__ __forceinline__ V3 operator*(V3 a, float k) { return v3(a.x * k, a.y * k, a.z * k); }
        __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                 // glm::dot(vec3): left to right
        __device__ __forceinline__ V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }                           // glm::normalize: v * inversesqrt(dot)
        __device__ __forceinline__ float gmax(float a, float b) { return (a < b) ? b : a; }                               // glm::max
        __device__ __forceinline__ float gmin(float a, float b) { return (b < a) ? b : a; }                               // glm::min

        // mat4 * vec4(p, 1): (m0*x + m1*y) + (m2*z + m3*w), per component (glm scalar path)
        __device__ __forceinline__ float4 mul_point(const float* __restrict__ m, V3 p)
        {
            float4 r;
            r.x = (m[0] * p.x + m[4] * p.y) + (m[8] * p.z + m[12] * 1.0f);
            r.y = (m[1] * p.x + m[5] * p.y) + (m[9] * p.z + m[13] * 1.0f);
            r.z = (m[2] * p.x + m[6] * p.y) + (m[10] * p.z + m[14] * 1.0f);
            r.w = (m[3] * p.x + m[7] * p.y) + (m[11] * p.z + m[15] * 1.0f);
            return r;
        }

        __device__ __forceinline__ bool finite3(float a, float b, float c) { return isfinite(a) && isfinite(b) && isfinite(c); }

        // ---- set-up: blinn_phong_vertex_shader x 3 (:48-58), Canvas::clip_to_screen (shs_renderer.hpp:822-831), the `area <= 0`
        // cull (:219-220) and the P-independent half of Canvas::barycentric_coordinate (shs_renderer.hpp:803-820).
        __global__ void __launch_bounds__(128) legacy_setup_kernel(const LegacyDraw d, LegacyTri* __restrict__ out)
        {
            const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
            if (t >= d.n_tris) return;
            LegacyTri r;
            r.valid = 0;
            float sx[3], sy[3], sz[3];
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                uint32_t vi = 3u * t + (uint32_t)k;
                if (d.indices) vi = d.indices[vi];
                if (vi >= d.n_positions || vi >= d.n_normals) { ok = false; vi = 0; } // the demo's loader always emits both streams
                const V3 p = v3(d.positions[3 * vi], d.positions[3 * vi + 1], d.positions[3 * vi + 2]);
                const V3 n = v3(d.normals[3 * vi], d.normals[3 * vi + 1], d.normals[3 * vi + 2]);
                const float4 clip = mul_point(d.mvp, p);
                const float4 wp = mul_point(d.model, p);
                r.world[k][0] = wp.x; r.world[k][1] = wp.y; r.world[k][2] = wp.z;
                // normalize(mat3(transpose(inverse(model))) * n): the matrix is per draw (host), mat3 * vec3 sums left to right
                const float* nm = d.normal_matrix;
                const V3 nn = normalize(v3(nm[0] * n.x + nm[3] * n.y + nm[6] * n.z, nm[1] * n.x + nm[4] * n.y + nm[7] * n.z, nm[2] * n.x + nm[5] * n.y + nm[8] * n.z));
                r.normal[k][0] = nn.x; r.normal[k][1] = nn.y; r.normal[k][2] = nn.z;
                const float ndx = clip.x / clip.w, ndy = clip.y / clip.w, ndz = clip.z / clip.w;
                sx[k] = (ndx + 1.0f) * 0.5f * (float)(d.W - 1);
                sy[k] = (1.0f - ndy) * 0.5f * (float)(d.H - 1);
                sz[k] = ndz;
            }
            // Non-finite screen coordinates (w == 0, overflow) are undefined behaviour in the reference (it casts the bounding box to
            // int, :222-224); such triangles are dropped here.
            ok = ok && finite3(sx[0], sx[1], sx[2]) && finite3(sy[0], sy[1], sy[2]);
            const float area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sy[1] - sy[0]) * (sx[2] - sx[0]);
            if (ok && !(area <= 0.0f))
            {
                const float v0x = sx[1] - sx[0], v0y = sy[1] - sy[0], v1x = sx[2] - sx[0], v1y = sy[2] - sy[0];
                const float d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
                const float denom = d00 * d11 - d01 * d01;
                // `std::abs(denom) < 1e-5` compares against a DOUBLE literal
                if (!((double)fabsf(denom) < 1e-5))
                {
                    r.valid = 1;
                    r.ax = sx[0]; r.ay = sy[0];
                    r.v0x = v0x; r.v0y = v0y; r.v1x = v1x; r.v1y = v1y;
                    r.d00 = d00; r.d01 = d01; r.d11 = d11; r.denom = denom;
                    r.z[0] = sz[0]; r.z[1] = sz[1]; r.z[2] = sz[2];
                    r.minx = fminf(sx[0], fminf(sx[1], sx[2])); r.maxx = fmaxf(sx[0], fmaxf(sx[1], sx[2]));
                    r.miny = fminf(sy[0], fminf(sy[1], sy[2])); r.maxy = fmaxf(sy[0], fmaxf(sy[1], sy[2]));
                }
            }
            out[t] = r;
        }

        struct LegacyStage { float ax, ay, v0x, v0y, v1x, v1y, d00, d01, d11, denom, z0, z1, z2, minx, maxx, miny, maxy; uint32_t tri; };
        constexpr int LEGACY_TILE = 16;
        constexpr int LEGACY_CHUNK = 256;

        // Canvas::barycentric_coordinate for sample point P = (px + 0.5, py + 0.5), shs_renderer.hpp:809-819
        __device__ __forceinline__ bool legacy_bary(const LegacyStage& s, float Px, float Py, float& u, float& v, float& w)
        {
            const float v2x = Px - s.ax, v2y = Py - s.ay;
            const float d20 = v2x * s.v0x + v2y * s.v0y;
            const float d21 = v2x * s.v1x + v2y * s.v1y;
            v = (s.d11 * d20 - s.d01 * d21) / s.denom;
            w = (s.d00 * d21 - s.d01 * d20) / s.denom;
            u = 1.0f - v - w;
            return !(u < 0.0f || v < 0.0f || w < 0.0f);
        }

        // The pixel range a job tile [t0, t1] tests along one axis for a triangle spanning [lo, hi] (:206-212, :223-224):
        // bboxmin = max(tile_min, min(tile_max, lo)), bboxmax = min(tile_max, max(tile_min, hi)), both cast to int.
        __device__ __forceinline__ void legacy_job_range(float lo, float hi, int t0, int t1, int& i0, int& i1)
        {
            const float f0 = (float)t0, f1 = (float)t1;
            i0 = (int)gmax(f0, gmin(f1, lo));
            i1 = (int)gmin(f1, gmax(f0, hi));
        }

        // Does any job tile J overlapping the pixel rectangle T = [tx0, tx1] x [ty0, ty1] test one of T's pixels for a triangle whose float box
        // is [lox, hix] x [loy, hiy]?  J tests the columns [i0, i1] = [(int)max(j0, min(j1, lox)), (int)min(j1, max(j0, hix))]
        // (legacy_job_range), a non-empty interval inside J; with [a0, a1] = T's columns inside J and all operands non-negative,
        //     i0 <= a1  <=>  max(j0, min(j1, lox)) < a1 + 1  <=>  lox < a1 + 1  or  a1 == j1      (j0 <= a1 always)
        //     i1 >= a0  <=>  min(j1, max(j0, hix)) >= a0     <=>  hix >= a0     or  a0 == j0      (j1 >= a0 always)
        // -- the same decisions as evaluating legacy_job_range and intersecting, in four float compares per axis.
        __device__ __forceinline__ bool legacy_tile_tests_box(const LegacyDraw& d, int tx0, int tx1, int ty0, int ty1, float lox, float hix, float loy, float hiy)
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

        // one triangle at one pixel of job tile [jx0, jx1] x [jy0, jy1]: the job's pixel loops (:223-224), the inside test (:226-227), the depth (:230)
        __device__ __forceinline__ bool legacy_probe(const LegacyStage& s, int px, int py, int jx0, int jx1, int jy0, int jy1, float& z)
        {
            // px in [ix0, ix1] of legacy_job_range(minx, maxx, jx0, jx1), in the compare form derived at legacy_tile_tests_box
            if (!((s.minx < (float)(px + 1) || px == jx1) && (s.maxx >= (float)px || px == jx0))) return false;
            if (!((s.miny < (float)(py + 1) || py == jy1) && (s.maxy >= (float)py || py == jy0))) return false;
            float u, v, w;
            if (!legacy_bary(s, (float)px + 0.5f, (float)py + 0.5f, u, v, w)) return false;
            z = u * s.z0 + v * s.z1 + w * s.z2;
            return true;
        }

        // ---- raster + shade: one CTA per 16x16 tile, one pixel per thread (screen space, y down)
        __global__ void __launch_bounds__(LEGACY_TILE * LEGACY_TILE) legacy_raster_kernel(const LegacyDraw d, const LegacyTri* __restrict__ tris,
                                                                                            uchar4* __restrict__ canvas, float* __restrict__ zbuf)
        {
            // two staging buffers: a chunk needs two barriers (counts published; records published) and none at its end, because the next
            // chunk fills the OTHER buffer and the one after that is behind two more barriers
            __shared__ LegacyStage s_tri[2][LEGACY_CHUNK];
            __shared__ uint32_t s_warp_cnt[2][LEGACY_TILE * LEGACY_TILE / 32];
            const int tx0 = blockIdx.x * LEGACY_TILE, ty0 = blockIdx.y * LEGACY_TILE;
            const int tx1 = min(tx0 + LEGACY_TILE, d.W) - 1, ty1 = min(ty0 + LEGACY_TILE, d.H) - 1; // last pixel of this CTA's tile
            const int px = tx0 + (int)(threadIdx.x % LEGACY_TILE), py = ty0 + (int)(threadIdx.x / LEGACY_TILE);
            const bool inside = px < d.W && py < d.H;
            const float Px = (float)px + 0.5f, Py = (float)py + 0.5f;
            // the job tile (the demo's 80x80 unit of work) this pixel belongs to
            const int jx0 = (px / d.job_w) * d.job_w, jx1 = min(jx0 + d.job_w, d.W) - 1;
            const int jy0 = (py / d.job_h) * d.job_h, jy1 = min(jy0 + d.job_h, d.H) - 1;
            float best_z = inside ? zbuf[(size_t)py * d.W + px] : 0.0f;
            uint32_t best_tri = 0xFFFFFFFFu;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const bool corner = (px == jx0 || px == jx1) && (py == jy0 || py == jy1); // a corner texel of its job tile

            int buf = 0;
            for (uint32_t base = 0; base < d.n_tris; base += LEGACY_CHUNK, buf ^= 1)
            {
                const uint32_t t = base + threadIdx.x;
                bool keep = false;
                float minx = 0.0f, maxx = 0.0f, miny = 0.0f, maxy = 0.0f;
                if (t < d.n_tris && tris[t].valid)
                {
                    // the filter reads the box only (16 of the record's 156 bytes); kept triangles fetch the rest below
                    minx = tris[t].minx; maxx = tris[t].maxx; miny = tris[t].miny; maxy = tris[t].maxy;
                    // does any job tile overlapping this CTA's 16x16 tile test a pixel of the 16x16 tile for this triangle?
                    keep = legacy_tile_tests_box(d, tx0, tx1, ty0, ty1, minx, maxx, miny, maxy);
                }
                // order-preserving compaction (ties are broken by triangle index, candidates are visited in draw order)
                const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) s_warp_cnt[buf][warp] = __popc(ballot);
                __syncthreads();
                uint32_t before = 0, n = 0;
#pragma unroll
                for (int i = 0; i < LEGACY_TILE * LEGACY_TILE / 32; ++i) { const uint32_t c = s_warp_cnt[buf][i]; if (i < warp) before += c; n += c; }
                if (n == 0) continue; // CTA-uniform: nothing of this chunk reaches the tile
                if (keep)
                {
                    const LegacyTri& r = tris[t];
                    LegacyStage& s = s_tri[buf][before + __popc(ballot & ((1u << lane) - 1u))];
                    s.ax = r.ax; s.ay = r.ay; s.v0x = r.v0x; s.v0y = r.v0y; s.v1x = r.v1x; s.v1y = r.v1y;
                    s.d00 = r.d00; s.d01 = r.d01; s.d11 = r.d11; s.denom = r.denom;
                    s.z0 = r.z[0]; s.z1 = r.z[1]; s.z2 = r.z[2];
                    s.minx = minx; s.maxx = maxx; s.miny = miny; s.maxy = maxy; s.tri = t;
                }
                __syncthreads();
                if (inside && !corner)
                {
                    for (uint32_t i = 0; i < n; ++i)
                    {
                        const LegacyStage& s = s_tri[buf][i];
                        float z;
                        if (legacy_probe(s, px, py, jx0, jx1, jy0, jy1, z) && z < best_z) { best_z = z; best_tri = s.tri; } // ZBuffer::test_and_set_depth: strict LESS, in order
                    }
                }
                // The corner pixels of a job tile are tested for EVERY triangle that lies outside the job tile in both directions
                // (legacy_job_range clamps such a box to the corner texel): one lane walking ~all triangles of the draw is the kernel's
                // critical path.  Their warp probes 32 triangles at a time for them and applies the few hits in draw order.
                unsigned hot = __ballot_sync(0xffffffffu, inside && corner);
                while (hot)
                {
                    const int src = __ffs(hot) - 1;
                    hot &= hot - 1u;
                    const int cpx = __shfl_sync(0xffffffffu, px, src), cpy = __shfl_sync(0xffffffffu, py, src);
                    const int cjx0 = __shfl_sync(0xffffffffu, jx0, src), cjx1 = __shfl_sync(0xffffffffu, jx1, src);
                    const int cjy0 = __shfl_sync(0xffffffffu, jy0, src), cjy1 = __shfl_sync(0xffffffffu, jy1, src);
                    float cz = __shfl_sync(0xffffffffu, best_z, src);
                    uint32_t ctri = __shfl_sync(0xffffffffu, best_tri, src);
                    for (uint32_t i0 = 0; i0 < n; i0 += 32u)
                    {
                        const uint32_t i = i0 + (uint32_t)lane;
                        float z = 0.0f;
                        const bool cand = i < n && legacy_probe(s_tri[buf][i], cpx, cpy, cjx0, cjx1, cjy0, cjy1, z);
                        unsigned m = __ballot_sync(0xffffffffu, cand);
                        while (m)
                        {
                            const int l = __ffs(m) - 1;
                            m &= m - 1u;
                            const float zl = __shfl_sync(0xffffffffu, z, l);
                            if (zl < cz) { cz = zl; ctri = s_tri[buf][i0 + (uint32_t)l].tri; }
                        }
                    }
                    if (lane == src) { best_z = cz; best_tri = ctri; }
                }
            }
            if (!inside || best_tri == 0xFFFFFFFFu) return;
            zbuf[(size_t)py * d.W + px] = best_z;

            // ---- the winner, once: re-derive its barycentrics with the same operations, interpolate AFFINELY (:234-236), shade (:64-96)
            const LegacyTri r = tris[best_tri];
            LegacyStage s;
            s.ax = r.ax; s.ay = r.ay; s.v0x = r.v0x; s.v0y = r.v0y; s.v1x = r.v1x; s.v1y = r.v1y;
            s.d00 = r.d00; s.d01 = r.d01; s.d11 = r.d11; s.denom = r.denom;
            float u, v, w;
            legacy_bary(s, Px, Py, u, v, w);
            const V3 n0 = v3(r.normal[0][0], r.normal[0][1], r.normal[0][2]), n1 = v3(r.normal[1][0], r.normal[1][1], r.normal[1][2]),
                     n2 = v3(r.normal[2][0], r.normal[2][1], r.normal[2][2]);
            const V3 w0 = v3(r.world[0][0], r.world[0][1], r.world[0][2]), w1 = v3(r.world[1][0], r.world[1][1], r.world[1][2]),
                     w2 = v3(r.world[2][0], r.world[2][1], r.world[2][2]);
            const V3 in_normal = normalize((u * n0 + v * n1) + w * n2);
            const V3 world_pos = (u * w0 + v * w1) + w * w2;

            const V3 norm = normalize(in_normal);
            const V3 light_dir = normalize(v3(-d.light_dir[0], -d.light_dir[1], -d.light_dir[2]));
            const V3 view_dir = normalize(v3(d.camera_pos[0], d.camera_pos[1], d.camera_pos[2]) - world_pos);
            const float ambient = 0.15f * 1.0f;
            const float diff = gmax(dot(norm, light_dir), 0.0f);
            const float diffuse = diff * 1.0f;
            const V3 halfway = normalize(light_dir + view_dir);
            const float spec = powf(gmax(dot(norm, halfway), 0.0f), 64.0f);
            const float specular = (0.5f * spec) * 1.0f;
            const V3 object_color = v3((float)d.color[0] / 255.0f, (float)d.color[1] / 255.0f, (float)d.color[2] / 255.0f);
            const float lit = (ambient + diffuse) + specular;
            V3 result = v3(lit, lit, lit) * object_color;
            result = v3(gmin(gmax(result.x, 0.0f), 1.0f), gmin(gmax(result.y, 0.0f), 1.0f), gmin(gmax(result.z, 0.0f), 1.0f));
            // (uint8_t)(c * 255): truncation; Canvas::draw_pixel_screen_space flips y (shs_renderer.hpp:792-796)
            const uchar4 c = make_uchar4((unsigned char)(result.x * 255.0f), (unsigned char)(result.y * 255.0f), (unsigned char)(result.z * 255.0f), 255);
            canvas[(size_t)((d.H - 1) - py) * d.W + px] = c;
        }
    }

    void launch_legacy_draw(const LegacyDraw& d, LegacyTri* tris, uchar4* canvas, float* zbuf, cudaStream_t s, uint64_t* launches)
    {
        if (d.n_tris == 0 || d.W <= 0 || d.H <= 0) return;
        legacy_setup_kernel<<<(d.n_tris + 127) / 128, 128, 0, s>>>(d, tris);
        const dim3 grid((unsigned)((d.W + LEGACY_TILE - 1) / LEGACY_TILE), (unsigned)((d.H + LEGACY_TILE - 1) / LEGACY_TILE));
        legacy_raster_kernel<<<grid, LEGACY_TILE * LEGACY_TILE, 0, s>>>(d, tris, canvas, zbuf);
        if (launches) *launches += 2;
    }
}
