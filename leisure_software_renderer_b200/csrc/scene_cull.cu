// scene_cull.cu -- scene-level culling steps upstream of draw submission: objects against the camera
// frustum (cull_vs_frustum, geometry/jolt_culling.hpp:279-306: classes, the ordered visible list, the four counters) and per-object
// light selection (collect_object_lights, lighting/light_runtime.hpp:592-616).  Arithmetic in scene_cull_core.cuh; --fmad=false.
// The software-occlusion pass of culling_software.hpp is serial in the ORDER of objects (each object's test reads the depth the
// previous ones wrote); one persistent CTA walks that order and spreads the work of each object over its threads.
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        struct Planes { float p[24]; };

        __global__ void __launch_bounds__(256) classify_objects_kernel(const float* __restrict__ bounds10, uint32_t n, const Planes planes, uint8_t* __restrict__ classes)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n) return;
            float b[10];
            for (int k = 0; k < 10; ++k) b[k] = bounds10[(size_t)i * 10 + k];
            classes[i] = (uint8_t)sc::classify_object(b, planes.p);
        }

        // one CTA: the visible list keeps object order (CullResult::visible_indices is filled by a serial loop), so the compaction is
        // an ordered ballot scan over 1024 objects per round; counts[0..3] = tested, outside, intersecting, inside
        __global__ void __launch_bounds__(1024) compact_visible_kernel(const uint8_t* __restrict__ classes, uint32_t n, uint32_t* __restrict__ visible, uint32_t* __restrict__ counts)
        {
            __shared__ uint32_t s_warp[32];
            __shared__ uint32_t s_base;
            __shared__ uint32_t s_cls[3];
            if (threadIdx.x == 0) { s_base = 0; s_cls[0] = s_cls[1] = s_cls[2] = 0; }
            __syncthreads();
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (uint32_t base = 0; base < n; base += 1024)
            {
                const uint32_t i = base + threadIdx.x;
                const int c = (i < n) ? (int)classes[i] : -1;
                const bool vis = c == sc::INTERSECTING || c == sc::INSIDE;
                const unsigned ballot = __ballot_sync(0xffffffffu, vis);
                if (lane == 0) s_warp[warp] = __popc(ballot);
                if (c >= 0) atomicAdd(&s_cls[c], 1u);
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    uint32_t acc = s_base;
                    for (int w = 0; w < 32; ++w) { const uint32_t k = s_warp[w]; s_warp[w] = acc; acc += k; }
                    s_base = acc;
                }
                __syncthreads();
                if (vis) visible[s_warp[warp] + __popc(ballot & ((1u << lane) - 1u))] = i;
                __syncthreads();
            }
            if (threadIdx.x == 0) { counts[0] = n; counts[1] = s_cls[0]; counts[2] = s_cls[1]; counts[3] = s_cls[2]; counts[4] = s_base; }
        }

        // ---- one WARP per object.  The selection's slot layout depends on the order in which the affecting lights arrive
        // (add_light_candidate, light_runtime.hpp:263-289), so that part stays serial; what is order-free -- reading a light's
        // record, the light-affects-object test, the squared distance -- is done 32 candidates at a time, one per lane, and the
        // survivors are fed to the (redundantly held, register-resident) selection in ascending lane order = the reference's visit order.
        __device__ __forceinline__ void warp_consider(sc::Selection& sel, bool live, uint32_t li, const float* __restrict__ records, const float* box, float cx, float cy, float cz, int mode)
        {
            float d2 = 0.0f;
            const bool hit = live && sc::light_candidate(records + (size_t)li * sc::LIGHT_RECORD_FLOATS, box, cx, cy, cz, mode, d2);
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m)
            {
                const int src = __ffs(m) - 1;
                m &= m - 1u;
                sc::selection_insert(sel, __shfl_sync(0xffffffffu, li, src), __shfl_sync(0xffffffffu, d2, src));
            }
        }

        __device__ __forceinline__ void warp_store_selection(const sc::Selection& sel, uint32_t o, uint32_t lane, uint32_t* __restrict__ out_counts, uint32_t* __restrict__ out_idx,
                                                             float* __restrict__ out_d2)
        {
            if (lane == 0) out_counts[o] = sel.count;
#pragma unroll
            for (uint32_t k = 0; k < sc::LIGHT_SELECTION_CAPACITY; ++k)
                if (lane == k) { out_idx[(size_t)o * sc::LIGHT_SELECTION_CAPACITY + k] = sel.idx[k]; out_d2[(size_t)o * sc::LIGHT_SELECTION_CAPACITY + k] = sel.d2[k]; }
        }

        __global__ void __launch_bounds__(128) collect_object_lights_kernel(const float* __restrict__ boxes6, uint32_t n_objects, const uint32_t* __restrict__ visible,
                                                                             uint32_t n_visible, const float* __restrict__ records, uint32_t n_lights, int mode,
                                                                             uint32_t* __restrict__ out_counts, uint32_t* __restrict__ out_idx, float* __restrict__ out_d2)
        {
            const uint32_t o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
            if (o >= n_objects) return; // whole warps leave together
            float box[6];
            for (int k = 0; k < 6; ++k) box[k] = boxes6[(size_t)o * 6 + k];
            const float cx = 0.5f * (box[0] + box[3]), cy = 0.5f * (box[1] + box[4]), cz = 0.5f * (box[2] + box[5]);
            sc::Selection sel;
            sc::selection_clear(sel);
            for (uint32_t v0 = 0; v0 < n_visible; v0 += 32u)
            {
                const uint32_t v = v0 + lane;
                const uint32_t li = v < n_visible ? visible[v] : 0xFFFFFFFFu;
                warp_consider(sel, li < n_lights, li, records, box, cx, cy, cz, mode); // entries >= n_lights are skipped (:604-606)
            }
            warp_store_selection(sel, o, lane, out_counts, out_idx, out_d2);
        }

        __global__ void __launch_bounds__(256) scene_range_init_kernel(uint32_t tiles, float z_near, float z_far, uint32_t* __restrict__ kmin, uint32_t* __restrict__ kmax, uint32_t* __restrict__ has)
        {
            const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
            if (t >= tiles) return;
            kmin[t] = sc::depth_key(z_far);
            kmax[t] = sc::depth_key(z_near);
            has[t] = 0u;
        }

        struct ViewMats { float view[16], view_proj[16]; };

        // one thread per visible object: project, then fold its depth range into every tile of its rectangle
        __global__ void __launch_bounds__(128) scene_range_scatter_kernel(const float* __restrict__ boxes6, uint32_t n_objects, const uint32_t* __restrict__ visible, uint32_t n_visible,
                                                                           const ViewMats m, float z_near, float z_far, uint32_t tiles_x, uint32_t tiles_y,
                                                                           uint32_t* __restrict__ kmin, uint32_t* __restrict__ kmax, uint32_t* __restrict__ has)
        {
            const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
            if (v >= n_visible) return;
            const uint32_t o = visible[v];
            if (o >= n_objects) return;
            float box[6];
            for (int k = 0; k < 6; ++k) box[k] = boxes6[(size_t)o * 6 + k];
            sc::TileRect r;
            if (!sc::project_object(box, m.view, m.view_proj, z_near, z_far, tiles_x, tiles_y, r)) return;
            const uint32_t lo = sc::depth_key(r.min_depth), hi = sc::depth_key(r.max_depth);
            for (uint32_t ty = r.ty0; ty <= r.ty1; ++ty)
                for (uint32_t tx = r.tx0; tx <= r.tx1; ++tx)
                {
                    const uint32_t t = ty * tiles_x + tx;
                    atomicMin(&kmin[t], lo);
                    atomicMax(&kmax[t], hi);
                    has[t] = 1u;
                }
        }

        __global__ void __launch_bounds__(256) scene_range_finish_kernel(uint32_t tiles, float z_near, float z_far, const uint32_t* __restrict__ kmin, const uint32_t* __restrict__ kmax,
                                                                          const uint32_t* __restrict__ has, float* __restrict__ out_min, float* __restrict__ out_max)
        {
            const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
            if (t >= tiles) return;
            sc::finish_tile(has[t], sc::key_depth(kmin[t]), sc::key_depth(kmax[t]), z_near, z_far, out_min[t], out_max[t]);
        }
    }

    void launch_cull_objects(const float* bounds10, uint32_t n, const float planes24[24], uint8_t* classes, uint32_t* visible, uint32_t* counts5, cudaStream_t s, uint64_t* launches)
    {
        Planes pl;
        for (int i = 0; i < 24; ++i) pl.p[i] = planes24[i];
        if (n) classify_objects_kernel<<<(n + 255) / 256, 256, 0, s>>>(bounds10, n, pl, classes);
        compact_visible_kernel<<<1, 1024, 0, s>>>(classes, n, visible, counts5);
        if (launches) *launches += n ? 2 : 1;
    }

    void launch_collect_object_lights(const float* boxes6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const float* records, uint32_t n_lights, int mode,
                                      uint32_t* out_counts, uint32_t* out_idx, float* out_d2, cudaStream_t s, uint64_t* launches)
    {
        if (n_objects == 0) return;
        collect_object_lights_kernel<<<(n_objects + 3) / 4, 128, 0, s>>>(boxes6, n_objects, visible, n_visible, records, n_lights, mode, out_counts, out_idx, out_d2);
        if (launches) *launches += 1;
    }
}

namespace shsb
{
    // build_tile_view_depth_range_from_scene; scratch3: 3 x tiles uint32 (min keys, max keys, has-depth flags)
    void launch_scene_tile_depth_range(const float* boxes6, uint32_t n_objects, const uint32_t* visible, uint32_t n_visible, const float view[16], const float view_proj[16],
                                       float z_near, float z_far, uint32_t tiles_x, uint32_t tiles_y, uint32_t* scratch3, float* out_min, float* out_max, cudaStream_t s, uint64_t* launches)
    {
        const uint32_t tiles = tiles_x * tiles_y;
        if (tiles == 0) return;
        uint32_t *kmin = scratch3, *kmax = scratch3 + tiles, *has = scratch3 + 2 * (size_t)tiles;
        ViewMats m;
        for (int i = 0; i < 16; ++i) { m.view[i] = view[i]; m.view_proj[i] = view_proj[i]; }
        scene_range_init_kernel<<<(tiles + 255) / 256, 256, 0, s>>>(tiles, z_near, z_far, kmin, kmax, has);
        if (n_visible) scene_range_scatter_kernel<<<(n_visible + 127) / 128, 128, 0, s>>>(boxes6, n_objects, visible, n_visible, m, z_near, z_far, tiles_x, tiles_y, kmin, kmax, has);
        scene_range_finish_kernel<<<(tiles + 255) / 256, 256, 0, s>>>(tiles, z_near, z_far, kmin, kmax, has, out_min, out_max);
        if (launches) *launches += n_visible ? 3 : 2;
    }
}

namespace shsb
{
    namespace
    {
        struct SelectArgs { float view[16], view_proj[16]; sc::BinGrid grid; };

        // one WARP per object (see collect_object_lights_kernel).  A bin lists a light at most once, so the 32 entries a warp reads from one
        // bin are distinct lights: their "seen" bits are tested together, set together, and the fresh ones considered in entry order.
        // seen: ceil(n_lights / 32) words per object -- in shared memory when the CTA's four objects fit (the launcher decides), else in the
        // zeroed global scratch (volatile reads: the bits are set by other lanes' atomics).
        __global__ void __launch_bounds__(128) select_object_lights_kernel(const float* __restrict__ boxes6, uint32_t n_objects, const SelectArgs a, const uint32_t* __restrict__ bin_counts,
                                                                            const uint32_t* __restrict__ bin_indices, const float* __restrict__ records, uint32_t n_lights, int mode,
                                                                            uint32_t* __restrict__ seen_global, uint32_t words_per_object, int seen_in_shared,
                                                                            uint32_t* __restrict__ out_counts, uint32_t* __restrict__ out_idx, float* __restrict__ out_d2,
                                                                            uint32_t* __restrict__ out_candidates)
        {
            extern __shared__ uint32_t s_seen[];
            const uint32_t o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
            if (o >= n_objects) return; // whole warps leave together
            float box[6];
            for (int k = 0; k < 6; ++k) box[k] = boxes6[(size_t)o * 6 + k];
            const float cx = 0.5f * (box[0] + box[3]), cy = 0.5f * (box[1] + box[4]), cz = 0.5f * (box[2] + box[5]);
            sc::Selection sel;
            sc::selection_clear(sel);
            sc::TileRect r;
            uint32_t tz0, tz1, n_candidates = 0;
            if (!sc::object_bin_range(box, a.view, a.view_proj, a.grid, r, tz0, tz1))
            {
                // fallback_light_scene_candidates: every visible light, in order
                for (uint32_t l0 = 0; l0 < n_lights; l0 += 32u) warp_consider(sel, l0 + lane < n_lights, l0 + lane, records, box, cx, cy, cz, mode);
                n_candidates = n_lights;
            }
            else
            {
                volatile uint32_t* seen;
                if (seen_in_shared)
                {
                    seen = s_seen + (size_t)(threadIdx.x >> 5) * words_per_object;
                    for (uint32_t w = lane; w < words_per_object; w += 32u) seen[w] = 0u;
                    __syncwarp();
                }
                else seen = seen_global + (size_t)o * words_per_object;
                const sc::BinGrid& g = a.grid;
                for (uint32_t tz = tz0; tz <= tz1; ++tz)
                    for (uint32_t ty = r.ty0; ty <= r.ty1; ++ty)
                        for (uint32_t tx = r.tx0; tx <= r.tx1; ++tx)
                        {
                            const uint32_t bin = tz * (g.bins_x * g.bins_y) + ty * g.bins_x + tx;
                            const uint32_t n = min(bin_counts[bin], g.max_per_bin);
                            for (uint32_t k0 = 0; k0 < n; k0 += 32u)
                            {
                                const uint32_t k = k0 + lane;
                                const uint32_t li = k < n ? bin_indices[(size_t)bin * g.max_per_bin + k] : 0xFFFFFFFFu;
                                const uint32_t bit = 1u << (li & 31u);
                                const bool fresh = li < n_lights && !(seen[li >> 5] & bit);
                                __syncwarp(); // every lane has read the bits before any is set
                                if (fresh) atomicOr(const_cast<uint32_t*>(&seen[li >> 5]), bit);
                                __syncwarp();
                                n_candidates += (uint32_t)__popc(__ballot_sync(0xffffffffu, fresh));
                                warp_consider(sel, fresh, li, records, box, cx, cy, cz, mode);
                            }
                        }
            }
            if (lane == 0) out_candidates[o] = n_candidates;
            warp_store_selection(sel, o, lane, out_counts, out_idx, out_d2);
        }
    }

    void launch_select_object_lights(const float* boxes6, uint32_t n_objects, const float view[16], const float view_proj[16], const sc::BinGrid& grid, const uint32_t* bin_counts,
                                     const uint32_t* bin_indices, const float* records, uint32_t n_lights, int mode, uint32_t* seen, uint32_t words_per_object, uint32_t* out_counts,
                                     uint32_t* out_idx, float* out_d2, uint32_t* out_candidates, cudaStream_t s, uint64_t* launches)
    {
        if (n_objects == 0) return;
        SelectArgs a;
        for (int i = 0; i < 16; ++i) { a.view[i] = view[i]; a.view_proj[i] = view_proj[i]; }
        a.grid = grid;
        const size_t shared = (size_t)4 * words_per_object * sizeof(uint32_t); // 4 warps = 4 objects per CTA
        const int in_shared = shared <= 40 * 1024;                             // up to 81 920 lights; beyond that the zeroed global scratch
        if (!in_shared) cudaMemsetAsync(seen, 0, (size_t)n_objects * words_per_object * 4, s);
        select_object_lights_kernel<<<(n_objects + 3) / 4, 128, in_shared ? shared : 0, s>>>(boxes6, n_objects, a, bin_counts, bin_indices, records, n_lights, mode, seen, words_per_object,
                                                                                            in_shared, out_counts, out_idx, out_d2, out_candidates);
        if (launches) *launches += 1;
    }
    namespace
    {
        struct OccParams
        {
            const float* boxes6; uint32_t n_objects;
            const uint32_t* sorted; uint32_t n_sorted;           // frustum-visible indices in the host's front-to-back order
            const uint32_t* object_mesh;                         // per object: index into mesh_table, 0xFFFFFFFF = no occluder mesh
            const float* object_models;                          // per object: model matrix (16 floats)
            const uint32_t* mesh_table;                          // per mesh: first index, index count, base vertex
            uint32_t n_meshes;
            const float* vertices; uint32_t n_vertices;          // local-space positions of all occluder meshes
            const uint32_t* indices; uint32_t n_indices;
            float view_proj[16];
            int width, height;
            float epsilon;
            uint32_t* depth_bits;                                // width * height, pre-filled with 1.0f
            uint8_t* occluded;                                   // per object
            uint32_t* visible;                                   // output list (sorted order)
            uint32_t* counts;                                    // [0] visible, [1] occluded
        };

        // project_aabb_to_screen_rect for every entry of the sorted list, in parallel: a rectangle depends on the object and the camera
        // only, not on the depth buffer
        __global__ void __launch_bounds__(256) occlusion_rects_kernel(const OccParams p, sc::OccRect* __restrict__ rects)
        {
            const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
            if (k >= p.n_sorted) return;
            const uint32_t idx = p.sorted[k];
            if (idx < p.n_objects) rects[k] = sc::occ_project_rect(p.boxes6 + (size_t)idx * 6, p.view_proj, p.width, p.height);
        }

        // run_software_occlusion_pass, geometry/culling_software.hpp:292-322, on ONE CTA of 32 warps.  The order of objects is serial,
        // but the depth buffer only ever gets NEARER (minimum writes), so "occluded" is monotone: an object hidden by the buffer as it
        // is now is still hidden when its turn comes.  Each round therefore tests the next 32 entries at once (one warp per entry,
        // lanes over the rectangle's texels): every entry before the first one that shows is occluded for good -- occluded objects
        // rasterise nothing, so the buffer that first showing object sees is the current one, and it is visible for good too.  It
        // rasterises its occluder mesh (rasterize_mesh_depth_transformed: a warp per triangle, lanes over the bbox texels, atomic
        // minimum on the depth's bit pattern -- depths are in [0, 1]), and the next round starts behind it.  A round thus settles one
        // visible object plus all occluded ones in front of it; the results are those of the serial loop.
        __global__ void __launch_bounds__(1024) software_occlusion_kernel(const OccParams p, const sc::OccRect* __restrict__ rects, uint8_t* __restrict__ settled)
        {
            __shared__ int s_state[32];       // per window entry: 0 occluded, 1 shows, 2 not an object (stale index), 3 large rectangle, not tested yet
            __shared__ uint32_t s_nvis, s_nocc;
            const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            constexpr int WARP_RECT_MAX = 2048; // texels a single warp scans; larger rectangles are scanned by the whole CTA
            if (tid == 0) { s_nvis = 0; s_nocc = 0; }
            uint32_t pos = 0;
            while (pos < p.n_sorted)
            {
                const uint32_t e = pos + (uint32_t)warp;
                int state = 2;
                uint32_t idx = 0xFFFFFFFFu;
                if (e < p.n_sorted)
                {
                    idx = p.sorted[e];
                    if (idx < p.n_objects)
                    {
                        const sc::OccRect r = rects[e];
                        const int rw = r.x_max - r.x_min + 1, total = r.valid ? rw * (r.y_max - r.y_min + 1) : 0;
                        if (settled[e]) state = 0;               // found hidden in an earlier round: hidden for good (the buffer only gets nearer)
                        else if (!r.valid) state = 1;            // an invalid rectangle is never occluded (:208)
                        else if (total > WARP_RECT_MAX) state = 3;
                        else
                        {
                            bool shows = false;
                            for (int base = 0; base < total && !shows; base += 32)
                            {
                                const int i = base + lane;
                                bool mine = false;
                                if (i < total)
                                    mine = sc::occ_texel_shows(r.z_near, __uint_as_float(__ldcg(p.depth_bits + (size_t)(r.y_min + i / rw) * p.width + (r.x_min + i % rw))), p.epsilon);
                                shows = __any_sync(0xffffffffu, mine);
                            }
                            state = shows ? 1 : 0;
                            if (!shows && lane == 0) settled[e] = 1;
                        }
                    }
                }
                if (lane == 0) s_state[warp] = state;
                __syncthreads();
                const uint32_t window = min(32u, p.n_sorted - pos);
                int first = 32;
                for (int w = 0; w < (int)window; ++w) // CTA-uniform walk in order; large rectangles are scanned by all threads, only while no earlier entry shows
                {
                    const int st = s_state[w];
                    if (st == 1) { first = w; break; }
                    if (st != 3) continue;
                    const sc::OccRect r = rects[pos + (uint32_t)w];
                    const int rw = r.x_max - r.x_min + 1, total = rw * (r.y_max - r.y_min + 1);
                    bool mine = false;
                    for (int i = tid; i < total && !mine; i += 1024)
                        mine = sc::occ_texel_shows(r.z_near, __uint_as_float(__ldcg(p.depth_bits + (size_t)(r.y_min + i / rw) * p.width + (r.x_min + i % rw))), p.epsilon);
                    const int shows = __syncthreads_or(mine ? 1 : 0);
                    if (shows) { first = w; break; }
                    if (tid == 0) { s_state[w] = 0; settled[pos + (uint32_t)w] = 1; }
                }
                __syncthreads();
                state = ((uint32_t)warp < window) ? s_state[warp] : 2;
                // entries in front of the first showing one are settled: occluded (or stale)
                if (lane == 0 && (uint32_t)warp < window && warp < first && state == 0) { p.occluded[idx] = 1; atomicAdd(&s_nocc, 1u); }
                if (first < 32)
                {
                    const uint32_t vidx = p.sorted[pos + (uint32_t)first];
                    if (tid == 0) { p.occluded[vidx] = 0; p.visible[s_nvis++] = vidx; }
                    const uint32_t m = p.object_mesh[vidx];
                    if (m < p.n_meshes)
                    {
                        const uint32_t firsti = p.mesh_table[3 * m], count = p.mesh_table[3 * m + 1], base_v = p.mesh_table[3 * m + 2];
                        const float* model = p.object_models + (size_t)vidx * 16;
                        for (uint32_t t = (uint32_t)warp * 3u; t + 2u < count; t += 32u * 3u) // `i + 2 < indices.size()`, :125
                        {
                            float xy[3][2], z[3];
                            bool ok = true;
                            for (int v = 0; v < 3; ++v)
                            {
                                const uint32_t vi = base_v + p.indices[firsti + t + v];
                                ok = ok && vi < p.n_vertices && sc::occ_project_vertex(model, p.vertices + (size_t)vi * 3, p.view_proj, p.width, p.height, xy[v], z[v]);
                            }
                            if (!ok) continue; // warp-uniform
                            const sc::OccTri tri = sc::occ_setup_triangle(xy[0], z[0], xy[1], z[1], xy[2], z[2], p.width, p.height);
                            if (!tri.valid) continue;
                            const int bw = tri.max_x - tri.min_x + 1, total = bw * (tri.max_y - tri.min_y + 1);
                            for (int i = lane; i < total; i += 32)
                            {
                                const int x = tri.min_x + i % bw, y = tri.min_y + i / bw;
                                float d;
                                if (sc::occ_texel_depth(tri, x, y, d)) atomicMin(p.depth_bits + (size_t)y * p.width + x, __float_as_uint(d == 0.0f ? 0.0f : d)); // -0 -> +0
                            }
                        }
                    }
                    pos += (uint32_t)first + 1u;
                }
                else pos += window;
                __syncthreads(); // the next round's tests must see this object's depths; s_state is rewritten
            }
            __syncthreads();
            if (tid == 0) { p.counts[0] = s_nvis; p.counts[1] = s_nocc; }
        }
    }

    void launch_software_occlusion(const float* boxes6, uint32_t n_objects, const uint32_t* sorted, uint32_t n_sorted, const uint32_t* object_mesh, const float* object_models,
                                   const uint32_t* mesh_table, uint32_t n_meshes, const float* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices,
                                   const float view_proj[16], int width, int height, float epsilon, float* depth, uint8_t* occluded, uint32_t* visible, uint32_t* counts2,
                                   void* rect_scratch, cudaStream_t s, uint64_t* launches)
    {
        sc::OccRect* rects = static_cast<sc::OccRect*>(rect_scratch); // n_sorted x (sizeof(sc::OccRect) = 24 bytes + 1 "settled" byte)
        OccParams p{};
        p.boxes6 = boxes6; p.n_objects = n_objects; p.sorted = sorted; p.n_sorted = n_sorted; p.object_mesh = object_mesh; p.object_models = object_models;
        p.mesh_table = mesh_table; p.n_meshes = n_meshes; p.vertices = vertices; p.n_vertices = n_vertices; p.indices = indices; p.n_indices = n_indices;
        for (int i = 0; i < 16; ++i) p.view_proj[i] = view_proj[i];
        p.width = width; p.height = height; p.epsilon = epsilon;
        p.depth_bits = reinterpret_cast<uint32_t*>(depth); p.occluded = occluded; p.visible = visible; p.counts = counts2;
        launch_fill_u32(p.depth_bits, 0x3F800000u, (size_t)width * (size_t)height, s, launches); // std::fill(occlusion_depth, 1.0f), :286
        if (n_sorted == 0) { cudaMemsetAsync(counts2, 0, 8, s); return; }
        uint8_t* settled = reinterpret_cast<uint8_t*>(rects + n_sorted);
        cudaMemsetAsync(settled, 0, n_sorted, s);
        occlusion_rects_kernel<<<(n_sorted + 255) / 256, 256, 0, s>>>(p, rects);
        software_occlusion_kernel<<<1, 1024, 0, s>>>(p, rects, settled);
        *launches += 2;
    }
}
