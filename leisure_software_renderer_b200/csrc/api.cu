// api.cu -- the C-ABI of include/shsb.h: context, device-resident resources / render targets and the
// pass entry points.  This is the host half of the path; it replaces the reference's job-system dispatch
// (job/parallel_for.hpp:23-59, job/thread_pool_job_system.hpp:26-110) with kernel launches on CUDA streams,
// and its std::vector render targets (gfx/rt_types.hpp:35-157) with HBM buffers.
//
// Frame pipelining.  A submission has a FRONT END (draw-list upload, vertex / clip / set-up, binning, light
// culling) that touches only per-frame transients, and the TILE KERNEL that touches the render targets.  The
// front end runs on a high-priority side stream in one of three transient arenas; the tile kernel runs on the
// context's main stream.  So the front ends of frames f+1 and f+2 overlap frame f's tile kernel, while everything that
// touches render targets (tile kernels, clears, tonemap, downloads, work the caller orders on shsb_stream())
// stays in submission order on the main stream.
//
// There is no CPU raster path in this file: every pass either launches the kernels or returns an error.
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unistd.h>
#include <vector>

#include "../../include/shsb.h"
#include "host_math.hpp"
#include "asset_loaders.hpp"
#include "shsb_dev.cuh"

using namespace shsb;
namespace hm = shsb_host;

namespace
{
    // stage events: 0 front-end begin, 1 after geometry, 2 after binning, 3 tile kernel begin, 4 tile kernel end,
    // 5 / 6 light cull or standalone tonemap begin / end
    constexpr int NUM_STAGE_EVENTS = 7;
    constexpr int TIMING_EVENTS_PER_FRAME = 5;
    constexpr int NUM_ARENAS = 6;       // maximum number of transient arenas; ctx->n_arenas of them rotate (SHSB_ARENAS, default 4):
                                        // the front end may run n_arenas - 1 frames ahead of the tile kernel
    constexpr int TILE_DONE_RING = 16;  // per-frame "tile kernel finished" events
    constexpr int MAX_TILE_STREAMS = 4; // render streams: [0] = the context's main stream, [1..] side streams for asynchronous frames

    struct MeshSlot
    {
        bool live = false;
        float* positions = nullptr;
        float* normals = nullptr;
        float* uvs = nullptr;
        uint32_t* indices = nullptr;
        uint32_t n_positions = 0, n_normals = 0, n_uvs = 0, n_indices = 0;
        hm::vec3f bmin{0, 0, 0}, bmax{0, 0, 0}; // local bounds (PassShadowMap mesh_bounds_cache, pass_shadow_map.hpp:96-110)
    };

    struct TexSlot
    {
        bool live = false;
        uchar4* texels = nullptr;
        int w = 0, h = 0;
    };

    struct RtSlot
    {
        bool live = false;
        int kind = 0, w = 0, h = 0;
        float zn = 0.1f, zf = 1000.0f;
        void* color = nullptr;    // float4 (HDR) or uchar4 (LDR)
        float* depth = nullptr;   // DEPTH_MOTION, SHADOW
        float2* motion = nullptr; // DEPTH_MOTION
        uint32_t* tri_id = nullptr;
        uint32_t* coverage = nullptr;
        cudaEvent_t read_done = nullptr; // last asynchronous download of this RT (copy stream)
        bool read_pending = false;
        bool motion_dirty = false;       // the motion plane may hold non-zero vectors (a pass that clears it has work to do)
        // hazard tracking across render streams: the last frame (run_frame) that used this target
        long long last_frame = -1;       // its frame number (-> ctx->ev_tile_done[last_frame % TILE_DONE_RING]), -1 = none
        int last_stream = 0;             // render stream that frame ran on
        bool frame_is_last = false;      // nothing on the main stream has touched the target since that frame
        unsigned long long main_seq = 0; // stamp of the last main-stream operation that touched it (ctx->main_seq)
        int affinity = -1;               // render stream its asynchronous frames go to
    };

    template <typename T>
    struct DevBuf
    {
        T* p = nullptr;
        size_t cap = 0;
    };

    template <typename T>
    struct PinnedBuf // pinned AND mapped: kernels read it in place through `dp` (zero-copy)
    {
        T* p = nullptr;
        T* dp = nullptr;
        size_t cap = 0;
    };

    // Per-frame transients (DESIGN.md "Data layout").  The arenas rotate so that front ends can fill the next
    // ones while the tile kernel of an earlier frame still reads its own.
    struct Arena
    {
        DevBuf<unsigned char> d_draw;   // [DevItem x n_items | uint2 block table], one H2D copy per frame
        DevBuf<RasterRec> d_rrecs;
        DevBuf<ShadeRec> d_srecs;
        DevBuf<uint2> d_clipq;
        DevBuf<uint32_t> d_tile_offset, d_tile_fill, d_tile_list, d_tile_order;
        // frame header, cleared by ONE memset: [0] rec_count [1] clipq_count [2] list_cursor [4..7] class_count |
        // [32 ...] DevStats[STAT_SHARDS] | tile_count[n_tiles + 1]
        DevBuf<uint32_t> d_hdr;
    };

    // Tile light lists (LightCullingRuntimePayload, pipeline/render_pass.hpp:32-50): one set per arena for the fused
    // frame, one for the standalone shsb_light_cull.
    struct LightLists
    {
        DevBuf<uint32_t> counts, indices, scratch;
        uint32_t w = 0, h = 0, ts = 0, max_per_tile = 0;
        long long last_reader = -1;     // frame number of the last tile kernel that read these lists
    };
    constexpr int LISTS_STANDALONE = NUM_ARENAS;
}

namespace
{
    // Fork-join helper for the host half of a submission (per-draw model / normal matrices in the reference's exact GLM
    // order are ~200 ns each: 2 ms for the 10 k draws of the 8K config).  It takes the place ThreadPoolJobSystem +
    // parallel_for_1d (job/thread_pool_job_system.hpp:26-110, job/parallel_for.hpp:23-59) have in the reference, for the one
    // loop that stays on the host.  Workers sleep on a condition variable between submissions; the caller works too.
    class HostPool
    {
    public:
        explicit HostPool(int n_workers)
        {
            for (int i = 0; i < n_workers; ++i) workers_.emplace_back([this, i] { run(i + 1); });
        }
        ~HostPool()
        {
            stop_.store(true);
            gen_.fetch_add(1);
            { std::lock_guard<std::mutex> lk(m_); }
            cv_.notify_all();
            for (std::thread& t : workers_) t.join();
        }
        int parts() const { return (int)workers_.size() + 1; }
        // fn(part, n_parts) is called once per part, part 0 on the calling thread; returns when all parts are done.
        // Workers spin for a couple of milliseconds after a job before they go to sleep: frames arrive every millisecond or
        // so, and waking a sleeping thread costs more than the whole job.
        void run_parts(const std::function<void(int, int)>& fn)
        {
            if (workers_.empty()) { fn(0, 1); return; }
            fn_ = &fn;
            pending_.store((int)workers_.size());
            gen_.fetch_add(1);
            { std::lock_guard<std::mutex> lk(m_); } // a worker between its predicate check and its wait holds m_
            cv_.notify_all();
            fn(0, parts());
            while (pending_.load() != 0) std::this_thread::yield();
            fn_ = nullptr;
        }

    private:
        void run(int part)
        {
            uint64_t seen = 0;
            for (;;)
            {
                const auto t0 = std::chrono::steady_clock::now();
                int spins = 0;
                while (gen_.load() == seen)
                {
#if defined(__x86_64__) || defined(__i386__)
                    __builtin_ia32_pause();
#else
                    std::this_thread::yield(); // aarch64 (Grace) hosts have no pause intrinsic of that name
#endif
                    if ((++spins & 63) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2))
                    {
                        std::unique_lock<std::mutex> lk(m_);
                        cv_.wait(lk, [&] { return gen_.load() != seen; });
                        break;
                    }
                }
                seen = gen_.load();
                if (stop_.load()) return;
                (*fn_)(part, parts());
                pending_.fetch_sub(1);
            }
        }
        std::vector<std::thread> workers_;
        std::mutex m_;
        std::condition_variable cv_;
        const std::function<void(int, int)>* fn_ = nullptr;
        std::atomic<int> pending_{0};
        std::atomic<uint64_t> gen_{0};
        std::atomic<bool> stop_{false};
    };
}

namespace
{
    // Per-item scratch of a scene submission (forward_common): filled in parallel, consumed in draw order.
    struct DrawPrep
    {
        uint8_t state;        // 0 not a draw, 1 preculled (sphere), 2 culled by the exact bounds test, 3 kept
        uint32_t hist_slot;   // position among the draws (the motion history is positional first)
        int64_t prev_src;     // index into hist_models, or -1
        hm::mat4f model;
        DevItem item;         // everything but tri_offset
    };
}

namespace
{
    // Sort-first frame assembly (shsb_gather_*): the root's assembly memory and its control block of step counters.
    struct GatherCtl
    {
        unsigned long long arrive[SHSB_GATHER_MAX_RANKS]; // arrive[r] = last step rank r committed
        unsigned long long released;                      // last step whose slot the root released
        unsigned long long pad[15];
    };
    struct GatherSlot
    {
        bool live = false, root = false, ipc = false;
        uint32_t n_ranks = 0, rank = 0, slots = 0;
        size_t slot_bytes = 0;
        unsigned char* base = nullptr;
        GatherCtl* ctl = nullptr;
        uint64_t begun = 0, committed = 0, waited = 0, released = 0; // highest step this rank has begun / committed / (root) waited for / released
    };
    constexpr unsigned long long GATHER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
}

struct shsb_context_t
{
    int device = 0;
    cudaStream_t stream = nullptr;          // main render stream == tile_streams[0]
    // Asynchronous frames into DIFFERENT render targets are independent: their tile kernels go to different render streams
    // so that the head of frame f+1 fills the SMs the tail of frame f leaves idle.  Hazards are tracked per render target.
    cudaStream_t tile_streams[MAX_TILE_STREAMS]{};
    int n_tile_streams = 2;                 // SHSB_TILE_STREAMS / shsb_set_tile_streams (1 = everything on the main stream)
    int affinity_next = 0;
    unsigned long long main_seq = 0;        // counts main-stream operations that touch render targets
    unsigned long long fence_seq = 0;       // main_seq at the last shsb_fence / shsb_stream
    unsigned long long side_synced[MAX_TILE_STREAMS]{}; // main_seq up to which each side stream is ordered behind the main stream
    long long side_last_frame[MAX_TILE_STREAMS] = {-1, -1, -1, -1}; // last frame submitted to each render stream
    bool side_joined[MAX_TILE_STREAMS] = {true, true, true, true};  // the main stream already waits for side_last_frame[k]
    cudaEvent_t ev_main_sync[MAX_TILE_STREAMS]{};
    // sort-first frame assembly
    std::vector<GatherSlot> gathers;
    cudaStream_t gather_stream = nullptr;
    uint32_t* h_gather_timeout = nullptr;   // pinned + mapped: raised by a flag wait that ran out of time
    uint32_t* d_gather_timeout = nullptr;
    std::string error;
    uint64_t launches = 0;

    std::vector<MeshSlot> meshes;
    std::vector<TexSlot> textures;
    std::vector<RtSlot> rts;
    DevBuf<DevMesh> d_meshes;
    DevBuf<DevTexture> d_textures;
    bool mesh_table_dirty = true, tex_table_dirty = true;
    float* d_srgb_lut = nullptr;

    // lights: a ring of buffers so that an upload for a later frame does not wait for the tile kernel still
    // reading the current records
    DevBuf<DevLightRec> d_lights[NUM_ARENAS];
    DevBuf<SmLight> d_smlights[NUM_ARENAS];   // digested at upload (light_prep_kernel)
    PinnedBuf<DevLightRec> h_lights[NUM_ARENAS]; // staging for callers whose records are in pageable memory
    cudaEvent_t lights_stage_done[NUM_ARENAS]{};
    bool lights_stage_busy[NUM_ARENAS]{};
    int lights_cur = 0;
    long long lights_last_user[NUM_ARENAS][MAX_TILE_STREAMS]; // per buffer and render stream: frame number of the last tile kernel that read it (-1 = none)
    uint32_t n_lights = 0;
    cudaEvent_t ev_lights_up = nullptr;       // last upload (front stream)
    cudaEvent_t ev_cull_main = nullptr;       // last standalone cull (main stream)
    bool cull_main_pending = false;
    bool lights_uploaded = false;
    LightLists lists[NUM_ARENAS + 1];
    int lists_cur = -1;                       // the set produced by the most recent cull, -1 = none

    // RenderHistoryState (core/context.hpp:84-94): last frame's model matrix per object, in draw order with its key;
    // looked up positionally first (the usual case: same scene, same order), through the map otherwise
    std::vector<uint64_t> hist_keys;
    bool hist_has_duplicates = false;       // two draws of the previous frame shared a motion key: the reference's map then holds the LAST one's model for both
    std::vector<uint64_t> key_set;          // scratch of the duplicate test (open addressing, a power of two of slots)
    std::vector<hm::mat4f> hist_models;
    std::unordered_map<uint64_t, size_t> hist_index;
    bool hist_index_valid = false;
    bool has_prev_frame = false;

    // per-frame transients
    Arena arena[NUM_ARENAS];
    int n_arenas = 4;
    long long frame_no = 0;
    cudaEvent_t ev_tile_done[TILE_DONE_RING]{};
    cudaEvent_t ev_front_done[NUM_ARENAS]{};
    // pinned staging ring: a frame's draw list is copied H2D asynchronously, so a slot may only be rewritten once
    // the copy that read it has completed (its event)
    static constexpr int STAGE_SLOTS = 8;
    PinnedBuf<unsigned char> h_draw[STAGE_SLOTS];
    cudaEvent_t stage_done[STAGE_SLOTS]{};
    bool stage_busy[STAGE_SLOTS]{};
    int stage_slot = 0;
    static constexpr size_t HDR_STATS = 32, HDR_TILE_COUNT = HDR_STATS + STAT_SHARDS * sizeof(DevStats) / 4; // u32 words
    DevStats* h_stats = nullptr;    // pinned
    double rec_growth = 1.0;        // multiplier learned from overflow reruns
    uint32_t* h_overflow = nullptr; // pinned + mapped, sticky: [0] a kernel of some frame had to drop a record / list / clip-queue entry, [1] / [2] its demand
    uint32_t* d_overflow = nullptr;
    size_t min_list_cap = 0, min_rec_cap = 0; // learned from overflows: what a frame of this context has needed

    cudaEvent_t ev[NUM_STAGE_EVENTS]{};
    bool ev_valid[NUM_STAGE_EVENTS]{};
    std::vector<DrawPrep> prep;
    HostPool* host_pool = nullptr;       // created on first use by a submission with >= 2048 draws
    int host_threads = 1;                // SHSB_HOST_THREADS, default clamp(hardware threads / 8, 1, 4): a share of an 8-GPU box
    bool stage_events = true;            // false while an asynchronous frame without timing history is submitted
    bool main_needs_lights = false;      // the render stream has not yet been ordered behind the last light upload

    // Frame submission as a CUDA graph: the frame's memset / H2D copy / kernels are stream-captured, an existing
    // executable graph of the same topology is updated in place (cudaGraphExecUpdate) and launched once.  The
    // light-culling kernels do not depend on geometry / binning, so in the fused Forward+ frame they are
    // captured on a second stream (fork / join) and run concurrently with those small launch-bound kernels.
    bool use_graph = true;
    bool pipeline = true;                // SHSB_NO_PIPELINE=1: front end on the main stream (no frame overlap)
    bool capturing = false;
    // One front-end stream (+ one for its light-cull branch) PER ARENA, high priority: the front ends of consecutive
    // frames are independent of each other, so they overlap instead of queueing behind one another.
    cudaStream_t front_streams[NUM_ARENAS]{};
    cudaStream_t cull_streams[NUM_ARENAS]{};
    cudaStream_t copy_stream = nullptr;  // asynchronous render-target downloads alternate between two streams so that
    cudaStream_t copy_stream2 = nullptr; // the next copy is already queued on the engine when one finishes
    int copy_flip = 0;
    cudaEvent_t ev_fork[NUM_ARENAS]{}, ev_join[NUM_ARENAS]{}, ev_frame_done = nullptr, ev_front_sync = nullptr;
    bool shadow_direct = true;           // SHSB_SHADOW_DIRECT=0: every shadow-pass triangle through the binned tile path
    bool area_lights = true;             // the current light set may hold rect / tube lights (decides the tile kernel's instantiation)
    int last_tile_mode = -1;             // instantiation of the last tile kernel launched: PROG * 10 + LIGHTS (tile_raster.cu), 0 = general
    bool fast_tile = true;               // SHSB_NO_FAST_TILE=1: the tile kernel's general instantiation for every frame (tile_raster.cu: launch_tile_raster)
    bool hiz = false;                    // SHSB_HIZ=1: hierarchical-Z early reject in the tile kernel for asynchronous frames without AOVs
    cudaGraphExec_t graph_exec[NUM_ARENAS][8]{}; // per arena (an executable graph cannot run concurrently with itself): [stage events][cull branch][shadow mode] -- one executable per topology, so that a sampled (timed) frame does not force a re-instantiation

    // depth-range / clustered light culling (shsb_light_cull_ex): per-tile view-depth ranges, per-slice NDC bounds, cluster bins
    DevBuf<float> d_range_min, d_range_max;   // last shsb_tile_depth_range result
    uint32_t range_w = 0, range_h = 0, range_ts = 0;
    DevBuf<float> d_range_up_min, d_range_up_max; // caller-provided ranges
    DevBuf<float2> d_slice_ndc;
    DevBuf<uint32_t> d_vis;                   // frustum-visible light list
    LightLists cluster_lists;
    uint32_t cluster_slices = 0;

    // post passes: scratch for in-place operation, the light-shaft luma plane, TAA colour history
    DevBuf<uchar4> d_post_scratch;
    DevBuf<float> d_post_luma;
    DevBuf<uchar4> d_taa_hist;      // TemporalAARuntimeState::history (core/context.hpp:101)
    DevBuf<LegacyTri> d_legacy_tris; // set-up records of the legacy tile-job variant (legacy.cu)
    DevBuf<uint8_t> d_sc_bytes;      // scratch of the scene-level culling calls (scene_cull.cu): inputs and outputs, one allocation
    DevBuf<uint8_t> d_fd_bytes;      // scratch of the flat-shaded draws (flat_draw.cu): draw / light tables, triangle records, work list, depth keys
    uint32_t fd_item_cap = 0;        // work-list capacity the last batches needed
    DevBuf<l2::RasterRec> d_l2_raster; // slot records of the legacy render-target demos (legacy2.cu)
    DevBuf<l2::BoxRec> d_l2_box;
    DevBuf<l2::ShadeRec> d_l2_shade;
    float* d_l2_rot = nullptr;         // (sin, cos) of the 2^24 PCSS kernel rotation angles, tabulated by the host's libm on first use (128 MB)
    struct IblSlot { bool live = false; float* irradiance = nullptr; float* prefiltered = nullptr; int irr_size = 0, n_mips = 0; int spec_size[l2::MAX_SPEC_MIPS] = {}; uint32_t spec_off[l2::MAX_SPEC_MIPS] = {}; };
    std::vector<IblSlot> ibls;        // EnvIBL of the legacy PBR demo
    int taa_w = 0, taa_h = 0;
    bool taa_valid = false;

    // host-side submit cost breakdown (microseconds, accumulated): [0] scene -> draw list, [1] staging copy,
    // [2] arena checks, [3] capture / enqueue, [4] graph update + launch, [5] frames
    double host_us[8]{};

    // optional per-frame stage timing history (4 events per frame, no host sync while recording)
    bool timing_on = false;
    long long timing_stride = 1;         // record the stage events of every n-th frame only
    std::vector<cudaEvent_t> timing_ev;
    size_t timing_used = 0;
};

namespace
{
    inline double now_us()
    {
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

    int fail(shsb_ctx c, int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        if (c) c->error = buf;
        return code;
    }

#define CK(call)                                                                                                \
    do                                                                                                          \
    {                                                                                                           \
        const cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(ctx, SHSB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

    void sync_all(shsb_ctx ctx)
    {
        for (cudaStream_t st : ctx->front_streams) if (st) cudaStreamSynchronize(st);
        for (cudaStream_t st : ctx->cull_streams) if (st) cudaStreamSynchronize(st);
        for (cudaStream_t st : ctx->tile_streams) if (st) cudaStreamSynchronize(st);
        if (ctx->gather_stream) cudaStreamSynchronize(ctx->gather_stream);
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->copy_stream2);
    }

    template <typename T>
    int ensure_dev(shsb_ctx ctx, DevBuf<T>& b, size_t n)
    {
        if (n <= b.cap) return SHSB_OK;
        const size_t want = std::max(n, b.cap + b.cap / 2);
        T* np = nullptr;
        const cudaError_t e = cudaMalloc(&np, want * sizeof(T));
        if (e != cudaSuccess) return fail(ctx, SHSB_E_OUT_OF_MEMORY, "cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e));
        if (b.p)
        {
            sync_all(ctx);
            cudaFree(b.p);
        }
        b.p = np;
        b.cap = want;
        return SHSB_OK;
    }

    template <typename T>
    int ensure_pinned(shsb_ctx ctx, PinnedBuf<T>& b, size_t n)
    {
        if (n <= b.cap) return SHSB_OK;
        const size_t want = std::max(n, b.cap + b.cap / 2);
        T* np = nullptr;
        T* ndp = nullptr;
        cudaError_t e = cudaHostAlloc(&np, want * sizeof(T), cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(&ndp, np, 0);
        if (e != cudaSuccess) return fail(ctx, SHSB_E_OUT_OF_MEMORY, "cudaHostAlloc(%zu bytes, mapped): %s", want * sizeof(T), cudaGetErrorString(e));
        if (b.p)
        {
            sync_all(ctx);
            cudaFreeHost(b.p);
        }
        b.p = np;
        b.dp = ndp;
        b.cap = want;
        return SHSB_OK;
    }

    // Handle lookup only.
    RtSlot* peek_rt(shsb_ctx ctx, shsb_rt h, int kind = 0)
    {
        if (h == 0 || h > ctx->rts.size()) return nullptr;
        RtSlot* r = &ctx->rts[h - 1];
        if (!r->live) return nullptr;
        if (kind && r->kind != kind) return nullptr;
        return r;
    }

    // The main stream is about to use target r: order it behind the side-stream frame that last used r, if any.
    void join_rt_to_main(shsb_ctx ctx, RtSlot* r)
    {
        if (r->last_frame >= 0 && r->last_stream > 0)
        {
            cudaStreamWaitEvent(ctx->stream, ctx->ev_tile_done[r->last_frame % TILE_DONE_RING], 0);
            r->last_stream = 0;
        }
        r->frame_is_last = false;
        r->main_seq = ++ctx->main_seq;
    }

    // Lookup for an operation that is about to touch the target ON THE MAIN STREAM (every entry point but the frame
    // submissions, which choose their own render stream): joins pending side-stream work on it, orders the main stream
    // behind an asynchronous download still reading it, and stamps it so that later side-stream frames order themselves
    // behind this operation.
    RtSlot* get_rt(shsb_ctx ctx, shsb_rt h, int kind = 0)
    {
        RtSlot* r = peek_rt(ctx, h, kind);
        if (!r) return nullptr;
        join_rt_to_main(ctx, r);
        if (r->read_pending)
        {
            if (cudaEventQuery(r->read_done) != cudaSuccess)
            {
                cudaGetLastError();
                cudaStreamWaitEvent(ctx->stream, r->read_done, 0);
            }
            r->read_pending = false;
        }
        return r;
    }

    MeshSlot* get_mesh(shsb_ctx ctx, shsb_mesh h)
    {
        if (h == 0 || h > ctx->meshes.size()) return nullptr;
        MeshSlot* m = &ctx->meshes[h - 1];
        return m->live ? m : nullptr;
    }

    size_t plane_bytes(const RtSlot& r, int plane, void** ptr)
    {
        const size_t n = (size_t)r.w * (size_t)r.h;
        switch (plane)
        {
        case SHSB_PLANE_COLOR:
            *ptr = r.color;
            if (r.kind == SHSB_RT_COLOR_HDR) return n * 16;
            if (r.kind == SHSB_RT_COLOR_LDR) return n * 4;
            return 0;
        case SHSB_PLANE_DEPTH: *ptr = r.depth; return r.depth ? n * 4 : 0;
        case SHSB_PLANE_MOTION: *ptr = r.motion; return r.motion ? n * 8 : 0;
        case SHSB_PLANE_TRI_ID: *ptr = r.tri_id; return r.tri_id ? n * 4 : 0;
        case SHSB_PLANE_COVERAGE: *ptr = r.coverage; return r.coverage ? n * 4 : 0;
        default: *ptr = nullptr; return 0;
        }
    }

    int sync_tables(shsb_ctx ctx)
    {
        if (ctx->mesh_table_dirty)
        {
            std::vector<DevMesh> t(std::max<size_t>(1, ctx->meshes.size()));
            for (size_t i = 0; i < ctx->meshes.size(); ++i)
            {
                const MeshSlot& m = ctx->meshes[i];
                t[i] = DevMesh{m.positions, m.normals, m.uvs, m.indices, m.n_positions, m.n_normals, m.n_uvs, m.n_indices};
            }
            if (int rc = ensure_dev(ctx, ctx->d_meshes, t.size())) return rc;
            CK(cudaMemcpyAsync(ctx->d_meshes.p, t.data(), t.size() * sizeof(DevMesh), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->mesh_table_dirty = false;
        }
        if (ctx->tex_table_dirty)
        {
            std::vector<DevTexture> t(std::max<size_t>(1, ctx->textures.size()));
            for (size_t i = 0; i < ctx->textures.size(); ++i) t[i] = DevTexture{ctx->textures[i].texels, ctx->textures[i].w, ctx->textures[i].h};
            if (int rc = ensure_dev(ctx, ctx->d_textures, t.size())) return rc;
            CK(cudaMemcpyAsync(ctx->d_textures.p, t.data(), t.size() * sizeof(DevTexture), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->tex_table_dirty = false;
        }
        return SHSB_OK;
    }

    struct FrameJob
    {
        FrameConst fc{};
        FrameBuffers fb{};
        uint32_t n_items = 0;
        uint64_t n_src_tris = 0;
        uint32_t n_blocks = 0;
        int lists_set = -1;     // light-list set the tile kernel reads (Forward+), -1 = none
        int tile_stream = 0;    // render stream of the tile kernel (0 = main)
        long long frame_no = -1; // out: the frame number the submission got
        // shadow pass: its set-up kernel rasterises small triangles straight into the depth plane, so the FRONT END touches a render
        // target here: it is ordered behind the main stream and starts by filling the plane with the clear value
        float* prefill_depth = nullptr;
        size_t prefill_n = 0;
    };

    void record_on(shsb_ctx ctx, cudaEvent_t e, cudaStream_t s)
    {
        // inside a stream capture a plain cudaEventRecord would only mark a dependency; External makes it a node
        if (ctx->capturing) cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);
        else cudaEventRecord(e, s);
    }

    // Stage events feed shsb_last_stage_ms / the timing history.  An asynchronous submission without the timing history has
    // no reader for them, and every extra command on the render stream is a separate fetch by the GPU's host interface
    // over PCIe -- microseconds each once a frame read-back saturates the link (profiles/r1_pcie_command_latency_s4.md).
    void record(shsb_ctx ctx, int i, cudaStream_t s)
    {
        if (!ctx->stage_events) { ctx->ev_valid[i] = false; return; }
        record_on(ctx, ctx->ev[i], s);
        ctx->ev_valid[i] = true;
        if (ctx->timing_on && i < TIMING_EVENTS_PER_FRAME)
        {
            if (ctx->timing_used == ctx->timing_ev.size())
            {
                cudaEvent_t e = nullptr;
                if (cudaEventCreate(&e) != cudaSuccess) return;
                ctx->timing_ev.push_back(e);
            }
            record_on(ctx, ctx->timing_ev[ctx->timing_used++], s);
        }
    }

    // Host half of cull_lights_tiled: inverse(view_proj) and the camera frustum planes (jolt_light_culling.hpp:150-153).
    struct CullJob
    {
        float planes[24];
        float inv_vp[16];
        uint32_t vw = 0, vh = 0, ts = 0, max_per_tile = 0;
        int own_first = 0, own_count = 0, own_stride = 1; // sort-first: light-tile rows whose lists are needed (count 0 = all)
    };

    int prepare_light_cull(shsb_ctx ctx, const float view_proj[16], uint32_t vw, uint32_t vh, uint32_t ts, uint32_t max_per_tile, CullJob& job)
    {
        if (!view_proj || vw == 0 || vh == 0 || ts == 0 || max_per_tile == 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "bad light-cull arguments");
        const hm::mat4f vp = hm::load(view_proj);
        const hm::mat4f inv = hm::inverse(vp);
        hm::frustum_planes(vp, job.planes);
        hm::store(inv, job.inv_vp);
        job.vw = vw; job.vh = vh; job.ts = ts; job.max_per_tile = max_per_tile;
        return SHSB_OK;
    }

    int ensure_lists(shsb_ctx ctx, LightLists& L, const CullJob& job)
    {
        const uint32_t tiles = ((job.vw + job.ts - 1) / job.ts) * ((job.vh + job.ts - 1) / job.ts);
        if (int rc = ensure_dev(ctx, L.counts, tiles)) return rc;
        if (int rc = ensure_dev(ctx, L.indices, (size_t)tiles * job.max_per_tile)) return rc;
        if (int rc = ensure_dev(ctx, L.scratch, std::max<size_t>(1, light_cull_scratch_words(ctx->n_lights, job.vw, job.vh, job.ts)))) return rc;
        if (int rc = ensure_dev(ctx, ctx->d_lights[ctx->lights_cur], 1)) return rc;
        return SHSB_OK;
    }

    void enqueue_light_cull(shsb_ctx ctx, const CullJob& job, int set, cudaStream_t s)
    {
        LightLists& L = ctx->lists[set];
        record(ctx, 5, s);
        launch_light_cull(ctx->d_lights[ctx->lights_cur].p, ctx->n_lights, job.planes, job.inv_vp, job.vw, job.vh, job.ts, job.max_per_tile, L.scratch.p,
                          L.counts.p, L.indices.p, s, &ctx->launches, job.own_first, job.own_count, job.own_stride);
        record(ctx, 6, s);
        L.w = job.vw; L.h = job.vh; L.ts = job.ts; L.max_per_tile = job.max_per_tile;
        ctx->lists_cur = set;
    }

    int main_wait_lights(shsb_ctx ctx);

    // An asynchronous submission (out_stats == NULL) cannot re-run itself when its arena turns out too small: the kernels raise
    // a sticky flag in mapped host memory instead, and the next submission / shsb_sync / synchronous download reports it --
    // after growing the capacities, so that re-submitting the frame succeeds.
    void learn_from_overflow(shsb_ctx ctx)
    {
        // the demand words lag the flag by at most one kernel (the tile kernel writes them): whatever is there is a lower bound
        ctx->min_list_cap = std::max(ctx->min_list_cap, (size_t)ctx->h_overflow[1] + (size_t)ctx->h_overflow[1] / 4);
        ctx->min_rec_cap = std::max(ctx->min_rec_cap, (size_t)ctx->h_overflow[2] + (size_t)ctx->h_overflow[2] / 4);
        ctx->rec_growth *= 4.0;
        std::memset(ctx->h_overflow, 0, 4 * sizeof(uint32_t));
    }

    int check_async_overflow(shsb_ctx ctx)
    {
        if (ctx->h_overflow && *ctx->h_overflow)
        {
            learn_from_overflow(ctx);
            return fail(ctx, SHSB_E_OVERFLOW, "an asynchronous frame submitted earlier overflowed its per-frame arena (records / tile lists / clip queue) and is missing "
                                              "geometry; capacities have been grown x4: submit it again (or pass out_stats to let the call re-run itself)");
        }
        return SHSB_OK;
    }

    // Submits one frame: front end (draw list staged in ctx->h_draw[stage_slot]) on the front stream in arena
    // frame_no % NUM_ARENAS, tile kernel on the main stream.
    int run_frame(shsb_ctx ctx, FrameJob& job, ShsbStats* out_stats, const CullJob* cull = nullptr)
    {
        FrameConst& fc = job.fc;
        fc.tiles_x = (fc.W + TILE - 1) / TILE;
        fc.tiles_y = (fc.H + TILE_H - 1) / TILE_H;
        fc.hiz = (ctx->hiz && !out_stats && !fc.write_aovs) ? 1 : 0; // hidden fragments a Hi-Z reject skips are not counted: only where nobody reads the counters
        const uint32_t n_tiles = (uint32_t)fc.tiles_x * (uint32_t)fc.tiles_y;
        if (fc.W > 65535 || fc.H > 65535) return fail(ctx, SHSB_E_UNSUPPORTED, "render target larger than 65535 pixels on a side");
        if ((job.n_src_tris + 1) * 8ull >= 0xFFFFFFFFull) return fail(ctx, SHSB_E_UNSUPPORTED, "more than 2^29 source triangles in one submission");
        if (int rc = sync_tables(ctx)) return rc;
        if (int rc = check_async_overflow(ctx)) return rc;
        struct StageEventScope { shsb_ctx c; ~StageEventScope() { c->stage_events = true; } } stage_scope{ctx};
        ctx->stage_events = out_stats != nullptr || (ctx->timing_on && ctx->frame_no % ctx->timing_stride == 0);

        for (int attempt = 0; attempt < 4; ++attempt)
        {
            const double t_a = now_us();
            const long long f = ctx->frame_no++;
            const int a = (int)(f % ctx->n_arenas);
            Arena& A = ctx->arena[a];
            // capacities: every source triangle may emit one record; clipped ones up to 7
            // the clip queue takes every source triangle (8 bytes each: it cannot overflow); records are sized for a quarter of the
            // triangles being clipped into the maximum of 7 fan triangles, x the growth learned from overflows
            const size_t clipq_cap = std::max<size_t>(4096, (size_t)job.n_src_tris);
            const size_t clipped_est = std::min<size_t>(job.n_src_tris, std::max<size_t>(4096, (size_t)((double)job.n_src_tris * 0.25 * ctx->rec_growth)));
            const size_t rec_cap = std::max<size_t>(std::max<size_t>(4096, ctx->min_rec_cap), (size_t)((double)job.n_src_tris + 6.0 * (double)clipped_est));
            const size_t list_cap = std::max<size_t>(std::max<size_t>((size_t)n_tiles + 65536, ctx->min_list_cap), (size_t)((double)rec_cap * 4.0 * ctx->rec_growth) + (size_t)n_tiles * 2);
            const size_t items_bytes = (size_t)job.n_items * sizeof(DevItem);
            const size_t draw_bytes = items_bytes + (size_t)job.n_blocks * sizeof(uint2);
            if (int rc = ensure_dev(ctx, A.d_rrecs, rec_cap)) return rc;
            if (!fc.shadow_mode) { if (int rc = ensure_dev(ctx, A.d_srecs, rec_cap)) return rc; }
            if (int rc = ensure_dev(ctx, A.d_clipq, clipq_cap)) return rc;
            if (int rc = ensure_dev(ctx, A.d_hdr, shsb_context_t::HDR_TILE_COUNT + n_tiles + 1)) return rc;
            if (int rc = ensure_dev(ctx, A.d_tile_order, (size_t)4 * n_tiles)) return rc;
            if (int rc = ensure_dev(ctx, A.d_tile_offset, n_tiles + 2)) return rc;
            if (int rc = ensure_dev(ctx, A.d_tile_fill, n_tiles + 1)) return rc;
            if (int rc = ensure_dev(ctx, A.d_tile_list, list_cap)) return rc;
            if (int rc = ensure_dev(ctx, A.d_draw, std::max<size_t>(16, draw_bytes + 16))) return rc;
            if (cull) { if (int rc = ensure_lists(ctx, ctx->lists[a], *cull)) return rc; }

            uint32_t* hdr = A.d_hdr.p;
            DevStats* d_stats = reinterpret_cast<DevStats*>(hdr + shsb_context_t::HDR_STATS);
            const double t_b = now_us();
            ctx->host_us[2] += t_b - t_a;

            Geometry g{};
            g.meshes = ctx->d_meshes.p;
            g.items = reinterpret_cast<const DevItem*>(A.d_draw.p);
            g.block_table = reinterpret_cast<const uint2*>(A.d_draw.p + items_bytes);
            g.n_blocks = job.n_blocks;
            g.rrecs = A.d_rrecs.p;
            g.srecs = A.d_srecs.p;
            g.rec_capacity = (uint32_t)std::min<size_t>(A.d_rrecs.cap, fc.shadow_mode ? A.d_rrecs.cap : A.d_srecs.cap);
            g.rec_count = hdr + 0;
            g.clip_queue = A.d_clipq.p;
            g.clipq_capacity = (uint32_t)A.d_clipq.cap;
            g.clipq_count = hdr + 1;
            g.list_cursor = hdr + 2;
            g.class_count = hdr + 4;
            g.tile_order = A.d_tile_order.p;
            g.tile_count = hdr + shsb_context_t::HDR_TILE_COUNT;
            g.tile_offset = A.d_tile_offset.p;
            g.tile_fill = A.d_tile_fill.p;
            g.tile_list = A.d_tile_list.p;
            g.list_capacity = (uint32_t)std::min<size_t>(A.d_tile_list.cap, 0xFFFFFFFFull);
            g.stats = d_stats;
            g.overflow_flag = ctx->d_overflow;
            if (cull) job.lists_set = a; // the lists this frame shades with are the ones its own cull branch produces
            if (fc.forward_plus)
            {
                const LightLists& L = ctx->lists[job.lists_set];
                fc.lights = ctx->d_lights[ctx->lights_cur].p;
                fc.sm_lights = ctx->d_smlights[ctx->lights_cur].p;
                fc.tile_counts = L.counts.p;
                fc.tile_indices = L.indices.p;
            }

            const int slot = ctx->stage_slot;
            const bool graph = ctx->use_graph;
            cudaStream_t s1 = ctx->tile_streams[job.tile_stream], sf = ctx->pipeline ? ctx->front_streams[a] : s1, sc = ctx->cull_streams[a];
            // the ring slot this frame records into belonged to frame f - TILE_DONE_RING: make "slot re-recorded => that frame
            // has finished" an invariant (it finished long ago; this is a query, not a wait, in steady state)
            if (f >= TILE_DONE_RING) CK(cudaEventSynchronize(ctx->ev_tile_done[f % TILE_DONE_RING]));
            cudaError_t err = cudaSuccess;
            auto ok = [&](cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; };

            // ---- what the front end must wait for: the tile kernel that last read this arena (NUM_ARENAS frames ago) and,
            // when it rebuilds light lists, the last tile kernel that read that set
            if (f >= ctx->n_arenas) ok(cudaStreamWaitEvent(sf, ctx->ev_tile_done[(f - ctx->n_arenas) % TILE_DONE_RING], 0));
            if (cull && ctx->lists[a].last_reader >= 0) ok(cudaStreamWaitEvent(sf, ctx->ev_tile_done[ctx->lists[a].last_reader % TILE_DONE_RING], 0));
            if (cull && ctx->lights_uploaded) ok(cudaStreamWaitEvent(sf, ctx->ev_lights_up, 0)); // the upload ran on some arena's front stream

            if (job.prefill_depth && sf != s1)
            {
                ok(cudaEventRecord(ctx->ev_front_sync, s1));
                ok(cudaStreamWaitEvent(sf, ctx->ev_front_sync, 0));
            }
            if (graph)
            {
                CK(cudaStreamBeginCapture(sf, cudaStreamCaptureModeThreadLocal));
                ctx->capturing = true;
            }
            if (cull)
            {
                ok(cudaEventRecord(ctx->ev_fork[a], sf));
                ok(cudaStreamWaitEvent(sc, ctx->ev_fork[a], 0));
                enqueue_light_cull(ctx, *cull, a, sc);
                ok(cudaEventRecord(ctx->ev_join[a], sc));
            }
            record(ctx, 0, sf);
            ok(cudaMemsetAsync(hdr, 0, (shsb_context_t::HDR_TILE_COUNT + n_tiles + 1) * sizeof(uint32_t), sf));
            if (job.prefill_depth)
            {
                // shadow->clear(1.0f), pass_shadow_map.hpp:55 -- on the forked stream, next to the draw-list upload (which waits on PCIe)
                ok(cudaEventRecord(ctx->ev_fork[a], sf));
                ok(cudaStreamWaitEvent(sc, ctx->ev_fork[a], 0));
                launch_fill_u32(reinterpret_cast<uint32_t*>(job.prefill_depth), 0x3F800000u, job.prefill_n, sc, &ctx->launches);
                ok(cudaEventRecord(ctx->ev_join[a], sc));
            }
            if (draw_bytes) launch_upload(A.d_draw.p, ctx->h_draw[slot].dp, draw_bytes, sf, &ctx->launches); // SM-driven: no copy engine on the frame path
            if (job.prefill_depth) ok(cudaStreamWaitEvent(sf, ctx->ev_join[a], 0));
            launch_geometry(fc, g, sf, &ctx->launches);
            record(ctx, 1, sf);
            if (cull) ok(cudaStreamWaitEvent(sf, ctx->ev_join[a], 0)); // alloc_kernel reads the tile light counts (scheduling classes)
            if (!fc.direct_depth) launch_binning(fc, g, sf, &ctx->launches); // the direct shadow pass has no tile stage
            record(ctx, 2, sf);
            ok(cudaGetLastError());
            const double t_c = now_us();
            ctx->host_us[3] += t_c - t_b;
            if (graph)
            {
                ctx->capturing = false;
                cudaGraph_t captured = nullptr;
                const cudaError_t ec = cudaStreamEndCapture(sf, &captured);
                if (err == cudaSuccess) err = ec;
                if (err == cudaSuccess)
                {
                    cudaGraphExec_t& exec = ctx->graph_exec[a][(ctx->stage_events ? 4 : 0) + (cull ? 2 : 0) + (fc.shadow_mode ? 1 : 0)];
                    if (exec)
                    {
                        cudaGraphExecUpdateResultInfo info{};
                        if (cudaGraphExecUpdate(exec, captured, &info) != cudaSuccess)
                        {
                            cudaGetLastError(); // different topology (e.g. no draws this frame): rebuild
                            cudaGraphExecDestroy(exec);
                            exec = nullptr;
                        }
                    }
                    if (!exec) err = cudaGraphInstantiate(&exec, captured, 0);
                    if (err == cudaSuccess) err = cudaGraphLaunch(exec, sf);
                }
                if (captured) cudaGraphDestroy(captured);
            }
            if (err != cudaSuccess) return fail(ctx, SHSB_E_CUDA, "frame submission failed: %s", cudaGetErrorString(err));
            CK(cudaEventRecord(ctx->ev_front_done[a], sf));
            if (draw_bytes)
            {
                CK(cudaEventRecord(ctx->stage_done[slot], sf));
                ctx->stage_busy[slot] = true;
            }
            // ---- tile kernel: main stream, after this frame's front end (and, by stream order, after every earlier
            // tile kernel and whatever the caller ordered on the main stream)
            if (fc.forward_plus && !cull) { if (int rc = main_wait_lights(ctx)) return rc; } // lists built earlier: the records are read directly
            CK(cudaStreamWaitEvent(s1, ctx->ev_front_done[a], 0));
            record(ctx, 3, s1);
            if (!fc.direct_depth) launch_tile_raster(fc, g, job.fb, ctx->d_textures.p, ctx->d_srgb_lut, s1, &ctx->launches, ctx->fast_tile, &ctx->last_tile_mode);
            record(ctx, 4, s1);
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->ev_tile_done[f % TILE_DONE_RING], s1));
            if (fc.forward_plus)
            {
                ctx->lights_last_user[ctx->lights_cur][job.tile_stream] = f;
                ctx->lists[job.lists_set].last_reader = f;
            }
            ctx->host_us[4] += now_us() - t_c;
            ctx->host_us[5] += 1.0;
            job.frame_no = f;
            ctx->side_last_frame[job.tile_stream] = f;
            ctx->side_joined[job.tile_stream] = job.tile_stream == 0;

            if (!out_stats) return SHSB_OK; // asynchronous submission; overflow would surface at the next stats read
            CK(cudaMemcpyAsync(ctx->h_stats, d_stats, sizeof(DevStats) * STAT_SHARDS, cudaMemcpyDeviceToHost, s1));
            CK(cudaStreamSynchronize(s1));
            DevStats st{};
            for (int k = 0; k < STAT_SHARDS; ++k)
            {
                const DevStats& sh = ctx->h_stats[k];
                st.tri_input += sh.tri_input; st.tri_after_clip += sh.tri_after_clip; st.tri_raster += sh.tri_raster;
                st.frag_covered += sh.frag_covered; st.frag_shaded += sh.frag_shaded;
                st.overflow_recs += sh.overflow_recs; st.overflow_lists += sh.overflow_lists; st.overflow_clipq += sh.overflow_clipq;
            }
            if (st.overflow_recs || st.overflow_lists || st.overflow_clipq)
            {
                learn_from_overflow(ctx); // arena too small for this scene: grow (x4, and at least to what this frame needed) and re-run the frame
                continue;
            }
            out_stats->tri_input += st.tri_input;
            out_stats->tri_after_clip += st.tri_after_clip;
            out_stats->tri_raster += st.tri_raster;
            out_stats->frag_covered += st.frag_covered;
            out_stats->frag_shaded += st.frag_shaded;
            return SHSB_OK;
        }
        return fail(ctx, SHSB_E_OUT_OF_MEMORY, "per-frame arena overflow persisted after regrowth");
    }

    // Sort-first partitions (ShsbFrameParams::own_row_*): the two conservative "can this draw reach an owned tile row?" tests live in
    // host_math.hpp (pure host arithmetic, property-tested on the CPU by tests/cpp/host_cull_test.cpp).
    using VpRows = hm::VpRows;

    inline hm::RowOwnership ownership(const FrameConst& fc) { return hm::RowOwnership{fc.H, TILE, fc.own_first, fc.own_count, fc.own_stride}; }

    bool item_touches_owned_rows(const FrameConst& fc, const hm::mat4f& viewproj, const hm::mat4f& model, const MeshSlot& mesh)
    {
        return hm::bounds_touch_owned_rows(ownership(fc), viewproj, model, mesh.bmin, mesh.bmax);
    }

    bool sphere_may_touch_owned_rows(const FrameConst& fc, const VpRows& v, const ShsbTransform& tr, const MeshSlot& mesh)
    {
        return hm::sphere_may_touch_owned_rows(ownership(fc), v, tr.pos, tr.scl, mesh.bmin, mesh.bmax);
    }

    // One draw's device record (everything but its place in the draw order).  Thread-safe: reads the context only.
    void fill_item(shsb_ctx ctx, DevItem& it, const hm::mat4f& model, const MeshSlot& mesh, uint32_t mesh_index,
                   const float base_color[3], float metallic, float roughness, float ao, uint32_t tex, const hm::mat4f* prev_model = nullptr)
    {
        it = DevItem{};
        hm::store(model, it.model);
        if (prev_model)
        {
            // curr_to_prev_model, rasterizer.hpp:296-307
            hm::mat4f c2p = hm::identity();
            if (std::fabs(hm::determinant(model)) > 1e-10f) c2p = hm::mul(*prev_model, hm::inverse(model));
            hm::store(c2p, it.c2p);
        }
        hm::normal_matrix(model, it.nrm);
        it.base_color[0] = base_color[0]; it.base_color[1] = base_color[1]; it.base_color[2] = base_color[2];
        it.metallic = metallic; it.roughness = roughness; it.ao = ao;
        it.tex = (tex >= 1 && tex <= ctx->textures.size() && ctx->textures[tex - 1].live) ? tex : 0u;
        it.mesh = mesh_index;
        it.tri_count = mesh.n_indices ? mesh.n_indices / 3 : mesh.n_positions / 3;
    }

    // Stages one draw into the host item / block tables.
    int stage_item(shsb_ctx ctx, std::vector<DevItem>& items, std::vector<uint2>& blocks, uint64_t& tri_cursor,
                   const hm::mat4f& model, const MeshSlot& mesh, uint32_t mesh_index,
                   const float base_color[3], float metallic, float roughness, float ao, uint32_t tex, const hm::mat4f* prev_model = nullptr)
    {
        DevItem it{};
        hm::store(model, it.model);
        if (prev_model)
        {
            // curr_to_prev_model, rasterizer.hpp:296-307
            hm::mat4f c2p = hm::identity();
            if (std::fabs(hm::determinant(model)) > 1e-10f) c2p = hm::mul(*prev_model, hm::inverse(model));
            hm::store(c2p, it.c2p);
        }
        hm::normal_matrix(model, it.nrm);
        it.base_color[0] = base_color[0]; it.base_color[1] = base_color[1]; it.base_color[2] = base_color[2];
        it.metallic = metallic; it.roughness = roughness; it.ao = ao;
        it.tex = (tex >= 1 && tex <= ctx->textures.size() && ctx->textures[tex - 1].live) ? tex : 0u;
        it.mesh = mesh_index;
        it.tri_offset = (uint32_t)tri_cursor;
        it.tri_count = mesh.n_indices ? mesh.n_indices / 3 : mesh.n_positions / 3;
        const uint32_t item_index = (uint32_t)items.size();
        items.push_back(it);
        for (uint32_t t = 0; t < it.tri_count; t += 128) blocks.push_back(make_uint2(item_index, t));
        tri_cursor += it.tri_count;
        return SHSB_OK;
    }

    // Copies the draw list ([items | blocks]) into the next slot of the pinned staging ring.
    int upload_staging(shsb_ctx ctx, const std::vector<DevItem>& items, const std::vector<uint2>& blocks)
    {
        const int slot = ctx->stage_slot = (ctx->stage_slot + 1) % shsb_context_t::STAGE_SLOTS;
        if (ctx->stage_busy[slot])
        {
            CK(cudaEventSynchronize(ctx->stage_done[slot])); // the H2D copy that last read this slot has finished
            ctx->stage_busy[slot] = false;
        }
        const size_t items_bytes = items.size() * sizeof(DevItem), blocks_bytes = blocks.size() * sizeof(uint2);
        if (int rc = ensure_pinned(ctx, ctx->h_draw[slot], std::max<size_t>(16, items_bytes + blocks_bytes + 16))) return rc;
        if (items_bytes) std::memcpy(ctx->h_draw[slot].p, items.data(), items_bytes);
        if (blocks_bytes) std::memcpy(ctx->h_draw[slot].p + items_bytes, blocks.data(), blocks_bytes);
        return SHSB_OK;
    }

    // A pass that is about to overwrite a render target must not overtake an asynchronous download of it.  A download
    // that has already finished (the usual case with a few render-target sets in rotation) needs no command at all.
    void wait_pending_read(shsb_ctx ctx, RtSlot* r, cudaStream_t s = nullptr)
    {
        if (r && r->read_pending)
        {
            if (cudaEventQuery(r->read_done) != cudaSuccess)
            {
                cudaGetLastError();
                cudaStreamWaitEvent(s ? s : ctx->stream, r->read_done, 0);
            }
            r->read_pending = false;
        }
    }

    // Orders render stream k behind everything that last touched the given targets: frames on OTHER render streams (their
    // "tile done" events), main-stream operations (one lazily recorded event per side stream), asynchronous downloads.
    int order_frame_on(shsb_ctx ctx, int k, RtSlot* const* targets, int n)
    {
        cudaStream_t s = ctx->tile_streams[k];
        bool need_main = k != 0 && ctx->fence_seq > ctx->side_synced[k];
        for (int i = 0; i < n; ++i)
        {
            RtSlot* r = targets[i];
            if (!r) continue;
            if (r->last_frame >= 0 && r->last_stream != k) CK(cudaStreamWaitEvent(s, ctx->ev_tile_done[r->last_frame % TILE_DONE_RING], 0));
            if (k != 0 && r->main_seq > ctx->side_synced[k]) need_main = true;
            wait_pending_read(ctx, r, s);
        }
        if (need_main)
        {
            CK(cudaEventRecord(ctx->ev_main_sync[k], ctx->stream));
            CK(cudaStreamWaitEvent(s, ctx->ev_main_sync[k], 0));
            ctx->side_synced[k] = ctx->main_seq;
        }
        return SHSB_OK;
    }

    void frame_used_targets(int k, long long frame_no, RtSlot* const* targets, int n)
    {
        for (int i = 0; i < n; ++i)
            if (targets[i]) { targets[i]->last_frame = frame_no; targets[i]->last_stream = k; targets[i]->frame_is_last = true; }
    }

    // Main stream waits for the last frame of every side stream (shsb_fence, shsb_stream, shsb_sync callers that go on
    // to order their own work on the main stream).
    void join_side_streams(shsb_ctx ctx)
    {
        for (int k = 1; k < MAX_TILE_STREAMS; ++k)
            if (!ctx->side_joined[k] && ctx->side_last_frame[k] >= 0)
            {
                cudaStreamWaitEvent(ctx->stream, ctx->ev_tile_done[ctx->side_last_frame[k] % TILE_DONE_RING], 0);
                ctx->side_joined[k] = true;
            }
    }

    // Orders the render stream behind the last light upload (which ran on a front-end stream).  Needed by work on the
    // render stream that reads the light records directly: a standalone cull, or the tile kernel of a frame whose lists
    // were built earlier.  A frame that culls itself is ordered through its own front end instead.
    int main_wait_lights(shsb_ctx ctx)
    {
        if (ctx->main_needs_lights)
        {
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_lights_up, 0));
            ctx->main_needs_lights = false;
        }
        return SHSB_OK;
    }

    int shader_from_params(const ShsbFrameParams* fp)
    {
        int shader = (fp->shading_model == SHSB_SHADING_BLINN_PHONG) ? SHSB_SHADER_BLINN_PHONG : SHSB_SHADER_PBR_MR; // pass_pbr_forward.hpp:100-108
        if (fp->debug_view == SHSB_DEBUG_ALBEDO) shader = SHSB_SHADER_DEBUG_ALBEDO;
        else if (fp->debug_view == SHSB_DEBUG_NORMAL) shader = SHSB_SHADER_DEBUG_NORMAL;
        else if (fp->debug_view == SHSB_DEBUG_DEPTH) shader = SHSB_SHADER_DEBUG_DEPTH;
        return shader;
    }

    int ensure_aovs(shsb_ctx ctx, RtSlot* r)
    {
        const size_t n = (size_t)r->w * r->h;
        if (!r->tri_id) CK(cudaMalloc(&r->tri_id, n * 4));
        if (!r->coverage) CK(cudaMalloc(&r->coverage, n * 4));
        return SHSB_OK;
    }

    void fill_camera_sun(FrameConst& fc, const ShsbScene* s)
    {
        std::memcpy(fc.viewproj, s->cam_viewproj, 64);
        std::memcpy(fc.camera_pos, s->cam_pos, 12);
        std::memcpy(fc.sun_dir, s->sun_dir_ws, 12);
        std::memcpy(fc.sun_color, s->sun_color, 12);
        fc.sun_intensity = s->sun_intensity;
    }

    int forward_common(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp, shsb_rt hdr_rt, shsb_rt depth_rt, shsb_rt shadow_rt,
                       const float* shadow_lvp, int preserve_depth, bool depth_only, shsb_rt ldr_rt, ShsbStats* out_stats,
                       const CullJob* cull = nullptr)
    {
        if (!scene || !fp) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "scene / frame params are null");
        if (scene->n_items && !scene->items) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "scene->items is null");
        RtSlot* hdr = depth_only ? nullptr : peek_rt(ctx, hdr_rt, SHSB_RT_COLOR_HDR);
        if (!depth_only && !hdr) return fail(ctx, SHSB_E_INVALID_HANDLE, "hdr_rt is not a live RT_ColorHDR");
        RtSlot* dm = depth_rt ? peek_rt(ctx, depth_rt, SHSB_RT_DEPTH_MOTION) : nullptr;
        if (depth_rt && !dm) return fail(ctx, SHSB_E_INVALID_HANDLE, "depth_motion_rt is not a live RT_ColorDepthMotion");
        if (depth_only && !dm) return fail(ctx, SHSB_E_INVALID_HANDLE, "depth prepass needs a depth target");
        // the reference ignores a motion RT whose size differs from the HDR target (pass_pbr_forward.hpp:87,111)
        if (hdr && dm && (dm->w != hdr->w || dm->h != hdr->h)) dm = nullptr;
        const int W = hdr ? hdr->w : dm->w, H = hdr ? hdr->h : dm->h;
        RtSlot* ldr = ldr_rt ? peek_rt(ctx, ldr_rt, SHSB_RT_COLOR_LDR) : nullptr;
        if (ldr_rt && (!ldr || ldr->w != W || ldr->h != H)) return fail(ctx, SHSB_E_INVALID_HANDLE, "ldr_rt is not a live RT_ColorLDR of the HDR target's size");
        RtSlot* sh = (!depth_only && shadow_rt) ? peek_rt(ctx, shadow_rt, SHSB_RT_SHADOW) : nullptr;
        RtSlot* const dm_arg = depth_rt ? peek_rt(ctx, depth_rt, SHSB_RT_DEPTH_MOTION) : nullptr; // tracked even when its size mismatch makes the pass ignore it

        // Render stream of the tile kernel.  A synchronous submission (statistics) and a frame that shades with lists built
        // earlier on the main stream stay on the main stream; an asynchronous frame goes to the stream its primary target
        // has an affinity to (assigned round-robin at first use), so frames into different target sets overlap while
        // frames into the same set keep their order without any cross-stream command.
        FrameJob job;
        const bool lists_from_main = !depth_only && fp->light_culling && ctx->n_lights > 0 && !cull;
        if (!out_stats && ctx->pipeline && ctx->n_tile_streams > 1 && !lists_from_main)
        {
            RtSlot* prim = hdr ? hdr : dm;
            if (prim->affinity < 0 || prim->affinity >= ctx->n_tile_streams) prim->affinity = ctx->affinity_next++ % ctx->n_tile_streams;
            job.tile_stream = prim->affinity;
        }
        RtSlot* const used[5] = {hdr, dm, ldr, sh, dm_arg != dm ? dm_arg : nullptr};
        if (int rc = order_frame_on(ctx, job.tile_stream, used, 5)) return rc;
        if (job.tile_stream == 0) ctx->main_seq++; // a main-stream frame is a main-stream operation for fences

        FrameConst& fc = job.fc;
        fill_camera_sun(fc, scene);
        fc.W = W; fc.H = H;
        fc.zn = dm ? dm->zn : 0.1f;
        fc.zf = dm ? dm->zf : 1000.0f;
        fc.shader_id = depth_only ? SHSB_SHADER_DEPTH_ONLY : shader_from_params(fp);
        fc.cull_mode = fp->cull_mode;
        fc.front_face_ccw = fp->front_face_ccw;
        if (fp->own_row_count > 0)
        {
            if (fp->own_row_first < 0 || fp->own_row_stride < fp->own_row_count) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "own_row_first must be >= 0 and own_row_stride >= own_row_count");
            fc.own_first = fp->own_row_first; fc.own_count = fp->own_row_count; fc.own_stride = fp->own_row_stride;
        }
        fc.has_depth = dm ? 1 : 0;
        fc.linear_depth = (dm && (dm->zf > dm->zn + 1e-6f)) ? 1 : 0;
        fc.load_depth = (!depth_only && preserve_depth) ? 1 : 0;
        fc.load_color = 0;
        fc.write_aovs = fp->write_aovs;
        fc.shadow_mode = 0;

        if (fp->shadow_enable && sh && shadow_lvp) // pass_pbr_forward.hpp:185-194
        {
            fc.shadow_map = sh->depth;
            fc.shadow_w = sh->w; fc.shadow_h = sh->h;
            std::memcpy(fc.light_viewproj, shadow_lvp, 64);
            fc.bias_const = fp->shadow_bias_const;
            fc.bias_slope = fp->shadow_bias_slope;
            fc.pcf_radius = fp->shadow_pcf_radius;
            fc.pcf_step = fp->shadow_pcf_step;
            fc.shadow_strength = fp->shadow_strength;
        }
        if (!depth_only && fp->light_culling && ctx->n_lights > 0)
        {
            const LightLists* cur = ctx->lists_cur >= 0 ? &ctx->lists[ctx->lists_cur] : nullptr;
            if (!cull && (!cur || cur->w != (uint32_t)W || cur->h != (uint32_t)H))
                return fail(ctx, SHSB_E_INVALID_ARGUMENT, "light_culling is on but shsb_light_cull has not been run for a %dx%d viewport", W, H);
            const uint32_t ts = cull ? cull->ts : cur->ts, mx = cull ? cull->max_per_tile : cur->max_per_tile;
            fc.forward_plus = 1;
            fc.n_lights = ctx->n_lights;
            fc.area_lights = ctx->area_lights ? 1 : 0;
            job.lists_set = ctx->lists_cur; // run_frame points the kernels at the set (its own when it culls)
            fc.light_tile_size = ts;
            fc.max_per_tile = mx;
            fc.light_tiles_x = (W + ts - 1) / ts;
            fc.light_tiles_y = (H + ts - 1) / ts;
        }
        if (ldr)
        {
            fc.fuse_tonemap = 1;
            fc.exposure = std::max(0.0001f, fp->exposure);            // pass_tonemap.hpp:49-50
            fc.inv_gamma = 1.0f / std::max(0.001f, fp->gamma);
        }

        // motion vectors + history (pass_pbr_forward.hpp:87-98, 143-166, 212-213); never in the depth pre-pass (pass_adapters.hpp:521)
        const bool lit_pass = !depth_only;
        const bool write_motion = lit_pass && dm && dm->motion && fp->motion_vectors_enable != 0;
        if (lit_pass && dm && dm->motion)
        {
            fc.write_motion = write_motion ? 1 : 0;
            fc.clear_motion = (write_motion || dm->motion_dirty) ? 1 : 0; // both branches of the reference clear the plane
            job.fb.motion = (write_motion || dm->motion_dirty) ? dm->motion : nullptr;
            dm->motion_dirty = write_motion || (fp->own_row_count > 0 && dm->motion_dirty); // a partition cleans only its own rows
            const float* pvp = ctx->has_prev_frame ? scene->cam_prev_viewproj : scene->cam_viewproj;
            std::memcpy(fc.prev_viewproj, pvp, 64);
        }
        // sky model background (Scene::sky)
        if (lit_pass && scene->sky_kind != SHSB_SKY_NONE)
        {
            if (scene->sky_kind != SHSB_SKY_PROCEDURAL && scene->sky_kind != SHSB_SKY_CUBEMAP)
                return fail(ctx, SHSB_E_UNSUPPORTED, "sky kind %d is not one of the reference's sky models (procedural, cubemap)", scene->sky_kind);
            fc.sky_kind = scene->sky_kind;
            hm::store(hm::inverse(hm::load(scene->cam_viewproj)), fc.inv_viewproj);
            const hm::vec3f sun = hm::normalize(hm::vec3f{scene->sky_sun_dir_ws[0], scene->sky_sun_dir_ws[1], scene->sky_sun_dir_ws[2]});
            fc.sky_sun[0] = sun.x; fc.sky_sun[1] = sun.y; fc.sky_sun[2] = sun.z;
            fc.sky_intensity = scene->sky_intensity;
            if (scene->sky_kind == SHSB_SKY_CUBEMAP)
            {
                bool valid = true; // CubemapData::valid(), sky/cubemap_sky.hpp:26-33: an invalid cubemap samples black
                for (int i = 0; i < 6; ++i)
                {
                    const uint32_t t = scene->sky_faces[i];
                    if (t == 0 || t > ctx->textures.size() || !ctx->textures[t - 1].live) valid = false;
                    else fc.sky_faces[i] = t - 1;
                }
                if (!valid) { fc.sky_kind = SHSB_SKY_CUBEMAP; fc.sky_intensity = 0.0f; for (int i = 0; i < 6; ++i) fc.sky_faces[i] = 0; if (ctx->textures.empty()) return fail(ctx, SHSB_E_INVALID_HANDLE, "cubemap sky without any texture uploaded"); }
            }
        }

        job.fb.hdr = hdr ? (float4*)hdr->color : nullptr;
        job.fb.depth = dm ? dm->depth : nullptr;
        job.fb.ldr = ldr ? (uchar4*)ldr->color : nullptr;
        if (fp->write_aovs)
        {
            RtSlot* host = hdr ? hdr : dm;
            if (int rc = ensure_aovs(ctx, host)) return rc;
            job.fb.aov_tri_id = host->tri_id;
            job.fb.aov_coverage = host->coverage;
        }

        const double t_items = now_us();
        std::vector<DevItem> items;
        std::vector<uint2> blocks;
        items.reserve(scene->n_items);
        uint64_t tri_cursor = 0;
        std::vector<uint64_t> next_keys;
        std::vector<hm::mat4f> next_models;
        if (lit_pass) { next_keys.reserve(scene->n_items); next_models.reserve(scene->n_items); }
        // rows of cam.viewproj that produce clip.y and clip.w, with the norms of their xyz parts (sphere_may_touch_owned_rows)
        const VpRows vp_rows = hm::vp_rows_of(scene->cam_viewproj);
        // Motion vectors need last frame's model matrix of every item, so a partition that writes motion builds every matrix;
        // one that does not may drop a draw before its matrix exists, and then leaves the history empty (the next frame that
        // asks for motion vectors starts like a first frame) rather than half-filled.
        const bool precull = fc.own_count > 0 && !write_motion;
        bool history_partial = false;

        // ---- pass 0 (serial, integer work only): which items are draws at all, their motion keys and history slots
        using Prep = DrawPrep;
        std::vector<Prep>& prep = ctx->prep;
        prep.resize(scene->n_items);
        uint32_t n_draws = 0;
        const bool want_prev = lit_pass && write_motion && ctx->has_prev_frame;
        for (uint32_t i = 0; i < scene->n_items; ++i)
        {
            const ShsbRenderItem& it = scene->items[i];
            Prep& p = prep[i];
            p.state = 0;
            if (!it.visible) continue;
            const MeshSlot* mesh = get_mesh(ctx, it.mesh);
            if (!mesh || mesh->n_positions == 0 || mesh->n_indices == 0) continue; // MeshData::empty(), resources/mesh.hpp:32-35
            p.state = 3;
            p.hist_slot = n_draws++;
            p.prev_src = -1;
            if (lit_pass)
            {
                // motion key, pass_pbr_forward.hpp:143-148 (the reference's material handle is 0 for "no material"; items with a
                // material resolved on the caller's side are told apart by object_id or by their position)
                uint64_t key = it.object_id;
                if (key == 0)
                {
                    key = ((uint64_t)it.mesh << 32) ^ (uint64_t)it.has_material ^ ((uint64_t)i + 1u); // has_material carries RenderItem::mat (0 = none)
                    if (key == 0) key = 1;
                }
                next_keys.push_back(key);
                if (want_prev)
                {
                    const size_t k = p.hist_slot;
                    // same position, same key: the usual frame-to-frame case, no map needed -- unless the previous frame held the key more
                    // than once, where prev_model_by_object[key] is the LAST such draw's model for every one of them
                    if (!ctx->hist_has_duplicates && k < ctx->hist_keys.size() && ctx->hist_keys[k] == key) p.prev_src = (int64_t)k;
                    else
                    {
                        if (!ctx->hist_index_valid)
                        {
                            ctx->hist_index.clear();
                            for (size_t q = 0; q < ctx->hist_keys.size(); ++q) ctx->hist_index[ctx->hist_keys[q]] = q; // a later duplicate key wins, like the map assignment
                            ctx->hist_index_valid = true;
                        }
                        const auto f = ctx->hist_index.find(key);
                        if (f != ctx->hist_index.end()) p.prev_src = (int64_t)f->second;
                    }
                }
            }
        }

        // ---- pass 1 (parallel over items for large scenes): culling tests and the per-draw matrices, all in the reference's
        // operation order (host_math.hpp); items are independent of each other here
        const hm::mat4f cam_vp = hm::load(scene->cam_viewproj);
        std::atomic<uint32_t> next_chunk{0};
        constexpr uint32_t CHUNK = 64; // items are claimed in small chunks: kept draws (200 ns) cluster in index ranges, dropped ones cost 40 ns
        auto prepare = [&](int, int)
        {
            for (;;)
            {
            const uint32_t lo = next_chunk.fetch_add(1) * CHUNK;
            if (lo >= scene->n_items) break;
            const uint32_t hi = std::min(scene->n_items, lo + CHUNK);
            for (uint32_t i = lo; i < hi; ++i)
            {
                Prep& p = prep[i];
                if (p.state == 0) continue;
                const ShsbRenderItem& it = scene->items[i];
                const MeshSlot& mesh = ctx->meshes[it.mesh - 1];
                if (precull && !sphere_may_touch_owned_rows(fc, vp_rows, it.tr, mesh)) { p.state = 1; continue; } // cheap test first: no trigonometry, no matrix products
                p.model = hm::model_from_transform(it.tr.pos, it.tr.rot_euler, it.tr.scl);
                if (fc.own_count > 0 && !item_touches_owned_rows(fc, cam_vp, p.model, mesh)) { p.state = 2; continue; }
                const hm::mat4f prev_model = p.prev_src >= 0 ? ctx->hist_models[(size_t)p.prev_src] : p.model;
                const float def_color[3] = {0.8f, 0.5f, 0.2f}; // pass_pbr_forward.hpp:179-184
                const hm::mat4f* prev = (lit_pass && write_motion) ? &prev_model : nullptr;
                if (it.has_material) fill_item(ctx, p.item, p.model, mesh, it.mesh - 1, it.base_color, it.metallic, it.roughness, it.ao, it.base_color_tex, prev);
                else fill_item(ctx, p.item, p.model, mesh, it.mesh - 1, def_color, 0.1f, 0.5f, 1.0f, 0u, prev);
            }
            }
        };
        if (n_draws >= 2048 && ctx->host_threads > 1)
        {
            if (!ctx->host_pool) ctx->host_pool = new HostPool(ctx->host_threads - 1);
            ctx->host_pool->run_parts(prepare);
        }
        else prepare(0, 1);

        // ---- pass 2 (serial): draw order -- triangle offsets (culled draws keep their place: depth ties and the triangle-id
        // AOV are defined by the whole scene's order), the item and block tables, the motion history
        for (uint32_t i = 0; i < scene->n_items; ++i)
        {
            Prep& p = prep[i];
            if (p.state == 0) continue;
            const MeshSlot& mesh = ctx->meshes[scene->items[i].mesh - 1];
            const uint32_t tri_count = mesh.n_indices ? mesh.n_indices / 3 : mesh.n_positions / 3;
            if (p.state == 1) history_partial = true;              // dropped before its model matrix existed
            else if (lit_pass) next_models.push_back(p.model);
            if (p.state != 3) { tri_cursor += tri_count; continue; }
            p.item.tri_offset = (uint32_t)tri_cursor;
            const uint32_t item_index = (uint32_t)items.size();
            items.push_back(p.item);
            for (uint32_t t = 0; t < p.item.tri_count; t += 128) blocks.push_back(make_uint2(item_index, t));
            tri_cursor += p.item.tri_count;
        }
        const double t_stage = now_us();
        ctx->host_us[0] += t_stage - t_items;
        if (int rc = upload_staging(ctx, items, blocks)) return rc;
        ctx->host_us[1] += now_us() - t_stage;
        job.n_items = (uint32_t)items.size();
        job.n_blocks = (uint32_t)blocks.size();
        job.n_src_tris = tri_cursor;
        const int rc = run_frame(ctx, job, out_stats, cull);
        if (rc != SHSB_OK) return rc; // a failed submission leaves the motion history as it was
        frame_used_targets(job.tile_stream, job.frame_no, used, 5);
        // Context::history (pass_pbr_forward.hpp:212-213): committed once the frame is on its way
        if (lit_pass && history_partial)
        {
            ctx->hist_keys.clear();
            ctx->hist_models.clear();
            ctx->hist_index_valid = false;
            ctx->hist_has_duplicates = false;
            ctx->has_prev_frame = false;
        }
        else if (lit_pass)
        {
            // does any key occur twice?  (object ids are the caller's; derived keys can collide as well)
            bool dup = false;
            if (next_keys.size() > 1)
            {
                size_t slots = 16;
                while (slots < next_keys.size() * 2) slots <<= 1;
                ctx->key_set.assign(slots, 0ull); // keys are never 0 (a derived 0 becomes 1; object_id 0 means "derive")
                for (const uint64_t key : next_keys)
                {
                    size_t at = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 17) & (slots - 1);
                    while (ctx->key_set[at] != 0ull && ctx->key_set[at] != key) at = (at + 1) & (slots - 1);
                    if (ctx->key_set[at] == key) { dup = true; break; }
                    ctx->key_set[at] = key;
                }
            }
            ctx->hist_has_duplicates = dup;
            ctx->hist_keys.swap(next_keys);
            ctx->hist_models.swap(next_models);
            ctx->hist_index_valid = false;
            ctx->has_prev_frame = true;
        }
        return rc;
    }
}

// ======================================================================================== C ABI
extern "C" {

SHSB_API const char* shsb_version(void)
{
    return "shsb 0.1 (sm_100a; geometry/light-cull: --fmad=false; raster: exact intrinsics)";
}

SHSB_API int32_t shsb_context_create(int32_t device_ordinal, shsb_ctx* out_ctx)
{
    if (!out_ctx) return SHSB_E_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_ordinal < 0 || device_ordinal >= n)
    {
        cudaGetLastError();
        return SHSB_E_NO_DEVICE;
    }
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return SHSB_E_NO_DEVICE;
    shsb_ctx ctx = new shsb_context_t();
    ctx->device = device_ordinal;
    for (auto& row : ctx->lights_last_user) for (long long& u : row) u = -1;
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (const char* e = std::getenv("SHSB_ARENAS")) ctx->n_arenas = std::min(NUM_ARENAS, std::max(2, std::atoi(e)));
    if (const char* e = std::getenv("SHSB_NO_PIPELINE")) ctx->pipeline = !(e[0] == '1');
    ctx->host_threads = std::min(4, std::max(1, (int)std::thread::hardware_concurrency() / 8));
    if (const char* e = std::getenv("SHSB_HOST_THREADS")) ctx->host_threads = std::min(32, std::max(1, std::atoi(e)));
    if (const char* e = std::getenv("SHSB_FRONT_PRIORITY")) { if (e[0] == '0') prio_greatest = prio_least; }
    bool ok = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_least) == cudaSuccess;
    ctx->tile_streams[0] = ctx->stream;
    for (int i = 1; ok && i < MAX_TILE_STREAMS; ++i)
    {
        ok = cudaStreamCreateWithPriority(&ctx->tile_streams[i], cudaStreamNonBlocking, prio_least) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_main_sync[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (const char* e = std::getenv("SHSB_TILE_STREAMS")) ctx->n_tile_streams = std::min(MAX_TILE_STREAMS, std::max(1, std::atoi(e)));
    // the front end is small and latency-bound: at high priority its CTAs slot in between the tile kernel's
    for (int i = 0; ok && i < NUM_ARENAS; ++i)
    {
        ok = cudaStreamCreateWithPriority(&ctx->front_streams[i], cudaStreamNonBlocking, prio_greatest) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&ctx->cull_streams[i], cudaStreamNonBlocking, prio_greatest) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_lights_up, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_cull_main, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < TILE_DONE_RING; ++i) ok = cudaEventCreateWithFlags(&ctx->ev_tile_done[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < NUM_ARENAS; ++i) ok = cudaEventCreateWithFlags(&ctx->ev_front_done[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < NUM_ARENAS; ++i) ok = cudaEventCreateWithFlags(&ctx->lights_stage_done[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_frame_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_front_sync, cudaEventDisableTiming) == cudaSuccess;
    if (const char* e = std::getenv("SHSB_SHADOW_DIRECT")) ctx->shadow_direct = !(e[0] == '0');
    if (const char* e = std::getenv("SHSB_HIZ")) ctx->hiz = e[0] == '1';
    if (const char* e = std::getenv("SHSB_NO_FAST_TILE")) ctx->fast_tile = e[0] != '1';
    if (const char* e = std::getenv("SHSB_NO_GRAPH")) ctx->use_graph = !(e[0] == '1');
    ok = ok && cudaHostAlloc(&ctx->h_stats, sizeof(DevStats) * STAT_SHARDS, cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc(&ctx->h_overflow, 4 * sizeof(uint32_t), cudaHostAllocMapped) == cudaSuccess;
    if (ok) { std::memset(ctx->h_overflow, 0, 4 * sizeof(uint32_t)); ok = cudaHostGetDevicePointer(&ctx->d_overflow, ctx->h_overflow, 0) == cudaSuccess; }
    ok = ok && cudaMalloc(&ctx->d_srgb_lut, 256 * sizeof(float)) == cudaSuccess;
    for (int i = 0; ok && i < NUM_STAGE_EVENTS; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    for (int i = 0; ok && i < shsb_context_t::STAGE_SLOTS; ++i) ok = cudaEventCreateWithFlags(&ctx->stage_done[i], cudaEventDisableTiming) == cudaSuccess;
    if (ok)
    {
        // srgb_to_linear_rgb (shader/builtin_shaders.hpp:25-31) evaluated by the host libm, as the reference does per tap
        float lut[256];
        for (int i = 0; i < 256; ++i) lut[i] = std::pow((float)i / 255.0f, 2.2f);
        ok = cudaMemcpy(ctx->d_srgb_lut, lut, sizeof(lut), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok)
    {
        delete ctx;
        return SHSB_E_CUDA;
    }
    *out_ctx = ctx;
    return SHSB_OK;
}

SHSB_API int32_t shsb_context_destroy(shsb_ctx ctx)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    delete ctx->host_pool;
    ctx->host_pool = nullptr;
    for (auto& m : ctx->meshes) { cudaFree(m.positions); cudaFree(m.normals); cudaFree(m.uvs); cudaFree(m.indices); }
    for (auto& t : ctx->textures) cudaFree(t.texels);
    for (auto& r : ctx->rts) { cudaFree(r.color); cudaFree(r.depth); cudaFree(r.motion); cudaFree(r.tri_id); cudaFree(r.coverage); }
    cudaFree(ctx->d_meshes.p); cudaFree(ctx->d_textures.p); cudaFree(ctx->d_srgb_lut);
    cudaFree(ctx->d_range_min.p); cudaFree(ctx->d_range_max.p); cudaFree(ctx->d_range_up_min.p); cudaFree(ctx->d_range_up_max.p);
    cudaFree(ctx->d_slice_ndc.p); cudaFree(ctx->d_vis.p); cudaFree(ctx->cluster_lists.counts.p); cudaFree(ctx->cluster_lists.indices.p);
    cudaFree(ctx->d_post_scratch.p); cudaFree(ctx->d_post_luma.p); cudaFree(ctx->d_taa_hist.p); cudaFree(ctx->d_legacy_tris.p);
    cudaFree(ctx->d_sc_bytes.p);
    cudaFree(ctx->d_fd_bytes.p);
    cudaFree(ctx->d_l2_raster.p); cudaFree(ctx->d_l2_box.p); cudaFree(ctx->d_l2_shade.p); cudaFree(ctx->d_l2_rot);
    for (auto& e : ctx->ibls) { cudaFree(e.irradiance); cudaFree(e.prefiltered); }
    for (auto& l : ctx->d_lights) cudaFree(l.p);
    for (auto& l : ctx->d_smlights) cudaFree(l.p);
    for (auto& l : ctx->h_lights) cudaFreeHost(l.p);
    for (cudaEvent_t e : ctx->lights_stage_done) if (e) cudaEventDestroy(e);
    for (LightLists& L : ctx->lists) { cudaFree(L.counts.p); cudaFree(L.indices.p); cudaFree(L.scratch.p); }
    for (Arena& A : ctx->arena)
    {
        cudaFree(A.d_draw.p); cudaFree(A.d_rrecs.p); cudaFree(A.d_srecs.p); cudaFree(A.d_clipq.p); cudaFree(A.d_hdr.p);
        cudaFree(A.d_tile_order.p); cudaFree(A.d_tile_offset.p); cudaFree(A.d_tile_fill.p); cudaFree(A.d_tile_list.p);
    }
    cudaFreeHost(ctx->h_stats);
    cudaFreeHost(ctx->h_overflow);
    for (cudaEvent_t e : ctx->ev_tile_done) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_front_done) if (e) cudaEventDestroy(e);
    if (ctx->ev_lights_up) cudaEventDestroy(ctx->ev_lights_up);
    if (ctx->ev_cull_main) cudaEventDestroy(ctx->ev_cull_main);
    for (int i = 0; i < shsb_context_t::STAGE_SLOTS; ++i)
    {
        cudaFreeHost(ctx->h_draw[i].p);
        if (ctx->stage_done[i]) cudaEventDestroy(ctx->stage_done[i]);
    }
    for (int i = 0; i < NUM_STAGE_EVENTS; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (cudaEvent_t e : ctx->timing_ev) cudaEventDestroy(e);
    for (auto& row : ctx->graph_exec) for (cudaGraphExec_t e : row) if (e) cudaGraphExecDestroy(e);
    for (cudaEvent_t e : ctx->ev_fork) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_join) if (e) cudaEventDestroy(e);
    if (ctx->ev_frame_done) cudaEventDestroy(ctx->ev_frame_done);
    if (ctx->ev_front_sync) cudaEventDestroy(ctx->ev_front_sync);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->copy_stream2) { cudaStreamSynchronize(ctx->copy_stream2); cudaStreamDestroy(ctx->copy_stream2); }
    for (auto& r : ctx->rts) if (r.read_done) cudaEventDestroy(r.read_done);
    for (cudaStream_t st : ctx->cull_streams) if (st) cudaStreamDestroy(st);
    for (cudaStream_t st : ctx->front_streams) if (st) cudaStreamDestroy(st);
    for (size_t i = 0; i < ctx->gathers.size(); ++i) if (ctx->gathers[i].live) shsb_gather_destroy(ctx, (shsb_gather)(i + 1));
    if (ctx->gather_stream) cudaStreamDestroy(ctx->gather_stream);
    if (ctx->h_gather_timeout) cudaFreeHost(ctx->h_gather_timeout);
    for (int i = 1; i < MAX_TILE_STREAMS; ++i)
    {
        if (ctx->tile_streams[i]) cudaStreamDestroy(ctx->tile_streams[i]);
        if (ctx->ev_main_sync[i]) cudaEventDestroy(ctx->ev_main_sync[i]);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SHSB_OK;
}

SHSB_API const char* shsb_last_error_string(shsb_ctx ctx) { return ctx ? ctx->error.c_str() : "null context"; }

SHSB_API int32_t shsb_sync(shsb_ctx ctx)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    for (cudaStream_t st : ctx->front_streams) CK(cudaStreamSynchronize(st));
    for (cudaStream_t st : ctx->tile_streams) CK(cudaStreamSynchronize(st));
    if (ctx->gather_stream) CK(cudaStreamSynchronize(ctx->gather_stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->copy_stream2));
    if (ctx->h_gather_timeout && *ctx->h_gather_timeout) return fail(ctx, SHSB_E_TIMEOUT, "a frame-assembly wait timed out (a rank never committed / the root never released a step)");
    return check_async_overflow(ctx);
}

SHSB_API int32_t shsb_fence(shsb_ctx ctx)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    join_side_streams(ctx);                  // main stream: behind every frame submitted so far
    ctx->fence_seq = ++ctx->main_seq;        // side streams: their next frame orders itself behind the main stream's tail
    CK(cudaGetLastError());
    return SHSB_OK;
}

SHSB_API int32_t shsb_stream(shsb_ctx ctx, void** out_stream)
{
    if (!ctx || !out_stream) return SHSB_E_INVALID_ARGUMENT;
    if (int rc = shsb_fence(ctx)) return rc;
    *out_stream = (void*)ctx->stream;
    return SHSB_OK;
}

SHSB_API int32_t shsb_set_tile_streams(shsb_ctx ctx, int32_t n)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (n < 1 || n > MAX_TILE_STREAMS) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "render-stream count must be 1..%d", MAX_TILE_STREAMS);
    ctx->n_tile_streams = n;
    for (RtSlot& r : ctx->rts) r.affinity = -1;
    ctx->affinity_next = 0;
    return SHSB_OK;
}

SHSB_API int32_t shsb_last_tile_kernel(shsb_ctx ctx, int32_t* out_mode)
{
    if (!ctx || !out_mode) return SHSB_E_INVALID_ARGUMENT;
    *out_mode = ctx->last_tile_mode;
    return SHSB_OK;
}

SHSB_API int32_t shsb_launch_count(shsb_ctx ctx, uint64_t* out_count)
{
    if (!ctx || !out_count) return SHSB_E_INVALID_ARGUMENT;
    *out_count = ctx->launches;
    return SHSB_OK;
}

// ---------------------------------------------------------------------------------------- resources
SHSB_API int32_t shsb_mesh_upload(shsb_ctx ctx, const float* positions, uint32_t n_positions, const float* normals, uint32_t n_normals,
                                  const float* uvs, uint32_t n_uvs, const uint32_t* indices, uint32_t n_indices, shsb_mesh* out_mesh)
{
    if (!ctx || !out_mesh) return SHSB_E_INVALID_ARGUMENT;
    if (n_positions && !positions) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "positions is null");
    if ((n_normals && !normals) || (n_uvs && !uvs) || (n_indices && !indices)) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "null attribute stream with non-zero count");
    CK(cudaSetDevice(ctx->device));
    MeshSlot m;
    m.live = true;
    m.n_positions = n_positions; m.n_normals = n_normals; m.n_uvs = n_uvs; m.n_indices = n_indices;
    auto up = [&](auto** dst, const void* src, size_t bytes) -> cudaError_t {
        if (!bytes) { *dst = nullptr; return cudaSuccess; }
        cudaError_t e = cudaMalloc((void**)dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    };
    // a failing step releases what the earlier ones allocated
    auto step = [&](cudaError_t e) -> int {
        if (e == cudaSuccess) return SHSB_OK;
        cudaFree(m.positions); cudaFree(m.normals); cudaFree(m.uvs); cudaFree(m.indices);
        cudaGetLastError();
        return fail(ctx, e == cudaErrorMemoryAllocation ? SHSB_E_OUT_OF_MEMORY : SHSB_E_CUDA, "mesh upload: %s", cudaGetErrorString(e));
    };
    if (int rc = step(up(&m.positions, positions, (size_t)n_positions * 12))) return rc;
    if (int rc = step(up(&m.normals, normals, (size_t)n_normals * 12))) return rc;
    if (int rc = step(up(&m.uvs, uvs, (size_t)n_uvs * 8))) return rc;
    if (int rc = step(up(&m.indices, indices, (size_t)n_indices * 4))) return rc;
    if (int rc = step(cudaStreamSynchronize(ctx->stream))) return rc;
    if (n_positions)
    {
        hm::vec3f mn{3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx{-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
        for (uint32_t v = 0; v < n_positions; ++v)
        {
            const float* p = positions + (size_t)v * 3;
            mn = {hm::gmin(mn.x, p[0]), hm::gmin(mn.y, p[1]), hm::gmin(mn.z, p[2])};
            mx = {hm::gmax(mx.x, p[0]), hm::gmax(mx.y, p[1]), hm::gmax(mx.z, p[2])};
        }
        m.bmin = mn; m.bmax = mx;
    }
    ctx->meshes.push_back(m);
    ctx->mesh_table_dirty = true;
    *out_mesh = (shsb_mesh)ctx->meshes.size();
    return SHSB_OK;
}

SHSB_API int32_t shsb_mesh_destroy(shsb_ctx ctx, shsb_mesh mesh)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    MeshSlot* m = get_mesh(ctx, mesh);
    if (!m) return fail(ctx, SHSB_E_INVALID_HANDLE, "mesh handle %u is not live", mesh);
    sync_all(ctx);
    cudaFree(m->positions); cudaFree(m->normals); cudaFree(m->uvs); cudaFree(m->indices);
    *m = MeshSlot{};
    ctx->mesh_table_dirty = true;
    return SHSB_OK;
}

SHSB_API int32_t shsb_texture_upload(shsb_ctx ctx, const uint8_t* rgba, int32_t w, int32_t h, shsb_tex* out_tex)
{
    if (!ctx || !out_tex || !rgba || w <= 0 || h <= 0) return ctx ? fail(ctx, SHSB_E_INVALID_ARGUMENT, "bad texture arguments") : SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    TexSlot t;
    t.live = true; t.w = w; t.h = h;
    CK(cudaMalloc(&t.texels, (size_t)w * h * 4));
    CK(cudaMemcpyAsync(t.texels, rgba, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->textures.push_back(t);
    ctx->tex_table_dirty = true;
    *out_tex = (shsb_tex)ctx->textures.size();
    return SHSB_OK;
}

SHSB_API int32_t shsb_texture_destroy(shsb_ctx ctx, shsb_tex tex)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (tex == 0 || tex > ctx->textures.size() || !ctx->textures[tex - 1].live) return fail(ctx, SHSB_E_INVALID_HANDLE, "texture handle %u is not live", tex);
    sync_all(ctx);
    cudaFree(ctx->textures[tex - 1].texels);
    ctx->textures[tex - 1] = TexSlot{};
    ctx->tex_table_dirty = true;
    return SHSB_OK;
}

// ---------------------------------------------------------------------------------------- on-disk fixtures
SHSB_API int32_t shsb_mesh_load_obj(shsb_ctx ctx, const char* path, shsb_mesh* out_mesh)
{
    if (!ctx || !path || !out_mesh) return SHSB_E_INVALID_ARGUMENT;
    shsb_loaders::ObjMesh m;
    const std::string err = shsb_loaders::load_obj(path, m);
    if (!err.empty()) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "%s", err.c_str());
    const uint32_t nv = (uint32_t)(m.positions.size() / 3);
    return shsb_mesh_upload(ctx, m.positions.data(), nv, m.normals.data(), nv, m.uvs.data(), nv, m.indices.data(), (uint32_t)m.indices.size(), out_mesh);
}

SHSB_API int32_t shsb_texture_load_png(shsb_ctx ctx, const char* path, int32_t flip_y, shsb_tex* out_tex)
{
    if (!ctx || !path || !out_tex) return SHSB_E_INVALID_ARGUMENT;
    std::vector<unsigned char> rgba;
    int w = 0, h = 0;
    const std::string err = shsb_loaders::load_png(path, flip_y != 0, rgba, w, h);
    if (!err.empty()) return fail(ctx, err.find("not supported") != std::string::npos || err.find("unsupported") != std::string::npos ? SHSB_E_UNSUPPORTED : SHSB_E_INVALID_ARGUMENT, "%s", err.c_str());
    return shsb_texture_upload(ctx, rgba.data(), w, h, out_tex);
}

SHSB_API int32_t shsb_mesh_info(shsb_ctx ctx, shsb_mesh mesh, uint32_t out_counts4[4])
{
    if (!ctx || !out_counts4) return SHSB_E_INVALID_ARGUMENT;
    MeshSlot* m = get_mesh(ctx, mesh);
    if (!m) return fail(ctx, SHSB_E_INVALID_HANDLE, "mesh handle %u is not live", mesh);
    out_counts4[0] = m->n_positions; out_counts4[1] = m->n_normals; out_counts4[2] = m->n_uvs; out_counts4[3] = m->n_indices;
    return SHSB_OK;
}

SHSB_API int32_t shsb_mesh_download(shsb_ctx ctx, shsb_mesh mesh, float* positions, float* normals, float* uvs, uint32_t* indices)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    MeshSlot* m = get_mesh(ctx, mesh);
    if (!m) return fail(ctx, SHSB_E_INVALID_HANDLE, "mesh handle %u is not live", mesh);
    CK(cudaSetDevice(ctx->device));
    if (positions && m->n_positions) CK(cudaMemcpy(positions, m->positions, (size_t)m->n_positions * 12, cudaMemcpyDeviceToHost));
    if (normals && m->n_normals) CK(cudaMemcpy(normals, m->normals, (size_t)m->n_normals * 12, cudaMemcpyDeviceToHost));
    if (uvs && m->n_uvs) CK(cudaMemcpy(uvs, m->uvs, (size_t)m->n_uvs * 8, cudaMemcpyDeviceToHost));
    if (indices && m->n_indices) CK(cudaMemcpy(indices, m->indices, (size_t)m->n_indices * 4, cudaMemcpyDeviceToHost));
    return SHSB_OK;
}

SHSB_API int32_t shsb_texture_info(shsb_ctx ctx, shsb_tex tex, int32_t out_wh2[2])
{
    if (!ctx || !out_wh2) return SHSB_E_INVALID_ARGUMENT;
    if (tex == 0 || tex > ctx->textures.size() || !ctx->textures[tex - 1].live) return fail(ctx, SHSB_E_INVALID_HANDLE, "texture handle %u is not live", tex);
    out_wh2[0] = ctx->textures[tex - 1].w; out_wh2[1] = ctx->textures[tex - 1].h;
    return SHSB_OK;
}

SHSB_API int32_t shsb_texture_download(shsb_ctx ctx, shsb_tex tex, uint8_t* rgba, size_t bytes)
{
    if (!ctx || !rgba) return SHSB_E_INVALID_ARGUMENT;
    if (tex == 0 || tex > ctx->textures.size() || !ctx->textures[tex - 1].live) return fail(ctx, SHSB_E_INVALID_HANDLE, "texture handle %u is not live", tex);
    const TexSlot& t = ctx->textures[tex - 1];
    if (bytes != (size_t)t.w * t.h * 4) return fail(ctx, SHSB_E_SIZE_MISMATCH, "texture is %zu bytes, caller passed %zu", (size_t)t.w * t.h * 4, bytes);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(rgba, t.texels, bytes, cudaMemcpyDeviceToHost));
    return SHSB_OK;
}

// ---------------------------------------------------------------------------------------- render targets
SHSB_API int32_t shsb_rt_create(shsb_ctx ctx, int32_t kind, int32_t w, int32_t h, float zn, float zf, shsb_rt* out_rt)
{
    if (!ctx || !out_rt) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (w <= 0 || h <= 0 || kind < SHSB_RT_COLOR_HDR || kind > SHSB_RT_SHADOW) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "bad render-target description");
    RtSlot r;
    r.live = true; r.kind = kind; r.w = w; r.h = h; r.zn = zn; r.zf = zf;
    const size_t n = (size_t)w * h;
    // a failing step releases what the earlier ones allocated
    auto step = [&](cudaError_t e, const char* what) -> int {
        if (e == cudaSuccess) return SHSB_OK;
        cudaFree(r.color); cudaFree(r.depth); cudaFree(r.motion);
        cudaGetLastError();
        return fail(ctx, e == cudaErrorMemoryAllocation ? SHSB_E_OUT_OF_MEMORY : SHSB_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    };
    if (kind == SHSB_RT_COLOR_HDR)
    {
        if (int rc = step(cudaMalloc(&r.color, n * 16), "render target colour plane")) return rc;
        launch_fill_f4((float4*)r.color, make_float4(0, 0, 0, 1), n, ctx->stream, &ctx->launches); // RT_ColorHDR ctor, rt_types.hpp:85
    }
    else if (kind == SHSB_RT_COLOR_LDR)
    {
        if (int rc = step(cudaMalloc(&r.color, n * 4), "render target colour plane")) return rc;
        launch_fill_u32((uint32_t*)r.color, 0xFF000000u, n, ctx->stream, &ctx->launches);          // {0,0,0,255}, rt_types.hpp:68
    }
    else if (kind == SHSB_RT_DEPTH_MOTION)
    {
        if (int rc = step(cudaMalloc(&r.depth, n * 4), "render target depth plane")) return rc;
        if (int rc = step(cudaMalloc(&r.motion, n * 8), "render target motion plane")) return rc;
        launch_fill_u32((uint32_t*)r.depth, 0x3F800000u, n, ctx->stream, &ctx->launches);          // depth 1.0, rt_types.hpp:144
        if (int rc = step(cudaMemsetAsync(r.motion, 0, n * 8, ctx->stream), "render target motion clear")) return rc;
    }
    else
    {
        if (int rc = step(cudaMalloc(&r.depth, n * 4), "render target depth plane")) return rc;
        launch_fill_u32((uint32_t*)r.depth, 0x3F800000u, n, ctx->stream, &ctx->launches);          // rt_shadow.hpp:26-29
    }
    if (int rc = step(cudaGetLastError(), "render target clear")) return rc;
    ctx->rts.push_back(r);
    *out_rt = (shsb_rt)ctx->rts.size();
    return SHSB_OK;
}

SHSB_API int32_t shsb_rt_destroy(shsb_ctx ctx, shsb_rt rt)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* r = peek_rt(ctx, rt);
    if (!r) return fail(ctx, SHSB_E_INVALID_HANDLE, "render target %u is not live", rt);
    sync_all(ctx);
    cudaFree(r->color); cudaFree(r->depth); cudaFree(r->motion); cudaFree(r->tri_id); cudaFree(r->coverage);
    if (r->read_done) cudaEventDestroy(r->read_done);
    *r = RtSlot{};
    return SHSB_OK;
}

SHSB_API int32_t shsb_rt_clear(shsb_ctx ctx, shsb_rt rt, int32_t plane, const void* value)
{
    if (!ctx || !value) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* r = get_rt(ctx, rt);
    if (!r) return fail(ctx, SHSB_E_INVALID_HANDLE, "render target %u is not live", rt);
    void* p = nullptr;
    const size_t bytes = plane_bytes(*r, plane, &p);
    if (!bytes) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "render target %u has no plane %d", rt, plane);
    const size_t n = (size_t)r->w * r->h;
    if (bytes == n * 16) { float4 v; std::memcpy(&v, value, 16); launch_fill_f4((float4*)p, v, n, ctx->stream, &ctx->launches); }
    else if (bytes == n * 8) { uint32_t v[2]; std::memcpy(v, value, 8); if (v[0] == v[1]) launch_fill_u32((uint32_t*)p, v[0], n * 2, ctx->stream, &ctx->launches); else return fail(ctx, SHSB_E_UNSUPPORTED, "motion clear needs x == y"); }
    else { uint32_t v; std::memcpy(&v, value, 4); launch_fill_u32((uint32_t*)p, v, n, ctx->stream, &ctx->launches); }
    if (plane == SHSB_PLANE_MOTION) r->motion_dirty = true;
    CK(cudaGetLastError());
    return SHSB_OK;
}

SHSB_API int32_t shsb_rt_upload(shsb_ctx ctx, shsb_rt rt, int32_t plane, const void* src, size_t bytes)
{
    if (!ctx || !src) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* r = get_rt(ctx, rt);
    if (!r) return fail(ctx, SHSB_E_INVALID_HANDLE, "render target %u is not live", rt);
    void* p = nullptr;
    const size_t want = plane_bytes(*r, plane, &p);
    if (!want) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "render target %u has no plane %d", rt, plane);
    if (bytes != want) return fail(ctx, SHSB_E_SIZE_MISMATCH, "plane is %zu bytes, caller passed %zu", want, bytes);
    CK(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (plane == SHSB_PLANE_MOTION) r->motion_dirty = true;
    return SHSB_OK;
}

SHSB_API int32_t shsb_rt_download(shsb_ctx ctx, shsb_rt rt, int32_t plane, void* dst, size_t bytes)
{
    if (!ctx || !dst) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* r = get_rt(ctx, rt);
    if (!r) return fail(ctx, SHSB_E_INVALID_HANDLE, "render target %u is not live", rt);
    void* p = nullptr;
    const size_t want = plane_bytes(*r, plane, &p);
    if (!want) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "render target %u has no plane %d", rt, plane);
    if (bytes != want) return fail(ctx, SHSB_E_SIZE_MISMATCH, "plane is %zu bytes, caller passed %zu", want, bytes);
    CK(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return check_async_overflow(ctx); // the pixels just read may come from an asynchronous frame that dropped geometry
}

SHSB_API int32_t shsb_rt_download_async(shsb_ctx ctx, shsb_rt rt, int32_t plane, void* dst_pinned, size_t bytes)
{
    if (!ctx || !dst_pinned) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* r = peek_rt(ctx, rt);
    if (!r) return fail(ctx, SHSB_E_INVALID_HANDLE, "render target %u is not live", rt);
    void* p = nullptr;
    const size_t want = plane_bytes(*r, plane, &p);
    if (!want) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "render target %u has no plane %d", rt, plane);
    if (bytes != want) return fail(ctx, SHSB_E_SIZE_MISMATCH, "plane is %zu bytes, caller passed %zu", want, bytes);
    if (!r->read_done) CK(cudaEventCreateWithFlags(&r->read_done, cudaEventDisableTiming));
    // the copy stream waits for whatever wrote the target last, then copies while later frames render
    cudaStream_t cs = (ctx->copy_flip ^= 1) ? ctx->copy_stream : ctx->copy_stream2;
    if (r->frame_is_last && r->last_frame >= 0)
    {
        // a frame's tile kernel, which already recorded an event on its render stream: no further command on a (critical) render stream
        CK(cudaStreamWaitEvent(cs, ctx->ev_tile_done[r->last_frame % TILE_DONE_RING], 0));
    }
    else
    {
        // main-stream work touched it since (that work joined any side-stream frame when it looked the target up)
        CK(cudaEventRecord(ctx->ev_frame_done, ctx->stream));
        CK(cudaStreamWaitEvent(cs, ctx->ev_frame_done, 0));
    }
    if (r->read_pending) CK(cudaStreamWaitEvent(cs, r->read_done, 0)); // an earlier reader on another stream: the one event then covers both
    CK(cudaMemcpyAsync(dst_pinned, p, bytes, cudaMemcpyDeviceToHost, cs));
    CK(cudaEventRecord(r->read_done, cs));
    r->read_pending = true;
    return SHSB_OK;
}

SHSB_API int32_t shsb_rt_device_ptr(shsb_ctx ctx, shsb_rt rt, int32_t plane, void** out_ptr, size_t* out_bytes)
{
    if (!ctx || !out_ptr) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* r = get_rt(ctx, rt);
    if (!r) return fail(ctx, SHSB_E_INVALID_HANDLE, "render target %u is not live", rt);
    void* p = nullptr;
    const size_t bytes = plane_bytes(*r, plane, &p);
    if (!bytes) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "render target %u has no plane %d", rt, plane);
    *out_ptr = p;
    if (out_bytes) *out_bytes = bytes;
    return SHSB_OK;
}

// ---------------------------------------------------------------------------------------- host helpers
SHSB_API int32_t shsb_model_from_transform(const ShsbTransform* tr, float out_model[16])
{
    if (!tr || !out_model) return SHSB_E_INVALID_ARGUMENT;
    hm::store(hm::model_from_transform(tr->pos, tr->rot_euler, tr->scl), out_model);
    return SHSB_OK;
}

SHSB_API int32_t shsb_camera_viewproj(const float eye[3], const float target[3], const float up[3],
                                      float fovy_radians, float aspect, float znear, float zfar, float out_viewproj[16])
{
    if (!eye || !target || !up || !out_viewproj) return SHSB_E_INVALID_ARGUMENT;
    const hm::mat4f view = hm::look_at_lh({eye[0], eye[1], eye[2]}, {target[0], target[1], target[2]}, {up[0], up[1], up[2]});
    const hm::mat4f proj = hm::perspective_lh_no(fovy_radians, aspect, znear, zfar);
    hm::store(hm::mul(proj, view), out_viewproj);
    return SHSB_OK;
}

// ---------------------------------------------------------------------------------------- legacy tile-job variant (row L1)
SHSB_API int32_t shsb_legacy_camera(const float position[3], float horizontal_angle_deg, float vertical_angle_deg, float out_view[16], float out_proj[16])
{
    if (!position || !out_view || !out_proj) return SHSB_E_INVALID_ARGUMENT;
    // Camera3D::update, shs_renderer.hpp:1223-1236.  The reference calls the unqualified C functions cos / sin on float
    // arguments: with <cmath> alone those are the DOUBLE functions, the products are formed in double and narrowed by the
    // glm::vec3 constructor (checked bit for bit against the compiled reference in tests/test_legacy_cpu.py).
    const float deg = 0.01745329251994329576923690768489f;
    const float va = vertical_angle_deg * deg, ha = horizontal_angle_deg * deg;
    hm::vec3f dir{(float)(std::cos((double)va) * std::sin((double)ha)), (float)std::sin((double)va), (float)(std::cos((double)va) * std::cos((double)ha))};
    dir = hm::normalize(dir);
    const hm::vec3f world_up{0.0f, 1.0f, 0.0f};
    const hm::vec3f right = hm::normalize(hm::cross(world_up, dir));
    const hm::vec3f up = hm::normalize(hm::cross(dir, right));
    const hm::vec3f pos{position[0], position[1], position[2]};
    // field_of_view is a run-time member in the reference: tan() must be libm's at run time, not the compiler's folded constant
    volatile float field_of_view = 60.0f;
    hm::store(hm::perspective_lh_no(field_of_view * deg, 4.0f / 3.0f, 0.1f, 1000.0f), out_proj);
    hm::store(hm::look_at_lh(pos, pos + dir, up), out_view);
    return SHSB_OK;
}

SHSB_API int32_t shsb_legacy_world_matrix(const float position[3], const float scale[3], float rotation_angle_deg, float out_model[16])
{
    if (!position || !scale || !out_model) return SHSB_E_INVALID_ARGUMENT;
    const hm::mat4f t = hm::translate(hm::identity(), {position[0], position[1], position[2]});
    const hm::mat4f r = hm::rotate(hm::identity(), rotation_angle_deg * 0.01745329251994329576923690768489f, {0.0f, 1.0f, 0.0f});
    const hm::mat4f sc = hm::scale(hm::identity(), {scale[0], scale[1], scale[2]});
    hm::store(hm::mul(hm::mul(t, r), sc), out_model);
    return SHSB_OK;
}

SHSB_API int32_t shsb_legacy_mvp(const float proj[16], const float view[16], const float model[16], float out_mvp[16])
{
    if (!proj || !view || !model || !out_mvp) return SHSB_E_INVALID_ARGUMENT;
    hm::store(hm::mul(hm::mul(hm::load(proj), hm::load(view)), hm::load(model)), out_mvp);
    return SHSB_OK;
}

SHSB_API int32_t shsb_legacy_draw_blinn_phong(shsb_ctx ctx, shsb_mesh mesh_h, const ShsbLegacyUniforms* u, shsb_rt canvas_rt, shsb_rt zbuffer_rt)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (!u) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "uniforms are null");
    CK(cudaSetDevice(ctx->device));
    const MeshSlot* mesh = get_mesh(ctx, mesh_h);
    if (!mesh) return fail(ctx, SHSB_E_INVALID_HANDLE, "mesh %u is not live", mesh_h);
    RtSlot* canvas = get_rt(ctx, canvas_rt, SHSB_RT_COLOR_LDR);
    if (!canvas) return fail(ctx, SHSB_E_INVALID_HANDLE, "canvas_ldr is not a live RT_ColorLDR");
    RtSlot* zb = get_rt(ctx, zbuffer_rt);
    if (!zb || !zb->depth || (zb->kind != SHSB_RT_SHADOW && zb->kind != SHSB_RT_DEPTH_MOTION))
        return fail(ctx, SHSB_E_INVALID_HANDLE, "zbuffer is not a live target with a depth plane (RT_ShadowDepth / RT_ColorDepthMotion)");
    if (zb->w != canvas->w || zb->h != canvas->h) return fail(ctx, SHSB_E_SIZE_MISMATCH, "canvas is %dx%d, z-buffer %dx%d", canvas->w, canvas->h, zb->w, zb->h);
    if (u->job_tile_w < 0 || u->job_tile_h < 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "job tile size must be positive (0 = the demo's 80)");
    if (mesh->n_normals < mesh->n_positions)
        return fail(ctx, SHSB_E_INVALID_ARGUMENT, "the legacy vertex shader needs a normal per position (ModelGeometry emits both streams)");
    LegacyDraw d{};
    d.positions = mesh->positions;
    d.normals = mesh->normals;
    d.indices = mesh->n_indices ? mesh->indices : nullptr;
    d.n_positions = mesh->n_positions;
    d.n_normals = mesh->n_normals;
    d.n_tris = (mesh->n_indices ? mesh->n_indices : mesh->n_positions) / 3;
    d.W = canvas->w; d.H = canvas->h;
    d.job_w = u->job_tile_w ? u->job_tile_w : 80;
    d.job_h = u->job_tile_h ? u->job_tile_h : 80;
    std::memcpy(d.mvp, u->mvp, 64);
    std::memcpy(d.model, u->model, 64);
    // glm::mat3(glm::transpose(glm::inverse(u.model))) (:55): the 4x4 cofactor inverse, transposed, upper-left 3x3
    const hm::mat4f inv = hm::inverse(hm::load(u->model));
    const float* c = reinterpret_cast<const float*>(&inv);
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row) d.normal_matrix[col * 3 + row] = c[row * 4 + col]; // transpose(inv)[col][row] = inv[row][col]
    std::memcpy(d.light_dir, u->light_dir, 12);
    std::memcpy(d.camera_pos, u->camera_pos, 12);
    std::memcpy(d.color, u->color, 4);
    if (d.n_tris == 0) return SHSB_OK;
    wait_pending_read(ctx, canvas);
    wait_pending_read(ctx, zb);
    if (int rc = ensure_dev(ctx, ctx->d_legacy_tris, d.n_tris)) return rc;
    launch_legacy_draw(d, ctx->d_legacy_tris.p, (uchar4*)canvas->color, zb->depth, ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
    return SHSB_OK;
}


// ---------------------------------------------------------------------------------------- legacy render-target demos (rows L2, L3)
namespace
{
    int legacy2_fill_mesh(shsb_ctx ctx, shsb_mesh mesh_h, bool camera, l2::Draw& d)
    {
        const MeshSlot* mesh = get_mesh(ctx, mesh_h);
        if (!mesh) return fail(ctx, SHSB_E_INVALID_HANDLE, "mesh %u is not live", mesh_h);
        if (camera && mesh->n_normals < mesh->n_positions)
            return fail(ctx, SHSB_E_INVALID_ARGUMENT, "the legacy vertex shader needs a normal per position (ModelGeometry emits both streams)");
        d.positions = mesh->positions;
        d.normals = mesh->normals;
        d.uvs = mesh->n_uvs ? mesh->uvs : nullptr;
        d.indices = mesh->n_indices ? mesh->indices : nullptr;
        d.n_positions = mesh->n_positions;
        d.n_normals = mesh->n_normals;
        d.n_uvs = mesh->n_uvs;
        d.n_tris = (mesh->n_indices ? mesh->n_indices : mesh->n_positions) / 3;
        return SHSB_OK;
    }

    int legacy2_launch(shsb_ctx ctx, const l2::Draw& d, uchar4* canvas, float* zbuf, float2* velocity)
    {
        if (d.n_tris == 0) return SHSB_OK;
        const size_t slots = legacy2_slots(d);
        if (int rc = ensure_dev(ctx, ctx->d_l2_raster, slots)) return rc;
        if (int rc = ensure_dev(ctx, ctx->d_l2_box, slots)) return rc;
        if (int rc = ensure_dev(ctx, ctx->d_l2_shade, d.mode == l2::MODE_SHADOW ? 1 : slots)) return rc;
        launch_legacy2_draw(d, ctx->d_l2_raster.p, ctx->d_l2_box.p, ctx->d_l2_shade.p, canvas, zbuf, velocity, ctx->stream, &ctx->launches);
        CK(cudaGetLastError());
        return SHSB_OK;
    }

    // rotate2 (hello_shadow_mapping_soft.cpp:320-325) turns the PCSS kernels by ang = hash01(seed) * 6.2831853f with std::cos / std::sin
    // of a float, i.e. the platform libm's cosf / sinf, which no device function reproduces bit for bit (glibc's are not correctly
    // rounded, and its x86-64 build picks an FMA or a non-FMA variant at load time).  hash01 has a 24-bit numerator, so there are 2^24
    // angles: the host evaluates the SAME libm calls the reference makes for all of them once per context (a few hundred ms over the
    // host's threads, 128 MB on the device) and the kernels look the pair up (legacy2_core.cuh: rotation()).
    int ensure_l2_rotation(shsb_ctx ctx)
    {
        if (ctx->d_l2_rot) return SHSB_OK;
        constexpr uint32_t N = 0x01000000u;
        std::vector<float> table((size_t)N * 2);
        const unsigned hw = std::thread::hardware_concurrency();
        const unsigned n_threads = std::max(1u, std::min(hw ? hw : 1u, 32u));
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_threads; ++t)
            pool.emplace_back([&table, t, n_threads]() {
                const uint32_t lo = (uint32_t)((uint64_t)N * t / n_threads), hi = (uint32_t)((uint64_t)N * (t + 1) / n_threads);
                for (uint32_t k = lo; k < hi; ++k)
                {
                    const float ang = (float(k) / float(0x01000000u)) * 6.2831853f; // hash01(seed) * 6.2831853f, :316-319, :360
                    table[2 * (size_t)k] = std::sin(ang);
                    table[2 * (size_t)k + 1] = std::cos(ang);
                }
            });
        for (auto& th : pool) th.join();
        float* dev = nullptr;
        if (cudaMalloc((void**)&dev, table.size() * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return fail(ctx, SHSB_E_OUT_OF_MEMORY, "PCSS rotation table (128 MB)"); }
        if (cudaMemcpy(dev, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return fail(ctx, SHSB_E_CUDA, "PCSS rotation table upload"); }
        ctx->d_l2_rot = dev;
        return SHSB_OK;
    }

    // the parts of a camera-pass draw that the soft-shadow and the PBR demo share
    int legacy2_fill_camera(shsb_ctx ctx, shsb_mesh mesh_h, const ShsbLegacy2Uniforms* u, shsb_rt shadow_rt, shsb_rt canvas_rt, shsb_rt zbuffer_rt,
                            bool need_motion, l2::Draw& d, RtSlot*& canvas, RtSlot*& zb)
    {
        if (!u) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "uniforms are null");
        if (int rc = legacy2_fill_mesh(ctx, mesh_h, true, d)) return rc;
        canvas = get_rt(ctx, canvas_rt, SHSB_RT_COLOR_LDR);
        if (!canvas) return fail(ctx, SHSB_E_INVALID_HANDLE, "canvas_ldr is not a live RT_ColorLDR");
        zb = get_rt(ctx, zbuffer_rt);
        if (!zb || !zb->depth || (zb->kind != SHSB_RT_DEPTH_MOTION && (need_motion || zb->kind != SHSB_RT_SHADOW)))
            return fail(ctx, SHSB_E_INVALID_HANDLE, need_motion ? "depth_motion is not a live RT_ColorDepthMotion" : "zbuffer is not a live target with a depth plane (RT_ShadowDepth / RT_ColorDepthMotion)");
        if (zb->w != canvas->w || zb->h != canvas->h) return fail(ctx, SHSB_E_SIZE_MISMATCH, "canvas is %dx%d, z-buffer %dx%d", canvas->w, canvas->h, zb->w, zb->h);
        if (u->job_tile_w < 0 || u->job_tile_h < 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "job tile size must be positive (0 = the demo's 160)");
        d.W = canvas->w; d.H = canvas->h;
        d.job_w = u->job_tile_w ? u->job_tile_w : 160;
        d.job_h = u->job_tile_h ? u->job_tile_h : 160;
        std::memcpy(d.mvp, u->mvp, 64); std::memcpy(d.prev_mvp, u->prev_mvp, 64); std::memcpy(d.model, u->model, 64); std::memcpy(d.mv, u->mv, 64);
        std::memcpy(d.normal_mat, u->normal_mat, 36); std::memcpy(d.light_vp, u->light_vp, 64);
        std::memcpy(d.light_dir, u->light_dir_world, 12); std::memcpy(d.camera_pos, u->camera_pos, 12); std::memcpy(d.color, u->base_color, 4);
        d.use_texture = u->use_texture;
        if (u->albedo)
        {
            if (u->albedo > ctx->textures.size() || !ctx->textures[u->albedo - 1].live) return fail(ctx, SHSB_E_INVALID_HANDLE, "texture handle %u is not live", u->albedo);
            const TexSlot& t = ctx->textures[u->albedo - 1];
            d.tex = reinterpret_cast<const unsigned char*>(t.texels); d.tex_w = t.w; d.tex_h = t.h;
        }
        if (shadow_rt)
        {
            RtSlot* sh = get_rt(ctx, shadow_rt, SHSB_RT_SHADOW);
            if (!sh) return fail(ctx, SHSB_E_INVALID_HANDLE, "shadow_map is not a live RT_ShadowDepth");
            if (sh == zb) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "the shadow map and the z-buffer are the same target");
            wait_pending_read(ctx, sh);
            d.shadow = sh->depth; d.sm_w = sh->w; d.sm_h = sh->h;
        }
        return SHSB_OK;
    }
}

SHSB_API int32_t shsb_legacy2_shadow_draw(shsb_ctx ctx, shsb_mesh mesh_h, const float model[16], const float light_vp[16], int32_t job_tile_w, int32_t job_tile_h,
                                          shsb_rt shadow_rt)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (!model || !light_vp) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "model / light_vp are null");
    if (job_tile_w < 0 || job_tile_h < 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "job tile size must be positive (0 = the demo's 160)");
    CK(cudaSetDevice(ctx->device));
    l2::Draw d{};
    d.mode = l2::MODE_SHADOW;
    if (int rc = legacy2_fill_mesh(ctx, mesh_h, false, d)) return rc;
    RtSlot* sh = get_rt(ctx, shadow_rt, SHSB_RT_SHADOW);
    if (!sh) return fail(ctx, SHSB_E_INVALID_HANDLE, "shadow_map is not a live RT_ShadowDepth");
    d.W = sh->w; d.H = sh->h;
    d.job_w = job_tile_w ? job_tile_w : 160;
    d.job_h = job_tile_h ? job_tile_h : 160;
    hm::store(hm::mul(hm::load(light_vp), hm::load(model)), d.light_model); // u.light_vp * u.model * vec4: the matrix product first (:779)
    wait_pending_read(ctx, sh);
    return legacy2_launch(ctx, d, nullptr, sh->depth, nullptr);
}

SHSB_API int32_t shsb_legacy2_draw_softshadow(shsb_ctx ctx, shsb_mesh mesh_h, const ShsbLegacy2Uniforms* u, shsb_rt shadow_rt, shsb_rt canvas_rt, shsb_rt zbuffer_rt)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    l2::Draw d{};
    d.mode = l2::MODE_SOFTSHADOW;
    RtSlot *canvas = nullptr, *zb = nullptr;
    if (int rc = legacy2_fill_camera(ctx, mesh_h, u, shadow_rt, canvas_rt, zbuffer_rt, false, d, canvas, zb)) return rc;
    if (d.shadow)
    {
        if (int rc = ensure_l2_rotation(ctx)) return rc;
        d.rot = ctx->d_l2_rot;
    }
    wait_pending_read(ctx, canvas);
    wait_pending_read(ctx, zb);
    return legacy2_launch(ctx, d, (uchar4*)canvas->color, zb->depth, nullptr);
}

SHSB_API int32_t shsb_legacy3_ibl_upload(shsb_ctx ctx, const float* irradiance, int32_t irr_size, const float* prefiltered, const int32_t* spec_sizes, int32_t n_mips,
                                         shsb_ibl* out_ibl)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (!out_ibl || !irradiance || !prefiltered || !spec_sizes || irr_size <= 0 || n_mips <= 0 || n_mips > l2::MAX_SPEC_MIPS)
        return fail(ctx, SHSB_E_INVALID_ARGUMENT, "bad IBL arguments (1 <= n_mips <= %d)", l2::MAX_SPEC_MIPS);
    CK(cudaSetDevice(ctx->device));
    shsb_context_t::IblSlot e;
    e.live = true; e.irr_size = irr_size; e.n_mips = n_mips;
    size_t off = 0;
    for (int m = 0; m < n_mips; ++m)
    {
        if (spec_sizes[m] <= 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "prefiltered mip %d has size %d", m, spec_sizes[m]);
        e.spec_size[m] = spec_sizes[m];
        e.spec_off[m] = (uint32_t)off;
        off += (size_t)6 * spec_sizes[m] * spec_sizes[m] * 3;
        if (off > 0xFFFFFFFFull) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "prefiltered chain too large");
    }
    const size_t irr_bytes = (size_t)6 * irr_size * irr_size * 3 * sizeof(float);
    // a failing step releases what the earlier ones allocated
    auto step = [&](cudaError_t err) -> int {
        if (err == cudaSuccess) return SHSB_OK;
        cudaFree(e.irradiance); cudaFree(e.prefiltered);
        cudaGetLastError();
        return fail(ctx, err == cudaErrorMemoryAllocation ? SHSB_E_OUT_OF_MEMORY : SHSB_E_CUDA, "IBL upload: %s", cudaGetErrorString(err));
    };
    if (int rc = step(cudaMalloc(&e.irradiance, irr_bytes))) return rc;
    if (int rc = step(cudaMalloc(&e.prefiltered, off * sizeof(float)))) return rc;
    if (int rc = step(cudaMemcpyAsync(e.irradiance, irradiance, irr_bytes, cudaMemcpyHostToDevice, ctx->stream))) return rc;
    if (int rc = step(cudaMemcpyAsync(e.prefiltered, prefiltered, off * sizeof(float), cudaMemcpyHostToDevice, ctx->stream))) return rc;
    if (int rc = step(cudaStreamSynchronize(ctx->stream))) return rc;
    ctx->ibls.push_back(e);
    *out_ibl = (shsb_ibl)ctx->ibls.size();
    return SHSB_OK;
}

SHSB_API int32_t shsb_legacy3_ibl_destroy(shsb_ctx ctx, shsb_ibl ibl)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (ibl == 0 || ibl > ctx->ibls.size() || !ctx->ibls[ibl - 1].live) return fail(ctx, SHSB_E_INVALID_HANDLE, "IBL handle %u is not live", ibl);
    sync_all(ctx);
    cudaFree(ctx->ibls[ibl - 1].irradiance);
    cudaFree(ctx->ibls[ibl - 1].prefiltered);
    ctx->ibls[ibl - 1] = shsb_context_t::IblSlot{};
    return SHSB_OK;
}

SHSB_API int32_t shsb_legacy3_draw_pbr(shsb_ctx ctx, shsb_mesh mesh_h, const ShsbLegacy2Uniforms* u, shsb_rt shadow_rt, shsb_ibl ibl, shsb_rt canvas_rt, shsb_rt depth_motion_rt)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    l2::Draw d{};
    d.mode = l2::MODE_PBR;
    RtSlot *canvas = nullptr, *dm = nullptr;
    if (int rc = legacy2_fill_camera(ctx, mesh_h, u, shadow_rt, canvas_rt, depth_motion_rt, true, d, canvas, dm)) return rc;
    d.metallic = u->metallic; d.roughness = u->roughness; d.ao = u->ao;
    d.ibl_diffuse = u->ibl_diffuse_intensity; d.ibl_specular = u->ibl_specular_intensity; d.ibl_reflection = u->ibl_reflection_strength;
    if (ibl)
    {
        if (ibl > ctx->ibls.size() || !ctx->ibls[ibl - 1].live) return fail(ctx, SHSB_E_INVALID_HANDLE, "IBL handle %u is not live", ibl);
        const shsb_context_t::IblSlot& e = ctx->ibls[ibl - 1];
        d.irradiance = e.irradiance; d.irr_size = e.irr_size; d.prefiltered = e.prefiltered; d.n_mips = e.n_mips;
        std::memcpy(d.spec_size, e.spec_size, sizeof(d.spec_size));
        std::memcpy(d.spec_off, e.spec_off, sizeof(d.spec_off));
    }
    wait_pending_read(ctx, canvas);
    wait_pending_read(ctx, dm);
    dm->motion_dirty = true;
    return legacy2_launch(ctx, d, (uchar4*)canvas->color, dm->depth, dm->motion);
}

// ---------------------------------------------------------------------------------------- scene-level culling
namespace
{
    inline size_t sc_align(size_t n) { return (n + 255) & ~(size_t)255; }
}

SHSB_API int32_t shsb_cull_objects_frustum(shsb_ctx ctx, const float* bounds10, uint32_t n, const float view_proj[16], uint8_t* out_classes, uint32_t* out_visible,
                                           uint32_t out_counts5[5])
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if ((n && (!bounds10 || !out_classes || !out_visible)) || !view_proj || !out_counts5) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(ctx->device));
    float planes[24];
    hm::frustum_planes(hm::load(view_proj), planes); // extract_frustum_planes, on the host like the light-list builders' pre-filter
    const size_t o_bounds = 0, o_classes = sc_align((size_t)n * 40), o_visible = o_classes + sc_align(n), o_counts = o_visible + sc_align((size_t)n * 4);
    if (int rc = ensure_dev(ctx, ctx->d_sc_bytes, o_counts + 256)) return rc;
    uint8_t* base = ctx->d_sc_bytes.p;
    if (n) CK(cudaMemcpyAsync(base + o_bounds, bounds10, (size_t)n * 40, cudaMemcpyHostToDevice, ctx->stream));
    launch_cull_objects((const float*)(base + o_bounds), n, planes, base + o_classes, (uint32_t*)(base + o_visible), (uint32_t*)(base + o_counts), ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_counts5, base + o_counts, 20, cudaMemcpyDeviceToHost, ctx->stream));
    if (n) CK(cudaMemcpyAsync(out_classes, base + o_classes, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (out_counts5[4]) CK(cudaMemcpy(out_visible, base + o_visible, (size_t)out_counts5[4] * 4, cudaMemcpyDeviceToHost));
    return SHSB_OK;
}

SHSB_API int32_t shsb_software_occlusion(shsb_ctx ctx, const float* object_aabbs6, uint32_t n_objects, const uint32_t* frustum_visible, uint32_t n_visible,
                                         const uint32_t* object_mesh, const float* object_models16, const uint32_t* mesh_table3, uint32_t n_meshes,
                                         const float* occluder_vertices, uint32_t n_vertices, const uint32_t* occluder_indices, uint32_t n_indices,
                                         const float view[16], const float view_proj[16], int32_t occ_w, int32_t occ_h, float depth_epsilon, int32_t enable_occlusion,
                                         uint8_t* out_occluded, uint32_t* out_visible, uint32_t out_counts4[4], float* out_depth)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if ((n_objects && (!object_aabbs6 || !out_occluded)) || (n_visible && (!frustum_visible || !out_visible)) || !view || !view_proj || !out_counts4)
        return fail(ctx, SHSB_E_INVALID_ARGUMENT, "null argument");
    if (n_objects && (!object_mesh || !object_models16)) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "object_mesh / object_models16 are null");
    if ((n_meshes && !mesh_table3) || (n_vertices && !occluder_vertices) || (n_indices && !occluder_indices)) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "null occluder mesh data");
    if (occ_w <= 0 || occ_h <= 0 || (size_t)occ_w * (size_t)occ_h > (1u << 26)) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "bad occlusion buffer size %d x %d", occ_w, occ_h);
    for (uint32_t m = 0; m < n_meshes; ++m)
        if ((uint64_t)mesh_table3[3 * m] + mesh_table3[3 * m + 1] > n_indices) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "mesh %u reaches beyond the index array", m);
    CK(cudaSetDevice(ctx->device));
    if (n_objects) std::memset(out_occluded, 0, n_objects);
    out_counts4[0] = n_objects; out_counts4[1] = n_visible; out_counts4[2] = 0; out_counts4[3] = 0;
    if (!enable_occlusion)
    {
        // :270-284: nothing to rasterise or test -- the visible list is the frustum list minus out-of-range entries
        uint32_t nv = 0;
        for (uint32_t k = 0; k < n_visible; ++k) if (frustum_visible[k] < n_objects) out_visible[nv++] = frustum_visible[k];
        out_counts4[2] = nv;
        out_counts4[1] = std::max(out_counts4[1], nv); // normalize_culling_stats
        out_counts4[3] = out_counts4[1] - nv;
        if (out_depth) std::fill(out_depth, out_depth + (size_t)occ_w * occ_h, 1.0f); // the buffer is not touched by the reference in this mode; 1.0 = cleared
        return SHSB_OK;
    }
    // front-to-back order: the reference's std::sort with its comparator (:289-297), same library, same ties
    std::vector<float> key(n_objects);
    for (uint32_t i = 0; i < n_objects; ++i) key[i] = sc::occ_view_depth(object_aabbs6 + (size_t)i * 6, view);
    std::vector<uint32_t> sorted(frustum_visible, frustum_visible + n_visible);
    std::sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) {
        if (a >= n_objects) return false;
        if (b >= n_objects) return true;
        return key[a] < key[b];
    });
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += sc_align(std::max<size_t>(bytes, 4)); return o; };
    const size_t o_boxes = take((size_t)n_objects * 24), o_sorted = take((size_t)n_visible * 4), o_omesh = take((size_t)n_objects * 4), o_models = take((size_t)n_objects * 64),
                 o_table = take((size_t)n_meshes * 12), o_verts = take((size_t)n_vertices * 12), o_idx = take((size_t)n_indices * 4), o_depth = take((size_t)occ_w * occ_h * 4),
                 o_occ = take(n_objects), o_vis = take((size_t)n_visible * 4), o_counts = take(8), o_rects = take((size_t)n_visible * 32);
    if (int rc = ensure_dev(ctx, ctx->d_sc_bytes, off + 256)) return rc;
    uint8_t* base = ctx->d_sc_bytes.p;
    cudaStream_t st = ctx->stream;
    auto up = [&](size_t o, const void* src, size_t bytes) -> cudaError_t { return bytes ? cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
    CK(up(o_boxes, object_aabbs6, (size_t)n_objects * 24));
    CK(up(o_sorted, sorted.data(), (size_t)n_visible * 4));
    CK(up(o_omesh, object_mesh, (size_t)n_objects * 4));
    CK(up(o_models, object_models16, (size_t)n_objects * 64));
    CK(up(o_table, mesh_table3, (size_t)n_meshes * 12));
    CK(up(o_verts, occluder_vertices, (size_t)n_vertices * 12));
    CK(up(o_idx, occluder_indices, (size_t)n_indices * 4));
    CK(cudaMemsetAsync(base + o_occ, 0, std::max<size_t>(n_objects, 4), st));
    launch_software_occlusion((const float*)(base + o_boxes), n_objects, (const uint32_t*)(base + o_sorted), n_visible, (const uint32_t*)(base + o_omesh), (const float*)(base + o_models),
                              (const uint32_t*)(base + o_table), n_meshes, (const float*)(base + o_verts), n_vertices, (const uint32_t*)(base + o_idx), n_indices, view_proj, occ_w, occ_h,
                              depth_epsilon, (float*)(base + o_depth), base + o_occ, (uint32_t*)(base + o_vis), (uint32_t*)(base + o_counts), base + o_rects, st, &ctx->launches);
    CK(cudaGetLastError());
    uint32_t c2[2] = {0, 0};
    CK(cudaMemcpyAsync(c2, base + o_counts, 8, cudaMemcpyDeviceToHost, st));
    if (n_objects) CK(cudaMemcpyAsync(out_occluded, base + o_occ, n_objects, cudaMemcpyDeviceToHost, st));
    if (out_depth) CK(cudaMemcpyAsync(out_depth, base + o_depth, (size_t)occ_w * occ_h * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (c2[0]) CK(cudaMemcpy(out_visible, base + o_vis, (size_t)c2[0] * 4, cudaMemcpyDeviceToHost));
    out_counts4[2] = c2[0];
    out_counts4[1] = std::max(out_counts4[1], c2[0]); // normalize_culling_stats, culling_runtime.hpp:59-75
    out_counts4[3] = out_counts4[1] - c2[0];
    return SHSB_OK;
}

SHSB_API int32_t shsb_collect_object_lights(shsb_ctx ctx, const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible_lights, uint32_t n_visible,
                                            const void* records160, uint32_t n_lights, int32_t cull_mode, uint32_t* out_counts, uint32_t* out_indices8, float* out_dist2_8)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if ((n_objects && (!object_aabbs6 || !out_counts || !out_indices8 || !out_dist2_8)) || (n_visible && !visible_lights) || (n_lights && !records160))
        return fail(ctx, SHSB_E_INVALID_ARGUMENT, "null argument");
    if (cull_mode < SHSB_LIGHT_OBJECT_CULL_NONE || cull_mode > SHSB_LIGHT_OBJECT_CULL_VOLUME_AABB) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "unknown LightObjectCullMode %d", cull_mode);
    if (n_objects == 0) return SHSB_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t o_boxes = 0, o_vis = sc_align((size_t)n_objects * 24), o_recs = o_vis + sc_align((size_t)n_visible * 4), o_counts = o_recs + sc_align((size_t)n_lights * 160);
    const size_t o_idx = o_counts + sc_align((size_t)n_objects * 4), o_d2 = o_idx + sc_align((size_t)n_objects * 32), total = o_d2 + sc_align((size_t)n_objects * 32);
    if (int rc = ensure_dev(ctx, ctx->d_sc_bytes, total)) return rc;
    uint8_t* base = ctx->d_sc_bytes.p;
    CK(cudaMemcpyAsync(base + o_boxes, object_aabbs6, (size_t)n_objects * 24, cudaMemcpyHostToDevice, ctx->stream));
    if (n_visible) CK(cudaMemcpyAsync(base + o_vis, visible_lights, (size_t)n_visible * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n_lights) CK(cudaMemcpyAsync(base + o_recs, records160, (size_t)n_lights * 160, cudaMemcpyHostToDevice, ctx->stream));
    launch_collect_object_lights((const float*)(base + o_boxes), n_objects, (const uint32_t*)(base + o_vis), n_visible, (const float*)(base + o_recs), n_lights, cull_mode,
                                 (uint32_t*)(base + o_counts), (uint32_t*)(base + o_idx), (float*)(base + o_d2), ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_counts, base + o_counts, (size_t)n_objects * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_indices8, base + o_idx, (size_t)n_objects * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_dist2_8, base + o_d2, (size_t)n_objects * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SHSB_OK;
}

SHSB_API int32_t shsb_tile_depth_range_from_scene(shsb_ctx ctx, const float* object_aabbs6, uint32_t n_objects, const uint32_t* visible_objects, uint32_t n_visible,
                                                  const float view[16], const float view_proj[16], uint32_t vw, uint32_t vh, uint32_t tile_size, float z_near, float z_far)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if ((n_objects && !object_aabbs6) || (n_visible && !visible_objects) || !view || !view_proj) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "null argument");
    if (vw == 0 || vh == 0 || tile_size == 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "viewport %ux%u, tile size %u", vw, vh, tile_size);
    CK(cudaSetDevice(ctx->device));
    const uint32_t tiles_x = (vw + tile_size - 1) / tile_size, tiles_y = (vh + tile_size - 1) / tile_size;
    const size_t tiles = (size_t)tiles_x * tiles_y;
    if (int rc = ensure_dev(ctx, ctx->d_range_min, tiles)) return rc;
    if (int rc = ensure_dev(ctx, ctx->d_range_max, tiles)) return rc;
    const size_t o_boxes = 0, o_vis = sc_align((size_t)n_objects * 24), o_scratch = o_vis + sc_align((size_t)n_visible * 4), total = o_scratch + sc_align(tiles * 12);
    if (int rc = ensure_dev(ctx, ctx->d_sc_bytes, total)) return rc;
    uint8_t* base = ctx->d_sc_bytes.p;
    if (n_objects) CK(cudaMemcpyAsync(base + o_boxes, object_aabbs6, (size_t)n_objects * 24, cudaMemcpyHostToDevice, ctx->stream));
    if (n_visible) CK(cudaMemcpyAsync(base + o_vis, visible_objects, (size_t)n_visible * 4, cudaMemcpyHostToDevice, ctx->stream));
    launch_scene_tile_depth_range((const float*)(base + o_boxes), n_objects, (const uint32_t*)(base + o_vis), n_objects ? n_visible : 0, view, view_proj, z_near, z_far, tiles_x, tiles_y,
                                  (uint32_t*)(base + o_scratch), ctx->d_range_min.p, ctx->d_range_max.p, ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream)); // the host arrays may go away after the call
    ctx->range_w = vw; ctx->range_h = vh; ctx->range_ts = tile_size;
    return SHSB_OK;
}

SHSB_API int32_t shsb_select_object_lights_from_bins(shsb_ctx ctx, const float* object_aabbs6, uint32_t n_objects, const float view[16], const float view_proj[16], int32_t clustered,
                                                     float z_near, float z_far, const void* records160, uint32_t n_lights, int32_t cull_mode, uint32_t* out_counts,
                                                     uint32_t* out_indices8, float* out_dist2_8, uint32_t* out_candidates)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if ((n_objects && (!object_aabbs6 || !out_counts || !out_indices8 || !out_dist2_8 || !out_candidates)) || !view || !view_proj || (n_lights && !records160))
        return fail(ctx, SHSB_E_INVALID_ARGUMENT, "null argument");
    if (cull_mode < SHSB_LIGHT_OBJECT_CULL_NONE || cull_mode > SHSB_LIGHT_OBJECT_CULL_VOLUME_AABB) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "unknown LightObjectCullMode %d", cull_mode);
    const LightLists& L = clustered ? ctx->cluster_lists : ctx->lists[LISTS_STANDALONE];
    if (!L.ts || !L.counts.p || !L.indices.p || (clustered && !ctx->cluster_slices))
        return fail(ctx, SHSB_E_INVALID_ARGUMENT, clustered ? "no cluster bins: call shsb_light_cull_ex(SHSB_LIGHT_CULL_CLUSTERED) first" : "no tile light lists: call shsb_light_cull / shsb_light_cull_ex first");
    if (n_objects == 0) return SHSB_OK;
    CK(cudaSetDevice(ctx->device));
    sc::BinGrid g{};
    g.bins_x = (L.w + L.ts - 1) / L.ts;
    g.bins_y = (L.h + L.ts - 1) / L.ts;
    g.bins_z = clustered ? ctx->cluster_slices : 1u;
    g.clustered = clustered ? 1 : 0;
    g.z_near = std::max(z_near, 1e-4f);            // LightBinCullingData::z_near / z_far, light_culling_runtime.hpp:278-279
    g.z_far = std::max(z_far, g.z_near + 1e-3f);
    g.max_per_bin = L.max_per_tile;
    const uint32_t words = (n_lights + 31u) / 32u + 1u;
    const size_t o_boxes = 0, o_recs = sc_align((size_t)n_objects * 24), o_seen = o_recs + sc_align((size_t)n_lights * 160), o_counts = o_seen + sc_align((size_t)n_objects * words * 4);
    const size_t o_idx = o_counts + sc_align((size_t)n_objects * 4), o_d2 = o_idx + sc_align((size_t)n_objects * 32), o_cand = o_d2 + sc_align((size_t)n_objects * 32);
    const size_t total = o_cand + sc_align((size_t)n_objects * 4);
    if (int rc = ensure_dev(ctx, ctx->d_sc_bytes, total)) return rc;
    uint8_t* base = ctx->d_sc_bytes.p;
    CK(cudaMemcpyAsync(base + o_boxes, object_aabbs6, (size_t)n_objects * 24, cudaMemcpyHostToDevice, ctx->stream));
    if (n_lights) CK(cudaMemcpyAsync(base + o_recs, records160, (size_t)n_lights * 160, cudaMemcpyHostToDevice, ctx->stream));
    launch_select_object_lights((const float*)(base + o_boxes), n_objects, view, view_proj, g, L.counts.p, L.indices.p, (const float*)(base + o_recs), n_lights, cull_mode,
                                (uint32_t*)(base + o_seen), words, (uint32_t*)(base + o_counts), (uint32_t*)(base + o_idx), (float*)(base + o_d2), (uint32_t*)(base + o_cand),
                                ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_counts, base + o_counts, (size_t)n_objects * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_indices8, base + o_idx, (size_t)n_objects * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_dist2_8, base + o_d2, (size_t)n_objects * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_candidates, base + o_cand, (size_t)n_objects * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SHSB_OK;
}

// ---------------------------------------------------------------------------------------- flat-shaded mesh draws (sw_render/debug_draw.hpp)
namespace
{
    // One batch of flat-shaded draws into (canvas, depth): see flat_draw.cu.  mode: fd::MODE_BLINN_PHONG / fd::MODE_MULTI_LIGHT.
    int flat_draw_batch(shsb_ctx ctx, int mode, const ShsbFlatDraw* draws, uint32_t n_draws, const float view_proj[16], const float camera_pos[3], const float light_dir_ws[3],
                        const ShsbLightProperties* lights, uint32_t n_lights, shsb_rt canvas_rt, shsb_rt depth_rt)
    {
        if (!ctx) return SHSB_E_INVALID_ARGUMENT;
        if ((n_draws && !draws) || !view_proj || !camera_pos || (mode == fd::MODE_BLINN_PHONG && !light_dir_ws) || (n_lights && !lights))
            return fail(ctx, SHSB_E_INVALID_ARGUMENT, "null argument");
        CK(cudaSetDevice(ctx->device));
        RtSlot* canvas = get_rt(ctx, canvas_rt, SHSB_RT_COLOR_LDR);
        if (!canvas) return fail(ctx, SHSB_E_INVALID_HANDLE, "canvas_ldr is not a live RT_ColorLDR");
        RtSlot* zb = get_rt(ctx, depth_rt);
        if (!zb || !zb->depth) return fail(ctx, SHSB_E_INVALID_HANDLE, "depth is not a live target with a depth plane (RT_ShadowDepth / RT_ColorDepthMotion)");
        if (zb->w != canvas->w || zb->h != canvas->h) return fail(ctx, SHSB_E_SIZE_MISMATCH, "canvas is %dx%d, depth buffer %dx%d", canvas->w, canvas->h, zb->w, zb->h);
        if (canvas->w > 65535 * 64 || canvas->h > 65535 * 64) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "canvas too large");

        fd::BatchDesc bd{};
        std::memcpy(bd.view_proj, view_proj, 64);
        std::memcpy(bd.camera, camera_pos, 12);
        if (mode == fd::MODE_BLINN_PHONG)
        {
            // const glm::vec3 L = glm::normalize(-light_dir_ws), debug_draw.hpp:165
            const fd::V3 L = fd::glm_normalize(-fd::v3(light_dir_ws));
            bd.L[0] = L.x; bd.L[1] = L.y; bd.L[2] = L.z;
        }
        bd.W = canvas->w; bd.H = canvas->h; bd.mode = mode; bd.n_lights = n_lights;

        std::vector<fd::DrawRec> recs;
        recs.reserve(n_draws);
        uint64_t n_tris = 0;
        for (uint32_t i = 0; i < n_draws; ++i)
        {
            const ShsbFlatDraw& d = draws[i];
            const MeshSlot* mesh = get_mesh(ctx, d.mesh);
            if (!mesh) return fail(ctx, SHSB_E_INVALID_HANDLE, "draw %u: mesh handle %u is not live", i, d.mesh);
            if (d.selection_count > 8) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "draw %u: a LightSelection holds at most 8 lights", i);
            // for (i = 0; i + 2 < indices.size(); i += 3): whole triangles of the index list only (debug_draw.hpp:166)
            const uint32_t tris = mesh->n_indices / 3;
            if (tris == 0) continue;
            fd::DrawRec r{};
            r.positions = mesh->positions; r.indices = mesh->indices; r.n_positions = mesh->n_positions; r.n_tris = tris;
            r.tri_base = (uint32_t)n_tris; r.selection_count = d.selection_count;
            std::memcpy(r.model, d.model, 64); std::memcpy(r.base, d.base_color, 12); std::memcpy(r.selection, d.selection, 32);
            recs.push_back(r);
            n_tris += tris;
            if (n_tris >= 0xFFFFFFFEull) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "more than 2^32 - 2 triangles in one batch");
        }
        if (recs.empty()) return SHSB_OK;
        bd.n_draws = (uint32_t)recs.size(); bd.n_tris = (uint32_t)n_tris;

        const size_t n_px = (size_t)canvas->w * canvas->h;
        const size_t tri_bytes = flat_tri_record_bytes();
        if (ctx->fd_item_cap < bd.n_tris + 4096u) ctx->fd_item_cap = bd.n_tris + bd.n_tris / 2 + 4096u;
        wait_pending_read(ctx, canvas);
        wait_pending_read(ctx, zb);
        for (int attempt = 0; attempt < 2; ++attempt)
        {
            const size_t o_draws = 0, o_lights = sc_align(recs.size() * sizeof(fd::DrawRec)), o_tris = o_lights + sc_align((size_t)n_lights * sizeof(fd::LightProps)),
                         o_colours = o_tris + sc_align((size_t)bd.n_tris * tri_bytes), o_items = o_colours + sc_align((size_t)bd.n_tris * 4),
                         o_total = o_items + sc_align((size_t)ctx->fd_item_cap * 8), o_zkey = o_total + 256, total = o_zkey + n_px * 8;
            if (int rc = ensure_dev(ctx, ctx->d_fd_bytes, total)) return rc;
            uint8_t* base = ctx->d_fd_bytes.p;
            CK(cudaMemcpyAsync(base + o_draws, recs.data(), recs.size() * sizeof(fd::DrawRec), cudaMemcpyHostToDevice, ctx->stream));
            if (n_lights) CK(cudaMemcpyAsync(base + o_lights, lights, (size_t)n_lights * sizeof(fd::LightProps), cudaMemcpyHostToDevice, ctx->stream));
            launch_flat_draw_batch(bd, (const fd::DrawRec*)(base + o_draws), (const fd::LightProps*)(base + o_lights), base + o_tris, (uint32_t*)(base + o_colours),
                                   (uint2*)(base + o_items), ctx->fd_item_cap, (uint32_t*)(base + o_total), (unsigned long long*)(base + o_zkey), zb->depth, true, ctx->stream,
                                   &ctx->launches);
            CK(cudaGetLastError());
            uint32_t n_items = 0;
            CK(cudaMemcpyAsync(&n_items, base + o_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (n_items > ctx->fd_item_cap)
            {
                if (attempt) return fail(ctx, SHSB_E_OVERFLOW, "flat draw: work list still too small (%u > %u)", n_items, ctx->fd_item_cap);
                ctx->fd_item_cap = n_items + n_items / 4; // the set-up pass told how many 64 x 64 chunks the batch covers: run it again with room
                continue;
            }
            launch_flat_raster_resolve(bd, base + o_tris, (const uint32_t*)(base + o_colours), (const uint2*)(base + o_items), n_items, (unsigned long long*)(base + o_zkey),
                                       zb->depth, (uchar4*)canvas->color, ctx->stream, &ctx->launches);
            CK(cudaGetLastError());
            break;
        }
        return SHSB_OK;
    }
}

SHSB_API int32_t shsb_flat_draw_blinn_phong(shsb_ctx ctx, const ShsbFlatDraw* draws, uint32_t n_draws, const float view_proj[16], const float camera_pos[3],
                                            const float light_dir_ws[3], shsb_rt canvas_ldr, shsb_rt depth)
{
    return flat_draw_batch(ctx, fd::MODE_BLINN_PHONG, draws, n_draws, view_proj, camera_pos, light_dir_ws, nullptr, 0, canvas_ldr, depth);
}

SHSB_API int32_t shsb_flat_draw_multi_light(shsb_ctx ctx, const ShsbFlatDraw* draws, uint32_t n_draws, const float view_proj[16], const float camera_pos[3],
                                            const ShsbLightProperties* lights, uint32_t n_lights, shsb_rt canvas_ldr, shsb_rt depth)
{
    return flat_draw_batch(ctx, fd::MODE_MULTI_LIGHT, draws, n_draws, view_proj, camera_pos, nullptr, lights, n_lights, canvas_ldr, depth);
}

// ---------------------------------------------------------------------------------------- passes
SHSB_API int32_t shsb_rasterize_mesh(shsb_ctx ctx, shsb_mesh mesh_h, int32_t shader_id, const ShsbUniforms* u,
                                     shsb_rt hdr_rt, shsb_rt depth_motion_rt, const ShsbRasterCfg* cfg, ShsbStats* out_stats)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (!u || !cfg) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "uniforms / config are null");
    if (shader_id < 0 || shader_id >= SHSB_SHADER_COUNT)
        return fail(ctx, SHSB_E_UNSUPPORTED_SHADER, "shader id %d is not a builtin program; host std::function shaders cannot run on the device", shader_id);
    CK(cudaSetDevice(ctx->device));
    RtSlot* hdr = get_rt(ctx, hdr_rt, SHSB_RT_COLOR_HDR);
    if (!hdr) return fail(ctx, SHSB_E_INVALID_HANDLE, "hdr_rt is not a live RT_ColorHDR"); // rasterizer.hpp:190
    RtSlot* dm = depth_motion_rt ? get_rt(ctx, depth_motion_rt, SHSB_RT_DEPTH_MOTION) : nullptr;
    if (depth_motion_rt && !dm) return fail(ctx, SHSB_E_INVALID_HANDLE, "depth_motion_rt is not a live RT_ColorDepthMotion");
    if (dm && (dm->w != hdr->w || dm->h != hdr->h)) return fail(ctx, SHSB_E_SIZE_MISMATCH, "depth target size differs from the HDR target");
    const MeshSlot* mesh = get_mesh(ctx, mesh_h);
    if (!mesh) return fail(ctx, SHSB_E_INVALID_HANDLE, "mesh handle %u is not live", mesh_h);
    if (mesh->n_positions == 0) return SHSB_OK; // rasterizer.hpp:191

    FrameJob job;
    FrameConst& fc = job.fc;
    std::memcpy(fc.viewproj, u->viewproj, 64);
    std::memcpy(fc.camera_pos, u->camera_pos, 12);
    std::memcpy(fc.sun_dir, u->light_dir_ws, 12);
    std::memcpy(fc.sun_color, u->light_color, 12);
    fc.sun_intensity = u->light_intensity;
    fc.W = hdr->w; fc.H = hdr->h;
    fc.zn = dm ? dm->zn : 0.1f;
    fc.zf = dm ? dm->zf : 1000.0f;
    fc.shader_id = shader_id;
    fc.cull_mode = cfg->cull_mode;
    fc.front_face_ccw = cfg->front_face_ccw;
    fc.has_depth = dm ? 1 : 0;
    fc.linear_depth = (dm && (dm->zf > dm->zn + 1e-6f)) ? 1 : 0;
    fc.load_depth = 1;
    fc.load_color = 1;
    fc.write_aovs = cfg->write_aovs;
    RtSlot* sh = u->shadow_map ? get_rt(ctx, u->shadow_map, SHSB_RT_SHADOW) : nullptr;
    if (u->shadow_map && !sh) return fail(ctx, SHSB_E_INVALID_HANDLE, "uniforms.shadow_map is not a live RT_ShadowDepth");
    if (sh)
    {
        fc.shadow_map = sh->depth;
        fc.shadow_w = sh->w; fc.shadow_h = sh->h;
        std::memcpy(fc.light_viewproj, u->light_viewproj, 64);
        fc.bias_const = u->shadow_bias_const;
        fc.bias_slope = u->shadow_bias_slope;
        fc.pcf_radius = u->shadow_pcf_radius;
        fc.pcf_step = u->shadow_pcf_step;
        fc.shadow_strength = u->shadow_strength;
    }
    job.fb.hdr = (float4*)hdr->color;
    job.fb.depth = dm ? dm->depth : nullptr;
    const bool write_motion = dm && dm->motion && u->enable_motion_vectors != 0; // rasterizer.hpp:295
    if (write_motion)
    {
        fc.write_motion = 1;
        std::memcpy(fc.prev_viewproj, u->prev_viewproj, 64);
        job.fb.motion = dm->motion;
        dm->motion_dirty = true;
    }
    if (cfg->write_aovs)
    {
        if (int rc = ensure_aovs(ctx, hdr)) return rc;
        job.fb.aov_tri_id = hdr->tri_id;
        job.fb.aov_coverage = hdr->coverage;
    }
    std::vector<DevItem> items;
    std::vector<uint2> blocks;
    uint64_t tri_cursor = 0;
    const hm::mat4f prev_model = hm::load(u->prev_model);
    stage_item(ctx, items, blocks, tri_cursor, hm::load(u->model), *mesh, mesh_h - 1, u->base_color, u->metallic, u->roughness, u->ao, u->base_color_tex,
               write_motion ? &prev_model : nullptr);
    if (int rc = upload_staging(ctx, items, blocks)) return rc;
    job.n_items = 1;
    job.n_blocks = (uint32_t)blocks.size();
    job.n_src_tris = tri_cursor;
    return run_frame(ctx, job, out_stats);
}

SHSB_API int32_t shsb_pass_pbr_forward(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp, shsb_rt hdr_rt, shsb_rt depth_motion_rt,
                                       shsb_rt shadow_rt, const float* shadow_light_viewproj, int32_t preserve_existing_depth, ShsbStats* out_stats)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (out_stats) *out_stats = ShsbStats{}; // ctx.debug.tri_* = 0, pass_pbr_forward.hpp:53-55
    return forward_common(ctx, scene, fp, hdr_rt, depth_motion_rt, shadow_rt, shadow_light_viewproj, preserve_existing_depth, false, 0, out_stats);
}

SHSB_API int32_t shsb_pass_depth_prepass(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp, shsb_rt depth_motion_rt, ShsbStats* out_stats)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (out_stats) *out_stats = ShsbStats{};
    return forward_common(ctx, scene, fp, 0, depth_motion_rt, 0, nullptr, 0, true, 0, out_stats);
}

SHSB_API int32_t shsb_frame_forward_plus(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp, shsb_rt hdr_rt, shsb_rt depth_motion_rt,
                                         shsb_rt ldr_rt, ShsbStats* out_stats)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (!scene || !fp) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "scene / frame params are null");
    CK(cudaSetDevice(ctx->device));
    RtSlot* hdr = peek_rt(ctx, hdr_rt, SHSB_RT_COLOR_HDR);
    if (!hdr) return fail(ctx, SHSB_E_INVALID_HANDLE, "hdr_rt is not a live RT_ColorHDR");
    CullJob cull;
    const bool with_cull = fp->light_culling && ctx->n_lights > 0;
    if (with_cull)
    {
        if (int rc = prepare_light_cull(ctx, scene->cam_viewproj, (uint32_t)hdr->w, (uint32_t)hdr->h, std::max(1u, fp->tile_size), std::max(1u, fp->max_lights_per_tile), cull)) return rc;
        if (fp->own_row_count > 0 && cull.ts == (uint32_t)TILE && fp->own_row_first >= 0 && fp->own_row_stride >= fp->own_row_count)
        {
            // sort-first partition: only the owned tile rows' lists are ever read (light tile == raster tile)
            cull.own_first = fp->own_row_first; cull.own_count = fp->own_row_count; cull.own_stride = fp->own_row_stride;
        }
    }
    if (out_stats) *out_stats = ShsbStats{};
    return forward_common(ctx, scene, fp, hdr_rt, depth_motion_rt, 0, nullptr, 0, false, ldr_rt, out_stats, with_cull ? &cull : nullptr);
}

SHSB_API int32_t shsb_pass_shadow_map(shsb_ctx ctx, const ShsbScene* scene, const ShsbFrameParams* fp, shsb_rt shadow_rt, float out_light_viewproj[16])
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (!scene || !fp) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "scene / frame params are null");
    CK(cudaSetDevice(ctx->device));
    RtSlot* sh = get_rt(ctx, shadow_rt, SHSB_RT_SHADOW);
    if (!sh) return fail(ctx, SHSB_E_INVALID_HANDLE, "shadow_rt is not a live RT_ShadowDepth");
    if (!fp->shadow_enable) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "pass.shadow.enable is false (pass_shadow_map.hpp:50)");

    // scene AABB over the transformed corners of each caster's local bounds, pass_shadow_map.hpp:80-131
    hm::Aabb box;
    bool any = false;
    std::vector<DevItem> items;
    std::vector<uint2> blocks;
    uint64_t tri_cursor = 0;
    const float white[3] = {1, 1, 1};
    for (uint32_t i = 0; i < scene->n_items; ++i)
    {
        const ShsbRenderItem& it = scene->items[i];
        if (!it.visible || !it.casts_shadow) continue;
        const MeshSlot* mesh = get_mesh(ctx, it.mesh);
        if (mesh && mesh->n_positions)
        {
            const hm::mat4f model = hm::model_from_transform(it.tr.pos, it.tr.rot_euler, it.tr.scl);
            for (int c = 0; c < 8; ++c)
            {
                const hm::vec4f p = hm::mul_v(model, {(c & 1) ? mesh->bmax.x : mesh->bmin.x, (c & 2) ? mesh->bmax.y : mesh->bmin.y, (c & 4) ? mesh->bmax.z : mesh->bmin.z, 1.0f});
                box.expand({p.x, p.y, p.z});
            }
            stage_item(ctx, items, blocks, tri_cursor, model, *mesh, it.mesh - 1, white, 0, 0, 1, 0);
        }
        else box.expand({it.tr.pos[0], it.tr.pos[1], it.tr.pos[2]});
        any = true;
    }
    if (!any) { box.expand({-1, -1, -1}); box.expand({1, 1, 1}); }
    const hm::mat4f lvp = hm::light_camera_viewproj({scene->sun_dir_ws[0], scene->sun_dir_ws[1], scene->sun_dir_ws[2]}, box, 10.0f, (unsigned)std::max(sh->w, 1));
    if (out_light_viewproj) hm::store(lvp, out_light_viewproj);

    FrameJob job;
    FrameConst& fc = job.fc;
    hm::store(lvp, fc.viewproj);
    fc.W = sh->w; fc.H = sh->h;
    fc.shadow_mode = 1;
    fc.shader_id = SHSB_SHADER_DEPTH_ONLY;
    fc.load_depth = 0; // shadow->clear(1.0f), pass_shadow_map.hpp:55
    job.fb.depth = sh->depth;
    if (ctx->shadow_direct)
    {
        // small and medium triangles (nearly all of a shadow map's) are rasterised by the set-up kernel into the pre-cleared plane,
        // what is larger than 64x64 texels by a chunked second kernel: no bins, no tile kernel
        fc.direct_depth = sh->depth;
        fc.load_depth = 1;
        job.prefill_depth = sh->depth;
        job.prefill_n = (size_t)sh->w * (size_t)sh->h;
    }
    if (int rc = upload_staging(ctx, items, blocks)) return rc;
    job.n_items = (uint32_t)items.size();
    job.n_blocks = (uint32_t)blocks.size();
    job.n_src_tris = tri_cursor;
    ShsbStats st{};
    return run_frame(ctx, job, &st);
}

SHSB_API int32_t shsb_history_reset(shsb_ctx ctx)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    ctx->hist_keys.clear();
    ctx->hist_models.clear();
    ctx->hist_has_duplicates = false;
    ctx->hist_index.clear();
    ctx->hist_index_valid = false;
    ctx->has_prev_frame = false;
    return SHSB_OK;
}

SHSB_API int32_t shsb_pass_tonemap(shsb_ctx ctx, shsb_rt hdr_rt, shsb_rt ldr_rt, float exposure, float gamma)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* hdr = get_rt(ctx, hdr_rt, SHSB_RT_COLOR_HDR);
    RtSlot* ldr = get_rt(ctx, ldr_rt, SHSB_RT_COLOR_LDR);
    if (!hdr || !ldr) return fail(ctx, SHSB_E_INVALID_HANDLE, "tonemap needs a live RT_ColorHDR and RT_ColorLDR");
    if (hdr->w != ldr->w || hdr->h != ldr->h) return fail(ctx, SHSB_E_SIZE_MISMATCH, "tonemap targets differ in size"); // reference crops to min(w,h); not needed on this path
    record(ctx, 5, ctx->stream);
    launch_tonemap((const float4*)hdr->color, (uchar4*)ldr->color, hdr->w * hdr->h, std::max(0.0001f, exposure), 1.0f / std::max(0.001f, gamma), ctx->stream, &ctx->launches);
    record(ctx, 6, ctx->stream);
    CK(cudaGetLastError());
    return SHSB_OK;
}

// ---------------------------------------------------------------------------------------- post passes
namespace
{
    // copy_ldr of the reference (pass_motion_blur.hpp:187-199, pass_light_shafts.hpp:59-66): cropped copy, no-op in place
    int copy_ldr(shsb_ctx ctx, RtSlot* src, RtSlot* dst, int w, int h)
    {
        if (src == dst) return SHSB_OK;
        wait_pending_read(ctx, dst);
        launch_copy_ldr((const uchar4*)src->color, src->w, (uchar4*)dst->color, dst->w, w, h, ctx->stream, &ctx->launches);
        CK(cudaGetLastError());
        return SHSB_OK;
    }
}

SHSB_API int32_t shsb_pass_motion_blur(shsb_ctx ctx, const ShsbMotionBlurParams* p, shsb_rt input_ldr, shsb_rt output_ldr, shsb_rt depth_motion_rt)
{
    if (!ctx || !p) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* src = get_rt(ctx, input_ldr, SHSB_RT_COLOR_LDR);
    RtSlot* dst = get_rt(ctx, output_ldr, SHSB_RT_COLOR_LDR);
    RtSlot* mot = get_rt(ctx, depth_motion_rt, SHSB_RT_DEPTH_MOTION);
    if (!src || !dst || !mot) return fail(ctx, SHSB_E_INVALID_HANDLE, "motion blur needs live input / output RT_ColorLDR and an RT_ColorDepthMotion"); // :49
    const int w = std::min({src->w, dst->w, mot->w}), h = std::min({src->h, dst->h, mot->h}); // :51-53
    if (w <= 0 || h <= 0) return SHSB_OK;
    if (!p->enable) return copy_ldr(ctx, src, dst, w, h); // :56-60
    PostMotionBlur a{};
    a.src = (const uchar4*)src->color; a.src_w = src->w;
    a.motion = mot->motion; a.depth = mot->depth; a.mot_w = mot->w;
    a.w = w; a.h = h;
    a.samples = std::clamp(p->samples, 4, 32);                                   // :79
    a.strength = std::max(0.0f, p->strength);                                    // :80
    a.max_vel = std::max(1.0f, p->max_velocity_px);                              // :81
    a.min_vel = std::max(0.0f, p->min_velocity_px);                              // :82
    a.depth_eps = std::max(0.0f, p->depth_reject);                               // :83
    a.dt_scale = std::clamp(std::max(p->dt, 1e-4f) * 60.0f, 0.5f, 2.5f);         // :84
    const bool in_place = (src == dst);
    if (in_place)
    {
        if (int rc = ensure_dev(ctx, ctx->d_post_scratch, (size_t)w * h)) return rc;
        a.dst = ctx->d_post_scratch.p; a.dst_w = w;
    }
    else
    {
        wait_pending_read(ctx, dst);
        a.dst = (uchar4*)dst->color; a.dst_w = dst->w;
    }
    launch_motion_blur(a, ctx->stream, &ctx->launches);
    if (in_place)
    {
        wait_pending_read(ctx, dst);
        launch_copy_ldr(ctx->d_post_scratch.p, w, (uchar4*)dst->color, dst->w, w, h, ctx->stream, &ctx->launches); // :166-183
    }
    CK(cudaGetLastError());
    return SHSB_OK;
}

SHSB_API int32_t shsb_pass_light_shafts(shsb_ctx ctx, const ShsbLightShaftsParams* p, shsb_rt input_ldr, shsb_rt output_ldr, shsb_rt depth_like_rt)
{
    if (!ctx || !p) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* src = get_rt(ctx, input_ldr, SHSB_RT_COLOR_LDR);
    RtSlot* dst = get_rt(ctx, output_ldr, SHSB_RT_COLOR_LDR);
    if (!src || !dst) return fail(ctx, SHSB_E_INVALID_HANDLE, "light shafts need live input / output RT_ColorLDR"); // :50
    RtSlot* dep = depth_like_rt ? get_rt(ctx, depth_like_rt, SHSB_RT_DEPTH_MOTION) : nullptr;
    if (depth_like_rt && !dep) return fail(ctx, SHSB_E_INVALID_HANDLE, "depth_like_rt %u is not a live RT_ColorDepthMotion", depth_like_rt);
    const int w = std::min(src->w, dst->w), h = std::min(src->h, dst->h);
    if (w <= 0 || h <= 0) return SHSB_OK;
    if (!p->enable) return copy_ldr(ctx, src, dst, w, h); // :53-67

    // sun position on screen, pass_light_shafts.hpp:77-93 (host arithmetic in the reference's order)
    float sun_u = 0.5f, sun_v = 0.2f;
    bool sun_valid = false;
    {
        const hm::vec3f sun_pos{p->cam_pos[0] + (-p->sun_dir_ws[0]) * 100.0f, p->cam_pos[1] + (-p->sun_dir_ws[1]) * 100.0f, p->cam_pos[2] + (-p->sun_dir_ws[2]) * 100.0f};
        const hm::vec4f clip = hm::mul_v(hm::load(p->cam_viewproj), hm::vec4f{sun_pos.x, sun_pos.y, sun_pos.z, 1.0f});
        if (std::abs(clip.w) > 1e-6f)
        {
            const float nx = clip.x / clip.w, ny = clip.y / clip.w, nz = clip.z / clip.w;
            sun_u = nx * 0.5f + 0.5f;
            sun_v = ny * 0.5f + 0.5f;
            sun_valid = (clip.w > 0.0f) && (nz >= -1.0f && nz <= 1.0f) && (sun_u >= 0.0f && sun_u <= 1.0f) && (sun_v >= 0.0f && sun_v <= 1.0f);
        }
    }
    if (!sun_valid) return copy_ldr(ctx, src, dst, w, h); // :96-108

    PostLightShafts a{};
    a.src = (const uchar4*)src->color; a.src_w = src->w;
    a.w = w; a.h = h;
    a.steps = std::max(8, p->steps);                    // :135
    a.density = std::max(0.0f, p->density);             // :136
    a.weight = std::max(0.0f, p->weight);               // :137
    a.decay = std::clamp(p->decay, 0.0f, 1.0f);         // :138
    a.sun_u = sun_u; a.sun_v = sun_v;
    if (int rc = ensure_dev(ctx, ctx->d_post_luma, (size_t)w * h)) return rc;
    const bool use_depth = dep && dep->w == w && dep->h == h; // :165
    const bool in_place = (src == dst);
    if (in_place)
    {
        if (int rc = ensure_dev(ctx, ctx->d_post_scratch, (size_t)w * h)) return rc;
        a.dst = ctx->d_post_scratch.p; a.dst_w = w;
    }
    else
    {
        wait_pending_read(ctx, dst);
        a.dst = (uchar4*)dst->color; a.dst_w = dst->w;
    }
    launch_light_shafts(a, use_depth ? dep->depth : nullptr, use_depth ? dep->w : 0, ctx->d_post_luma.p, ctx->stream, &ctx->launches);
    if (in_place)
    {
        wait_pending_read(ctx, dst);
        launch_copy_ldr(ctx->d_post_scratch.p, w, (uchar4*)dst->color, dst->w, w, h, ctx->stream, &ctx->launches); // :197-211
    }
    CK(cudaGetLastError());
    return SHSB_OK;
}

SHSB_API int32_t shsb_pass_taa(shsb_ctx ctx, shsb_rt ldr_rt)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    RtSlot* ldr = get_rt(ctx, ldr_rt, SHSB_RT_COLOR_LDR);
    if (!ldr) return fail(ctx, SHSB_E_INVALID_HANDLE, "TAA needs a live RT_ColorLDR"); // pass_adapters.hpp:1444-1445
    const size_t count = (size_t)ldr->w * ldr->h;
    if (ctx->taa_w != ldr->w || ctx->taa_h != ldr->h) // :1451-1457
    {
        if (int rc = ensure_dev(ctx, ctx->d_taa_hist, count)) return rc;
        ctx->taa_w = ldr->w;
        ctx->taa_h = ldr->h;
        ctx->taa_valid = false;
    }
    if (!ctx->taa_valid) // :1461-1469: seed the history, leave the frame untouched
    {
        CK(cudaMemcpyAsync(ctx->d_taa_hist.p, ldr->color, count * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->launches += 1; // a device operation on the render stream (see shsb_rt_download_async)
        ctx->taa_valid = true;
        return SHSB_OK;
    }
    wait_pending_read(ctx, ldr);
    const float blend = 0.12f, keep = 1.0f - blend; // :1459-1460
    launch_taa((uchar4*)ldr->color, ctx->d_taa_hist.p, count, keep, blend, ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
    return SHSB_OK;
}

SHSB_API int32_t shsb_taa_reset(shsb_ctx ctx)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    ctx->taa_w = ctx->taa_h = 0;
    ctx->taa_valid = false;
    return SHSB_OK;
}

SHSB_API int32_t shsb_lights_upload(shsb_ctx ctx, const void* records, uint32_t n_lights)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (n_lights && !records) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "records is null");
    CK(cudaSetDevice(ctx->device));
    // The records go into the buffer the in-flight frames are NOT reading, on the front stream: the upload for
    // frame f+1 overlaps frame f's tile kernel.  It waits only for the last tile kernel / standalone cull that read
    // that buffer; later culls are ordered behind it on the front stream, and the main stream waits for it too.
    const int b = (ctx->lights_cur + 1) % ctx->n_arenas;
    cudaStream_t su = ctx->front_streams[ctx->frame_no % ctx->n_arenas]; // the stream the next frame's front end will use
    if (int rc = ensure_dev(ctx, ctx->d_lights[b], std::max(1u, n_lights))) return rc;
    if (int rc = ensure_dev(ctx, ctx->d_smlights[b], std::max(1u, n_lights))) return rc;
    for (int k = 0; k < MAX_TILE_STREAMS; ++k) // frames on different render streams finish in any order: wait for the last reader on each
        if (ctx->lights_last_user[b][k] >= 0) CK(cudaStreamWaitEvent(su, ctx->ev_tile_done[ctx->lights_last_user[b][k] % TILE_DONE_RING], 0));
    if (ctx->cull_main_pending) { CK(cudaStreamWaitEvent(su, ctx->ev_cull_main, 0)); ctx->cull_main_pending = false; }
    if (n_lights)
    {
        // The SMs pull the records straight from host memory (zero-copy) while digesting them: no copy-engine H2D that
        // could queue behind a frame read-back.  Pinned caller memory is read in place (it must stay unchanged until the
        // upload has run, like any cudaMemcpyAsync source); pageable memory is staged through a pinned ring first.
        const DevLightRec* src = nullptr;
        cudaPointerAttributes attr{};
        bool host_readable = true;
        if (cudaPointerGetAttributes(&attr, records) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            src = static_cast<const DevLightRec*>(attr.devicePointer);
        else if (cudaPointerGetAttributes(&attr, records) == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged))
        {
            src = static_cast<const DevLightRec*>(records);
            host_readable = attr.type == cudaMemoryTypeManaged;
        }
        else
        {
            cudaGetLastError();
            if (ctx->lights_stage_busy[b]) { CK(cudaEventSynchronize(ctx->lights_stage_done[b])); ctx->lights_stage_busy[b] = false; }
            if (int rc = ensure_pinned(ctx, ctx->h_lights[b], n_lights)) return rc;
            std::memcpy(ctx->h_lights[b].p, records, (size_t)n_lights * sizeof(DevLightRec));
            src = ctx->h_lights[b].dp;
            ctx->lights_stage_busy[b] = true;
        }
        // rect / tube lights go through the general per-record evaluation in the tile kernel; a set without any lets the frame run an
        // instantiation that does not carry that code (one word per record read here; device-resident records are not looked at)
        bool area = !host_readable;
        if (host_readable)
        {
            const DevLightRec* h = static_cast<const DevLightRec*>(records);
            for (uint32_t i = 0; i < n_lights && !area; ++i) area = h[i].type_shape_flags[0] > 2u;
        }
        ctx->area_lights = area;
        launch_light_prep(src, ctx->d_lights[b].p, ctx->d_smlights[b].p, n_lights, su, &ctx->launches);
        CK(cudaGetLastError());
        if (ctx->lights_stage_busy[b]) CK(cudaEventRecord(ctx->lights_stage_done[b], su));
    }
    CK(cudaEventRecord(ctx->ev_lights_up, su));
    ctx->lights_uploaded = true;
    ctx->main_needs_lights = true; // ordered lazily: a frame that culls reaches the records through its front end (main_wait_lights)
    ctx->lights_cur = b;
    for (long long& u : ctx->lights_last_user[b]) u = -1;
    ctx->n_lights = n_lights;
    ctx->lists_cur = -1;
    return SHSB_OK;
}

SHSB_API int32_t shsb_light_cull(shsb_ctx ctx, const float view_proj[16], uint32_t vw, uint32_t vh, uint32_t ts, uint32_t max_per_tile)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    CullJob job;
    if (int rc = prepare_light_cull(ctx, view_proj, vw, vh, ts, max_per_tile, job)) return rc;
    if (int rc = ensure_lists(ctx, ctx->lists[LISTS_STANDALONE], job)) return rc;
    if (int rc = main_wait_lights(ctx)) return rc;
    // main stream: ordered behind every tile kernel that read the standalone set
    enqueue_light_cull(ctx, job, LISTS_STANDALONE, ctx->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_cull_main, ctx->stream));
    ctx->cull_main_pending = true;
    return SHSB_OK;
}

namespace
{
    int tile_depth_range_common(shsb_ctx ctx, shsb_rt depth_rt, uint32_t tile_size, bool ndc01, float z_near, float z_far)
    {
        if (!ctx || tile_size == 0) return SHSB_E_INVALID_ARGUMENT;
        CK(cudaSetDevice(ctx->device));
        RtSlot* dm = ndc01 ? get_rt(ctx, depth_rt) : get_rt(ctx, depth_rt, SHSB_RT_DEPTH_MOTION);
        if (!dm || !dm->depth) return fail(ctx, SHSB_E_INVALID_HANDLE, ndc01 ? "tile depth range needs a live target with a depth plane" : "tile depth range needs a live RT_ColorDepthMotion");
        const size_t tiles = (size_t)((dm->w + tile_size - 1) / tile_size) * ((dm->h + tile_size - 1) / tile_size);
        if (int rc = ensure_dev(ctx, ctx->d_range_min, tiles)) return rc;
        if (int rc = ensure_dev(ctx, ctx->d_range_max, tiles)) return rc;
        launch_tile_depth_range(dm->depth, dm->w, dm->h, tile_size, ndc01 ? z_near : dm->zn, ndc01 ? z_far : dm->zf, ndc01 ? 1 : 0, ctx->d_range_min.p, ctx->d_range_max.p, ctx->stream,
                                &ctx->launches);
        CK(cudaGetLastError());
        ctx->range_w = (uint32_t)dm->w; ctx->range_h = (uint32_t)dm->h; ctx->range_ts = tile_size;
        return SHSB_OK;
    }
}

SHSB_API int32_t shsb_tile_depth_range(shsb_ctx ctx, shsb_rt depth_motion_rt, uint32_t tile_size)
{
    return tile_depth_range_common(ctx, depth_motion_rt, tile_size, false, 0.0f, 0.0f);
}

SHSB_API int32_t shsb_tile_depth_range_ndc01(shsb_ctx ctx, shsb_rt depth_rt, uint32_t tile_size, float z_near, float z_far)
{
    return tile_depth_range_common(ctx, depth_rt, tile_size, true, z_near, z_far);
}

SHSB_API int32_t shsb_tile_depth_range_download(shsb_ctx ctx, float* out_min, float* out_max, size_t n_tiles)
{
    if (!ctx || !out_min || !out_max) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->range_ts) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "no tile depth ranges: call shsb_tile_depth_range first");
    const size_t tiles = (size_t)((ctx->range_w + ctx->range_ts - 1) / ctx->range_ts) * ((ctx->range_h + ctx->range_ts - 1) / ctx->range_ts);
    if (n_tiles != tiles) return fail(ctx, SHSB_E_SIZE_MISMATCH, "caller passed %zu tiles, ranges have %zu", n_tiles, tiles);
    CK(cudaMemcpyAsync(out_min, ctx->d_range_min.p, tiles * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_max, ctx->d_range_max.p, tiles * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SHSB_OK;
}

SHSB_API int32_t shsb_light_cull_ex(shsb_ctx ctx, const ShsbLightCullDesc* d, const float* range_min, const float* range_max)
{
    if (!ctx || !d) return SHSB_E_INVALID_ARGUMENT;
    if (d->mode == SHSB_LIGHT_CULL_TILED) return shsb_light_cull(ctx, d->view_proj, d->viewport_w, d->viewport_h, d->tile_size, d->max_per_bin);
    if (d->mode < 0 || d->mode > SHSB_LIGHT_CULL_CLUSTERED) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "unknown light culling mode %d", d->mode);
    CK(cudaSetDevice(ctx->device));
    CullJob job;
    if (int rc = prepare_light_cull(ctx, d->view_proj, d->viewport_w, d->viewport_h, d->tile_size, d->max_per_bin, job)) return rc;
    const uint32_t tiles = ((job.vw + job.ts - 1) / job.ts) * ((job.vh + job.ts - 1) / job.ts);
    const bool clustered = d->mode == SHSB_LIGHT_CULL_CLUSTERED;
    const uint32_t slices = clustered ? d->depth_slices : 1u;
    if (slices == 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "depth_slices must be > 0");
    const float* dev_min = nullptr;
    const float* dev_max = nullptr;
    if (!clustered)
    {
        if ((range_min == nullptr) != (range_max == nullptr)) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "range_min and range_max must both be given or both be NULL");
        if (range_min)
        {
            if (int rc = ensure_dev(ctx, ctx->d_range_up_min, tiles)) return rc;
            if (int rc = ensure_dev(ctx, ctx->d_range_up_max, tiles)) return rc;
            CK(cudaMemcpyAsync(ctx->d_range_up_min.p, range_min, (size_t)tiles * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_range_up_max.p, range_max, (size_t)tiles * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream)); // the caller's arrays may be pageable and short-lived
            dev_min = ctx->d_range_up_min.p; dev_max = ctx->d_range_up_max.p;
        }
        else
        {
            if (ctx->range_ts != job.ts || ctx->range_w != job.vw || ctx->range_h != job.vh)
                return fail(ctx, SHSB_E_INVALID_ARGUMENT, "no device tile depth ranges for a %ux%u viewport with %u-px tiles: call shsb_tile_depth_range first", job.vw, job.vh, job.ts);
            dev_min = ctx->d_range_min.p; dev_max = ctx->d_range_max.p;
        }
    }
    else
    {
        // slice boundaries on the host (std::log / std::exp of the host libm, exactly like the reference: jolt_light_culling.hpp:373-382)
        std::vector<float2> ndc(slices);
        const float log_ratio = std::log(d->z_far / d->z_near);
        for (uint32_t cz = 0; cz < slices; ++cz)
        {
            const float slice_near = d->z_near * std::exp(log_ratio * static_cast<float>(cz) / static_cast<float>(slices));
            const float slice_far = d->z_near * std::exp(log_ratio * static_cast<float>(cz + 1) / static_cast<float>(slices));
            ndc[cz] = make_float2(hm::ndc_from_view_depth_lh_no(slice_near, d->z_near, d->z_far), hm::ndc_from_view_depth_lh_no(slice_far, d->z_near, d->z_far));
        }
        if (int rc = ensure_dev(ctx, ctx->d_slice_ndc, slices)) return rc;
        CK(cudaMemcpyAsync(ctx->d_slice_ndc.p, ndc.data(), (size_t)slices * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    LightLists& L = clustered ? ctx->cluster_lists : ctx->lists[LISTS_STANDALONE];
    const size_t bins = (size_t)tiles * slices;
    if (int rc = ensure_dev(ctx, L.counts, bins)) return rc;
    if (int rc = ensure_dev(ctx, L.indices, bins * job.max_per_tile)) return rc;
    if (int rc = ensure_dev(ctx, ctx->d_vis, (size_t)ctx->n_lights + 1)) return rc;
    if (int rc = ensure_dev(ctx, ctx->d_lights[ctx->lights_cur], 1)) return rc;
    if (int rc = main_wait_lights(ctx)) return rc;
    record(ctx, 5, ctx->stream);
    launch_light_cull_cells(ctx->d_lights[ctx->lights_cur].p, ctx->n_lights, job.planes, job.inv_vp, job.vw, job.vh, job.ts, job.max_per_tile, d->mode, slices,
                            dev_min, dev_max, ctx->d_slice_ndc.p, d->z_near, d->z_far, ctx->d_vis.p, L.counts.p, L.indices.p, ctx->stream, &ctx->launches);
    record(ctx, 6, ctx->stream);
    CK(cudaGetLastError());
    L.w = job.vw; L.h = job.vh; L.ts = job.ts; L.max_per_tile = job.max_per_tile;
    if (clustered) ctx->cluster_slices = slices;
    else
    {
        ctx->lists_cur = LISTS_STANDALONE;
        CK(cudaEventRecord(ctx->ev_cull_main, ctx->stream));
        ctx->cull_main_pending = true;
    }
    return SHSB_OK;
}

SHSB_API int32_t shsb_cluster_lists_download(shsb_ctx ctx, uint32_t* counts, size_t n_counts, uint32_t* indices, size_t n_indices)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->cluster_slices) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "no cluster bins: call shsb_light_cull_ex(SHSB_LIGHT_CULL_CLUSTERED) first");
    const LightLists& L = ctx->cluster_lists;
    const size_t bins = (size_t)((L.w + L.ts - 1) / L.ts) * ((L.h + L.ts - 1) / L.ts) * ctx->cluster_slices;
    if (counts && n_counts != bins) return fail(ctx, SHSB_E_SIZE_MISMATCH, "counts has %zu entries, there are %zu bins", n_counts, bins);
    if (indices && n_indices != bins * L.max_per_tile) return fail(ctx, SHSB_E_SIZE_MISMATCH, "indices has %zu entries, expected %zu", n_indices, bins * L.max_per_tile);
    if (counts) CK(cudaMemcpyAsync(counts, L.counts.p, bins * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (indices) CK(cudaMemcpyAsync(indices, L.indices.p, bins * L.max_per_tile * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SHSB_OK;
}

SHSB_API int32_t shsb_light_lists_download(shsb_ctx ctx, uint32_t* counts, size_t n_counts, uint32_t* indices, size_t n_indices)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (ctx->lists_cur < 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "no tile light lists: call shsb_light_cull first");
    const LightLists& L = ctx->lists[ctx->lists_cur];
    const size_t tiles = (size_t)((L.w + L.ts - 1) / L.ts) * ((L.h + L.ts - 1) / L.ts);
    if (counts && n_counts != tiles) return fail(ctx, SHSB_E_SIZE_MISMATCH, "counts has %zu entries, lists have %zu tiles", n_counts, tiles);
    if (indices && n_indices != tiles * L.max_per_tile) return fail(ctx, SHSB_E_SIZE_MISMATCH, "indices has %zu entries, expected %zu", n_indices, tiles * L.max_per_tile);
    sync_all(ctx); // a fused frame builds its lists on its front-end streams
    if (counts) CK(cudaMemcpyAsync(counts, L.counts.p, tiles * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (indices) CK(cudaMemcpyAsync(indices, L.indices.p, tiles * L.max_per_tile * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SHSB_OK;
}

// ---------------------------------------------------------------------------------------- sort-first frame assembly
namespace
{
    GatherSlot* get_gather(shsb_ctx ctx, shsb_gather g)
    {
        if (g == 0 || g > ctx->gathers.size() || !ctx->gathers[g - 1].live) return nullptr;
        return &ctx->gathers[g - 1];
    }

    int gather_prepare(shsb_ctx ctx)
    {
        CK(cudaSetDevice(ctx->device));
        if (!ctx->gather_stream) CK(cudaStreamCreateWithFlags(&ctx->gather_stream, cudaStreamNonBlocking));
        if (!ctx->h_gather_timeout)
        {
            CK(cudaHostAlloc(&ctx->h_gather_timeout, sizeof(uint32_t), cudaHostAllocMapped));
            *ctx->h_gather_timeout = 0u;
            CK(cudaHostGetDevicePointer(&ctx->d_gather_timeout, ctx->h_gather_timeout, 0));
        }
        if (*ctx->h_gather_timeout) return fail(ctx, SHSB_E_TIMEOUT, "a frame-assembly wait timed out (a rank never committed / the root never released a step)");
        return SHSB_OK;
    }
}

SHSB_API int32_t shsb_gather_create(shsb_ctx ctx, uint32_t n_ranks, uint32_t slots, size_t slot_bytes, shsb_gather* out_gather, ShsbGatherExport* out_export)
{
    if (!ctx || !out_gather || !out_export) return SHSB_E_INVALID_ARGUMENT;
    if (n_ranks == 0 || n_ranks > SHSB_GATHER_MAX_RANKS || slots == 0 || slot_bytes == 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "bad frame-assembly description");
    if (int rc = gather_prepare(ctx)) return rc;
    GatherSlot g;
    g.live = true; g.root = true; g.n_ranks = n_ranks; g.rank = 0; g.slots = slots; g.slot_bytes = slot_bytes;
    if (cudaMalloc(&g.base, slot_bytes * slots) != cudaSuccess) { cudaGetLastError(); return fail(ctx, SHSB_E_OUT_OF_MEMORY, "assembly memory: %zu bytes", slot_bytes * slots); }
    if (cudaMalloc(&g.ctl, sizeof(GatherCtl)) != cudaSuccess) { cudaGetLastError(); cudaFree(g.base); return fail(ctx, SHSB_E_OUT_OF_MEMORY, "assembly control block"); }
    CK(cudaMemset(g.ctl, 0, sizeof(GatherCtl)));
    CK(cudaDeviceSynchronize());
    *out_export = ShsbGatherExport{};
    cudaIpcMemHandle_t hm{}, hc{};
    // legacy IPC handles: valid for other processes; ranks of this process use the raw pointers below
    if (cudaIpcGetMemHandle(&hm, g.base) == cudaSuccess && cudaIpcGetMemHandle(&hc, g.ctl) == cudaSuccess)
    {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(out_export->mem_handle, &hm, 64);
        std::memcpy(out_export->ctl_handle, &hc, 64);
    }
    else cudaGetLastError();
    out_export->mem_ptr = (uint64_t)(uintptr_t)g.base;
    out_export->ctl_ptr = (uint64_t)(uintptr_t)g.ctl;
    out_export->slot_bytes = slot_bytes;
    out_export->n_ranks = n_ranks; out_export->slots = slots;
    out_export->root_device = ctx->device;
    out_export->root_pid = (int32_t)getpid();
    ctx->gathers.push_back(g);
    *out_gather = (shsb_gather)ctx->gathers.size();
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_open(shsb_ctx ctx, const ShsbGatherExport* exp, uint32_t rank, shsb_gather* out_gather)
{
    if (!ctx || !exp || !out_gather) return SHSB_E_INVALID_ARGUMENT;
    if (rank == 0 || rank >= exp->n_ranks || exp->n_ranks > SHSB_GATHER_MAX_RANKS || exp->slots == 0 || exp->slot_bytes == 0)
        return fail(ctx, SHSB_E_INVALID_ARGUMENT, "bad frame-assembly export / rank %u of %u", rank, exp->n_ranks);
    if (int rc = gather_prepare(ctx)) return rc;
    GatherSlot g;
    g.live = true; g.root = false; g.n_ranks = exp->n_ranks; g.rank = rank; g.slots = exp->slots; g.slot_bytes = (size_t)exp->slot_bytes;
    if (exp->root_pid == (int32_t)getpid())
    {
        // same process (several contexts of one host program): the allocations are directly addressable (unified addressing)
        if (exp->root_device != ctx->device)
        {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->device, exp->root_device));
            if (!can) return fail(ctx, SHSB_E_UNSUPPORTED, "device %d cannot access device %d's memory (no peer-to-peer path)", ctx->device, exp->root_device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(exp->root_device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, SHSB_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        g.base = (unsigned char*)(uintptr_t)exp->mem_ptr;
        g.ctl = (GatherCtl*)(uintptr_t)exp->ctl_ptr;
    }
    else
    {
        cudaIpcMemHandle_t hm{}, hc{};
        std::memcpy(&hm, exp->mem_handle, 64);
        std::memcpy(&hc, exp->ctl_handle, 64);
        void* pm = nullptr; void* pc = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&pm, hm, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&pc, hc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); if (pm) cudaIpcCloseMemHandle(pm); return fail(ctx, SHSB_E_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
        g.base = (unsigned char*)pm;
        g.ctl = (GatherCtl*)pc;
        g.ipc = true;
    }
    ctx->gathers.push_back(g);
    *out_gather = (shsb_gather)ctx->gathers.size();
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_destroy(shsb_ctx ctx, shsb_gather gather)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    GatherSlot* g = get_gather(ctx, gather);
    if (!g) return fail(ctx, SHSB_E_INVALID_HANDLE, "frame-assembly handle %u is not live", gather);
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    if (g->root) { cudaFree(g->base); cudaFree(g->ctl); }
    else if (g->ipc) { cudaIpcCloseMemHandle(g->base); cudaIpcCloseMemHandle(g->ctl); }
    *g = GatherSlot{};
    return SHSB_OK;
}

SHSB_API int32_t shsb_frame_gather(shsb_ctx ctx, shsb_gather gather, uint64_t step, shsb_rt rt, int32_t plane, size_t src_offset, size_t bytes, size_t dst_offset)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (int rc = gather_prepare(ctx)) return rc;
    GatherSlot* g = get_gather(ctx, gather);
    if (!g) return fail(ctx, SHSB_E_INVALID_HANDLE, "frame-assembly handle %u is not live", gather);
    RtSlot* r = peek_rt(ctx, rt);
    if (!r) return fail(ctx, SHSB_E_INVALID_HANDLE, "render target %u is not live", rt);
    void* p = nullptr;
    const size_t have = plane_bytes(*r, plane, &p);
    if (!have) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "render target %u has no plane %d", rt, plane);
    if (src_offset > have || bytes > have - src_offset) return fail(ctx, SHSB_E_SIZE_MISMATCH, "plane is %zu bytes, push asks for [%zu, %zu)", have, src_offset, src_offset + bytes);
    if (dst_offset > g->slot_bytes || bytes > g->slot_bytes - dst_offset) return fail(ctx, SHSB_E_SIZE_MISMATCH, "slot is %zu bytes, push targets [%zu, %zu)", g->slot_bytes, dst_offset, dst_offset + bytes);
    if (step == 0 || step < g->begun || step <= g->committed) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "steps are numbered from 1 and never go back (step %llu after %llu)", (unsigned long long)step, (unsigned long long)g->begun);
    cudaStream_t gs = ctx->gather_stream;
    if (step > g->begun)
    {
        // first push of a new step: its slot was last used by step - slots, which the root must have released
        if (step > g->slots)
        {
            if (g->root)
            {
                if (g->released < step - g->slots) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "step %llu reuses the slot of step %llu, which has not been released", (unsigned long long)step, (unsigned long long)(step - g->slots));
            }
            else launch_gather_wait(&g->ctl->released, 1, step - g->slots, GATHER_TIMEOUT_NS, ctx->d_gather_timeout, gs, &ctx->launches);
        }
        g->begun = step;
    }
    // the push starts when whatever wrote the target last has finished (same rule as shsb_rt_download_async)
    if (r->frame_is_last && r->last_frame >= 0) CK(cudaStreamWaitEvent(gs, ctx->ev_tile_done[r->last_frame % TILE_DONE_RING], 0));
    else
    {
        CK(cudaEventRecord(ctx->ev_frame_done, ctx->stream));
        CK(cudaStreamWaitEvent(gs, ctx->ev_frame_done, 0));
    }
    if (!r->read_done) CK(cudaEventCreateWithFlags(&r->read_done, cudaEventDisableTiming));
    if (r->read_pending) CK(cudaStreamWaitEvent(gs, r->read_done, 0));
    unsigned char* dst = g->base + (size_t)((step - 1) % g->slots) * g->slot_bytes + dst_offset;
    if (bytes) CK(cudaMemcpyAsync(dst, (const unsigned char*)p + src_offset, bytes, cudaMemcpyDeviceToDevice, gs)); // copy engine; peer write over NVLink for a remote root
    CK(cudaEventRecord(r->read_done, gs));
    r->read_pending = true; // a later pass that writes the target waits for the push
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_commit(shsb_ctx ctx, shsb_gather gather, uint64_t step)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (int rc = gather_prepare(ctx)) return rc;
    GatherSlot* g = get_gather(ctx, gather);
    if (!g) return fail(ctx, SHSB_E_INVALID_HANDLE, "frame-assembly handle %u is not live", gather);
    if (step == 0 || step <= g->committed || step < g->begun) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "commit of step %llu after step %llu", (unsigned long long)step, (unsigned long long)std::max(g->committed, g->begun));
    if (step > g->begun && step > g->slots && !g->root)
        launch_gather_wait(&g->ctl->released, 1, step - g->slots, GATHER_TIMEOUT_NS, ctx->d_gather_timeout, ctx->gather_stream, &ctx->launches); // a step without pushes still keeps the ring order
    launch_gather_signal(&g->ctl->arrive[g->rank], step, ctx->gather_stream, &ctx->launches);
    CK(cudaGetLastError());
    g->begun = g->committed = step;
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_wait(shsb_ctx ctx, shsb_gather gather, uint64_t step)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (int rc = gather_prepare(ctx)) return rc;
    GatherSlot* g = get_gather(ctx, gather);
    if (!g || !g->root) return fail(ctx, SHSB_E_INVALID_HANDLE, "frame-assembly handle %u is not a live root handle", gather);
    if (step == 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "steps are numbered from 1");
    launch_gather_wait(g->ctl->arrive, g->n_ranks, step, GATHER_TIMEOUT_NS, ctx->d_gather_timeout, ctx->gather_stream, &ctx->launches);
    CK(cudaGetLastError());
    g->waited = std::max<uint64_t>(g->waited, step);
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_release(shsb_ctx ctx, shsb_gather gather, uint64_t step)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    if (int rc = gather_prepare(ctx)) return rc;
    GatherSlot* g = get_gather(ctx, gather);
    if (!g || !g->root) return fail(ctx, SHSB_E_INVALID_HANDLE, "frame-assembly handle %u is not a live root handle", gather);
    if (step == 0 || step <= g->released) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "release of step %llu after step %llu", (unsigned long long)step, (unsigned long long)g->released);
    if (g->waited < step) { if (int rc = shsb_gather_wait(ctx, gather, step)) return rc; } // releasing a step nobody waited for would let ranks overwrite pushes still in flight
    launch_gather_signal(&g->ctl->released, step, ctx->gather_stream, &ctx->launches);
    CK(cudaGetLastError());
    g->released = step;
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_device_ptr(shsb_ctx ctx, shsb_gather gather, uint64_t step, void** out_ptr)
{
    if (!ctx || !out_ptr) return SHSB_E_INVALID_ARGUMENT;
    GatherSlot* g = get_gather(ctx, gather);
    if (!g || !g->root) return fail(ctx, SHSB_E_INVALID_HANDLE, "frame-assembly handle %u is not a live root handle", gather);
    if (step == 0) return fail(ctx, SHSB_E_INVALID_ARGUMENT, "steps are numbered from 1");
    *out_ptr = g->base + (size_t)((step - 1) % g->slots) * g->slot_bytes;
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_download_async(shsb_ctx ctx, shsb_gather gather, uint64_t step, size_t offset, void* dst_pinned, size_t bytes)
{
    if (!ctx || !dst_pinned) return SHSB_E_INVALID_ARGUMENT;
    if (int rc = gather_prepare(ctx)) return rc;
    GatherSlot* g = get_gather(ctx, gather);
    if (!g || !g->root) return fail(ctx, SHSB_E_INVALID_HANDLE, "frame-assembly handle %u is not a live root handle", gather);
    if (step == 0 || offset > g->slot_bytes || bytes > g->slot_bytes - offset) return fail(ctx, SHSB_E_SIZE_MISMATCH, "slot is %zu bytes, read asks for [%zu, %zu)", g->slot_bytes, offset, offset + bytes);
    if (g->waited < step) { if (int rc = shsb_gather_wait(ctx, gather, step)) return rc; }
    CK(cudaMemcpyAsync(dst_pinned, g->base + (size_t)((step - 1) % g->slots) * g->slot_bytes + offset, bytes, cudaMemcpyDeviceToHost, ctx->gather_stream));
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_download(shsb_ctx ctx, shsb_gather gather, uint64_t step, size_t offset, void* dst, size_t bytes)
{
    if (int rc = shsb_gather_download_async(ctx, gather, step, offset, dst, bytes)) return rc;
    CK(cudaStreamSynchronize(ctx->gather_stream));
    if (*ctx->h_gather_timeout) return fail(ctx, SHSB_E_TIMEOUT, "a frame-assembly wait timed out (a rank never committed step %llu)", (unsigned long long)step);
    return SHSB_OK;
}

SHSB_API int32_t shsb_gather_stream(shsb_ctx ctx, void** out_stream)
{
    if (!ctx || !out_stream) return SHSB_E_INVALID_ARGUMENT;
    if (int rc = gather_prepare(ctx)) return rc;
    *out_stream = (void*)ctx->gather_stream;
    return SHSB_OK;
}

SHSB_API int32_t shsb_timing_enable(shsb_ctx ctx, int32_t enable)
{
    if (!ctx) return SHSB_E_INVALID_ARGUMENT;
    ctx->timing_on = enable != 0;
    ctx->timing_stride = std::max(1, enable);
    ctx->timing_used = 0;
    return SHSB_OK;
}

SHSB_API int32_t shsb_timing_collect(shsb_ctx ctx, float* out_ms, size_t cap_frames, size_t* out_frames)
{
    if (!ctx || !out_frames) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    for (cudaStream_t st : ctx->tile_streams) CK(cudaStreamSynchronize(st));
    for (cudaStream_t st : ctx->front_streams) CK(cudaStreamSynchronize(st));
    const size_t frames = std::min(ctx->timing_used / TIMING_EVENTS_PER_FRAME, cap_frames);
    for (size_t f = 0; f < frames && out_ms; ++f)
    {
        // front-end begin, after geometry, after binning (front stream) | tile kernel begin, end (main stream)
        const cudaEvent_t* e = &ctx->timing_ev[f * TIMING_EVENTS_PER_FRAME];
        float g = 0, b = 0, r = 0, t = 0;
        cudaEventElapsedTime(&g, e[0], e[1]);
        cudaEventElapsedTime(&b, e[1], e[2]);
        cudaEventElapsedTime(&r, e[3], e[4]);
        cudaEventElapsedTime(&t, e[0], e[4]);
        out_ms[f * 4 + 0] = g; out_ms[f * 4 + 1] = b; out_ms[f * 4 + 2] = r; out_ms[f * 4 + 3] = t;
    }
    *out_frames = frames;
    ctx->timing_used = 0;
    return SHSB_OK;
}

/* Debug: like shsb_timing_collect but absolute times -- per frame the 5 stage events (front-end begin, after
 * geometry, after binning, tile begin, tile end) in milliseconds since the first recorded event. */
SHSB_API int32_t shsb_timing_collect_abs(shsb_ctx ctx, float* out_ms, size_t cap_frames, size_t* out_frames)
{
    if (!ctx || !out_frames) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    for (cudaStream_t st : ctx->front_streams) CK(cudaStreamSynchronize(st));
    for (cudaStream_t st : ctx->tile_streams) CK(cudaStreamSynchronize(st));
    const size_t frames = std::min(ctx->timing_used / TIMING_EVENTS_PER_FRAME, cap_frames);
    for (size_t f = 0; f < frames && out_ms; ++f)
        for (int k = 0; k < TIMING_EVENTS_PER_FRAME; ++k)
        {
            float t = 0;
            cudaEventElapsedTime(&t, ctx->timing_ev[0], ctx->timing_ev[f * TIMING_EVENTS_PER_FRAME + k]);
            out_ms[f * TIMING_EVENTS_PER_FRAME + k] = t;
        }
    *out_frames = frames;
    ctx->timing_used = 0;
    return SHSB_OK;
}

SHSB_API int32_t shsb_host_submit_us(shsb_ctx ctx, double out_us[8], int32_t reset)
{
    if (!ctx || !out_us) return SHSB_E_INVALID_ARGUMENT;
    for (int i = 0; i < 8; ++i) out_us[i] = ctx->host_us[i];
    if (reset) for (double& v : ctx->host_us) v = 0.0;
    return SHSB_OK;
}

SHSB_API int32_t shsb_last_stage_ms(shsb_ctx ctx, float out_ms[8])
{
    if (!ctx || !out_ms) return SHSB_E_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    for (cudaStream_t st : ctx->front_streams) CK(cudaStreamSynchronize(st));
    for (cudaStream_t st : ctx->tile_streams) CK(cudaStreamSynchronize(st));
    for (int i = 0; i < 8; ++i) out_ms[i] = 0.0f;
    auto span = [&](int a, int b) -> float {
        float ms = 0.0f;
        if (ctx->ev_valid[a] && ctx->ev_valid[b] && cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]) == cudaSuccess) return ms;
        cudaGetLastError();
        return 0.0f;
    };
    out_ms[0] = span(0, 1);
    out_ms[1] = span(1, 2);
    out_ms[2] = span(3, 4);
    out_ms[3] = span(5, 6); // light cull or standalone tonemap, whichever ran last
    out_ms[5] = span(0, 4);
    return SHSB_OK;
}

} // extern "C"
