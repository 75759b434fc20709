// tests/cpp/gather_test.cpp -- a two-RANK sort-first frame through include/shsb.h ONLY (no Python, no torch, no NCCL):
// the process forks; parent = rank 0 (root), child = rank 1, each with its own context (GPU 0 and GPU 1 when the box has two,
// otherwise both on GPU 0).  Every rank renders its half of a 4-camera batch AND its band of one split frame, pushes the LDR
// pixels into the root's assembly memory (CUDA IPC handle shipped over a pipe, copy-engine peer writes), commits; the root waits
// on the device, downloads the slot and compares it with its own one-at-a-time rendering of the same cameras / frame.
// Exit codes: 0 = pass, 77 = no CUDA device (skipped), anything else = failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

#include "shsb.h"

#define CHECK(call)                                                                                              \
    do                                                                                                           \
    {                                                                                                            \
        const int32_t rc_ = (call);                                                                              \
        if (rc_ != SHSB_OK)                                                                                      \
        {                                                                                                        \
            std::printf("rank %d: %s -> %d (%s)\n", g_rank, #call, rc_, g_ctx ? shsb_last_error_string(g_ctx) : "");  \
            std::fflush(stdout);                                                                                 \
            _exit(1);                                                                                            \
        }                                                                                                        \
    } while (0)

static int g_rank = 0;
static shsb_ctx g_ctx = nullptr;

static const int W = 256, H = 160, CAMS = 4; // 10 tile rows

struct Rank
{
    shsb_mesh mesh = 0;
    shsb_rt hdr[2]{}, dm[2]{}, ldr[2]{};
    ShsbRenderItem items[3]{};
    ShsbFrameParams fp{};
};

static void setup(Rank& r)
{
    // a few triangles in front of the cameras
    const float pos[] = {-2, -1, 0, 2, -1, 0, 0, 2, 0, -1.5f, -0.5f, 1, 1.5f, -0.5f, 1, 0, 1.5f, -1};
    const float nrm[] = {0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0.6f, -0.8f, 0, 0.6f, -0.8f, 0, 0.6f, -0.8f};
    const float uv[] = {0, 0, 1, 0, 0.5f, 1, 0, 0, 1, 0, 0.5f, 1};
    const uint32_t idx[] = {0, 2, 1, 3, 5, 4, 0, 1, 2, 3, 4, 5};
    CHECK(shsb_mesh_upload(g_ctx, pos, 6, nrm, 6, uv, 6, idx, 12, &r.mesh));
    for (int k = 0; k < 2; ++k)
    {
        CHECK(shsb_rt_create(g_ctx, SHSB_RT_COLOR_HDR, W, H, 0, 0, &r.hdr[k]));
        CHECK(shsb_rt_create(g_ctx, SHSB_RT_DEPTH_MOTION, W, H, 0.1f, 100.0f, &r.dm[k]));
        CHECK(shsb_rt_create(g_ctx, SHSB_RT_COLOR_LDR, W, H, 0, 0, &r.ldr[k]));
    }
    for (int i = 0; i < 3; ++i)
    {
        ShsbRenderItem& it = r.items[i];
        std::memset(&it, 0, sizeof(it));
        it.mesh = r.mesh; it.visible = 1; it.casts_shadow = 1; it.has_material = 1;
        it.tr.pos[0] = (float)(i - 1) * 2.5f; it.tr.pos[1] = 0.2f * (float)i; it.tr.pos[2] = (float)i;
        it.tr.rot_euler[1] = 0.7f * (float)i;
        it.tr.scl[0] = it.tr.scl[1] = it.tr.scl[2] = 1.0f + 0.25f * (float)i;
        it.base_color[0] = 0.9f - 0.3f * (float)i; it.base_color[1] = 0.4f + 0.2f * (float)i; it.base_color[2] = 0.3f;
        it.metallic = 0.2f * (float)i; it.roughness = 0.35f + 0.2f * (float)i; it.ao = 1.0f;
    }
    std::memset(&r.fp, 0, sizeof(r.fp));
    r.fp.shading_model = SHSB_SHADING_PBR_METAL_ROUGH; r.fp.cull_mode = SHSB_CULL_NONE; r.fp.front_face_ccw = 1;
    r.fp.exposure = 1.2f; r.fp.gamma = 2.2f; r.fp.tile_size = 16; r.fp.max_lights_per_tile = 64;
}

static void scene_of_camera(const Rank& r, int cam, ShsbScene& s)
{
    std::memset(&s, 0, sizeof(s));
    s.items = r.items; s.n_items = 3;
    const float a = 6.2831853f * (float)cam / (float)CAMS;
    const float eye[3] = {9.0f * std::sin(a), 3.0f, -9.0f * std::cos(a)}, target[3] = {0, 0, 0}, up[3] = {0, 1, 0};
    CHECK(shsb_camera_viewproj(eye, target, up, 1.0f, (float)W / (float)H, 0.1f, 100.0f, s.cam_viewproj));
    std::memcpy(s.cam_prev_viewproj, s.cam_viewproj, 64);
    std::memcpy(s.cam_pos, eye, 12);
    s.sun_dir_ws[0] = -0.4f; s.sun_dir_ws[1] = -1.0f; s.sun_dir_ws[2] = 0.3f;
    s.sun_color[0] = 1.0f; s.sun_color[1] = 0.96f; s.sun_color[2] = 0.9f;
    s.sun_intensity = 2.5f;
}

int main()
{
    int pipe_exp[2], pipe_ack[2];
    if (pipe(pipe_exp) || pipe(pipe_ack)) return 2;
    std::fflush(stdout);
    const pid_t child = fork();
    g_rank = child == 0 ? 1 : 0;

    // device: rank r on GPU r when there are two, both on GPU 0 otherwise (probe by creating a context)
    int32_t rc = shsb_context_create(g_rank, &g_ctx);
    int dev = g_rank;
    if (rc != SHSB_OK) { rc = shsb_context_create(0, &g_ctx); dev = 0; }
    if (rc == SHSB_E_NO_DEVICE)
    {
        if (g_rank == 0) { std::printf("SKIP: no CUDA device (there is no CPU fallback)\n"); int st; waitpid(child, &st, 0); }
        return 77;
    }
    if (rc != SHSB_OK) { std::printf("rank %d: context_create -> %d\n", g_rank, rc); return 1; }

    Rank r;
    setup(r);
    const size_t frame_bytes = (size_t)W * H * 4;
    const size_t slot_bytes = (CAMS + 1) * frame_bytes; // 4 cameras + one split frame
    shsb_gather g = 0;
    if (g_rank == 0)
    {
        ShsbGatherExport exp;
        CHECK(shsb_gather_create(g_ctx, 2, 2, slot_bytes, &g, &exp));
        if (write(pipe_exp[1], &exp, sizeof(exp)) != (ssize_t)sizeof(exp)) return 2;
    }
    else
    {
        ShsbGatherExport exp;
        if (read(pipe_exp[0], &exp, sizeof(exp)) != (ssize_t)sizeof(exp)) return 2;
        CHECK(shsb_gather_open(g_ctx, &exp, 1, &g));
    }

    // the split frame: camera 0, tile rows [0, 6) from the top -> rank 0, [6, 10) -> rank 1
    const int cuts[3] = {0, 6, 10};
    const int STEPS = 5; // 2 slots: the ring wraps twice
    for (int step = 1; step <= STEPS; ++step)
    {
        int k = 0;
        for (int cam = g_rank; cam < CAMS; cam += 2, ++k)
        {
            ShsbScene s;
            scene_of_camera(r, (cam + step) % CAMS, s);
            CHECK(shsb_frame_forward_plus(g_ctx, &s, &r.fp, r.hdr[k & 1], r.dm[k & 1], r.ldr[k & 1], nullptr));
            CHECK(shsb_frame_gather(g_ctx, g, (uint64_t)step, r.ldr[k & 1], SHSB_PLANE_COLOR, 0, frame_bytes, (size_t)cam * frame_bytes));
        }
        {
            ShsbScene s;
            scene_of_camera(r, step % CAMS, s);
            ShsbFrameParams f = r.fp;
            f.own_row_first = cuts[g_rank]; f.own_row_count = cuts[g_rank + 1] - cuts[g_rank]; f.own_row_stride = 10;
            CHECK(shsb_frame_forward_plus(g_ctx, &s, &f, r.hdr[k & 1], r.dm[k & 1], r.ldr[k & 1], nullptr));
            const int y0 = H - 16 * cuts[g_rank + 1], y1 = H - 16 * cuts[g_rank]; // bottom-up pixel rows of the band
            const size_t off = (size_t)y0 * W * 4, n = (size_t)(y1 - y0) * W * 4;
            CHECK(shsb_frame_gather(g_ctx, g, (uint64_t)step, r.ldr[k & 1], SHSB_PLANE_COLOR, off, n, (size_t)CAMS * frame_bytes + off));
        }
        CHECK(shsb_gather_commit(g_ctx, g, (uint64_t)step));
        if (g_rank == 0)
        {
            std::vector<unsigned char> got(slot_bytes), want(frame_bytes);
            CHECK(shsb_gather_download(g_ctx, g, (uint64_t)step, 0, got.data(), slot_bytes));
            CHECK(shsb_gather_release(g_ctx, g, (uint64_t)step));
            for (int cam = 0; cam <= CAMS; ++cam) // CAMS = the split frame (camera `step`)
            {
                ShsbScene s;
                scene_of_camera(r, cam < CAMS ? (cam + step) % CAMS : step % CAMS, s);
                ShsbStats st;
                CHECK(shsb_frame_forward_plus(g_ctx, &s, &r.fp, r.hdr[0], r.dm[0], r.ldr[0], &st));
                CHECK(shsb_rt_download(g_ctx, r.ldr[0], SHSB_PLANE_COLOR, want.data(), frame_bytes));
                if (st.frag_shaded == 0) { std::printf("step %d camera %d: nothing was drawn\n", step, cam); return 1; }
                if (std::memcmp(got.data() + (size_t)cam * frame_bytes, want.data(), frame_bytes) != 0)
                {
                    size_t bad = 0;
                    for (size_t i = 0; i < frame_bytes; ++i) bad += got[(size_t)cam * frame_bytes + i] != want[i];
                    std::printf("FAIL step %d %s %d: %zu bytes differ\n", step, cam < CAMS ? "camera" : "split frame", cam, bad);
                    return 1;
                }
            }
        }
    }
    CHECK(shsb_sync(g_ctx));
    if (g_rank == 1)
    {
        char ok = 0;
        if (read(pipe_ack[0], &ok, 1) != 1) return 2; // keep the mapping alive until the root is done
        CHECK(shsb_gather_destroy(g_ctx, g));
        shsb_context_destroy(g_ctx);
        _exit(0);
    }
    const char ok = 1;
    if (write(pipe_ack[1], &ok, 1) != 1) return 2;
    int st = 0;
    waitpid(child, &st, 0);
    CHECK(shsb_gather_destroy(g_ctx, g));
    shsb_context_destroy(g_ctx);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { std::printf("rank 1 failed (status %d)\n", st); return 1; }
    std::printf("gather_test: %d steps x (%d cameras + 1 split frame) assembled over 2 ranks (rank 1 on GPU %d): OK\n", STEPS, CAMS, dev);
    return 0;
}
