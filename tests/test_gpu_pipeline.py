"""GPU tests of the pipelined frame submission (`-m gpu`): the front end of later frames overlaps the tile kernel
of earlier ones in rotating arenas, light uploads go to a ring of buffers, tile light lists exist once per arena.
None of that may change a single bit of any frame."""
import numpy as np
import pytest

import harness
from leisure_software_renderer_b200 import capi, scenes

pytestmark = pytest.mark.gpu


def _download(gpu, g):
    return (gpu.rt_download(g.hdr).view(np.uint32), gpu.rt_download(g.dm, capi.PLANE_DEPTH).view(np.uint32), gpu.rt_download(g.ldr))


def test_async_frames_with_changing_lights_and_cameras(gpu):
    """Eight asynchronous frames in flight (no statistics => no host synchronisation), each with its own light set
    and camera, each into its own render targets, equal one-at-a-time synchronous renders bit for bit."""
    base = scenes.scene_small(w=320, h=200, lights=48, tex=True)
    cams = scenes.camera_ring(base, 8, radius=9.0, height=4.0)
    light_sets = [scenes.make_lights(36, 12, (-6, 0.2, -6), (6, 3.0, 6), seed=s) for s in range(8)]
    g = harness.GpuScene(gpu, base)
    extra = [(gpu.rt_create(capi.RT_COLOR_HDR, base.w, base.h), gpu.rt_create(capi.RT_DEPTH_MOTION, base.w, base.h, base.zn, base.zf),
              gpu.rt_create(capi.RT_COLOR_LDR, base.w, base.h)) for _ in range(8)]
    try:
        fp = capi.FrameParams.from_buffer_copy(base.fp)
        fp.light_culling = 1
        ref = []
        for cam, lights in zip(cams, light_sets):                       # one at a time, synchronised by the stats read
            gpu.lights_upload(lights.view(np.uint8))
            gpu.frame_forward_plus(cam.scene, fp, g.hdr, g.dm, g.ldr)
            ref.append(_download(gpu, g))
        for (hdr, dm, ldr), cam, lights in zip(extra, cams, light_sets):  # all in flight
            gpu.lights_upload(lights.view(np.uint8))
            gpu.frame_forward_plus(cam.scene, fp, hdr, dm, ldr, want_stats=False)
        gpu.sync()
        for k, (hdr, dm, ldr) in enumerate(extra):
            got = (gpu.rt_download(hdr).view(np.uint32), gpu.rt_download(dm, capi.PLANE_DEPTH).view(np.uint32), gpu.rt_download(ldr))
            for a, b, what in zip(got, ref[k], ("hdr", "depth", "ldr")):
                assert np.array_equal(a, b), f"frame {k}: {what} differs between pipelined and one-at-a-time submission"
    finally:
        for rts in extra:
            for rt in rts:
                gpu.rt_destroy(rt)
        g.release()


def test_same_targets_back_to_back(gpu):
    """Asynchronous frames into the SAME render targets: the last one wins, exactly."""
    base = scenes.scene_small(w=200, h=120, lights=24)
    cams = scenes.camera_ring(base, 5, radius=9.0, height=4.0)
    g = harness.GpuScene(gpu, base)
    try:
        fp = capi.FrameParams.from_buffer_copy(base.fp)
        fp.light_culling = 1
        gpu.frame_forward_plus(cams[-1].scene, fp, g.hdr, g.dm, g.ldr)
        want = _download(gpu, g)
        for cam in cams:
            gpu.frame_forward_plus(cam.scene, fp, g.hdr, g.dm, g.ldr, want_stats=False)
        gpu.sync()
        for a, b in zip(_download(gpu, g), want):
            assert np.array_equal(a, b)
    finally:
        g.release()


def test_fused_and_separate_passes_interleaved(gpu):
    """A standalone shsb_light_cull + PassPBRForward between fused frames reads ITS lists, and a fused frame that
    follows does not disturb the lists a still-running separate pass reads."""
    sd = scenes.scene_small(w=256, h=160, lights=40)
    other = scenes.camera_ring(sd, 4, radius=9.0, height=4.0)[1]
    g = harness.GpuScene(gpu, sd)
    try:
        fp = capi.FrameParams.from_buffer_copy(sd.fp)
        fp.light_culling = 1
        gpu.frame_forward_plus(sd.scene, fp, g.hdr, g.dm, g.ldr)
        fused = _download(gpu, g)
        for _ in range(3):
            gpu.frame_forward_plus(other.scene, fp, g.hdr, g.dm, g.ldr, want_stats=False)   # other camera, arena lists
            gpu.light_cull(sd.viewproj, sd.w, sd.h, fp.tile_size, fp.max_lights_per_tile)  # standalone lists
            gpu.frame_forward_plus(other.scene, fp, g.hdr, g.dm, g.ldr, want_stats=False)   # does not invalidate them ...
            gpu.light_cull(sd.viewproj, sd.w, sd.h, fp.tile_size, fp.max_lights_per_tile)  # ... (latest cull is what a pass uses)
            gpu.pass_pbr_forward(sd.scene, fp, g.hdr, g.dm)
            gpu.pass_tonemap(g.hdr, g.ldr, fp.exposure, fp.gamma)
            for a, b in zip(_download(gpu, g), fused):
                assert np.array_equal(a, b)
    finally:
        g.release()


def test_async_download_overlaps_next_frames(gpu):
    """shsb_rt_download_async: the copy of frame k is taken before frame k+1 overwrites the target."""
    import torch
    base = scenes.scene_small(w=320, h=200, lights=16)
    cams = scenes.camera_ring(base, 4, radius=9.0, height=4.0)
    g = harness.GpuScene(gpu, base)
    try:
        fp = capi.FrameParams.from_buffer_copy(base.fp)
        fp.light_culling = 1
        want = []
        for cam in cams:
            gpu.frame_forward_plus(cam.scene, fp, g.hdr, g.dm, g.ldr)
            want.append(gpu.rt_download(g.ldr))
        bufs = [torch.empty(base.w * base.h * 4, dtype=torch.uint8).pin_memory() for _ in cams]
        for cam, buf in zip(cams, bufs):
            gpu.frame_forward_plus(cam.scene, fp, g.hdr, g.dm, g.ldr, want_stats=False)
            gpu.rt_download_async(g.ldr, capi.PLANE_COLOR, buf.data_ptr(), buf.numel())
        gpu.sync()
        for buf, w in zip(bufs, want):
            assert np.array_equal(buf.numpy().reshape(w.shape), w)
    finally:
        g.release()


@pytest.mark.parametrize("name", ["c3", "c4", "c5"])
def test_full_size_configs_properties(gpu, name):
    """BASELINE configs[2..4] at full size: determinism, fused == separate passes, counters consistent with the images."""
    sd = {"c3": scenes.scene_c3, "c4": scenes.scene_c4, "c5": scenes.scene_c5}[name]()
    g = harness.GpuScene(gpu, sd)
    try:
        fp = capi.FrameParams.from_buffer_copy(sd.fp)
        lvp = None
        if name == "c3":
            lvp = gpu.pass_shadow_map(sd.scene, fp, g.shadow)
            sh = gpu.rt_download(g.shadow, capi.PLANE_DEPTH)
            assert 0.0 <= float(sh.min()) < 1.0 and float(sh.max()) == 1.0 and np.count_nonzero(sh < 1.0) > sh.size // 20
            lvp2 = gpu.pass_shadow_map(sd.scene, fp, g.shadow)
            assert np.array_equal(lvp, lvp2) and np.array_equal(gpu.rt_download(g.shadow, capi.PLANE_DEPTH).view(np.uint32), sh.view(np.uint32))

        def render():
            if name == "c3":
                st = gpu.pass_pbr_forward(sd.scene, fp, g.hdr, g.dm, g.shadow, lvp)
                gpu.pass_tonemap(g.hdr, g.ldr, fp.exposure, fp.gamma)
            else:
                st = gpu.frame_forward_plus(sd.scene, fp, g.hdr, g.dm, g.ldr)
            return st.as_dict(), gpu.rt_download(g.dm, capi.PLANE_DEPTH), gpu.rt_download(g.ldr)

        st1, dep1, ldr1 = render()
        st2, dep2, ldr2 = render()
        assert st1 == st2 and np.array_equal(dep1.view(np.uint32), dep2.view(np.uint32)) and np.array_equal(ldr1, ldr2)
        assert st1["tri_input"] == sd.n_triangles and st1["tri_input"] >= st1["tri_after_clip"] * 0 and st1["tri_raster"] > 0
        covered = dep1 < 1.0
        assert int(covered.sum()) == st1["frag_shaded"] <= st1["frag_covered"]
        assert float(dep1.min()) >= 0.0 and float(dep1.max()) == 1.0
        if name != "c3":
            # fused frame == PassPBRForward (+ standalone cull) followed by PassTonemap
            if fp.light_culling:
                gpu.light_cull(sd.viewproj, sd.w, sd.h, fp.tile_size, fp.max_lights_per_tile)
            gpu.pass_pbr_forward(sd.scene, fp, g.hdr, g.dm)
            gpu.pass_tonemap(g.hdr, g.ldr, fp.exposure, fp.gamma)
            assert np.array_equal(gpu.rt_download(g.ldr), ldr1)
            assert np.array_equal(gpu.rt_download(g.dm, capi.PLANE_DEPTH).view(np.uint32), dep1.view(np.uint32))
    finally:
        g.release()


# ---------------------------------------------------------------------------------- render streams (hazards per render target)
def _target_sets(gpu, base, n):
    return [(gpu.rt_create(capi.RT_COLOR_HDR, base.w, base.h), gpu.rt_create(capi.RT_DEPTH_MOTION, base.w, base.h, base.zn, base.zf),
             gpu.rt_create(capi.RT_COLOR_LDR, base.w, base.h)) for _ in range(n)]


@pytest.mark.parametrize("n_streams", [1, 2, 3, 4])
def test_render_streams_frames_into_different_targets(gpu, n_streams):
    """Asynchronous frames into different target sets are spread over n render streams and may overlap on the device;
    every frame equals its one-at-a-time synchronous rendering bit for bit, for every stream count."""
    base = scenes.scene_small(w=300, h=180, lights=40, tex=True)
    cams = scenes.camera_ring(base, 10, radius=9.0, height=4.0)
    g = harness.GpuScene(gpu, base)
    sets = _target_sets(gpu, base, 5)
    try:
        fp = capi.FrameParams.from_buffer_copy(base.fp)
        fp.light_culling = 1
        want = []
        for cam in cams:
            gpu.frame_forward_plus(cam.scene, fp, g.hdr, g.dm, g.ldr)
            want.append(_download(gpu, g))
        gpu.set_tile_streams(n_streams)
        for rnd in range(2):                                           # 10 frames over 5 sets: every set is written twice
            for k in range(5):
                hdr, dm, ldr = sets[k]
                gpu.frame_forward_plus(cams[rnd * 5 + k].scene, fp, hdr, dm, ldr, want_stats=False)
        gpu.sync()
        for k, (hdr, dm, ldr) in enumerate(sets):
            got = (gpu.rt_download(hdr).view(np.uint32), gpu.rt_download(dm, capi.PLANE_DEPTH).view(np.uint32), gpu.rt_download(ldr))
            for a, b, what in zip(got, want[5 + k], ("hdr", "depth", "ldr")):
                assert np.array_equal(a, b), f"{n_streams} streams, set {k}: {what} differs"
    finally:
        gpu.set_tile_streams(2)
        for rts in sets:
            for rt in rts:
                gpu.rt_destroy(rt)
        g.release()


def test_main_stream_operations_see_side_stream_frames_and_vice_versa(gpu):
    """A frame on a side render stream followed by main-stream work on the same targets (tonemap, clear, a synchronous
    frame that keeps the depth) and main-stream work followed by a side-stream frame: results are those of submission order."""
    base = scenes.scene_small(w=260, h=150, lights=0)
    cams = scenes.camera_ring(base, 4, radius=9.0, height=4.0)
    g = harness.GpuScene(gpu, base)
    sets = _target_sets(gpu, base, 4)
    try:
        fp = capi.FrameParams.from_buffer_copy(base.fp)
        fp.light_culling = 0
        want = []
        for cam in cams:                                                # reference: everything synchronous on one target set
            gpu.pass_pbr_forward(cam.scene, fp, g.hdr, g.dm)
            gpu.pass_tonemap(g.hdr, g.ldr, 1.3, 2.0)
            want.append((gpu.rt_download(g.hdr).view(np.uint32), gpu.rt_download(g.ldr)))
        gpu.set_tile_streams(4)
        for (hdr, dm, ldr), cam in zip(sets, cams):                     # side-stream lit pass, then main-stream tonemap of its output
            gpu.pass_pbr_forward(cam.scene, fp, hdr, dm, want_stats=False)
            gpu.pass_tonemap(hdr, ldr, 1.3, 2.0)
        for k, (hdr, dm, ldr) in enumerate(sets):
            assert np.array_equal(gpu.rt_download(ldr), want[k][1]), f"set {k}: tonemap ran ahead of the side-stream frame"
            assert np.array_equal(gpu.rt_download(hdr).view(np.uint32), want[k][0])
        # main-stream clear of the depth plane to 0.5, then a side-stream frame that PRESERVES depth: only nearer fragments land
        half = np.float32(0.5)
        gpu.rt_clear(g.dm, capi.PLANE_DEPTH, half)
        ref = gpu.pass_pbr_forward(cams[0].scene, fp, g.hdr, g.dm, preserve_existing_depth=True).as_dict()
        want_depth = gpu.rt_download(g.dm, capi.PLANE_DEPTH).view(np.uint32)
        want_hdr = gpu.rt_download(g.hdr).view(np.uint32)
        assert ref["frag_covered"] > 0
        for hdr, dm, ldr in sets:
            gpu.pass_pbr_forward(cams[0].scene, fp, hdr, dm)             # same colour history as the reference targets
            gpu.rt_clear(dm, capi.PLANE_DEPTH, half)
            gpu.pass_pbr_forward(cams[0].scene, fp, hdr, dm, preserve_existing_depth=True, want_stats=False)
        for k, (hdr, dm, ldr) in enumerate(sets):
            assert np.array_equal(gpu.rt_download(dm, capi.PLANE_DEPTH).view(np.uint32), want_depth), f"set {k}: the frame overtook the clear"
        # asynchronous read-back of a side-stream frame, then the next frame into the same targets must not overtake the copy
        import torch
        bufs = [torch.empty(base.w * base.h * 4, dtype=torch.uint8).pin_memory() for _ in range(8)]
        fused = []
        for cam in cams:
            gpu.frame_forward_plus(cam.scene, fp, g.hdr, g.dm, g.ldr)
            fused.append(gpu.rt_download(g.ldr))
        for i in range(8):
            hdr, dm, ldr = sets[i % 2]
            gpu.frame_forward_plus(cams[i % 4].scene, fp, hdr, dm, ldr, want_stats=False)
            gpu.rt_download_async(ldr, capi.PLANE_COLOR, bufs[i].data_ptr(), bufs[i].numel())
        gpu.sync()
        for i in range(8):
            assert np.array_equal(bufs[i].numpy().reshape(fused[0].shape), fused[i % 4]), f"read-back {i} torn or stale"
        assert want_hdr is not None
    finally:
        gpu.set_tile_streams(2)
        for rts in sets:
            for rt in rts:
                gpu.rt_destroy(rt)
        g.release()


def test_fence_orders_caller_work_on_the_main_stream(gpu):
    """shsb_fence: an event the caller records on shsb_stream() after the fence covers frames that ran on side streams."""
    import torch
    base = scenes.scene_small(w=320, h=200, lights=32)
    cams = scenes.camera_ring(base, 6, radius=9.0, height=4.0)
    g = harness.GpuScene(gpu, base)
    sets = _target_sets(gpu, base, 3)
    try:
        fp = capi.FrameParams.from_buffer_copy(base.fp)
        fp.light_culling = 1
        want = []
        for cam in cams[:3]:
            gpu.frame_forward_plus(cam.scene, fp, g.hdr, g.dm, g.ldr)
            want.append(gpu.rt_download(g.ldr))
        gpu.set_tile_streams(3)
        stream = torch.cuda.ExternalStream(gpu.stream(), device=0)
        views = []
        for hdr, dm, ldr in sets:
            ptr, nbytes = gpu.rt_device_ptr(ldr, capi.PLANE_COLOR)

            class _Cai:
                __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
            views.append(torch.as_tensor(_Cai(), device="cuda:0"))
        for (hdr, dm, ldr), cam in zip(sets, cams[:3]):
            gpu.frame_forward_plus(cam.scene, fp, hdr, dm, ldr, want_stats=False)
        gpu.fence()
        with torch.cuda.stream(stream):
            copies = [v.clone() for v in views]                        # caller's own work on the main stream
        stream.synchronize()
        for k in range(3):
            assert np.array_equal(copies[k].cpu().numpy().reshape(want[k].shape), want[k]), f"set {k}: the caller's copy ran ahead of the frame"
    finally:
        gpu.set_tile_streams(2)
        for rts in sets:
            for rt in rts:
                gpu.rt_destroy(rt)
        g.release()


def test_asynchronous_overflow_is_reported_not_silent():
    """an asynchronous submission (no statistics) that overflows its per-frame arena used to drop triangles with
    SHSB_OK.  Twelve screen-filling triangles at 4096 x 4096 need 786 k tile-list entries against the 147 k a fresh context
    sizes for 12 source triangles: the frame raises the sticky flag, the next shsb_sync returns SHSB_E_OVERFLOW (10) once and
    grows the capacities, and the re-submitted frame equals the synchronous rendering bit for bit."""
    from leisure_software_renderer_b200.renderer import Context
    ctx = Context(0)
    try:
        n = 12
        pos = np.zeros((n * 3, 3), np.float32)
        for k in range(n):
            z = 0.5 + 0.1 * k
            pos[3 * k:3 * k + 3] = [[-40, -40, z], [40, -40, z], [0, 60, z]]
        nrm = np.tile(np.array([[0, 0, -1]], np.float32), (n * 3, 1))
        uv = np.zeros((n * 3, 2), np.float32)
        idx = np.arange(n * 3, dtype=np.uint32)
        sd = scenes.scene_small(w=4096, h=4096, n_inst=0).with_camera((0.0, 0.0, -2.0), (0.0, 0.0, 10.0))   # looking down +z at the triangles
        mesh = ctx.mesh_upload(pos, nrm, uv, idx)
        item = capi.RenderItem()
        capi.set_f(item.tr.scl, (1, 1, 1)); capi.set_f(item.tr.pos, (0, 0, 6))
        item.mesh, item.visible, item.casts_shadow, item.has_material = mesh, 1, 1, 1
        capi.set_f(item.base_color, (0.7, 0.3, 0.2)); item.metallic, item.roughness, item.ao = 0.1, 0.6, 1.0
        import ctypes as C
        items = (capi.RenderItem * 1)(item)
        sc = capi.Scene.from_buffer_copy(sd.scene)
        sc.items, sc.n_items = C.cast(items, C.POINTER(capi.RenderItem)), 1
        fp = capi.FrameParams.from_buffer_copy(sd.fp)
        fp.light_culling, fp.cull_mode = 0, capi.CULL_NONE
        hdr, dm = ctx.rt_create(capi.RT_COLOR_HDR, 4096, 4096), ctx.rt_create(capi.RT_DEPTH_MOTION, 4096, 4096, sd.zn, sd.zf)
        ctx.pass_pbr_forward(sc, fp, hdr, dm, want_stats=False)          # overflows: no error yet, it is asynchronous
        with pytest.raises(capi.ShsbError, match="status 10"):
            ctx.sync()
        ctx.sync()                                                       # reported once
        ctx.pass_pbr_forward(sc, fp, hdr, dm, want_stats=False)          # capacities were grown: fits now
        ctx.sync()
        got = (ctx.rt_download(hdr).view(np.uint32), ctx.rt_download(dm, capi.PLANE_DEPTH).view(np.uint32))
        st = ctx.pass_pbr_forward(sc, fp, hdr, dm).as_dict()
        assert st["tri_input"] == n and st["frag_shaded"] > 4096 * 4096 // 2
        assert np.array_equal(got[0], ctx.rt_download(hdr).view(np.uint32)) and np.array_equal(got[1], ctx.rt_download(dm, capi.PLANE_DEPTH).view(np.uint32))
    finally:
        ctx.close()


def test_hierarchical_z_reject_changes_no_pixel(monkeypatch):
    """SHSB_HIZ=1: asynchronous frames skip, per 8x4-pixel block, staged triangles whose conservative nearest depth is behind
    everything the block already holds.  A scene with heavy overdraw (a dense grid of Suzannes seen at a grazing angle, plus a frame
    that PRESERVES a pre-pass depth) must come out bit-identical in depth, HDR and LDR with and without it."""
    from leisure_software_renderer_b200.renderer import Context
    base = scenes.scene_c5(w=640, h=360, nx=24, nz=24).with_camera((0, 0.6, -38), (0, 0.4, 40))   # grazing: rows of Suzannes behind each other, overdraw 4
    out = {}
    monkeypatch.setenv("SHSB_NO_FAST_TILE", "1")   # Hi-Z lives in the general instantiation only: compare like with like (see launch_tile_raster)
    for flag in ("0", "1"):
        monkeypatch.setenv("SHSB_HIZ", flag)
        ctx = Context(0)
        try:
            g = harness.GpuScene(ctx, base)
            fp = capi.FrameParams.from_buffer_copy(base.fp)
            fp.light_culling = 1
            ctx.frame_forward_plus(base.scene, fp, g.hdr, g.dm, g.ldr, want_stats=False)
            ctx.sync()
            a = (ctx.rt_download(g.hdr).view(np.uint32), ctx.rt_download(g.dm, capi.PLANE_DEPTH).view(np.uint32), ctx.rt_download(g.ldr))
            fp.light_culling = 0
            ctx.pass_depth_prepass(base.scene, fp, g.dm)
            ctx.pass_pbr_forward(base.scene, fp, g.hdr, g.dm, preserve_existing_depth=True, want_stats=False)   # loads the depth plane: Hi-Z starts from it
            ctx.sync()
            b = (ctx.rt_download(g.hdr).view(np.uint32), ctx.rt_download(g.dm, capi.PLANE_DEPTH).view(np.uint32))
            st = ctx.frame_forward_plus(base.scene, fp, g.hdr, g.dm, g.ldr).as_dict()
            out[flag] = (a, b, st)
            g.release()
        finally:
            ctx.close()
    assert out["0"][2]["frag_covered"] > 1.5 * out["0"][2]["frag_shaded"] > 0, "the scene has no overdraw to reject"
    for x, y, what in zip(out["0"][0] + out["0"][1], out["1"][0] + out["1"][1], ("hdr", "depth", "ldr", "hdr (preserved depth)", "depth (preserved)")):
        assert np.array_equal(x, y), f"Hi-Z changed the {what} plane"


def test_fast_tile_instantiation_agrees_with_the_general_one(monkeypatch):
    """The plain Forward+ frame runs a specialised instantiation of the tile kernel (mode flags as compile-time constants, one per
    shading program); SHSB_NO_FAST_TILE=1 renders the same frame with the general one.  Depth must be equal bit for bit, LDR within
    1 LSB, HDR beyond 100 dB -- the two share their source; only the compiler's contraction of the colour arithmetic may differ."""
    from leisure_software_renderer_b200.renderer import Context
    for shading in (capi.SHADING_PBR, capi.SHADING_BLINN):
        sd = scenes.scene_small(w=320, h=200, shading=shading, n_inst=4, lights=160, seed=5)
        out = {}
        for flag in ("0", "1"):
            monkeypatch.setenv("SHSB_NO_FAST_TILE", flag)
            ctx = Context(0)
            try:
                g = harness.GpuScene(ctx, sd)
                fp = capi.FrameParams.from_buffer_copy(sd.fp)
                fp.light_culling = 1
                ctx.frame_forward_plus(sd.scene, fp, g.hdr, g.dm, g.ldr, want_stats=False)
                ctx.sync()
                out[flag] = (ctx.rt_download(g.hdr), ctx.rt_download(g.dm, capi.PLANE_DEPTH).view(np.uint32), ctx.rt_download(g.ldr))
                g.release()
            finally:
                ctx.close()
        assert np.array_equal(out["0"][1], out["1"][1]), "depth differs between the instantiations"
        assert int(np.abs(out["0"][2].astype(np.int32) - out["1"][2].astype(np.int32)).max()) <= 1
        psnr = harness.psnr(out["0"][0][..., :3], out["1"][0][..., :3])
        print(f"shading {shading}: HDR PSNR fast vs general {psnr:.1f} dB, LDR pixels off by 1: {int(np.count_nonzero((out['0'][2] != out['1'][2]).any(axis=-1)))}")
        assert psnr >= 100.0
