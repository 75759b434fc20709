"""Sort-first frame assembly in the product (shsb_gather_* / shsb_frame_gather): every "rank" -- here a second
context of this process on the same device, the C++ test (tests/cpp/gather_test.cpp) forks real processes and goes through CUDA
IPC -- pushes its cameras / its screen band into the root's assembly memory with copy-engine copies ordered behind its frames;
the root waits for the commits on the device, reads the assembled slot and releases it.  The assembled bytes must equal the
frames rendered one at a time on one context."""
import numpy as np
import pytest

import harness
from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context

pytestmark = pytest.mark.gpu


@pytest.fixture()
def second(gpu):
    ctx = Context(0)
    yield ctx
    ctx.close()


def _sets(ctx, base, n):
    return [(ctx.rt_create(capi.RT_COLOR_HDR, base.w, base.h), ctx.rt_create(capi.RT_DEPTH_MOTION, base.w, base.h, base.zn, base.zf),
             ctx.rt_create(capi.RT_COLOR_LDR, base.w, base.h)) for _ in range(n)]


def test_camera_batch_assembled_on_the_root(gpu, second):
    """8 cameras, rank r renders cameras c with c % 2 == r into 2 rotating target sets, 3 steps through 2 slots (so the ring wraps and
    rank 1 waits for the root's release on the device)."""
    base = scenes.scene_small(w=256, h=144, lights=32, tex=True)
    fb = base.w * base.h * 4
    ranks = [gpu, second]
    per_rank = [base, scenes.scene_small(w=256, h=144, lights=32, tex=True)]
    scn = [harness.GpuScene(c, sd) for c, sd in zip(ranks, per_rank)]
    rings = [scenes.camera_ring(sd, 8, radius=9.0, height=4.0) for sd in per_rank]
    for g, ring in zip(scn, rings):
        for cam in ring:
            g.remap(cam)                                              # each context's own asset handles
    cams = rings[0]
    sets = [_sets(c, base, 2) for c in ranks]
    fp = capi.FrameParams.from_buffer_copy(base.fp)
    fp.light_culling = 1
    g0 = g1 = None
    try:
        want = []
        for cam in cams:
            gpu.frame_forward_plus(cam.scene, fp, scn[0].hdr, scn[0].dm, scn[0].ldr)
            want.append(gpu.rt_download(scn[0].ldr).reshape(-1))
        g0, exp = gpu.gather_create(2, 2, 8 * fb)
        g1 = second.gather_open(bytes(exp), 1)
        handles = [g0, g1]
        for step in (1, 2, 3):
            order = list(range(8)) if step != 2 else list(range(7, -1, -1))          # camera -> position in the slot changes per step
            for r, ctx in enumerate(ranks):
                k = 0
                for c in range(8):
                    if c % 2 != r:
                        continue
                    hdr, dm, ldr = sets[r][k % 2]; k += 1
                    ctx.frame_forward_plus(rings[r][c].scene, fp, hdr, dm, ldr, want_stats=False)
                    ctx.frame_gather(handles[r], step, ldr, capi.PLANE_COLOR, 0, fb, order[c] * fb)
                ctx.gather_commit(handles[r], step)
            got = gpu.gather_download(g0, step, 0, 8 * fb)          # waits for both commits on the gather stream
            gpu.gather_release(g0, step)
            for c in range(8):
                assert np.array_equal(got[order[c] * fb:(order[c] + 1) * fb], want[c]), f"step {step}: camera {c} differs in the assembled slot"
    finally:
        if g1:
            second.gather_destroy(g1)
        if g0:
            gpu.gather_destroy(g0)
        for ctx, ss in zip(ranks, sets):
            for rts in ss:
                for rt in rts:
                    ctx.rt_destroy(rt)
        for s in scn:
            s.release()


def test_screen_bands_assembled_on_the_root(gpu, second):
    """One frame split into two bands of tile rows (ShsbFrameParams::own_row_*): each rank pushes only the rows it owns (one
    contiguous byte range: rows are contiguous in a row-major target); the assembled LDR frame equals the whole-frame rendering."""
    base = scenes.scene_small(w=320, h=208, lights=40)              # 13 tile rows: 7 + 6
    ranks = [gpu, second]
    per_rank = [base, scenes.scene_small(w=320, h=208, lights=40)]   # GpuScene shifts the asset handles of the SceneData it is given: one each
    scn = [harness.GpuScene(c, sd) for c, sd in zip(ranks, per_rank)]
    g0 = g1 = None
    try:
        fp = capi.FrameParams.from_buffer_copy(base.fp)
        fp.light_culling = 1
        gpu.frame_forward_plus(base.scene, fp, scn[0].hdr, scn[0].dm, scn[0].ldr)
        want = gpu.rt_download(scn[0].ldr).reshape(-1).copy()
        nbytes = base.w * base.h * 4
        g0, exp = gpu.gather_create(2, 1, nbytes)
        g1 = second.gather_open(exp, 1)
        handles = [g0, g1]
        cuts = [0, 7, 13]                                             # tile rows, counted from the TOP of the frame
        for step in (1, 2):
            for r, ctx in enumerate(ranks):
                f = capi.FrameParams.from_buffer_copy(fp)
                f.own_row_first, f.own_row_count, f.own_row_stride = cuts[r], cuts[r + 1] - cuts[r], 13
                ctx.frame_forward_plus(per_rank[r].scene, f, scn[r].hdr, scn[r].dm, scn[r].ldr, want_stats=False)
                # framebuffer rows are bottom-up: tile rows [t0, t1) from the top are pixel rows [H - 16 t1, H - 16 t0)
                y0, y1 = max(0, base.h - 16 * cuts[r + 1]), base.h - 16 * cuts[r]
                off, n = y0 * base.w * 4, (y1 - y0) * base.w * 4
                ctx.frame_gather(handles[r], step, scn[r].ldr, capi.PLANE_COLOR, off, n, off)
                ctx.gather_commit(handles[r], step)
            got = gpu.gather_download(g0, step, 0, nbytes)
            gpu.gather_release(g0, step)
            assert np.array_equal(got, want), f"step {step}: assembled frame differs in {int(np.count_nonzero(got != want))} bytes"
    finally:
        if g1:
            second.gather_destroy(g1)
        if g0:
            gpu.gather_destroy(g0)
        for s in scn:
            s.release()


def test_gather_argument_checks(gpu):
    g, exp = gpu.gather_create(2, 2, 4096)
    try:
        lib = gpu.lib
        rt = gpu.rt_create(capi.RT_COLOR_LDR, 16, 16)
        assert lib.shsb_frame_gather(gpu.h, g, 0, rt, capi.PLANE_COLOR, 0, 1024, 0) == 1          # steps start at 1
        assert lib.shsb_frame_gather(gpu.h, g, 1, rt, capi.PLANE_COLOR, 0, 2048, 0) == 7          # plane is 1024 bytes
        assert lib.shsb_frame_gather(gpu.h, g, 1, rt, capi.PLANE_COLOR, 0, 1024, 4000) == 7       # beyond the slot
        assert lib.shsb_frame_gather(gpu.h, g, 1, rt, capi.PLANE_DEPTH, 0, 4, 0) == 1             # no such plane
        assert lib.shsb_frame_gather(gpu.h, 99, 1, rt, capi.PLANE_COLOR, 0, 4, 0) == 2
        assert lib.shsb_gather_commit(gpu.h, g, 1) == 0 and lib.shsb_gather_commit(gpu.h, g, 1) == 1   # a step commits once
        assert lib.shsb_frame_gather(gpu.h, g, 1, rt, capi.PLANE_COLOR, 0, 4, 0) == 1             # and takes no pushes afterwards
        import ctypes as C
        out = C.c_uint32()
        assert lib.shsb_gather_open(gpu.h, C.byref(exp), 0, C.byref(out)) == 1                    # rank 0 is the root
        assert lib.shsb_gather_open(gpu.h, C.byref(exp), 2, C.byref(out)) == 1                    # 2 ranks: 0 and 1
        gpu.rt_destroy(rt)
    finally:
        gpu.gather_destroy(g)


def test_cpp_two_process_frame_through_the_c_abi_only():
    """tests/cpp/gather_test.cpp: parent = root, forked child = rank 1 (GPU 1 when present, else GPU 0), CUDA IPC handle over a pipe,
    camera batch + split frame assembled through include/shsb.h alone."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "_build", "gather_test")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(root, "tests", "cpp"), "_build/gather_test"], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=240)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout
