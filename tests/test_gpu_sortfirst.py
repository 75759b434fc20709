"""GPU tests (`-m gpu`) of the sort-first screen partition (ShsbFrameParams::own_row_*): the union of
the partitions' submissions -- one after the other into the same targets here, one per GPU in production -- is bit-identical
to one whole-frame submission in every plane, rows of other owners are never touched, fragment statistics add up."""
import numpy as np
import pytest

import harness
from leisure_software_renderer_b200 import capi, scenes, sortfirst

pytestmark = pytest.mark.gpu

PLANES = ((capi.PLANE_COLOR, "hdr"), (capi.PLANE_TRI_ID, "hdr"), (capi.PLANE_COVERAGE, "hdr"), (capi.PLANE_DEPTH, "dm"), (capi.PLANE_MOTION, "dm"),
          (capi.PLANE_COLOR, "ldr"))


def _snapshot(gpu, g, fused=True):
    """Every plane the submission writes (the LDR plane only exists in the fused frame: PassPBRForward alone does not tonemap)."""
    return [gpu.rt_download(getattr(g, rt), plane) for plane, rt in PLANES if fused or rt != "ldr"]


def _poison(gpu, g, sd):
    gpu.rt_upload(g.hdr, capi.PLANE_COLOR, np.full((sd.h, sd.w, 4), -7.0, np.float32))
    gpu.rt_upload(g.dm, capi.PLANE_DEPTH, np.full((sd.h, sd.w), 0.123, np.float32))
    gpu.rt_upload(g.dm, capi.PLANE_MOTION, np.full((sd.h, sd.w, 2), 5.0, np.float32))
    gpu.rt_upload(g.ldr, capi.PLANE_COLOR, np.full((sd.h, sd.w, 4), 9, np.uint8))


def _frame(gpu, g, sd, fp, fused):
    if fused:
        return gpu.frame_forward_plus(sd.scene, fp, g.hdr, g.dm, g.ldr).as_dict()
    if fp.light_culling:
        gpu.light_cull(sd.viewproj, sd.w, sd.h, fp.tile_size, fp.max_lights_per_tile)
    return gpu.pass_pbr_forward(sd.scene, fp, g.hdr, g.dm).as_dict()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("layout", ["bands3", "interleaved4x1", "interleaved2x3", "bands3_no_motion"])
def test_partitions_reassemble_the_frame_bit_exactly(gpu, layout, fused):
    # with motion vectors every draw's matrix is built (history); without them draws are dropped by the cheap sphere test first
    motion = layout != "bands3_no_motion"
    layout = layout.replace("_no_motion", "")
    sd = scenes.scene_small(w=333, h=207, lights=64, tex=True, motion=motion, sky="procedural", n_inst=6)   # 13 tile rows, ragged last row
    g = harness.GpuScene(gpu, sd)
    try:
        fp = capi.FrameParams.from_buffer_copy(sd.fp)
        fp.write_aovs = 1
        gpu.history_reset()
        whole_stats = _frame(gpu, g, sd, fp, fused)
        whole = _snapshot(gpu, g, fused)
        tile_rows = (sd.h + 15) // 16
        if layout == "bands3":
            parts = []
            for r in range(3):
                y0, y1 = sortfirst.band_of_rank(sd.h, 3, r)
                parts.append((y0 // 16, (y1 - y0 + 15) // 16, 1 << 20))
        elif layout == "interleaved4x1":
            parts = [(r, 1, 4) for r in range(4)]
        else:
            parts = [(r * 3, 3, 6) for r in range(2)]
        _poison(gpu, g, sd)
        frag = {"frag_covered": 0, "frag_shaded": 0}
        owned_rows = np.zeros(tile_rows, dtype=np.int32)
        for i, (first, count, stride) in enumerate(parts):
            before = _snapshot(gpu, g, fused)
            p = capi.FrameParams.from_buffer_copy(fp)
            p.own_row_first, p.own_row_count, p.own_row_stride = first, count, stride
            gpu.history_reset()
            st = _frame(gpu, g, sd, p, fused)
            assert 0 < st["tri_input"] <= whole_stats["tri_input"]   # draws that cannot reach the owned rows are skipped on the host
            for k in frag:
                frag[k] += st[k]
            mine = np.array([ty >= first and (ty - first) % stride < count for ty in range(tile_rows)])
            owned_rows += mine
            # rows this partition does not own are untouched in every plane (framebuffer row y = h-1 - row from the top)
            after = _snapshot(gpu, g, fused)
            top_row_tile = (sd.h - 1 - np.arange(sd.h)) // 16
            foreign = ~mine[top_row_tile]
            for a, b in zip(after, before):
                assert np.array_equal(a[foreign].view(np.uint8), b[foreign].view(np.uint8)), f"partition {i} wrote outside its rows"
        assert np.all(owned_rows == 1), "the layouts cover every tile row exactly once"
        for (plane, rt), a, b in zip([p for p in PLANES if fused or p[1] != "ldr"], _snapshot(gpu, g, fused), whole):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"plane {plane} of {rt} differs from the whole-frame submission"
        assert frag == {k: whole_stats[k] for k in frag}
    finally:
        g.release()


def test_partition_argument_checks(gpu):
    import ctypes as C
    sd = scenes.scene_small(w=64, h=48)
    g = harness.GpuScene(gpu, sd)
    try:
        fp = capi.FrameParams.from_buffer_copy(sd.fp)
        fp.own_row_first, fp.own_row_count, fp.own_row_stride = 0, 2, 1      # stride < count
        st = capi.Stats()
        assert gpu.lib.shsb_pass_pbr_forward(gpu.h, C.byref(sd.scene), C.byref(fp), g.hdr, g.dm, 0, None, 0, C.byref(st)) == 1
        fp.own_row_first, fp.own_row_count, fp.own_row_stride = 100, 1, 1      # owns nothing inside the frame: a valid no-op
        before = gpu.rt_download(g.hdr)
        assert gpu.lib.shsb_pass_pbr_forward(gpu.h, C.byref(sd.scene), C.byref(fp), g.hdr, g.dm, 0, None, 0, C.byref(st)) == 0
        assert np.array_equal(before, gpu.rt_download(g.hdr)) and st.frag_shaded == 0
    finally:
        g.release()


def test_full_size_8k_partition_matches_whole_frame(gpu):
    """BASELINE configs[4] (7680x4320, ~10 M triangles, 1024 lights): rank 3 of 8's interleaved stripes equal the same rows of the whole frame."""
    sd = scenes.scene_c5()
    g = harness.GpuScene(gpu, sd)
    try:
        fp = capi.FrameParams.from_buffer_copy(sd.fp)
        whole_stats = gpu.frame_forward_plus(sd.scene, fp, g.hdr, g.dm, g.ldr).as_dict()
        whole_ldr = gpu.rt_download(g.ldr)
        whole_depth = gpu.rt_download(g.dm, capi.PLANE_DEPTH)
        gpu.rt_upload(g.ldr, capi.PLANE_COLOR, np.zeros((sd.h, sd.w, 4), np.uint8))
        gpu.rt_upload(g.dm, capi.PLANE_DEPTH, np.zeros((sd.h, sd.w), np.float32))
        p = capi.FrameParams.from_buffer_copy(fp)
        p.own_row_first, p.own_row_count, p.own_row_stride = 3 * 2, 2, 8 * 2
        st = gpu.frame_forward_plus(sd.scene, p, g.hdr, g.dm, g.ldr).as_dict()
        ldr = gpu.rt_download(g.ldr)
        depth = gpu.rt_download(g.dm, capi.PLANE_DEPTH)
        top_row_tile = (sd.h - 1 - np.arange(sd.h)) // 16
        mine = (top_row_tile >= 6) & ((top_row_tile - 6) % 16 < 2)
        assert np.array_equal(ldr[mine], whole_ldr[mine]) and np.array_equal(depth[mine].view(np.uint32), whole_depth[mine].view(np.uint32))
        assert not ldr[~mine].any() and not depth[~mine].any()
        assert 0 < st["frag_shaded"] < whole_stats["frag_shaded"] and 0 < st["tri_input"] <= whole_stats["tri_input"]
    finally:
        g.release()
