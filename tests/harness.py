"""Shared helpers of the parity tests: run one scene through a CPU checker or through the CUDA path."""
from __future__ import annotations

import numpy as np

from leisure_software_renderer_b200 import capi
from oracle.bindings import HostAssets


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units-in-the-last-place between two float32 arrays of non-negative values."""
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    mse = float(np.mean((a - b) ** 2))
    peak = max(float(np.max(np.abs(b))), 1.0)
    return float("inf") if mse == 0.0 else 10.0 * np.log10(peak * peak / mse)


class CpuFrame:
    def __init__(self, hdr, depth, tri_id, coverage, stats, ldr=None, shadow=None, lvp=None, counts=None, indices=None, motion=None):
        self.hdr, self.depth, self.tri_id, self.coverage, self.stats = hdr, depth, tri_id, coverage, stats
        self.ldr, self.shadow, self.lvp, self.counts, self.indices, self.motion = ldr, shadow, lvp, counts, indices, motion


def cpu_forward(o, sd, depth=True, aov=True, shadow=False, forward_plus=False, preserve_depth=False, init_depth=None, tonemap=True,
                prev_models=None, motion=False):
    """PassShadowMap (optional) -> light cull (optional) -> PassPBRForward -> PassTonemap on a CPU checker.
    motion: also return the motion plane; prev_models: Context::history of the previous frame (None = first frame)."""
    A = HostAssets(sd.meshes, sd.textures)
    hdr = np.zeros((sd.h, sd.w, 4), np.float32)
    dep = (np.ones((sd.h, sd.w), np.float32) if init_depth is None else init_depth.copy()) if depth else None
    tri = np.full((sd.h, sd.w), capi.TRI_ID_NONE, np.uint32) if aov and o.kind == "port" else None
    cov = np.zeros((sd.h, sd.w), np.uint32) if aov and o.kind == "port" else None
    sh = lvp = None
    if shadow:
        sh, lvp = o.pass_shadow_map(A, sd.scene, sd.fp, sd.shadow_size, sd.shadow_size)
    mot = np.full((sd.h, sd.w, 2), 7.0, np.float32) if (motion and depth) else None   # the pass must clear it
    tgt = o.make_target(sd.w, sd.h, hdr, dep, shadow=sh, tri_id=tri, coverage=cov, zn=sd.zn, zf=sd.zf, motion=mot)
    counts = indices = None
    if forward_plus:
        counts, indices = o.light_cull(sd.lights, sd.viewproj, sd.w, sd.h, sd.fp.tile_size, sd.fp.max_lights_per_tile)
        st = o.pass_pbr_forward_plus(A, sd.scene, sd.fp, tgt, sd.lights, counts, indices, shadow_lvp=lvp, preserve_depth=preserve_depth)
    else:
        st = o.pass_pbr_forward(A, sd.scene, sd.fp, tgt, shadow_lvp=lvp, preserve_depth=preserve_depth, prev_models=prev_models)
    ldr = o.pass_tonemap(hdr, sd.fp.exposure, sd.fp.gamma) if tonemap else None
    return CpuFrame(hdr, dep, tri, cov, st.as_dict(), ldr, sh, lvp, counts, indices, mot)


class GpuScene:
    """Uploads a SceneData's assets once and keeps the handle mapping (handles are 1-based and sequential,
    like ResourceRegistry, so scene items can be used unchanged on a fresh context)."""

    def __init__(self, ctx, sd):
        self.ctx, self.sd = ctx, sd
        self.mesh_base = None
        handles = [ctx.mesh_upload(m["positions"], m.get("normals"), m.get("uvs"), m.get("indices")) for m in sd.meshes]
        tex = [ctx.texture_upload(t) for t in sd.textures]
        self.mesh_offset = handles[0] - 1 if handles else 0
        self.tex_offset = tex[0] - 1 if tex else 0
        # remap item handles if this context already held other assets
        self._remapped = []
        self.remap(sd)
        self.hdr = ctx.rt_create(capi.RT_COLOR_HDR, sd.w, sd.h)
        self.dm = ctx.rt_create(capi.RT_DEPTH_MOTION, sd.w, sd.h, sd.zn, sd.zf)
        self.ldr = ctx.rt_create(capi.RT_COLOR_LDR, sd.w, sd.h)
        self.shadow = ctx.rt_create(capi.RT_SHADOW, sd.shadow_size, sd.shadow_size) if sd.shadow_size else 0
        if sd.lights is not None:
            ctx.lights_upload(sd.lights.view(np.uint8))

    def remap(self, sd, sign=1):
        """Shifts the asset handles of a SceneData that uses this scene's meshes / textures (undone by release())."""
        if self.mesh_offset or self.tex_offset:
            for i in range(sd.scene.n_items):
                it = sd.scene.items[i]
                it.mesh += sign * self.mesh_offset
                if it.base_color_tex:
                    it.base_color_tex += sign * self.tex_offset
            for k in range(6):
                if sd.scene.sky_faces[k]:
                    sd.scene.sky_faces[k] += sign * self.tex_offset
        if sign > 0:
            self._remapped.append(sd)

    def release(self):
        # undo the handle remaps so the SceneData objects can be reused
        for sd in self._remapped:
            self.remap(sd, -1)
        self._remapped = []
        for rt in (self.hdr, self.dm, self.ldr, self.shadow):
            if rt:
                self.ctx.rt_destroy(rt)


def gpu_forward(ctx, sd, depth=True, aov=True, shadow=False, forward_plus=False, preserve_depth=False, init_depth=None, fused=False,
                prev_scene=None, motion=False):
    """prev_scene: a SceneData rendered first into the same targets (it becomes the context's history); without it the
    history is reset so that the frame is a first frame."""
    g = GpuScene(ctx, sd)
    try:
        ctx.history_reset()
        if prev_scene is not None:
            g.remap(prev_scene)
            pfp = capi.FrameParams.from_buffer_copy(prev_scene.fp)
            pfp.light_culling = 0
            ctx.pass_pbr_forward(prev_scene.scene, pfp, g.hdr, g.dm if depth else 0)
        if motion and depth:
            ctx.rt_upload(g.dm, capi.PLANE_MOTION, np.full((sd.h, sd.w, 2), 7.0, np.float32))  # the pass must clear it
        fp = capi.FrameParams.from_buffer_copy(sd.fp)
        fp.write_aovs = 1 if aov else 0
        fp.light_culling = 1 if forward_plus else 0
        lvp = None
        sh_img = None
        if shadow:
            lvp = ctx.pass_shadow_map(sd.scene, fp, g.shadow)
            sh_img = ctx.rt_download(g.shadow, capi.PLANE_DEPTH)
        if init_depth is not None:
            ctx.rt_upload(g.dm, capi.PLANE_DEPTH, init_depth)
        counts = indices = None
        if fused:
            st = ctx.frame_forward_plus(sd.scene, fp, g.hdr, g.dm if depth else 0, g.ldr)
            if forward_plus:
                counts, indices = ctx.light_lists_download()
        else:
            if forward_plus:
                ctx.light_cull(sd.viewproj, sd.w, sd.h, fp.tile_size, fp.max_lights_per_tile)
                counts, indices = ctx.light_lists_download()
            st = ctx.pass_pbr_forward(sd.scene, fp, g.hdr, g.dm if depth else 0, g.shadow if shadow else 0, lvp, preserve_depth)
            ctx.pass_tonemap(g.hdr, g.ldr, fp.exposure, fp.gamma)
        hdr = ctx.rt_download(g.hdr)
        dep = ctx.rt_download(g.dm, capi.PLANE_DEPTH) if depth else None
        ldr = ctx.rt_download(g.ldr)
        tri = ctx.rt_download(g.hdr, capi.PLANE_TRI_ID) if aov else None
        cov = ctx.rt_download(g.hdr, capi.PLANE_COVERAGE) if aov else None
        mot = ctx.rt_download(g.dm, capi.PLANE_MOTION) if (motion and depth) else None
        return CpuFrame(hdr, dep, tri, cov, st.as_dict(), ldr, sh_img, lvp, counts, indices, mot)
    finally:
        g.release()


def assert_frame_parity(gpu_f: CpuFrame, cpu_f: CpuFrame, depth=True, name=""):
    """The north_star gates: coverage masks, triangle IDs and light lists bit-exact; depth <= 1 ULP;
    colour <= 1 LSB per 8-bit channel; PSNR >= 60 dB on the float target."""
    for k in ("tri_input", "tri_after_clip", "tri_raster", "frag_covered"):
        assert gpu_f.stats[k] == cpu_f.stats[k], f"{name}: stats[{k}] gpu {gpu_f.stats[k]} != cpu {cpu_f.stats[k]}"
    if cpu_f.tri_id is not None and gpu_f.tri_id is not None:
        assert np.array_equal(gpu_f.tri_id != capi.TRI_ID_NONE, cpu_f.tri_id != capi.TRI_ID_NONE), f"{name}: coverage mask differs"
        nbad = int(np.count_nonzero(gpu_f.tri_id != cpu_f.tri_id))
        assert nbad == 0, f"{name}: {nbad} pixels have a different winning triangle id"
        assert np.array_equal(gpu_f.coverage, cpu_f.coverage), f"{name}: per-pixel fragment counts differ"
    if depth:
        u = ulp_diff(gpu_f.depth, cpu_f.depth)
        assert int(u.max()) <= 1, f"{name}: depth differs by {int(u.max())} ULP"
    if cpu_f.counts is not None:
        assert np.array_equal(gpu_f.counts, cpu_f.counts), f"{name}: tile light counts differ"
        m = gpu_f.indices.shape[1]
        valid = np.arange(m)[None, :] < np.minimum(cpu_f.counts, m)[:, None]
        assert np.array_equal(gpu_f.indices[valid], cpu_f.indices[valid]), f"{name}: tile light lists differ"
    if cpu_f.motion is not None and gpu_f.motion is not None:
        # exact arithmetic end to end: the velocity of the winning fragment, (0, 0) elsewhere
        assert np.array_equal(gpu_f.motion.view(np.uint32), cpu_f.motion.view(np.uint32)), \
            f"{name}: motion vectors differ at {int(np.count_nonzero(gpu_f.motion != cpu_f.motion))} components"
    p = psnr(gpu_f.hdr[..., :3], cpu_f.hdr[..., :3])
    assert p >= 60.0, f"{name}: HDR PSNR {p:.1f} dB < 60 dB"
    d = np.abs(gpu_f.ldr.astype(np.int32) - cpu_f.ldr.astype(np.int32))
    assert int(d.max()) <= 1, f"{name}: LDR differs by {int(d.max())} LSB at {np.argwhere(d == d.max())[:3]}"
    return {"psnr_hdr": p, "ldr_max_lsb": int(d.max()), "ldr_pixels_off_by_1": int(np.count_nonzero(d.max(axis=2)))}
