"""Thin, numpy-facing wrapper over the C-ABI (include/shsb.h).

Names follow the reference's pass / render-target vocabulary (RT_ColorHDR, PassPBRForward ...).
Everything here forwards to libshsb.so; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .capi import (FrameParams, RasterCfg, RenderItem, Scene, Stats, Transform, Uniforms)


def _check(lib, ctx, rc, what):
    if rc != capi.OK:
        msg = lib.shsb_last_error_string(ctx) if ctx else b""
        raise capi.ShsbError(f"{what} failed: status {rc}: {(msg or b'').decode(errors='replace')}")


class Context:
    """shsb_ctx: one CUDA device, one stream; replaces Context + ThreadPoolJobSystem of the reference."""

    def __init__(self, device: int = 0):
        self.lib = capi.load_library()
        self.h = C.c_void_p()
        rc = self.lib.shsb_context_create(device, C.byref(self.h))
        if rc != capi.OK:
            raise capi.ShsbError(f"shsb_context_create(device={device}) failed with status {rc} "
                                 "(3 = no usable CUDA device; there is no CPU fallback)")
        self._rt_shape = {}

    def close(self):
        if self.h:
            self.lib.shsb_context_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- resources
    def mesh_upload(self, positions, normals=None, uvs=None, indices=None) -> int:
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals if normals is not None else np.zeros((0, 3)), dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uvs if uvs is not None else np.zeros((0, 2)), dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(indices if indices is not None else np.zeros((0,)), dtype=np.uint32).reshape(-1)
        out = C.c_uint32()
        rc = self.lib.shsb_mesh_upload(self.h, capi.fptr(pos), len(pos), capi.fptr(nrm), len(nrm), capi.fptr(uv), len(uv),
                                       capi.u32ptr(idx), len(idx), C.byref(out))
        _check(self.lib, self.h, rc, "shsb_mesh_upload")
        return out.value

    def texture_upload(self, rgba) -> int:
        t = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = t.shape[0], t.shape[1]
        out = C.c_uint32()
        rc = self.lib.shsb_texture_upload(self.h, t.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, C.byref(out))
        _check(self.lib, self.h, rc, "shsb_texture_upload")
        return out.value

    # ---- render targets
    def rt_create(self, kind, w, h, zn=0.1, zf=1000.0) -> int:
        out = C.c_uint32()
        rc = self.lib.shsb_rt_create(self.h, kind, w, h, zn, zf, C.byref(out))
        _check(self.lib, self.h, rc, "shsb_rt_create")
        self._rt_shape[out.value] = (kind, w, h)
        return out.value

    def rt_destroy(self, rt):
        _check(self.lib, self.h, self.lib.shsb_rt_destroy(self.h, rt), "shsb_rt_destroy")
        self._rt_shape.pop(rt, None)

    def _plane_array(self, rt, plane):
        kind, w, h = self._rt_shape[rt]
        if plane == capi.PLANE_COLOR:
            return np.empty((h, w, 4), dtype=np.float32 if kind == capi.RT_COLOR_HDR else np.uint8)
        if plane == capi.PLANE_DEPTH:
            return np.empty((h, w), dtype=np.float32)
        if plane == capi.PLANE_MOTION:
            return np.empty((h, w, 2), dtype=np.float32)
        return np.empty((h, w), dtype=np.uint32)

    def rt_download(self, rt, plane=capi.PLANE_COLOR) -> np.ndarray:
        a = self._plane_array(rt, plane)
        rc = self.lib.shsb_rt_download(self.h, rt, plane, a.ctypes.data_as(C.c_void_p), a.nbytes)
        _check(self.lib, self.h, rc, "shsb_rt_download")
        return a

    def rt_download_into(self, rt, plane, dst_ptr, nbytes):
        _check(self.lib, self.h, self.lib.shsb_rt_download(self.h, rt, plane, C.c_void_p(dst_ptr), nbytes), "shsb_rt_download")

    def rt_download_async(self, rt, plane, dst_pinned_ptr, nbytes):
        """D2H into pinned memory on the copy stream; overlaps with later submissions, valid after sync()."""
        _check(self.lib, self.h, self.lib.shsb_rt_download_async(self.h, rt, plane, C.c_void_p(dst_pinned_ptr), nbytes), "shsb_rt_download_async")

    def rt_upload(self, rt, plane, array):
        a = np.ascontiguousarray(array)
        rc = self.lib.shsb_rt_upload(self.h, rt, plane, a.ctypes.data_as(C.c_void_p), a.nbytes)
        _check(self.lib, self.h, rc, "shsb_rt_upload")

    def rt_clear(self, rt, plane, value):
        v = np.ascontiguousarray(value)
        _check(self.lib, self.h, self.lib.shsb_rt_clear(self.h, rt, plane, v.ctypes.data_as(C.c_void_p)), "shsb_rt_clear")

    def rt_device_ptr(self, rt, plane):
        p = C.c_void_p()
        n = C.c_size_t()
        _check(self.lib, self.h, self.lib.shsb_rt_device_ptr(self.h, rt, plane, C.byref(p), C.byref(n)), "shsb_rt_device_ptr")
        return p.value, n.value

    # ---- on-disk fixtures (Wavefront OBJ / PNG)
    def mesh_load_obj(self, path: str) -> int:
        out = C.c_uint32()
        _check(self.lib, self.h, self.lib.shsb_mesh_load_obj(self.h, os.fsencode(path), C.byref(out)), "shsb_mesh_load_obj")
        return out.value

    def texture_load_png(self, path: str, flip_y=True) -> int:
        out = C.c_uint32()
        _check(self.lib, self.h, self.lib.shsb_texture_load_png(self.h, os.fsencode(path), int(bool(flip_y)), C.byref(out)), "shsb_texture_load_png")
        return out.value

    def mesh_download(self, mesh) -> dict:
        n = np.zeros(4, np.uint32)
        _check(self.lib, self.h, self.lib.shsb_mesh_info(self.h, mesh, capi.u32ptr(n)), "shsb_mesh_info")
        pos, nrm, uv, idx = np.zeros((n[0], 3), np.float32), np.zeros((n[1], 3), np.float32), np.zeros((n[2], 2), np.float32), np.zeros(n[3], np.uint32)
        _check(self.lib, self.h, self.lib.shsb_mesh_download(self.h, mesh, capi.fptr(pos), capi.fptr(nrm), capi.fptr(uv), capi.u32ptr(idx)), "shsb_mesh_download")
        return {"positions": pos, "normals": nrm, "uvs": uv, "indices": idx}

    def texture_download(self, tex) -> np.ndarray:
        wh = np.zeros(2, np.int32)
        _check(self.lib, self.h, self.lib.shsb_texture_info(self.h, tex, wh.ctypes.data_as(C.POINTER(C.c_int32))), "shsb_texture_info")
        out = np.zeros((int(wh[1]), int(wh[0]), 4), np.uint8)
        _check(self.lib, self.h, self.lib.shsb_texture_download(self.h, tex, out.ctypes.data_as(C.POINTER(C.c_uint8)), out.size), "shsb_texture_download")
        return out

    # ---- passes
    def rasterize_mesh(self, mesh, shader_id, uniforms: Uniforms, hdr_rt, depth_rt=0, cull_mode=capi.CULL_BACK,
                       front_face_ccw=True, write_aovs=False) -> Stats:
        cfg = RasterCfg(cull_mode, int(front_face_ccw), int(write_aovs), 0)
        st = Stats()
        rc = self.lib.shsb_rasterize_mesh(self.h, mesh, shader_id, C.byref(uniforms), hdr_rt, depth_rt, C.byref(cfg), C.byref(st))
        _check(self.lib, self.h, rc, "shsb_rasterize_mesh")
        return st

    def pass_pbr_forward(self, scene: Scene, fp: FrameParams, hdr_rt, depth_rt=0, shadow_rt=0, shadow_light_viewproj=None,
                         preserve_existing_depth=False, want_stats=True) -> Stats:
        """want_stats=False: asynchronous submission (no host synchronisation, may run on a side render stream)."""
        st = Stats()
        lvp = None
        if shadow_light_viewproj is not None:
            lvp = np.ascontiguousarray(shadow_light_viewproj, dtype=np.float32).reshape(16)
        rc = self.lib.shsb_pass_pbr_forward(self.h, C.byref(scene), C.byref(fp), hdr_rt, depth_rt, shadow_rt,
                                            capi.fptr(lvp) if lvp is not None else None, int(preserve_existing_depth), C.byref(st) if want_stats else None)
        _check(self.lib, self.h, rc, "shsb_pass_pbr_forward")
        return st

    def pass_depth_prepass(self, scene: Scene, fp: FrameParams, depth_rt) -> Stats:
        st = Stats()
        _check(self.lib, self.h, self.lib.shsb_pass_depth_prepass(self.h, C.byref(scene), C.byref(fp), depth_rt, C.byref(st)),
               "shsb_pass_depth_prepass")
        return st

    def pass_shadow_map(self, scene: Scene, fp: FrameParams, shadow_rt) -> np.ndarray:
        lvp = np.zeros(16, dtype=np.float32)
        _check(self.lib, self.h, self.lib.shsb_pass_shadow_map(self.h, C.byref(scene), C.byref(fp), shadow_rt, capi.fptr(lvp)),
               "shsb_pass_shadow_map")
        return lvp

    def history_reset(self):
        """Context::history.reset(): the next lit pass behaves like a first frame (zero object motion)."""
        _check(self.lib, self.h, self.lib.shsb_history_reset(self.h), "shsb_history_reset")

    def pass_tonemap(self, hdr_rt, ldr_rt, exposure=1.0, gamma=2.2):
        _check(self.lib, self.h, self.lib.shsb_pass_tonemap(self.h, hdr_rt, ldr_rt, exposure, gamma), "shsb_pass_tonemap")

    def pass_motion_blur(self, params: "capi.MotionBlurParams", input_ldr, output_ldr, depth_motion_rt):
        rc = self.lib.shsb_pass_motion_blur(self.h, C.byref(params), input_ldr, output_ldr, depth_motion_rt)
        _check(self.lib, self.h, rc, "shsb_pass_motion_blur")

    def pass_light_shafts(self, params: "capi.LightShaftsParams", input_ldr, output_ldr, depth_like_rt=0):
        rc = self.lib.shsb_pass_light_shafts(self.h, C.byref(params), input_ldr, output_ldr, depth_like_rt)
        _check(self.lib, self.h, rc, "shsb_pass_light_shafts")

    def pass_taa(self, ldr_rt):
        _check(self.lib, self.h, self.lib.shsb_pass_taa(self.h, ldr_rt), "shsb_pass_taa")

    def taa_reset(self):
        _check(self.lib, self.h, self.lib.shsb_taa_reset(self.h), "shsb_taa_reset")

    # ---- legacy tile-job variant (BASELINE configs[0] as shipped)
    def legacy_draw_blinn_phong(self, mesh, uniforms: "capi.LegacyUniforms", canvas_ldr, zbuffer):
        """One object of the legacy demo's RendererSystem::process on the device (canvas in shs::Canvas order, z in ZBuffer order)."""
        _check(self.lib, self.h, self.lib.shsb_legacy_draw_blinn_phong(self.h, mesh, C.byref(uniforms), canvas_ldr, zbuffer),
               "shsb_legacy_draw_blinn_phong")

    def legacy2_shadow_draw(self, mesh, model, light_vp, shadow_map, job_tile_w=0, job_tile_h=0):
        """One object of the legacy render-target demos' shadow pass (draw_triangle_tile_shadow over every job tile)."""
        m, lvp = (np.ascontiguousarray(a, dtype=np.float32).reshape(16) for a in (model, light_vp))
        _check(self.lib, self.h, self.lib.shsb_legacy2_shadow_draw(self.h, mesh, capi.fptr(m), capi.fptr(lvp), job_tile_w, job_tile_h, shadow_map),
               "shsb_legacy2_shadow_draw")

    def legacy2_draw_softshadow(self, mesh, uniforms: "capi.Legacy2Uniforms", shadow_map, canvas_ldr, zbuffer):
        """One object of the soft-shadow demo's lit pass (canvas and z-buffer in shs::Canvas order)."""
        _check(self.lib, self.h, self.lib.shsb_legacy2_draw_softshadow(self.h, mesh, C.byref(uniforms), shadow_map, canvas_ldr, zbuffer),
               "shsb_legacy2_draw_softshadow")

    def legacy3_ibl_upload(self, irradiance, prefiltered) -> int:
        """irradiance: (6, n, n, 3) float32; prefiltered: list of (6, n_m, n_m, 3) float32 mips."""
        irr = np.ascontiguousarray(irradiance, dtype=np.float32)
        sizes = np.array([m.shape[1] for m in prefiltered], np.int32)
        pre = np.concatenate([np.ascontiguousarray(m, dtype=np.float32).reshape(-1) for m in prefiltered])
        out = C.c_uint32()
        _check(self.lib, self.h, self.lib.shsb_legacy3_ibl_upload(self.h, capi.fptr(irr), irr.shape[1], capi.fptr(pre), sizes.ctypes.data_as(C.POINTER(C.c_int32)),
                                                                 len(sizes), C.byref(out)), "shsb_legacy3_ibl_upload")
        return out.value

    def legacy3_ibl_destroy(self, ibl):
        _check(self.lib, self.h, self.lib.shsb_legacy3_ibl_destroy(self.h, ibl), "shsb_legacy3_ibl_destroy")

    def legacy3_draw_pbr(self, mesh, uniforms: "capi.Legacy2Uniforms", shadow_map, ibl, canvas_ldr, depth_motion):
        """One object of the PBR / IBL demo's lit pass (canvas, depth and velocity in shs::Canvas order)."""
        _check(self.lib, self.h, self.lib.shsb_legacy3_draw_pbr(self.h, mesh, C.byref(uniforms), shadow_map, ibl, canvas_ldr, depth_motion),
               "shsb_legacy3_draw_pbr")

    def cull_objects_frustum(self, bounds10, view_proj):
        """cull_vs_frustum over objects given as (n, 10) sphere + AABB bounds: classes (n,) uint8, visible indices, counts
        (tested, outside, intersecting, inside, visible)."""
        b = np.ascontiguousarray(bounds10, dtype=np.float32).reshape(-1, 10)
        vp = np.ascontiguousarray(view_proj, dtype=np.float32).reshape(16)
        classes, visible, counts = np.zeros(len(b), np.uint8), np.zeros(max(1, len(b)), np.uint32), np.zeros(5, np.uint32)
        _check(self.lib, self.h, self.lib.shsb_cull_objects_frustum(self.h, capi.fptr(b), len(b), capi.fptr(vp), classes.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                                   capi.u32ptr(visible), capi.u32ptr(counts)), "shsb_cull_objects_frustum")
        return classes, visible[:int(counts[4])].copy(), counts

    def software_occlusion(self, object_aabbs, frustum_visible, object_mesh, models, mesh_table, vertices, indices, view, view_proj, occ_w, occ_h, eps=1e-4, enable=True):
        """run_software_occlusion_pass on the device: occluded flags (n,), the ordered visible list, counts (scene, frustum-visible,
        visible, occluded) and the occlusion depth buffer (occ_h, occ_w)."""
        a = np.ascontiguousarray(object_aabbs, dtype=np.float32).reshape(-1, 6)
        vis = np.ascontiguousarray(frustum_visible, dtype=np.uint32).reshape(-1)
        om = np.ascontiguousarray(object_mesh, dtype=np.uint32).reshape(-1)
        mo = np.ascontiguousarray(models, dtype=np.float32).reshape(-1, 16)
        mt = np.ascontiguousarray(mesh_table, dtype=np.uint32).reshape(-1, 3)
        vt = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        ix = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        v, vp = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (view, view_proj))
        occ, out_vis, counts = np.zeros(max(1, len(a)), np.uint8), np.zeros(max(1, len(vis)), np.uint32), np.zeros(4, np.uint32)
        depth = np.zeros((int(occ_h), int(occ_w)), np.float32)
        rc = self.lib.shsb_software_occlusion(self.h, capi.fptr(a), len(a), capi.u32ptr(vis), len(vis), capi.u32ptr(om), capi.fptr(mo), capi.u32ptr(mt), len(mt), capi.fptr(vt), len(vt),
                                              capi.u32ptr(ix), len(ix), capi.fptr(v), capi.fptr(vp), int(occ_w), int(occ_h), float(eps), int(bool(enable)),
                                              occ.ctypes.data_as(C.POINTER(C.c_uint8)), capi.u32ptr(out_vis), capi.u32ptr(counts), capi.fptr(depth))
        _check(self.lib, self.h, rc, "shsb_software_occlusion")
        return occ[:len(a)], out_vis[:int(counts[2])].copy(), counts, depth

    @staticmethod
    def _flat_draws(draws):
        """draws: iterable of dicts mesh, model (16,), base_color (3,), optional selection (<= 8 light indices)."""
        arr = (capi.FlatDraw * max(1, len(draws)))()
        for i, d in enumerate(draws):
            arr[i].mesh = int(d["mesh"])
            capi.set_f(arr[i].model, d["model"])
            capi.set_f(arr[i].base_color, d["base_color"])
            sel = [int(v) for v in d.get("selection", ())]
            arr[i].selection_count = int(d.get("selection_count", len(sel)))
            for k, v in enumerate(sel[:8]):
                arr[i].selection[k] = v
        return arr

    def flat_draw_blinn_phong(self, draws, view_proj, camera_pos, light_dir_ws, canvas_ldr, depth):
        """debug_draw::draw_mesh_blinn_phong_transformed for a batch of draws, in order, into (canvas_ldr, depth plane of `depth`)."""
        arr = self._flat_draws(draws)
        vp, cam, ld = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1) for a in (view_proj, camera_pos, light_dir_ws))
        _check(self.lib, self.h, self.lib.shsb_flat_draw_blinn_phong(self.h, arr, len(draws), capi.fptr(vp), capi.fptr(cam), capi.fptr(ld), canvas_ldr, depth),
               "shsb_flat_draw_blinn_phong")

    def flat_draw_multi_light(self, draws, view_proj, camera_pos, lights, canvas_ldr, depth):
        """draw_mesh_multi_light_transformed (the consumer of the per-object LightSelections) for a batch of draws, in order.
        lights: array of capi.LIGHT_PROPS_DTYPE."""
        arr = self._flat_draws(draws)
        vp, cam = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1) for a in (view_proj, camera_pos))
        li = np.ascontiguousarray(lights, dtype=capi.LIGHT_PROPS_DTYPE).reshape(-1)
        _check(self.lib, self.h, self.lib.shsb_flat_draw_multi_light(self.h, arr, len(draws), capi.fptr(vp), capi.fptr(cam), li.ctypes.data_as(C.c_void_p), len(li),
                                                                    canvas_ldr, depth), "shsb_flat_draw_multi_light")

    def collect_object_lights(self, object_aabbs, visible, records, cull_mode):
        """collect_object_lights per object: counts (n,), light indices (n, 8), squared distances (n, 8)."""
        a = np.ascontiguousarray(object_aabbs, dtype=np.float32).reshape(-1, 6)
        v = np.ascontiguousarray(visible, dtype=np.uint32).reshape(-1)
        r = np.ascontiguousarray(records).view(np.uint8).reshape(-1, capi.LIGHT_RECORD_BYTES)
        counts, idx, d2 = np.zeros(len(a), np.uint32), np.zeros((len(a), 8), np.uint32), np.zeros((len(a), 8), np.float32)
        _check(self.lib, self.h, self.lib.shsb_collect_object_lights(self.h, capi.fptr(a), len(a), capi.u32ptr(v), len(v), r.ctypes.data_as(C.c_void_p), len(r), int(cull_mode),
                                                                    capi.u32ptr(counts), capi.u32ptr(idx), capi.fptr(d2)), "shsb_collect_object_lights")
        return counts, idx, d2

    def tile_depth_range_from_scene(self, object_aabbs, visible_objects, view, view_proj, w, h, tile_size, z_near, z_far):
        """build_tile_view_depth_range_from_scene on the device; the ranges stay there for light_cull_ex and are returned as two arrays."""
        a = np.ascontiguousarray(object_aabbs, dtype=np.float32).reshape(-1, 6)
        v = np.ascontiguousarray(visible_objects, dtype=np.uint32).reshape(-1)
        mv, mvp = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (view, view_proj))
        _check(self.lib, self.h, self.lib.shsb_tile_depth_range_from_scene(self.h, capi.fptr(a), len(a), capi.u32ptr(v), len(v), capi.fptr(mv), capi.fptr(mvp), int(w), int(h),
                                                                          int(tile_size), float(z_near), float(z_far)), "shsb_tile_depth_range_from_scene")
        n = ((int(w) + int(tile_size) - 1) // int(tile_size)) * ((int(h) + int(tile_size) - 1) // int(tile_size))
        lo, hi = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        _check(self.lib, self.h, self.lib.shsb_tile_depth_range_download(self.h, capi.fptr(lo), capi.fptr(hi), n), "shsb_tile_depth_range_download")
        return lo, hi

    def select_object_lights_from_bins(self, object_aabbs, view, view_proj, clustered, z_near, z_far, records, cull_mode):
        """gather_light_scene_candidates_for_aabb + collect_object_lights per object over the context's last light bins:
        counts (n,), light indices (n, 8), squared distances (n, 8), candidate counts (n,)."""
        a = np.ascontiguousarray(object_aabbs, dtype=np.float32).reshape(-1, 6)
        mv, mvp = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (view, view_proj))
        r = np.ascontiguousarray(records).view(np.uint8).reshape(-1, capi.LIGHT_RECORD_BYTES)
        counts, idx, d2, cand = np.zeros(len(a), np.uint32), np.zeros((len(a), 8), np.uint32), np.zeros((len(a), 8), np.float32), np.zeros(len(a), np.uint32)
        _check(self.lib, self.h, self.lib.shsb_select_object_lights_from_bins(self.h, capi.fptr(a), len(a), capi.fptr(mv), capi.fptr(mvp), int(bool(clustered)), float(z_near),
                                                                             float(z_far), r.ctypes.data_as(C.c_void_p), len(r), int(cull_mode), capi.u32ptr(counts),
                                                                             capi.u32ptr(idx), capi.fptr(d2), capi.u32ptr(cand)), "shsb_select_object_lights_from_bins")
        return counts, idx, d2, cand

    def lights_upload(self, records: np.ndarray):
        r = np.ascontiguousarray(records).view(np.uint8).reshape(-1, capi.LIGHT_RECORD_BYTES)
        _check(self.lib, self.h, self.lib.shsb_lights_upload(self.h, r.ctypes.data_as(C.c_void_p), len(r)), "shsb_lights_upload")

    def light_cull(self, view_proj, w, h, tile_size=16, max_per_tile=128):
        vp = np.ascontiguousarray(view_proj, dtype=np.float32).reshape(16)
        _check(self.lib, self.h, self.lib.shsb_light_cull(self.h, capi.fptr(vp), w, h, tile_size, max_per_tile), "shsb_light_cull")
        self._tiles = ((w + tile_size - 1) // tile_size) * ((h + tile_size - 1) // tile_size)
        self._max_per_tile = max_per_tile

    def light_cull_ex(self, desc: "capi.LightCullDesc", range_min=None, range_max=None):
        """Depth-range / clustered bin builders; returns (counts, indices[bins, max_per_bin])."""
        lo = np.ascontiguousarray(range_min, dtype=np.float32).reshape(-1) if range_min is not None else None
        hi = np.ascontiguousarray(range_max, dtype=np.float32).reshape(-1) if range_max is not None else None
        if lo is not None:
            assert lo.size == desc.tiles() and hi is not None and hi.size == desc.tiles()
        rc = self.lib.shsb_light_cull_ex(self.h, C.byref(desc), capi.fptr(lo) if lo is not None else None, capi.fptr(hi) if hi is not None else None)
        _check(self.lib, self.h, rc, "shsb_light_cull_ex")
        bins, mx = desc.bins(), desc.max_per_bin
        if desc.mode == capi.LIGHT_CULL_CLUSTERED:
            counts = np.zeros(bins, dtype=np.uint32)
            indices = np.zeros(bins * mx, dtype=np.uint32)
            rc = self.lib.shsb_cluster_lists_download(self.h, capi.u32ptr(counts), counts.size, capi.u32ptr(indices), indices.size)
            _check(self.lib, self.h, rc, "shsb_cluster_lists_download")
            return counts, indices.reshape(bins, mx)
        self._tiles, self._max_per_tile = bins, mx
        return self.light_lists_download()

    def tile_depth_range(self, depth_motion_rt, tile_size=16):
        """Per-tile [min, max] view depth of the z-buffer (kept on the device for light_cull_ex); returns both arrays."""
        _check(self.lib, self.h, self.lib.shsb_tile_depth_range(self.h, depth_motion_rt, tile_size), "shsb_tile_depth_range")
        _, w, h = self._rt_shape[depth_motion_rt]
        n = ((w + tile_size - 1) // tile_size) * ((h + tile_size - 1) // tile_size)
        lo, hi = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        _check(self.lib, self.h, self.lib.shsb_tile_depth_range_download(self.h, capi.fptr(lo), capi.fptr(hi), n), "shsb_tile_depth_range_download")
        return lo, hi

    def tile_depth_range_ndc01(self, depth_rt, tile_size, z_near, z_far):
        """fp_stress_depth_reduce.comp on a depth plane that holds projective depth in [0, 1]; returns (min, max) view depth per tile."""
        _check(self.lib, self.h, self.lib.shsb_tile_depth_range_ndc01(self.h, depth_rt, tile_size, float(z_near), float(z_far)), "shsb_tile_depth_range_ndc01")
        _, w, h = self._rt_shape[depth_rt]
        n = ((w + tile_size - 1) // tile_size) * ((h + tile_size - 1) // tile_size)
        lo, hi = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        _check(self.lib, self.h, self.lib.shsb_tile_depth_range_download(self.h, capi.fptr(lo), capi.fptr(hi), n), "shsb_tile_depth_range_download")
        return lo, hi

    def light_lists_download(self):
        counts = np.zeros(self._tiles, dtype=np.uint32)
        indices = np.zeros(self._tiles * self._max_per_tile, dtype=np.uint32)
        rc = self.lib.shsb_light_lists_download(self.h, capi.u32ptr(counts), counts.size, capi.u32ptr(indices), indices.size)
        _check(self.lib, self.h, rc, "shsb_light_lists_download")
        return counts, indices.reshape(self._tiles, self._max_per_tile)

    def frame_forward_plus(self, scene: Scene, fp: FrameParams, hdr_rt, depth_rt, ldr_rt, want_stats=True):
        """want_stats=False submits asynchronously (no stats read-back, no host synchronisation) and returns None."""
        st = Stats() if want_stats else None
        rc = self.lib.shsb_frame_forward_plus(self.h, C.byref(scene), C.byref(fp), hdr_rt, depth_rt, ldr_rt, C.byref(st) if want_stats else None)
        _check(self.lib, self.h, rc, "shsb_frame_forward_plus")
        if fp.light_culling:
            _, w, h = self._rt_shape[hdr_rt]
            ts = max(1, fp.tile_size)
            self._tiles = ((w + ts - 1) // ts) * ((h + ts - 1) // ts)
            self._max_per_tile = max(1, fp.max_lights_per_tile)
        return st

    def sync(self):
        _check(self.lib, self.h, self.lib.shsb_sync(self.h), "shsb_sync")

    def fence(self):
        """Main stream behind every frame submitted so far; later frames behind the main stream's current tail."""
        _check(self.lib, self.h, self.lib.shsb_fence(self.h), "shsb_fence")

    def set_tile_streams(self, n: int):
        _check(self.lib, self.h, self.lib.shsb_set_tile_streams(self.h, int(n)), "shsb_set_tile_streams")

    # ---- sort-first frame assembly (include/shsb.h "sort-first frame assembly")
    def gather_create(self, n_ranks: int, slots: int, slot_bytes: int):
        """Root: returns (handle, export); bytes(export) goes to the other ranks."""
        g, exp = C.c_uint32(), capi.GatherExport()
        _check(self.lib, self.h, self.lib.shsb_gather_create(self.h, n_ranks, slots, slot_bytes, C.byref(g), C.byref(exp)), "shsb_gather_create")
        return g.value, exp

    def gather_open(self, export, rank: int) -> int:
        exp = export if isinstance(export, capi.GatherExport) else capi.GatherExport.from_buffer_copy(bytes(export))
        g = C.c_uint32()
        _check(self.lib, self.h, self.lib.shsb_gather_open(self.h, C.byref(exp), rank, C.byref(g)), "shsb_gather_open")
        return g.value

    def gather_destroy(self, g):
        _check(self.lib, self.h, self.lib.shsb_gather_destroy(self.h, g), "shsb_gather_destroy")

    def frame_gather(self, g, step, rt, plane, src_offset, nbytes, dst_offset):
        _check(self.lib, self.h, self.lib.shsb_frame_gather(self.h, g, step, rt, plane, src_offset, nbytes, dst_offset), "shsb_frame_gather")

    def gather_commit(self, g, step):
        _check(self.lib, self.h, self.lib.shsb_gather_commit(self.h, g, step), "shsb_gather_commit")

    def gather_wait(self, g, step):
        _check(self.lib, self.h, self.lib.shsb_gather_wait(self.h, g, step), "shsb_gather_wait")

    def gather_release(self, g, step):
        _check(self.lib, self.h, self.lib.shsb_gather_release(self.h, g, step), "shsb_gather_release")

    def gather_download(self, g, step, offset, nbytes) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        _check(self.lib, self.h, self.lib.shsb_gather_download(self.h, g, step, offset, out.ctypes.data_as(C.c_void_p), nbytes), "shsb_gather_download")
        return out

    def gather_download_async(self, g, step, offset, dst_ptr, nbytes):
        _check(self.lib, self.h, self.lib.shsb_gather_download_async(self.h, g, step, offset, C.c_void_p(dst_ptr), nbytes), "shsb_gather_download_async")

    def gather_stream(self) -> int:
        p = C.c_void_p()
        _check(self.lib, self.h, self.lib.shsb_gather_stream(self.h, C.byref(p)), "shsb_gather_stream")
        return p.value

    def stream(self) -> int:
        p = C.c_void_p()
        _check(self.lib, self.h, self.lib.shsb_stream(self.h, C.byref(p)), "shsb_stream")
        return p.value or 0

    def last_tile_kernel(self) -> int:
        """Instantiation of the tile kernel the last frame launched: 0 general, else PROGRAM * 10 + LIGHTS (include/shsb.h)."""
        m = C.c_int32()
        _check(self.lib, self.h, self.lib.shsb_last_tile_kernel(self.h, C.byref(m)), "shsb_last_tile_kernel")
        return m.value

    def launch_count(self) -> int:
        n = C.c_uint64()
        _check(self.lib, self.h, self.lib.shsb_launch_count(self.h, C.byref(n)), "shsb_launch_count")
        return n.value

    def host_submit_us(self, reset=True) -> np.ndarray:
        a = np.zeros(8, dtype=np.float64)
        _check(self.lib, self.h, self.lib.shsb_host_submit_us(self.h, a.ctypes.data_as(C.POINTER(C.c_double)), int(reset)), "shsb_host_submit_us")
        return a

    def timing_enable(self, on=True):
        _check(self.lib, self.h, self.lib.shsb_timing_enable(self.h, int(on)), "shsb_timing_enable")

    def timing_collect(self, max_frames=4096) -> np.ndarray:
        """(n_frames, 4) float32: vertex+setup, binning, tile raster+shade, total -- milliseconds."""
        a = np.zeros((max_frames, 4), dtype=np.float32)
        n = C.c_size_t()
        _check(self.lib, self.h, self.lib.shsb_timing_collect(self.h, capi.fptr(a), max_frames, C.byref(n)), "shsb_timing_collect")
        return a[: n.value]

    def timing_collect_abs(self, max_frames=4096) -> np.ndarray:
        """(n_frames, 5) float32 ms since the first event: front begin, after geometry, after binning, tile begin, tile end."""
        a = np.zeros((max_frames, 5), dtype=np.float32)
        n = C.c_size_t()
        _check(self.lib, self.h, self.lib.shsb_timing_collect_abs(self.h, capi.fptr(a), max_frames, C.byref(n)), "shsb_timing_collect_abs")
        return a[: n.value]

    def last_stage_ms(self):
        a = np.zeros(8, dtype=np.float32)
        _check(self.lib, self.h, self.lib.shsb_last_stage_ms(self.h, capi.fptr(a)), "shsb_last_stage_ms")
        return a


def model_from_transform(pos, rot, scl) -> np.ndarray:
    lib = capi.load_library()
    tr = Transform()
    capi.set_f(tr.pos, pos)
    capi.set_f(tr.rot_euler, rot)
    capi.set_f(tr.scl, scl)
    out = np.zeros(16, dtype=np.float32)
    lib.shsb_model_from_transform(C.byref(tr), capi.fptr(out))
    return out


def camera_viewproj(eye, target, up, fovy, aspect, zn, zf) -> np.ndarray:
    lib = capi.load_library()
    e = np.asarray(eye, dtype=np.float32)
    t = np.asarray(target, dtype=np.float32)
    u = np.asarray(up, dtype=np.float32)
    out = np.zeros(16, dtype=np.float32)
    lib.shsb_camera_viewproj(capi.fptr(e), capi.fptr(t), capi.fptr(u), fovy, aspect, zn, zf, capi.fptr(out))
    return out


def legacy_camera(position, horizontal_angle_deg=0.0, vertical_angle_deg=0.0):
    """(view, proj) of the legacy demo's Viewer / Camera3D (host-side helper of the C-ABI; needs no device)."""
    lib = capi.load_library()
    pos = np.ascontiguousarray(position, dtype=np.float32)
    view, proj = np.zeros(16, np.float32), np.zeros(16, np.float32)
    assert lib.shsb_legacy_camera(capi.fptr(pos), float(horizontal_angle_deg), float(vertical_angle_deg), capi.fptr(view), capi.fptr(proj)) == 0
    return view, proj


def legacy_world_matrix(position, scale, rotation_angle_deg=0.0):
    lib = capi.load_library()
    p, sc = np.ascontiguousarray(position, dtype=np.float32), np.ascontiguousarray(scale, dtype=np.float32)
    out = np.zeros(16, np.float32)
    assert lib.shsb_legacy_world_matrix(capi.fptr(p), capi.fptr(sc), float(rotation_angle_deg), capi.fptr(out)) == 0
    return out


def legacy_mvp(proj, view, model):
    lib = capi.load_library()
    a, b, c = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (proj, view, model))
    out = np.zeros(16, np.float32)
    assert lib.shsb_legacy_mvp(capi.fptr(a), capi.fptr(b), capi.fptr(c), capi.fptr(out)) == 0
    return out
