"""Dry run, on the CPU, of the HOST logic of the GPU tests that were written after the round's GPU budget was spent
(tests/test_zz_gpu_*.py): their bodies are executed against a stand-in context whose device calls are served by the g++ builds of
the device functions (tests/cpp/legacy2_emul.cpp, tests/cpp/scene_cull_emul.cpp) and, for the light bins, by the restatement.  This
checks the TEST CODE -- argument plumbing, array shapes, gates -- not the kernels; the product is not involved (no libshsb.so call)."""
import numpy as np
import pytest

import fuzz_cases
import test_legacy2_emul_cpu as le
import test_scene_cull_cpu as se
import test_zz_gpu_golden_round1b as zg
import test_zz_gpu_legacy2 as zl
import test_zz_gpu_scene_cull as zs
from leisure_software_renderer_b200 import capi
from oracle.bindings import L2Uniforms, L3Uniforms, Oracle


class StandIn:
    """The subset of renderer.Context the test_zz files use, over numpy arrays."""

    def __init__(self):
        self.rts, self.meshes, self.textures, self.ibls = {}, {}, {}, {}
        self.e2, self.e3, self.sc, self.port = le.Emul2("libm"), le.Emul3("libm"), se.Emul(), Oracle("port")
        self.lights, self.bins, self.ranges = None, {}, None

    # ---- resources
    def rt_create(self, kind, w, h, zn=0.1, zf=1000.0):
        h_ = len(self.rts) + 1
        self.rts[h_] = {"kind": kind, capi.PLANE_COLOR: np.zeros((h, w, 4), np.uint8), capi.PLANE_DEPTH: np.zeros((h, w), np.float32), capi.PLANE_MOTION: np.zeros((h, w, 2), np.float32)}
        return h_

    def rt_destroy(self, rt):
        del self.rts[rt]

    def rt_upload(self, rt, plane, array):
        self.rts[rt][plane][...] = np.asarray(array).reshape(self.rts[rt][plane].shape)

    def rt_clear(self, rt, plane, value):
        self.rts[rt][plane][...] = value

    def rt_download(self, rt, plane=capi.PLANE_COLOR):
        return self.rts[rt][plane].copy()

    def mesh_upload(self, positions, normals=None, uvs=None, indices=None):
        pos, nrm, uv = (None if a is None else np.asarray(a, np.float32) for a in (positions, normals, uvs))
        if indices is not None:
            idx = np.asarray(indices).reshape(-1)
            pos, nrm, uv = pos[idx], nrm[idx], (uv[idx] if uv is not None else None)
        self.meshes[len(self.meshes) + 1] = (pos, nrm, uv if uv is not None else np.zeros((len(pos), 2), np.float32))
        return len(self.meshes)

    def texture_upload(self, rgba):
        self.textures[len(self.textures) + 1] = np.ascontiguousarray(rgba, np.uint8)
        return len(self.textures)

    def legacy3_ibl_upload(self, irradiance, prefiltered):
        self.ibls[len(self.ibls) + 1] = (irradiance, prefiltered)
        return len(self.ibls)

    def legacy3_ibl_destroy(self, ibl):
        del self.ibls[ibl]

    # ---- legacy render-target demos
    def legacy2_shadow_draw(self, mesh, model, light_vp, shadow_map, job_tile_w=0, job_tile_h=0):
        self.e2.shadow_draw(self.meshes[mesh][0], model, light_vp, self.rts[shadow_map][capi.PLANE_DEPTH], job_tile_w or 160, job_tile_h or 160)

    def _common(self, u, cls):
        o = cls()
        for f in ("mvp", "model", "mv", "normal_mat", "light_vp", "light_dir_world", "camera_pos"):
            getattr(o, f)[:] = list(getattr(u, f))
        o.use_texture = u.use_texture
        return o

    def legacy2_draw_softshadow(self, mesh, u, shadow_map, canvas_ldr, zbuffer):
        o = self._common(u, L2Uniforms)
        o.base_color[:] = list(u.base_color)
        pos, nrm, uv = self.meshes[mesh]
        self.e2.camera_draw(pos, nrm, uv, o, self.rts[canvas_ldr][capi.PLANE_COLOR], self.rts[zbuffer][capi.PLANE_DEPTH], texture=self.textures.get(u.albedo),
                            shadow=self.rts[shadow_map][capi.PLANE_DEPTH] if shadow_map else None, tile_w=u.job_tile_w or 160, tile_h=u.job_tile_h or 160)

    def legacy3_draw_pbr(self, mesh, u, shadow_map, ibl, canvas_ldr, depth_motion):
        o = self._common(u, L3Uniforms)
        o.prev_mvp[:] = list(u.prev_mvp)
        o.base_color_srgb[:] = list(u.base_color)
        o.metallic, o.roughness, o.ao = u.metallic, u.roughness, u.ao
        o.ibl_diffuse_intensity, o.ibl_specular_intensity, o.ibl_reflection_strength = u.ibl_diffuse_intensity, u.ibl_specular_intensity, u.ibl_reflection_strength
        pos, nrm, uv = self.meshes[mesh]
        irr, pre = self.ibls[ibl] if ibl else (None, None)
        rt = self.rts[depth_motion]
        self.e3.camera_draw(pos, nrm, uv, o, self.rts[canvas_ldr][capi.PLANE_COLOR], rt[capi.PLANE_DEPTH], rt[capi.PLANE_MOTION], texture=self.textures.get(u.albedo),
                            shadow=self.rts[shadow_map][capi.PLANE_DEPTH] if shadow_map else None, irradiance=irr, prefiltered=pre, tile_w=u.job_tile_w or 160, tile_h=u.job_tile_h or 160)

    # ---- scene-level culling
    def cull_objects_frustum(self, bounds10, view_proj):
        return self.sc.cull_objects(bounds10, view_proj)

    def collect_object_lights(self, object_aabbs, visible, records, cull_mode):
        if not 0 <= cull_mode <= 2:
            raise capi.ShsbError("shsb_collect_object_lights failed: status 1: unknown LightObjectCullMode")
        return self.sc.collect_object_lights(object_aabbs, visible, records, cull_mode)

    def tile_depth_range_from_scene(self, object_aabbs, visible_objects, view, view_proj, w, h, tile_size, z_near, z_far):
        self.ranges = self.sc.tile_depth_range_from_scene(object_aabbs, visible_objects, view, view_proj, w, h, tile_size, z_near, z_far)
        return self.ranges

    def lights_upload(self, records):
        self.lights = np.array(records, copy=True)

    def light_cull_ex(self, desc, range_min=None, range_max=None):
        if desc.mode == capi.LIGHT_CULL_TILED_VIEW_DEPTH and range_min is None:
            range_min, range_max = self.ranges
        c, i = self.port.light_cull_ex(self.lights, desc, range_min, range_max)
        self.bins[desc.mode == capi.LIGHT_CULL_CLUSTERED] = (c, i, desc.tiles(), desc)
        return c, i

    def select_object_lights_from_bins(self, object_aabbs, view, view_proj, clustered, z_near, z_far, records, cull_mode):
        c, i, _, d = self.bins[bool(clustered)]
        tx, ty = (d.viewport_w + d.tile_size - 1) // d.tile_size, (d.viewport_h + d.tile_size - 1) // d.tile_size
        zn = float(np.float32(max(np.float32(z_near), np.float32(1e-4))))
        zf = float(np.float32(max(np.float32(z_far), np.float32(zn) + np.float32(1e-3))))
        return self.sc.select_object_lights_from_bins(object_aabbs, view, view_proj, (tx, ty, d.depth_slices if clustered else 1), clustered, zn, zf, c, i, records, cull_mode)


@pytest.fixture(scope="module")
def standin():
    return StandIn()


def test_dry_run_legacy2_gpu_tests(standin):
    for seed in (0, 5):
        zl.test_fuzz_legacy2_softshadow_parity(standin, seed)
        zl.test_fuzz_legacy3_pbr_parity(standin, seed)
    zl.test_legacy2_indexed_mesh_equals_soup(standin)


def test_dry_run_scene_cull_gpu_tests(standin):
    for seed in (0, 1, 4):
        zs.test_fuzz_object_culling_parity(standin, seed)
        zs.test_fuzz_object_light_selection_parity(standin, seed)
        zs.test_fuzz_scene_tile_depth_range_parity(standin, seed)
        zs.test_fuzz_bin_gather_and_selection_parity(standin, seed)
    zs.test_scene_cull_large_and_empty(standin)
    zs.test_scene_depth_ranges_feed_the_view_depth_bin_builder(standin)


def test_dry_run_golden_gpu_tests(standin):
    zg.test_legacy_render_target_demos_match_the_reference_fixture(standin)
    zg.test_scene_culling_matches_the_reference_fixture(standin)
    zg.test_light_lists_match_the_reference_fixture(standin)
