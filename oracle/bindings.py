"""oracle/bindings.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings of the two CPU checkers declared in oracle/oracle_abi.h:
  Oracle("port")      -> oracle/liboracle.so        (this repo's restatement, shso_*)
  Oracle("reference") -> oracle/_ref/libshs_ref.so  (the reference's own headers, shsref_*)
                         oracle/_ref/libshs_legacy_ref.so (the reference's legacy tile-job demo sources, shsref_legacy_*)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from leisure_software_renderer_b200 import capi
from leisure_software_renderer_b200.capi import FrameParams, RasterCfg, Scene, Stats, Transform, Uniforms

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_HERE, "liboracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libshs_ref.so")
REF_LEGACY_LIB = os.path.join(_HERE, "_ref", "libshs_legacy_ref.so")
REF_LEGACY2_LIB = os.path.join(_HERE, "_ref", "libshs_legacy2_ref.so")
REF_LEGACY3_LIB = os.path.join(_HERE, "_ref", "libshs_legacy3_ref.so")
REF_LIGHTCULL_LIB = os.path.join(_HERE, "_ref", "libshs_lightcull_ref.so")
REF_TAA_LIB = os.path.join(_HERE, "_ref", "libshs_taa_ref.so")


class Mesh(C.Structure):
    _fields_ = [("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)), ("uvs", C.POINTER(C.c_float)),
                ("indices", C.POINTER(C.c_uint32)), ("n_positions", C.c_uint32), ("n_normals", C.c_uint32),
                ("n_uvs", C.c_uint32), ("n_indices", C.c_uint32)]


class Texture(C.Structure):
    _fields_ = [("rgba", C.POINTER(C.c_uint8)), ("w", C.c_int32), ("h", C.c_int32)]


class Assets(C.Structure):
    _fields_ = [("meshes", C.POINTER(Mesh)), ("textures", C.POINTER(Texture)), ("n_meshes", C.c_uint32), ("n_textures", C.c_uint32)]


class Target(C.Structure):
    _fields_ = [("hdr", C.POINTER(C.c_float)), ("depth", C.POINTER(C.c_float)), ("shadow", C.POINTER(C.c_float)),
                ("tri_id", C.POINTER(C.c_uint32)), ("coverage", C.POINTER(C.c_uint32)),
                ("w", C.c_int32), ("h", C.c_int32), ("shadow_w", C.c_int32), ("shadow_h", C.c_int32),
                ("zn", C.c_float), ("zf", C.c_float), ("motion", C.POINTER(C.c_float))]


def _records_u8(records) -> np.ndarray:
    """CullingLightGPU records (structured or raw) as a contiguous (n, 160) uint8 array."""
    r = np.ascontiguousarray(records)
    return r.view(np.uint8).reshape(-1, 160)


def build(kind: str = "port") -> None:
    subprocess.run(["make", "-C", _HERE, "oracle" if kind == "port" else "ref"], check=True, capture_output=True)


def available(kind: str) -> bool:
    return os.path.exists(PORT_LIB if kind == "port" else REF_LIB)


def _tile_depth_range_ndc01(fn, depth, tile_size, z_near, z_far):
    d = np.ascontiguousarray(depth, dtype=np.float32)
    h, w = d.shape
    n = ((w + tile_size - 1) // tile_size) * ((h + tile_size - 1) // tile_size)
    lo, hi = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    rc = fn(capi.fptr(d), C.c_int32(w), C.c_int32(h), C.c_uint32(tile_size), C.c_float(z_near), C.c_float(z_far), capi.fptr(lo), capi.fptr(hi))
    assert rc == 0, rc
    return lo, hi


def glsl_depth_reduce_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libshs_glsl_a9_ref.so")) or os.path.isdir("/root/reference")


def glsl_tile_depth_range_ndc01(depth, tile_size, z_near, z_far):
    """shaders/vulkan/fp_stress_depth_reduce.comp, the shader's own text compiled as C++ (oracle/extract_glsl_a9.py +
    oracle/ref_glsl_a9_harness.cpp: shsglsl_tile_depth_range_ndc01)."""
    path = os.path.join(_HERE, "_ref", "libshs_glsl_a9_ref.so")
    if not os.path.exists(path):
        build("reference")
    return _tile_depth_range_ndc01(C.CDLL(path).shsglsl_tile_depth_range_ndc01, depth, tile_size, z_near, z_far)


class HostAssets:
    """Keeps numpy arrays alive behind a ShsoAssets block (1-based handles like ResourceRegistry)."""

    def __init__(self, meshes, textures=()):
        self._keep = []
        self.meshes = (Mesh * max(1, len(meshes)))()
        for i, m in enumerate(meshes):
            pos = np.ascontiguousarray(m["positions"], dtype=np.float32).reshape(-1, 3)
            nrm = np.ascontiguousarray(m.get("normals", np.zeros((0, 3))), dtype=np.float32).reshape(-1, 3)
            uv = np.ascontiguousarray(m.get("uvs", np.zeros((0, 2))), dtype=np.float32).reshape(-1, 2)
            idx = np.ascontiguousarray(m.get("indices", np.zeros((0,))), dtype=np.uint32).reshape(-1)
            self._keep += [pos, nrm, uv, idx]
            self.meshes[i] = Mesh(capi.fptr(pos), capi.fptr(nrm), capi.fptr(uv), capi.u32ptr(idx), len(pos), len(nrm), len(uv), len(idx))
        self.textures = (Texture * max(1, len(textures)))()
        for i, t in enumerate(textures):
            a = np.ascontiguousarray(t, dtype=np.uint8)
            self._keep.append(a)
            self.textures[i] = Texture(a.ctypes.data_as(C.POINTER(C.c_uint8)), a.shape[1], a.shape[0])
        self.block = Assets(self.meshes, self.textures, len(meshes), len(textures))


class Oracle:
    def __init__(self, kind: str = "port"):
        assert kind in ("port", "reference")
        self.kind = kind
        path = PORT_LIB if kind == "port" else REF_LIB
        if not os.path.exists(path):
            build(kind)
        self.lib = C.CDLL(path)
        self.prefix = "shso_" if kind == "port" else "shsref_"

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- helpers
    def model_from_transform(self, pos, rot, scl):
        tr = Transform()
        capi.set_f(tr.pos, pos); capi.set_f(tr.rot_euler, rot); capi.set_f(tr.scl, scl)
        out = np.zeros(16, dtype=np.float32)
        f = self.fn("model_from_transform"); f.restype = None
        f(C.byref(tr), capi.fptr(out))
        return out

    def camera_viewproj(self, eye, target, up, fovy, aspect, zn, zf):
        e, t, u = (np.asarray(v, dtype=np.float32) for v in (eye, target, up))
        out = np.zeros(16, dtype=np.float32)
        f = self.fn("camera_viewproj"); f.restype = None
        f(capi.fptr(e), capi.fptr(t), capi.fptr(u), C.c_float(fovy), C.c_float(aspect), C.c_float(zn), C.c_float(zf), capi.fptr(out))
        return out

    def pack_point_light(self, pos, rng, color, intensity, model=1, power=1.0, bias=0.05, cutoff=0.0, jolt_bounds=False):
        p, c = np.asarray(pos, dtype=np.float32), np.asarray(color, dtype=np.float32)
        out = np.zeros(160, dtype=np.uint8)
        f = self.fn("pack_point_light"); f.restype = None
        f(capi.fptr(p), C.c_float(rng), capi.fptr(c), C.c_float(intensity), C.c_uint32(model), C.c_float(power), C.c_float(bias),
          C.c_float(cutoff), C.c_int32(int(jolt_bounds)), out.ctypes.data_as(C.c_void_p))
        return out

    def pack_spot_light(self, pos, rng, color, intensity, direction, inner, outer, model=1, power=1.0, bias=0.05, cutoff=0.0):
        p, c, d = (np.asarray(v, dtype=np.float32) for v in (pos, color, direction))
        out = np.zeros(160, dtype=np.uint8)
        f = self.fn("pack_spot_light"); f.restype = None
        f(capi.fptr(p), C.c_float(rng), capi.fptr(c), C.c_float(intensity), capi.fptr(d), C.c_float(inner), C.c_float(outer),
          C.c_uint32(model), C.c_float(power), C.c_float(bias), C.c_float(cutoff), out.ctypes.data_as(C.c_void_p))
        return out

    def pack_rect_light(self, pos, rng, color, intensity, direction, right, half_x, half_y, flags=7, model=1, power=1.0, bias=0.05, cutoff=0.0):
        assert self.kind == "reference"
        p, c, d, r = (np.asarray(v, dtype=np.float32) for v in (pos, color, direction, right))
        out = np.zeros(160, dtype=np.uint8)
        f = self.lib.shsref_pack_rect_light; f.restype = None
        f(capi.fptr(p), C.c_float(rng), capi.fptr(c), C.c_float(intensity), capi.fptr(d), capi.fptr(r), C.c_float(half_x), C.c_float(half_y),
          C.c_uint32(flags), C.c_uint32(model), C.c_float(power), C.c_float(bias), C.c_float(cutoff), out.ctypes.data_as(C.c_void_p))
        return out

    def pack_tube_light(self, pos, rng, color, intensity, axis, half_length, radius, flags=7, model=1, power=1.0, bias=0.05, cutoff=0.0):
        assert self.kind == "reference"
        p, c, a = (np.asarray(v, dtype=np.float32) for v in (pos, color, axis))
        out = np.zeros(160, dtype=np.uint8)
        f = self.lib.shsref_pack_tube_light; f.restype = None
        f(capi.fptr(p), C.c_float(rng), capi.fptr(c), C.c_float(intensity), capi.fptr(a), C.c_float(half_length), C.c_float(radius),
          C.c_uint32(flags), C.c_uint32(model), C.c_float(power), C.c_float(bias), C.c_float(cutoff), out.ctypes.data_as(C.c_void_p))
        return out

    def set_threads(self, n):
        if self.kind == "reference":
            self.lib.shsref_set_threads(C.c_int32(n))

    # ---- target plumbing
    @staticmethod
    def make_target(w, h, hdr, depth=None, shadow=None, tri_id=None, coverage=None, zn=0.1, zf=1000.0, motion=None):
        t = Target()
        t.hdr = capi.fptr(hdr)
        t.depth = capi.fptr(depth) if depth is not None else None
        if shadow is not None:
            t.shadow = capi.fptr(shadow)
            t.shadow_h, t.shadow_w = shadow.shape
        t.tri_id = capi.u32ptr(tri_id) if tri_id is not None else None
        t.coverage = capi.u32ptr(coverage) if coverage is not None else None
        t.w, t.h, t.zn, t.zf = w, h, zn, zf
        t.motion = capi.fptr(motion) if motion is not None else None
        return t

    # ---- passes
    def rasterize_mesh(self, assets: HostAssets, mesh, shader_id, u: Uniforms, tgt: Target, cull_mode=capi.CULL_BACK,
                       front_face_ccw=True, key_base=0):
        cfg = RasterCfg(cull_mode, int(front_face_ccw), 0, 0)
        st = Stats()
        rc = self.fn("rasterize_mesh")(C.byref(assets.block), C.c_uint32(mesh), C.c_int32(shader_id), C.byref(u), C.byref(tgt),
                                       C.byref(cfg), C.c_uint32(key_base), C.byref(st))
        assert rc == 0, rc
        return st

    def pass_pbr_forward(self, assets: HostAssets, scene: Scene, fp: FrameParams, tgt: Target, shadow_lvp=None, preserve_depth=False,
                         prev_models=None):
        """prev_models: (n_items, 16) model matrices of the previous frame (Context::history), None = first frame."""
        st = Stats()
        lvp = np.ascontiguousarray(shadow_lvp, dtype=np.float32) if shadow_lvp is not None else None
        pm = np.ascontiguousarray(prev_models, dtype=np.float32) if prev_models is not None else None
        rc = self.fn("pass_pbr_forward_history")(C.byref(assets.block), C.byref(scene), C.byref(fp), C.byref(tgt),
                                                 capi.fptr(lvp) if lvp is not None else None, C.c_int32(int(preserve_depth)),
                                                 capi.fptr(pm) if pm is not None else None, C.byref(st))
        assert rc == 0, rc
        return st

    def pass_shadow_map(self, assets: HostAssets, scene: Scene, fp: FrameParams, sw, sh):
        shadow = np.zeros((sh, sw), dtype=np.float32)
        lvp = np.zeros(16, dtype=np.float32)
        rc = self.fn("pass_shadow_map")(C.byref(assets.block), C.byref(scene), C.byref(fp), capi.fptr(shadow), C.c_int32(sw), C.c_int32(sh), capi.fptr(lvp))
        assert rc == 0, rc
        return shadow, lvp

    def pass_tonemap(self, hdr, exposure=1.0, gamma=2.2):
        h, w = hdr.shape[:2]
        ldr = np.zeros((h, w, 4), dtype=np.uint8)
        rc = self.fn("pass_tonemap")(capi.fptr(hdr), C.c_int32(w), C.c_int32(h), C.c_float(exposure), C.c_float(gamma),
                                     ldr.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert rc == 0, rc
        return ldr

    def pass_motion_blur(self, params, src_ldr, motion, depth):
        h, w = src_ldr.shape[:2]
        src = np.ascontiguousarray(src_ldr, dtype=np.uint8)
        mot = np.ascontiguousarray(motion, dtype=np.float32)
        dep = np.ascontiguousarray(depth, dtype=np.float32)
        out = np.zeros((h, w, 4), dtype=np.uint8)
        u8 = C.POINTER(C.c_uint8)
        rc = self.fn("pass_motion_blur")(C.byref(params), src.ctypes.data_as(u8), capi.fptr(mot), capi.fptr(dep), C.c_int32(w), C.c_int32(h),
                                         out.ctypes.data_as(u8))
        assert rc == 0, rc
        return out

    def pass_light_shafts(self, params, src_ldr, depth=None):
        h, w = src_ldr.shape[:2]
        src = np.ascontiguousarray(src_ldr, dtype=np.uint8)
        dep = np.ascontiguousarray(depth, dtype=np.float32) if depth is not None else None
        out = np.zeros((h, w, 4), dtype=np.uint8)
        u8 = C.POINTER(C.c_uint8)
        rc = self.fn("pass_light_shafts")(C.byref(params), src.ctypes.data_as(u8), capi.fptr(dep) if dep is not None else None,
                                          C.c_int32(w), C.c_int32(h), out.ctypes.data_as(u8))
        assert rc == 0, rc
        return out

    # ---- restatement-only
    def light_cull(self, records, view_proj, w, h, tile_size=16, max_per_tile=128):
        assert self.kind == "port"
        r = _records_u8(records)
        vp = np.ascontiguousarray(view_proj, dtype=np.float32)
        tiles = ((w + tile_size - 1) // tile_size) * ((h + tile_size - 1) // tile_size)
        counts = np.zeros(tiles, dtype=np.uint32)
        indices = np.zeros((tiles, max_per_tile), dtype=np.uint32)
        rc = self.lib.shso_light_cull(r.ctypes.data_as(C.c_void_p), C.c_uint32(len(r)), capi.fptr(vp), C.c_uint32(w), C.c_uint32(h),
                                      C.c_uint32(tile_size), C.c_uint32(max_per_tile), capi.u32ptr(counts), capi.u32ptr(indices))
        assert rc == 0, rc
        return counts, indices

    def light_cull_ex(self, records, desc, range_min=None, range_max=None):
        assert self.kind == "port"
        r = _records_u8(records)
        lo = np.ascontiguousarray(range_min, dtype=np.float32).reshape(-1) if range_min is not None else None
        hi = np.ascontiguousarray(range_max, dtype=np.float32).reshape(-1) if range_max is not None else None
        bins, mx = desc.bins(), desc.max_per_bin
        counts = np.zeros(bins, dtype=np.uint32)
        indices = np.zeros((bins, mx), dtype=np.uint32)
        rc = self.lib.shso_light_cull_ex(r.ctypes.data_as(C.c_void_p), C.c_uint32(len(r)), C.byref(desc),
                                         capi.fptr(lo) if lo is not None else None, capi.fptr(hi) if hi is not None else None,
                                         capi.u32ptr(counts), capi.u32ptr(indices))
        assert rc == 0, rc
        return counts, indices

    def tile_depth_range(self, depth, tile_size, zn, zf):
        assert self.kind == "port"
        d = np.ascontiguousarray(depth, dtype=np.float32)
        h, w = d.shape
        n = ((w + tile_size - 1) // tile_size) * ((h + tile_size - 1) // tile_size)
        lo, hi = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        rc = self.lib.shso_tile_depth_range(capi.fptr(d), C.c_int32(w), C.c_int32(h), C.c_uint32(tile_size), C.c_float(zn), C.c_float(zf),
                                            capi.fptr(lo), capi.fptr(hi))
        assert rc == 0, rc
        return lo, hi

    def tile_depth_range_ndc01(self, depth, tile_size, z_near, z_far):
        """fp_stress_depth_reduce.comp restated (shso_tile_depth_range_ndc01); glsl_tile_depth_range_ndc01 below is the shader's own text."""
        assert self.kind == "port"
        return _tile_depth_range_ndc01(self.lib.shso_tile_depth_range_ndc01, depth, tile_size, z_near, z_far)

    def pass_taa(self, ldr, history, history_valid):
        """In place on both arrays (PassTemporalAAAdapter); returns them."""
        assert self.kind == "port"
        assert ldr.dtype == np.uint8 and history.dtype == np.uint8 and ldr.flags.c_contiguous and history.flags.c_contiguous
        u8 = C.POINTER(C.c_uint8)
        rc = self.lib.shso_pass_taa(ldr.ctypes.data_as(u8), history.ctypes.data_as(u8), C.c_int32(int(history_valid)), C.c_size_t(ldr.size // 4))
        assert rc == 0, rc
        return ldr, history

    def pass_pbr_forward_plus(self, assets, scene, fp, tgt, records, counts, indices, shadow_lvp=None, preserve_depth=False):
        assert self.kind == "port"
        st = Stats()
        r = _records_u8(records)
        lvp = np.ascontiguousarray(shadow_lvp, dtype=np.float32) if shadow_lvp is not None else None
        rc = self.lib.shso_pass_pbr_forward_plus(C.byref(assets.block), C.byref(scene), C.byref(fp), C.byref(tgt),
                                                 capi.fptr(lvp) if lvp is not None else None, C.c_int32(int(preserve_depth)),
                                                 r.ctypes.data_as(C.c_void_p), C.c_uint32(len(r)), capi.u32ptr(counts), capi.u32ptr(indices),
                                                 C.byref(st))
        assert rc == 0, rc
        return st

    def pass_depth_prepass(self, assets, scene, fp, tgt):
        assert self.kind == "port"
        st = Stats()
        rc = self.lib.shso_pass_depth_prepass(C.byref(assets.block), C.byref(scene), C.byref(fp), C.byref(tgt), C.byref(st))
        assert rc == 0, rc
        return st


class LegacyOracle:
    """The legacy tile-job rasterizer (BASELINE configs[0] as shipped) on the CPU:
    LegacyOracle("port") = oracle/oracle_legacy.cpp, LegacyOracle("reference") = the reference's own demo sources
    (hello_pipeline_blinn_phong_shading.cpp + shs_renderer.hpp) compiled by oracle/ref_legacy_harness.cpp."""

    def __init__(self, kind: str = "port"):
        assert kind in ("port", "reference")
        self.kind = kind
        path = PORT_LIB if kind == "port" else REF_LEGACY_LIB
        if not os.path.exists(path):
            build(kind)
        self.lib = C.CDLL(path)
        self.prefix = "shso_legacy_" if kind == "port" else "shsref_legacy_"

    @staticmethod
    def available(kind: str) -> bool:
        return os.path.exists(PORT_LIB if kind == "port" else REF_LEGACY_LIB)

    def fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        return f

    def camera(self, position, horizontal_angle_deg=0.0, vertical_angle_deg=0.0):
        pos = np.ascontiguousarray(position, dtype=np.float32)
        view, proj = np.zeros(16, np.float32), np.zeros(16, np.float32)
        f = self.fn("camera"); f.restype = None
        f(capi.fptr(pos), C.c_float(horizontal_angle_deg), C.c_float(vertical_angle_deg), capi.fptr(view), capi.fptr(proj))
        return view, proj

    def world_matrix(self, position, scale, rotation_angle_deg=0.0):
        p, s = np.ascontiguousarray(position, dtype=np.float32), np.ascontiguousarray(scale, dtype=np.float32)
        out = np.zeros(16, np.float32)
        f = self.fn("world_matrix"); f.restype = None
        f(capi.fptr(p), capi.fptr(s), C.c_float(rotation_angle_deg), capi.fptr(out))
        return out

    def mvp(self, proj, view, model):
        a, b, c = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (proj, view, model))
        out = np.zeros(16, np.float32)
        f = self.fn("mvp"); f.restype = None
        f(capi.fptr(a), capi.fptr(b), capi.fptr(c), capi.fptr(out))
        return out

    def draw(self, positions, normals, mvp, model, light_dir, camera_pos, color, canvas, zbuffer, tile_w=80, tile_h=80):
        """In place on canvas (H, W, 4) uint8 in shs::Canvas order and zbuffer (H, W) float32 in shs::ZBuffer order.
        positions / normals: (n_vertices, 3) float32 triangle soup (ModelGeometry::triangles / normals)."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        assert len(pos) == len(nrm)
        assert canvas.dtype == np.uint8 and canvas.flags.c_contiguous and zbuffer.dtype == np.float32 and zbuffer.flags.c_contiguous
        h, w = zbuffer.shape
        m0 = np.ascontiguousarray(mvp, dtype=np.float32).reshape(16)
        m1 = np.ascontiguousarray(model, dtype=np.float32).reshape(16)
        ld, cp = np.ascontiguousarray(light_dir, dtype=np.float32), np.ascontiguousarray(camera_pos, dtype=np.float32)
        col = np.ascontiguousarray(color, dtype=np.uint8)
        rc = self.fn("draw")(capi.fptr(pos), capi.fptr(nrm), C.c_uint32(len(pos)), capi.fptr(m0), capi.fptr(m1), capi.fptr(ld), capi.fptr(cp),
                             col.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int32(w), C.c_int32(h), C.c_int32(tile_w), C.c_int32(tile_h),
                             canvas.ctypes.data_as(C.POINTER(C.c_uint8)), capi.fptr(zbuffer))
        assert rc == 0, rc
        return canvas, zbuffer


class L2Uniforms(C.Structure):
    """struct Uniforms of the legacy soft-shadow demo (hello_shadow_mapping_soft.cpp:714-732) as plain data."""
    _fields_ = [("mvp", C.c_float * 16), ("model", C.c_float * 16), ("mv", C.c_float * 16), ("normal_mat", C.c_float * 9), ("light_vp", C.c_float * 16),
                ("light_dir_world", C.c_float * 3), ("camera_pos", C.c_float * 3), ("base_color", C.c_uint8 * 4), ("use_texture", C.c_int32)]


class Legacy2Oracle:
    """The legacy soft-shadow demo (config-3 flavour) on the CPU: "port" = oracle/oracle_legacy.cpp (second
    half), "reference" = hello_shadow_mapping_soft.cpp compiled by oracle/ref_legacy2_harness.cpp.  The CUDA path of this row (csrc/legacy2.cu) is compared with the port by tests/test_zz_gpu_legacy2.py."""

    def __init__(self, kind: str = "port"):
        assert kind in ("port", "reference")
        self.kind = kind
        path = PORT_LIB if kind == "port" else REF_LEGACY2_LIB
        if not os.path.exists(path):
            build(kind)
        self.lib = C.CDLL(path)
        self.prefix = "shso_l2_" if kind == "port" else "shsref_l2_"

    @staticmethod
    def available(kind: str) -> bool:
        return os.path.exists(PORT_LIB if kind == "port" else REF_LEGACY2_LIB)

    def shadow_draw(self, positions, model, light_vp, shadow, tile_w=160, tile_h=160):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        m, lvp = (np.ascontiguousarray(a, dtype=np.float32).reshape(16) for a in (model, light_vp))
        assert shadow.dtype == np.float32 and shadow.flags.c_contiguous
        h, w = shadow.shape
        rc = getattr(self.lib, self.prefix + "shadow_draw")(capi.fptr(pos), C.c_uint32(len(pos)), capi.fptr(m), capi.fptr(lvp), C.c_int32(w), C.c_int32(h),
                                                            C.c_int32(tile_w), C.c_int32(tile_h), capi.fptr(shadow))
        assert rc == 0, rc
        return shadow

    def camera_draw(self, positions, normals, uvs, uniforms: L2Uniforms, canvas, zbuffer, texture=None, shadow=None, tile_w=160, tile_h=160):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        assert len(pos) == len(nrm) == len(uv)
        assert canvas.dtype == np.uint8 and canvas.flags.c_contiguous and zbuffer.dtype == np.float32 and zbuffer.flags.c_contiguous
        h, w = zbuffer.shape
        u8 = C.POINTER(C.c_uint8)
        tex = np.ascontiguousarray(texture, dtype=np.uint8) if texture is not None else None
        sm = np.ascontiguousarray(shadow, dtype=np.float32) if shadow is not None else None
        rc = getattr(self.lib, self.prefix + "camera_draw")(
            capi.fptr(pos), capi.fptr(nrm), capi.fptr(uv), C.c_uint32(len(pos)), C.byref(uniforms),
            tex.ctypes.data_as(u8) if tex is not None else None, C.c_int32(tex.shape[1] if tex is not None else 0), C.c_int32(tex.shape[0] if tex is not None else 0),
            capi.fptr(sm) if sm is not None else None, C.c_int32(sm.shape[1] if sm is not None else 0), C.c_int32(sm.shape[0] if sm is not None else 0),
            C.c_int32(w), C.c_int32(h), C.c_int32(tile_w), C.c_int32(tile_h), canvas.ctypes.data_as(u8), capi.fptr(zbuffer))
        assert rc == 0, rc
        return canvas, zbuffer


class L3Uniforms(C.Structure):
    """struct Uniforms + MaterialPBR of the legacy PBR / IBL demo (hello_pbr.cpp:474-519) as plain data."""
    _fields_ = [("mvp", C.c_float * 16), ("prev_mvp", C.c_float * 16), ("model", C.c_float * 16), ("mv", C.c_float * 16), ("normal_mat", C.c_float * 9),
                ("light_vp", C.c_float * 16), ("light_dir_world", C.c_float * 3), ("camera_pos", C.c_float * 3), ("base_color_srgb", C.c_uint8 * 4),
                ("metallic", C.c_float), ("roughness", C.c_float), ("ao", C.c_float), ("use_texture", C.c_int32),
                ("ibl_diffuse_intensity", C.c_float), ("ibl_specular_intensity", C.c_float), ("ibl_reflection_strength", C.c_float)]


class Legacy3Oracle:
    """The legacy PBR / IBL demo (config-4 flavour) on the CPU: "port" = oracle/oracle_legacy.cpp (third part),
    "reference" = hello_pbr.cpp compiled by oracle/ref_legacy3_harness.cpp.  The CUDA path (csrc/legacy2.cu, MODE_PBR) is compared with the port by tests/test_zz_gpu_legacy2.py."""

    def __init__(self, kind: str = "port"):
        assert kind in ("port", "reference")
        self.kind = kind
        path = PORT_LIB if kind == "port" else REF_LEGACY3_LIB
        if not os.path.exists(path):
            build(kind)
        self.lib = C.CDLL(path)
        self.prefix = "shso_l3_" if kind == "port" else "shsref_l3_"

    @staticmethod
    def available(kind: str) -> bool:
        return os.path.exists(PORT_LIB if kind == "port" else REF_LEGACY3_LIB)

    def shadow_draw(self, positions, model, light_vp, shadow, tile_w=160, tile_h=160):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        m, lvp = (np.ascontiguousarray(a, dtype=np.float32).reshape(16) for a in (model, light_vp))
        assert shadow.dtype == np.float32 and shadow.flags.c_contiguous
        h, w = shadow.shape
        rc = getattr(self.lib, self.prefix + "shadow_draw")(capi.fptr(pos), C.c_uint32(len(pos)), capi.fptr(m), capi.fptr(lvp), C.c_int32(w), C.c_int32(h),
                                                            C.c_int32(tile_w), C.c_int32(tile_h), capi.fptr(shadow))
        assert rc == 0, rc
        return shadow

    def camera_draw(self, positions, normals, uvs, uniforms: L3Uniforms, canvas, zbuffer, velocity, texture=None, shadow=None, irradiance=None,
                    prefiltered=None, tile_w=160, tile_h=160):
        """In place on canvas (H, W, 4) uint8, zbuffer (H, W) float32 and velocity (H, W, 2) float32 (shs::Buffer order).
        irradiance: (6, n, n, 3) float32; prefiltered: list of (6, n_m, n_m, 3) float32 mips."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        assert len(pos) == len(nrm) == len(uv)
        for a, dt in ((canvas, np.uint8), (zbuffer, np.float32), (velocity, np.float32)):
            assert a.dtype == dt and a.flags.c_contiguous
        h, w = zbuffer.shape
        u8 = C.POINTER(C.c_uint8)
        tex = np.ascontiguousarray(texture, dtype=np.uint8) if texture is not None else None
        sm = np.ascontiguousarray(shadow, dtype=np.float32) if shadow is not None else None
        irr = np.ascontiguousarray(irradiance, dtype=np.float32) if irradiance is not None else None
        if prefiltered is not None:
            sizes = np.array([m.shape[1] for m in prefiltered], np.int32)
            pre = np.concatenate([np.ascontiguousarray(m, dtype=np.float32).reshape(-1) for m in prefiltered])
        else:
            sizes = pre = None
        rc = getattr(self.lib, self.prefix + "camera_draw")(
            capi.fptr(pos), capi.fptr(nrm), capi.fptr(uv), C.c_uint32(len(pos)), C.byref(uniforms),
            tex.ctypes.data_as(u8) if tex is not None else None, C.c_int32(tex.shape[1] if tex is not None else 0), C.c_int32(tex.shape[0] if tex is not None else 0),
            capi.fptr(sm) if sm is not None else None, C.c_int32(sm.shape[1] if sm is not None else 0), C.c_int32(sm.shape[0] if sm is not None else 0),
            capi.fptr(irr) if irr is not None else None, C.c_int32(irr.shape[1] if irr is not None else 0),
            capi.fptr(pre) if pre is not None else None, sizes.ctypes.data_as(C.POINTER(C.c_int32)) if sizes is not None else None,
            C.c_int32(len(sizes) if sizes is not None else 0),
            C.c_int32(w), C.c_int32(h), C.c_int32(tile_w), C.c_int32(tile_h), canvas.ctypes.data_as(u8), capi.fptr(zbuffer), capi.fptr(velocity))
        assert rc == 0, rc
        return canvas, zbuffer, velocity


class LightCullReference:
    """The reference's own light-list builders (lighting/jolt_light_culling.hpp:135-412) compiled with SHS_HAS_JOLT=1 against the
    JoltPhysics declaration shim (oracle/jolt_shim) by oracle/ref_lightcull_harness.cpp.  A light is its world AABB (SHS space)."""

    def __init__(self):
        if not os.path.exists(REF_LIGHTCULL_LIB):
            build("reference")
        self.lib = C.CDLL(REF_LIGHTCULL_LIB)

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_LIGHTCULL_LIB)

    def bounds(self, aabbs):
        """(n, 6) AABBs -> (n, 10): sphere centre xyz + radius, AABB min xyz, max xyz as SceneShape reports them."""
        a = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        out = np.zeros((len(a), 10), np.float32)
        assert self.lib.shsref_light_bounds(capi.fptr(a), C.c_uint32(len(a)), capi.fptr(out)) == 0
        return out

    def light_cull(self, aabbs, desc, range_min=None, range_max=None):
        a = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        lo = np.ascontiguousarray(range_min, dtype=np.float32).reshape(-1) if range_min is not None else None
        hi = np.ascontiguousarray(range_max, dtype=np.float32).reshape(-1) if range_max is not None else None
        bins, mx = desc.bins(), desc.max_per_bin
        counts = np.zeros(bins, dtype=np.uint32)
        indices = np.zeros((bins, mx), dtype=np.uint32)
        vp = np.ascontiguousarray(np.frombuffer(bytes(desc.view_proj), dtype=np.float32))
        rc = self.lib.shsref_light_cull(capi.fptr(a), C.c_uint32(len(a)), capi.fptr(vp), C.c_uint32(desc.viewport_w), C.c_uint32(desc.viewport_h),
                                        C.c_uint32(desc.tile_size), C.c_uint32(mx), C.c_int32(desc.mode), C.c_uint32(desc.depth_slices),
                                        C.c_float(desc.z_near), C.c_float(desc.z_far), capi.fptr(lo) if lo is not None else None,
                                        capi.fptr(hi) if hi is not None else None, C.c_uint32(len(lo) if lo is not None else 0),
                                        capi.u32ptr(counts), capi.u32ptr(indices))
        assert rc == 0, rc
        return counts, indices


def reference_pass_taa(ldr, history, history_valid):
    """The reference's own PassTemporalAAAdapter (pipeline/pass_adapters.hpp:1402-1491, compiled by oracle/ref_taa_harness.cpp against
    the JoltPhysics declaration shim); in place on ldr / history (H, W, 4) uint8 like Oracle.pass_taa."""
    if not os.path.exists(REF_TAA_LIB):
        build("reference")
    lib = C.CDLL(REF_TAA_LIB)
    assert ldr.dtype == np.uint8 and history.dtype == np.uint8 and ldr.flags.c_contiguous and history.flags.c_contiguous and ldr.shape == history.shape
    u8 = C.POINTER(C.c_uint8)
    rc = lib.shsref_pass_taa(ldr.ctypes.data_as(u8), history.ctypes.data_as(u8), C.c_int32(int(history_valid)), C.c_int32(ldr.shape[1]), C.c_int32(ldr.shape[0]))
    assert rc == 0, rc
    return ldr, history


class SceneCull:
    """Scene-level steps upstream of draw submission: "port" = oracle/oracle_scene_cull.cpp, "reference" = the
    reference's cull_vs_frustum / collect_object_lights compiled by oracle/ref_lightcull_harness.cpp (Jolt declaration shim)."""

    def __init__(self, kind="port"):
        assert kind in ("port", "reference")
        path = PORT_LIB if kind == "port" else REF_LIGHTCULL_LIB
        if not os.path.exists(path):
            build(kind)
        self.kind, self.lib, self.prefix = kind, C.CDLL(path), ("shso_" if kind == "port" else "shsref_")

    def cull_objects(self, bounds, view_proj):
        """bounds: (n, 10) sphere + AABB for the port, (n, 6) AABBs for the reference (it derives the sphere itself).
        Returns classes (n,) uint8, visible indices, counts (tested, outside, intersecting, inside, visible)."""
        b = np.ascontiguousarray(bounds, dtype=np.float32).reshape(-1, 10 if self.kind == "port" else 6)
        vp = np.ascontiguousarray(view_proj, dtype=np.float32).reshape(16)
        classes, visible, counts = np.zeros(len(b), np.uint8), np.zeros(max(1, len(b)), np.uint32), np.zeros(5, np.uint32)
        rc = getattr(self.lib, self.prefix + "cull_objects")(capi.fptr(b), C.c_uint32(len(b)), capi.fptr(vp), classes.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                             capi.u32ptr(visible), capi.u32ptr(counts))
        assert rc == 0, rc
        return classes, visible[:int(counts[4])].copy(), counts

    def collect_object_lights(self, object_aabbs, visible, records, cull_mode):
        a = np.ascontiguousarray(object_aabbs, dtype=np.float32).reshape(-1, 6)
        v = np.ascontiguousarray(visible, dtype=np.uint32).reshape(-1)
        r = _records_u8(records)
        counts, idx, d2 = np.zeros(len(a), np.uint32), np.zeros((len(a), 8), np.uint32), np.zeros((len(a), 8), np.float32)
        rc = getattr(self.lib, self.prefix + "collect_object_lights")(capi.fptr(a), C.c_uint32(len(a)), capi.u32ptr(v), C.c_uint32(len(v)), r.ctypes.data_as(C.c_void_p),
                                                                      C.c_uint32(len(r)), C.c_int32(cull_mode), capi.u32ptr(counts), capi.u32ptr(idx), capi.fptr(d2))
        assert rc == 0, rc
        return counts, idx, d2

    def tile_depth_range_from_scene(self, object_aabbs, visible, view, view_proj, w, h, tile_size, z_near, z_far):
        a = np.ascontiguousarray(object_aabbs, dtype=np.float32).reshape(-1, 6)
        v = np.ascontiguousarray(visible, dtype=np.uint32).reshape(-1)
        mv, mvp = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (view, view_proj))
        tiles = ((w + tile_size - 1) // tile_size) * ((h + tile_size - 1) // tile_size)
        lo, hi = np.zeros(tiles, np.float32), np.zeros(tiles, np.float32)
        rc = getattr(self.lib, self.prefix + "tile_depth_range_from_scene")(capi.fptr(a), C.c_uint32(len(a)), capi.u32ptr(v), C.c_uint32(len(v)), capi.fptr(mv), capi.fptr(mvp),
                                                                            C.c_uint32(w), C.c_uint32(h), C.c_uint32(tile_size), C.c_float(z_near), C.c_float(z_far),
                                                                            capi.fptr(lo), capi.fptr(hi))
        assert rc == 0, rc
        return lo, hi

    def select_object_lights_from_bins(self, object_aabbs, view, view_proj, bins_xyz, clustered, z_near, z_far, bin_counts, bin_indices, records, cull_mode):
        """port / device-function flavour: bins in the capped device layout (counts[bins], indices[bins, max_per_bin])."""
        assert self.kind == "port"
        a = np.ascontiguousarray(object_aabbs, dtype=np.float32).reshape(-1, 6)
        mv, mvp = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (view, view_proj))
        bc = np.ascontiguousarray(bin_counts, dtype=np.uint32).reshape(-1)
        bi = np.ascontiguousarray(bin_indices, dtype=np.uint32).reshape(len(bc), -1)
        r = _records_u8(records)
        counts, idx, d2, cand = np.zeros(len(a), np.uint32), np.zeros((len(a), 8), np.uint32), np.zeros((len(a), 8), np.float32), np.zeros(len(a), np.uint32)
        rc = getattr(self.lib, self.prefix + "select_object_lights_from_bins")(
            capi.fptr(a), C.c_uint32(len(a)), capi.fptr(mv), capi.fptr(mvp), C.c_uint32(bins_xyz[0]), C.c_uint32(bins_xyz[1]), C.c_uint32(bins_xyz[2]), C.c_int32(int(clustered)),
            C.c_float(z_near), C.c_float(z_far), C.c_uint32(bi.shape[1]), capi.u32ptr(bc), capi.u32ptr(bi), r.ctypes.data_as(C.c_void_p), C.c_uint32(len(r)), C.c_int32(cull_mode),
            capi.u32ptr(counts), capi.u32ptr(idx), capi.fptr(d2), capi.u32ptr(cand))
        assert rc == 0, rc
        return counts, idx, d2, cand

    def reference_select_object_lights(self, object_aabbs, view, view_proj, w, h, culling_mode, tile_size, depth_slices, z_near, z_far, range_min, range_max, light_aabbs, records, cull_mode):
        """reference flavour: build_light_bin_culling + gather + collect_object_lights from the reference's own headers."""
        assert self.kind == "reference"
        a = np.ascontiguousarray(object_aabbs, dtype=np.float32).reshape(-1, 6)
        la = np.ascontiguousarray(light_aabbs, dtype=np.float32).reshape(-1, 6)
        mv, mvp = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (view, view_proj))
        lo = np.ascontiguousarray(range_min, dtype=np.float32).reshape(-1) if range_min is not None else None
        hi = np.ascontiguousarray(range_max, dtype=np.float32).reshape(-1) if range_max is not None else None
        r = _records_u8(records)
        counts, idx, d2, cand = np.zeros(len(a), np.uint32), np.zeros((len(a), 8), np.uint32), np.zeros((len(a), 8), np.float32), np.zeros(len(a), np.uint32)
        rc = self.lib.shsref_select_object_lights_from_bins(
            capi.fptr(a), C.c_uint32(len(a)), capi.fptr(mv), capi.fptr(mvp), C.c_uint32(w), C.c_uint32(h), C.c_int32(culling_mode), C.c_uint32(tile_size), C.c_uint32(depth_slices),
            C.c_float(z_near), C.c_float(z_far), capi.fptr(lo) if lo is not None else None, capi.fptr(hi) if hi is not None else None, C.c_uint32(len(lo) if lo is not None else 0),
            capi.fptr(la), r.ctypes.data_as(C.c_void_p), C.c_uint32(len(r)), C.c_int32(cull_mode), capi.u32ptr(counts), capi.u32ptr(idx), capi.fptr(d2), capi.u32ptr(cand))
        assert rc == 0, rc
        return counts, idx, d2, cand


REF_GLSL_A9_LIB = os.path.join(_HERE, "_ref", "libshs_glsl_a9_ref.so")


class LocalLightEvaluator:
    """Row A9 for ONE surface point through either checker:
      kind "glsl" -> oracle/_ref/libshs_glsl_a9_ref.so: the reference's fp_stress_scene.frag text compiled as C++
                     (oracle/extract_glsl_a9.py + oracle/ref_glsl_a9_harness.cpp),
      kind "port" -> oracle/liboracle.so: this repo's restatement (oracle.cpp eval_local_light / forward_plus_lights)."""

    def __init__(self, kind: str):
        self.kind = kind
        if kind == "glsl":
            if not os.path.exists(REF_GLSL_A9_LIB):
                build("reference")
            self.lib, self.pfx = C.CDLL(REF_GLSL_A9_LIB), "shsglsl_"
        else:
            if not os.path.exists(PORT_LIB):
                build("port")
            self.lib, self.pfx = C.CDLL(PORT_LIB), "shso_"
        self._att = getattr(self.lib, self.pfx + "attenuation_quadratic")
        self._att.restype = C.c_float

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_GLSL_A9_LIB) or os.path.isdir("/root/reference")

    @staticmethod
    def _f3(v):
        return (C.c_float * 3)(*[float(x) for x in v])

    def eval_local_light(self, records, idx, P, N, V, albedo, metallic, roughness, technique=0):
        r = _records_u8(records)
        out = (C.c_float * 3)()
        getattr(self.lib, self.pfx + "eval_local_light")(r.ctypes.data_as(C.c_void_p), C.c_uint32(idx), self._f3(P), self._f3(N), self._f3(V), self._f3(albedo),
                                                          C.c_float(metallic), C.c_float(roughness), C.c_uint32(technique), out)
        return np.array(out[:], dtype=np.float32)

    def attenuation(self, distance, rng, model, power, bias, cutoff):
        return np.float32(self._att(C.c_float(distance), C.c_float(rng), C.c_uint32(model), C.c_float(power), C.c_float(bias), C.c_float(cutoff)))

    def light_loop(self, records, counts, indices, tiles_x, tiles_y, max_per_tile, tile_size, px, py, H, P, N, V, albedo, metallic, roughness, technique=0, culling_mode=1):
        r = _records_u8(records)
        cnt = np.ascontiguousarray(counts, dtype=np.uint32).reshape(-1)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        out = (C.c_float * 3)()
        tail = (C.c_int32(px), C.c_int32(py), C.c_int32(H), self._f3(P), self._f3(N), self._f3(V), self._f3(albedo), C.c_float(metallic), C.c_float(roughness),
                C.c_uint32(technique), out)
        head = (r.ctypes.data_as(C.c_void_p), C.c_uint32(len(r)), capi.u32ptr(cnt), capi.u32ptr(idx), C.c_uint32(tiles_x), C.c_uint32(tiles_y), C.c_uint32(max_per_tile),
                C.c_uint32(tile_size))
        if self.kind == "glsl":
            self.lib.shsglsl_local_light_loop(*head, C.c_uint32(culling_mode), C.c_uint32(1), None, C.c_float(0.1), C.c_float(100.0), *tail)
        else:
            assert culling_mode == 1, "the restatement walks tile lists"
            self.lib.shso_local_light_loop(*head, *tail)
        return np.array(out[:], dtype=np.float32)


REF_OCCLUSION_LIB = os.path.join(_HERE, "_ref", "libshs_occlusion_ref.so")


class SoftwareOcclusion:
    """run_software_occlusion_pass (geometry/culling_software.hpp:253-333) through a checker:
      "reference" -> oracle/_ref/libshs_occlusion_ref.so (the reference's own header, oracle/ref_occlusion_harness.cpp)
      "port"      -> oracle/liboracle.so (oracle_scene_cull.cpp: shso_software_occlusion)
      a path      -> a library exporting the same signature under `prefix` (the g++ build of the device functions)."""

    def __init__(self, kind="port", prefix=None):
        if kind == "reference":
            if not os.path.exists(REF_OCCLUSION_LIB):
                build("reference")
            self.lib, self.prefix = C.CDLL(REF_OCCLUSION_LIB), "shsref_"
        elif kind == "port":
            if not os.path.exists(PORT_LIB):
                build("port")
            self.lib, self.prefix = C.CDLL(PORT_LIB), "shso_"
        else:
            self.lib, self.prefix = C.CDLL(kind), prefix

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_OCCLUSION_LIB) or os.path.isdir("/root/reference")

    def run(self, sc, enable=True):
        """sc: dict of aabbs (n, 6), visible, object_mesh (n,), models (n, 16), mesh_table (m, 3), vertices (v, 3), indices, view, view_proj,
        occ_w, occ_h, eps.  Returns occluded (n,) uint8, visible list, counts4, depth (occ_h, occ_w)."""
        return run_software_occlusion(getattr(self.lib, self.prefix + "software_occlusion"), None, sc, enable)


def run_software_occlusion(fn, ctx_handle, sc, enable=True):
    a = np.ascontiguousarray(sc["aabbs"], dtype=np.float32).reshape(-1, 6)
    vis = np.ascontiguousarray(sc["visible"], dtype=np.uint32).reshape(-1)
    om = np.ascontiguousarray(sc["object_mesh"], dtype=np.uint32).reshape(-1)
    mo = np.ascontiguousarray(sc["models"], dtype=np.float32).reshape(-1, 16)
    mt = np.ascontiguousarray(sc["mesh_table"], dtype=np.uint32).reshape(-1, 3)
    vt = np.ascontiguousarray(sc["vertices"], dtype=np.float32).reshape(-1, 3)
    ix = np.ascontiguousarray(sc["indices"], dtype=np.uint32).reshape(-1)
    v, vp = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (sc["view"], sc["view_proj"]))
    w, h = int(sc["occ_w"]), int(sc["occ_h"])
    occ, out_vis, counts, depth = np.zeros(max(1, len(a)), np.uint8), np.zeros(max(1, len(vis)), np.uint32), np.zeros(4, np.uint32), np.zeros((h, w), np.float32)
    args = [capi.fptr(a), C.c_uint32(len(a)), capi.u32ptr(vis), C.c_uint32(len(vis)), capi.u32ptr(om), capi.fptr(mo), capi.u32ptr(mt), C.c_uint32(len(mt)), capi.fptr(vt),
            C.c_uint32(len(vt)), capi.u32ptr(ix), C.c_uint32(len(ix)), capi.fptr(v), capi.fptr(vp), C.c_int32(w), C.c_int32(h), C.c_float(sc.get("eps", 1e-4)), C.c_int32(int(enable)),
            occ.ctypes.data_as(C.POINTER(C.c_uint8)), capi.u32ptr(out_vis), capi.u32ptr(counts), capi.fptr(depth)]
    rc = fn(*([ctx_handle] if ctx_handle is not None else []), *args)
    assert rc == 0, rc
    return occ[:len(a)], out_vis[:int(counts[2])].copy(), counts, depth


# ---------------------------------------------------------------- flat-shaded mesh draws (the consumer of the light selections)
REF_FLAT_DRAW_LIB = os.path.join(_HERE, "_ref", "libshs_flat_draw_ref.so")


class FlatDraw:
    """debug_draw::draw_mesh_blinn_phong_transformed (mode 0, sw_render/debug_draw.hpp:153-203) and draw_mesh_multi_light_transformed
    (mode 1, exp-plumbing/hello_light_types_culling_sw.cpp:366-422) over a batch of draws through a checker:
      "reference" -> oracle/_ref/libshs_flat_draw_ref.so (the reference's own text, oracle/ref_flat_draw_harness.cpp)
      "port"      -> oracle/liboracle.so (oracle_flat_draw.cpp: shso_flat_draw)
      a path      -> a library exporting the same signature under `prefix` (the g++ build of the device functions)."""

    def __init__(self, kind="port", prefix=None):
        if kind == "reference":
            if not os.path.exists(REF_FLAT_DRAW_LIB):
                build("reference")
            self.lib, self.prefix = C.CDLL(REF_FLAT_DRAW_LIB), "shsref_"
        elif kind == "port":
            if not os.path.exists(PORT_LIB):
                build("port")
            self.lib, self.prefix = C.CDLL(PORT_LIB), "shso_"
        else:
            self.lib, self.prefix = C.CDLL(kind), prefix

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_FLAT_DRAW_LIB) or os.path.isdir("/root/reference")

    def run(self, sc, mode):
        """sc: dict of draw_mesh (d,), models (d, 16), base (d, 3), sel_counts (d,), sel_idx (d, 8), mesh_table (m, 3), vertices (v, 3),
        indices, view_proj, camera, light_dir, lights (capi.LIGHT_PROPS_DTYPE), W, H, canvas (H, W, 4) uint8, depth (H, W) float32.
        Returns the canvas and the depth buffer after the batch."""
        dm = np.ascontiguousarray(sc["draw_mesh"], dtype=np.uint32).reshape(-1)
        mo = np.ascontiguousarray(sc["models"], dtype=np.float32).reshape(-1, 16)
        ba = np.ascontiguousarray(sc["base"], dtype=np.float32).reshape(-1, 3)
        scn = np.ascontiguousarray(sc["sel_counts"], dtype=np.uint32).reshape(-1)
        six = np.ascontiguousarray(sc["sel_idx"], dtype=np.uint32).reshape(-1, 8)
        mt = np.ascontiguousarray(sc["mesh_table"], dtype=np.uint32).reshape(-1, 3)
        vt = np.ascontiguousarray(sc["vertices"], dtype=np.float32).reshape(-1, 3)
        ix = np.ascontiguousarray(sc["indices"], dtype=np.uint32).reshape(-1)
        vp, cam, ld = (np.ascontiguousarray(sc[k], dtype=np.float32).reshape(-1) for k in ("view_proj", "camera", "light_dir"))
        li = np.ascontiguousarray(sc["lights"], dtype=capi.LIGHT_PROPS_DTYPE).reshape(-1)
        w, h = int(sc["W"]), int(sc["H"])
        canvas = np.ascontiguousarray(sc["canvas"], dtype=np.uint8).reshape(h, w, 4).copy()
        depth = np.ascontiguousarray(sc["depth"], dtype=np.float32).reshape(h, w).copy()
        fn = getattr(self.lib, self.prefix + "flat_draw")
        fn.restype = C.c_int32
        rc = fn(C.c_int32(int(mode)), C.c_uint32(len(dm)), capi.u32ptr(dm), capi.fptr(mo), capi.fptr(ba), capi.u32ptr(scn), capi.u32ptr(six), capi.u32ptr(mt), C.c_uint32(len(mt)),
                capi.fptr(vt), C.c_uint32(len(vt)), capi.u32ptr(ix), C.c_uint32(len(ix)), capi.fptr(vp), capi.fptr(cam), capi.fptr(ld), li.ctypes.data_as(C.c_void_p), C.c_uint32(len(li)),
                C.c_int32(w), C.c_int32(h), canvas.ctypes.data_as(C.POINTER(C.c_uint8)), capi.fptr(depth))
        assert rc == 0, rc
        return canvas, depth
