"""Deterministic synthetic scenes of the shapes BASELINE.json names.

A scene is plain host data (numpy arrays + the POD structs of include/shsb.h): the same bytes are
fed to the CUDA path, to the CPU oracle and to the compiled reference.  Camera matrices and light
records are INPUTS of the path, so they are built here once (numpy float32) and shared.

Hash RNG = the reference's own pseudo_random01 (exp-plumbing/hello_light_types_culling_sw.cpp:156-165).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import FrameParams, RenderItem, Scene

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def pseudo_random01(seed) -> np.ndarray:
    x = np.asarray(seed, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    m = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return ((x & np.uint64(0x00FFFFFF)).astype(np.float32) / np.float32(0x01000000)).astype(np.float32)


def load_suzanne() -> dict:
    """Suzanne (967 triangles), derived from the reference fixture by tests/golden/make_assets.py."""
    z = np.load(os.path.join(_ASSETS, "suzanne.npz"))
    return {k: z[k] for k in ("positions", "normals", "uvs", "indices")}


def make_grid_plane(extent=64.0, n=32) -> dict:
    """(n x n)-quad floor in the XZ plane, CCW seen from +Y in the LH/y-up convention."""
    xs = np.linspace(-extent / 2, extent / 2, n + 1, dtype=np.float32)
    gx, gz = np.meshgrid(xs, xs, indexing="xy")
    pos = np.stack([gx.ravel(), np.zeros_like(gx).ravel(), gz.ravel()], axis=1).astype(np.float32)
    nrm = np.tile(np.array([[0, 1, 0]], np.float32), (len(pos), 1))
    uv = np.stack([(gx.ravel() / extent + 0.5) * 8.0, (gz.ravel() / extent + 0.5) * 8.0], axis=1).astype(np.float32)
    idx = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            b, c, d = a + 1, a + (n + 1), a + (n + 1) + 1
            idx += [a, b, c, b, d, c]
    return {"positions": pos, "normals": nrm, "uvs": uv, "indices": np.asarray(idx, np.uint32)}


def make_checker_texture(size=64, seed=7) -> np.ndarray:
    y, x = np.mgrid[0:size, 0:size]
    r = (pseudo_random01((y * size + x + seed).astype(np.uint64)) * 60).astype(np.int32)
    base = np.where(((x // 8) + (y // 8)) % 2 == 0, 200, 90).astype(np.int32)
    t = np.zeros((size, size, 4), np.uint8)
    t[..., 0] = np.clip(base + r, 0, 255)
    t[..., 1] = np.clip(base - r // 2 + 20, 0, 255)
    t[..., 2] = np.clip(255 - base + r, 0, 255)
    t[..., 3] = 255
    return t


# ---------------------------------------------------------------- camera (inputs; numpy float32)
def _norm(v):
    v = np.asarray(v, np.float32)
    return v / np.float32(math.sqrt(float(np.dot(v, v))))


def look_at_lh(eye, target, up) -> np.ndarray:
    eye = np.asarray(eye, np.float32)
    f = _norm(np.asarray(target, np.float32) - eye)
    s = _norm(np.cross(np.asarray(up, np.float32), f))
    u = np.cross(f, s)
    m = np.eye(4, dtype=np.float32)  # m[row, col]
    m[0, :3], m[1, :3], m[2, :3] = s, u, f
    m[0, 3], m[1, 3], m[2, 3] = -np.dot(s, eye), -np.dot(u, eye), -np.dot(f, eye)
    return m


def perspective_lh_no(fovy, aspect, zn, zf) -> np.ndarray:
    t = math.tan(fovy / 2.0)
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = (zf + zn) / (zf - zn)
    m[3, 2] = 1.0
    m[2, 3] = -(2.0 * zf * zn) / (zf - zn)
    return m


def camera_viewproj(eye, target, up, fovy, aspect, zn, zf) -> np.ndarray:
    """Column-major float32[16] like glm::mat4 (proj * view)."""
    vp = (perspective_lh_no(fovy, aspect, zn, zf) @ look_at_lh(eye, target, up)).astype(np.float32)
    return np.ascontiguousarray(vp.T).reshape(16)


# ---------------------------------------------------------------- lights (CullingLightGPU, 160 B)
LIGHT_DTYPE = np.dtype([
    ("position_range", np.float32, 4), ("color_intensity", np.float32, 4), ("direction_spot", np.float32, 4),
    ("axis_spot_outer", np.float32, 4), ("up_shape_x", np.float32, 4), ("shape_attenuation", np.float32, 4),
    ("type_shape_flags", np.uint32, 4), ("cull_sphere", np.float32, 4), ("cull_aabb_min", np.float32, 4),
    ("cull_aabb_max", np.float32, 4)])
assert LIGHT_DTYPE.itemsize == capi.LIGHT_RECORD_BYTES

LIGHT_POINT, LIGHT_SPOT = 1, 2
ATTEN_LINEAR, ATTEN_SMOOTH, ATTEN_INVERSE_SQUARE = 0, 1, 2


def pack_lights(pos, rng, color, intensity, kind, direction=None, inner=None, outer=None,
                atten_model=ATTEN_SMOOTH, atten_power=1.25, atten_bias=0.05, atten_cutoff=0.0, jolt_bounds=False) -> np.ndarray:
    """Layout of make_point_culling_light / make_spot_culling_light (lighting/light_types.hpp:327-379)."""
    n = len(pos)
    r = np.zeros(n, LIGHT_DTYPE)
    pos = np.asarray(pos, np.float32)
    rng = np.maximum(np.asarray(rng, np.float32), 0)
    kind = np.broadcast_to(np.asarray(kind, np.uint32), (n,))
    r["position_range"][:, :3] = pos
    r["position_range"][:, 3] = rng
    r["color_intensity"][:, :3] = np.maximum(np.asarray(color, np.float32), 0)
    r["color_intensity"][:, 3] = np.maximum(np.asarray(intensity, np.float32), 0)
    r["direction_spot"][:] = (0, -1, 0, 1)
    r["axis_spot_outer"][:] = (1, 0, 0, 0)
    r["up_shape_x"][:] = (0, 1, 0, 0)
    r["shape_attenuation"][:] = (0, max(atten_power, 0.001), max(atten_bias, 1e-5), max(atten_cutoff, 0.0))
    r["type_shape_flags"][:, 0] = kind
    r["type_shape_flags"][:, 1] = kind  # Sphere = 1 for points, Cone = 2 for spots
    r["type_shape_flags"][:, 2] = 7     # LightFlagsDefault
    r["type_shape_flags"][:, 3] = atten_model
    spot = kind == LIGHT_SPOT
    if spot.any():
        d = np.asarray(direction, np.float32)
        d = d / np.sqrt((d * d).sum(axis=1, keepdims=True)).astype(np.float32)
        r["direction_spot"][spot, :3] = d[spot]
        r["direction_spot"][spot, 3] = np.cos(np.asarray(inner, np.float32))[spot]
        r["axis_spot_outer"][spot, 3] = np.cos(np.asarray(outer, np.float32))[spot]
    r["cull_sphere"][:, :3] = pos
    r["cull_sphere"][:, 3] = rng * np.float32(math.sqrt(3.0)) if jolt_bounds else rng
    r["cull_aabb_min"][:, :3] = pos - rng[:, None]
    r["cull_aabb_max"][:, :3] = pos + rng[:, None]
    r["cull_aabb_min"][:, 3] = 1
    r["cull_aabb_max"][:, 3] = 1
    return r


def make_lights(n_point, n_spot, lo, hi, seed=0, range_lo=2.0, range_hi=5.0, jolt_bounds=False) -> np.ndarray:
    """Local lights uniform in the box [lo, hi]; ranges / intensities / cone angles in the spirit of
    exp-plumbing/hello_light_types_culling_sw.cpp:600-627."""
    n = n_point + n_spot
    i = np.arange(n, dtype=np.uint64) + np.uint64(seed * 7919)
    r = [pseudo_random01(i * np.uint64(k) + np.uint64(c)) for k, c in
         ((747796405, 13), (2891336453, 17), (1181783497, 19), (2246822519, 23), (3266489917, 29), (668265263, 31), (374761393, 37), (2654435761, 41))]
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    pos = np.stack([lo[k] + (hi[k] - lo[k]) * r[k] for k in range(3)], axis=1).astype(np.float32)
    rng = (range_lo + (range_hi - range_lo) * r[3]).astype(np.float32)
    palette = np.array([[1.0, 0.55, 0.30], [0.35, 0.70, 1.0], [0.60, 1.0, 0.45], [1.0, 0.35, 0.75], [1.0, 0.90, 0.50]], np.float32)
    color = palette[(np.arange(n) * 3) % len(palette)] * (0.82 + 0.30 * r[4])[:, None]
    intensity = (2.0 + 1.0 * r[5]).astype(np.float32)
    kind = np.where(np.arange(n) < n_point, LIGHT_POINT, LIGHT_SPOT).astype(np.uint32)
    direction = np.stack([r[6] * 2 - 1, -np.ones(n, np.float32) * 1.5, r[7] * 2 - 1], axis=1).astype(np.float32)
    inner = np.radians(12.0 + 8.0 * r[6]).astype(np.float32)
    outer = (inner + np.radians(8.0 + 12.0 * r[7])).astype(np.float32)
    return pack_lights(pos, rng, color.astype(np.float32), intensity, kind, direction, inner, outer, jolt_bounds=jolt_bounds)


# ---------------------------------------------------------------- scene container
@dataclass
class SceneData:
    name: str
    w: int
    h: int
    zn: float
    zf: float
    meshes: list
    textures: list
    items: list          # list of dicts: pos, rot, scl, mesh (1-based), material or None
    cam_pos: tuple
    cam_target: tuple
    fovy: float
    sun_dir: tuple
    sun_color: tuple
    sun_intensity: float
    fp: FrameParams = field(default_factory=capi.default_frame_params)
    lights: np.ndarray | None = None
    aspect: float | None = None
    shadow_size: int = 0
    viewproj: np.ndarray | None = None
    prev_viewproj: np.ndarray | None = None   # Camera::prev_viewproj (defaults to viewproj)
    sky: dict | None = None                    # {"kind": "procedural", "sun_dir": (x, y, z)} | {"kind": "cubemap", "faces": [6 texture handles], "intensity": f}

    def __post_init__(self):
        if self.viewproj is None:
            asp = self.aspect if self.aspect is not None else self.w / self.h
            self.viewproj = camera_viewproj(self.cam_pos, self.cam_target, (0, 1, 0), self.fovy, asp, self.zn, self.zf)
        self._items_c = (RenderItem * max(1, len(self.items)))()
        for i, it in enumerate(self.items):
            ri = self._items_c[i]
            capi.set_f(ri.tr.pos, it["pos"])
            capi.set_f(ri.tr.rot_euler, it.get("rot", (0, 0, 0)))
            capi.set_f(ri.tr.scl, it.get("scl", (1, 1, 1)))
            ri.mesh = it["mesh"]
            mat = it.get("material")
            ri.has_material = 1 if mat else 0
            if mat:
                capi.set_f(ri.base_color, mat["base_color"])
                ri.metallic, ri.roughness, ri.ao = mat["metallic"], mat["roughness"], mat.get("ao", 1.0)
                ri.base_color_tex = mat.get("tex", 0)
            ri.casts_shadow = 1 if it.get("casts_shadow", True) else 0
            ri.visible = 1 if it.get("visible", True) else 0
            ri.object_id = int(it.get("object_id", 0))
        self.scene = Scene()
        capi.set_f(self.scene.cam_viewproj, self.viewproj)
        capi.set_f(self.scene.cam_pos, self.cam_pos)
        capi.set_f(self.scene.sun_dir_ws, self.sun_dir)
        capi.set_f(self.scene.sun_color, self.sun_color)
        self.scene.sun_intensity = self.sun_intensity
        self.scene.n_items = len(self.items)
        self.scene.items = C.cast(self._items_c, C.POINTER(RenderItem))
        capi.set_f(self.scene.cam_prev_viewproj, self.prev_viewproj if self.prev_viewproj is not None else self.viewproj)
        if self.sky:
            if self.sky["kind"] == "procedural":
                self.scene.sky_kind = capi.SKY_PROCEDURAL
                capi.set_f(self.scene.sky_sun_dir_ws, self.sky.get("sun_dir", (0.4668, -0.3487, 0.8127)))
            else:
                self.scene.sky_kind = capi.SKY_CUBEMAP
                self.scene.sky_intensity = float(self.sky.get("intensity", 1.0))
                for k, t in enumerate(self.sky["faces"]):
                    self.scene.sky_faces[k] = int(t)

    @property
    def n_triangles(self) -> int:
        return int(sum(len(self.meshes[it["mesh"] - 1]["indices"]) // 3 for it in self.items if it.get("visible", True)))

    def with_camera(self, cam_pos, cam_target):
        return SceneData(self.name, self.w, self.h, self.zn, self.zf, self.meshes, self.textures, self.items, tuple(cam_pos),
                         tuple(cam_target), self.fovy, self.sun_dir, self.sun_color, self.sun_intensity, self.fp, self.lights,
                         self.aspect, self.shadow_size, sky=self.sky)

    def models(self, oracle) -> np.ndarray:
        """(n_items, 16) model matrices in items order (for Context::history in the CPU checkers)."""
        if not self.items:
            return np.zeros((0, 16), np.float32)
        return np.stack([oracle.model_from_transform(it["pos"], it.get("rot", (0, 0, 0)), it.get("scl", (1, 1, 1))) for it in self.items])

    def moved(self, dpos=(0.3, 0.0, -0.2), drot=(0.0, 0.15, 0.05), cam_pos=None, cam_target=None):
        """The next frame of an animation: every k-th item translated / rotated a little, optionally a new camera whose
        prev_viewproj is this frame's viewproj."""
        items = []
        for k, it in enumerate(self.items):
            it2 = dict(it)
            if k % 2 == 1:
                it2["pos"] = tuple(float(a) + float(b) * (1 + 0.25 * k) for a, b in zip(it["pos"], dpos))
                it2["rot"] = tuple(float(a) + float(b) for a, b in zip(it.get("rot", (0, 0, 0)), drot))
            items.append(it2)
        return SceneData(self.name + "_next", self.w, self.h, self.zn, self.zf, self.meshes, self.textures, items,
                         tuple(cam_pos) if cam_pos is not None else self.cam_pos, tuple(cam_target) if cam_target is not None else self.cam_target,
                         self.fovy, self.sun_dir, self.sun_color, self.sun_intensity, self.fp, self.lights, self.aspect, self.shadow_size,
                         prev_viewproj=self.viewproj, sky=self.sky)


_SUN_DIR = tuple(_norm((-0.35, -1.0, -0.25)))  # exp-plumbing/hello_pass_basics.cpp:669-671
_SUN_COLOR = (1.0, 0.97, 0.92)
_MATERIALS = [
    {"base_color": (240 / 255, 195 / 255, 75 / 255), "metallic": 0.95, "roughness": 0.20},  # gold (hello_pass_basics.cpp:703)
    {"base_color": (0.42, 0.44, 0.48), "metallic": 0.0, "roughness": 0.96},                 # plastic
    {"base_color": (0.80, 0.25, 0.20), "metallic": 0.28, "roughness": 0.44},
    {"base_color": (0.25, 0.55, 0.85), "metallic": 0.60, "roughness": 0.35},
]


def scene_c1(w=640, h=480) -> SceneData:
    """BASELINE config 1: one Suzanne, Blinn-Phong, 1 directional light, z-buffer, 640x480
    (placement per hello-3d-primitives/hello_pipeline_blinn_phong_shading.cpp:152-153,384; aspect 4/3)."""
    fp = capi.default_frame_params(shading_model=capi.SHADING_BLINN, shadow_enable=0)
    return SceneData("C1_suzanne_blinn", w, h, 0.1, 1000.0, [load_suzanne()], [],
                     [{"pos": (0, 0, 10), "rot": (0, math.pi, 0), "scl": (4, 4, 4), "mesh": 1,
                       "material": {"base_color": (60 / 255, 100 / 255, 200 / 255), "metallic": 0.0, "roughness": 0.5}}],
                     (0, 5, -20), (0, 5, -19), math.radians(60.0), tuple(_norm((-1, -0.4, 1))), (1, 1, 1), 1.0, fp, aspect=4 / 3)


def _suzanne_grid(nx, nz, spacing, scale=1.0, mesh=1, y=0.0):
    items = []
    for j in range(nz):
        for i in range(nx):
            k = j * nx + i
            rot = float(2.0 * math.pi * pseudo_random01(np.uint64(k * 2654435761 + 101)))
            items.append({"pos": ((i - (nx - 1) / 2) * spacing, y, (j - (nz - 1) / 2) * spacing), "rot": (0, rot, 0),
                          "scl": (scale, scale, scale), "mesh": mesh, "material": _MATERIALS[k % len(_MATERIALS)]})
    return items


def scene_c2(w=1920, h=1080, grid=10, n_point=768, n_spot=256, seed=0) -> SceneData:
    """BASELINE config 2: 1080p Forward+, grid x grid Suzanne instances (~100k tris), 1024 point/spot
    lights, 16-px tiles, <=128 lights per tile (frame/frame_params.hpp:83-84)."""
    items = _suzanne_grid(grid, grid, 3.0)
    ext = (grid - 1) * 3.0 / 2 + 1.5
    lights = make_lights(n_point, n_spot, (-ext, 0.5, -ext), (ext, 3.0, ext), seed=seed)
    fp = capi.default_frame_params(shading_model=capi.SHADING_PBR, shadow_enable=0, light_culling=1, tile_size=16, max_lights_per_tile=128)
    return SceneData(f"C2_forward_plus_{w}x{h}_{grid * grid}inst_{n_point + n_spot}lights", w, h, 0.1, 200.0, [load_suzanne()], [], items,
                     (0, 12, -28), (0, 0, 0), math.radians(60.0), _SUN_DIR, _SUN_COLOR, 2.2, fp, lights)


def scene_c3(w=2560, h=1440, nx=32, nz=33, shadow_size=4096) -> SceneData:
    """BASELINE config 3: shadow-map depth pass (4096^2) + PCF lit pass at 1440p, ~1M triangles."""
    items = [{"pos": (0, -1.0, 0), "mesh": 2, "material": _MATERIALS[1], "casts_shadow": False}] + _suzanne_grid(nx, nz, 3.0)
    fp = capi.default_frame_params(shading_model=capi.SHADING_PBR, shadow_enable=1)
    ext = max(nx, nz) * 3.0
    return SceneData(f"C3_shadow_{w}x{h}_{nx * nz}inst", w, h, 0.1, 400.0, [load_suzanne(), make_grid_plane(ext * 1.2, 64)], [], items,
                     (0, ext * 0.35, -ext * 0.75), (0, 0, 0), math.radians(60.0), _SUN_DIR, _SUN_COLOR, 2.2, fp, shadow_size=shadow_size)


def scene_c4(w=3840, h=2160, nx=44, nz=47) -> SceneData:
    """BASELINE config 4 (library flavour): PBR-MR with eval_fake_ibl and an albedo texture at 4K, ~2M triangles.
    Normal mapping has no reference semantics and the cubemap skybox is a 'next' row (8f-3)."""
    mats = [dict(m, tex=1) for m in _MATERIALS]
    items = _suzanne_grid(nx, nz, 3.0)
    for k, it in enumerate(items):
        it["material"] = mats[k % len(mats)]
    fp = capi.default_frame_params(shading_model=capi.SHADING_PBR, shadow_enable=0)
    ext = max(nx, nz) * 3.0
    return SceneData(f"C4_pbr_{w}x{h}_{nx * nz}inst", w, h, 0.1, 600.0, [load_suzanne()], [make_checker_texture(256)], items,
                     (0, ext * 0.35, -ext * 0.75), (0, 0, 0), math.radians(60.0), _SUN_DIR, _SUN_COLOR, 2.2, fp)


def scene_c5(w=7680, h=4320, nx=101, nz=102, n_point=768, n_spot=256) -> SceneData:
    """BASELINE config 5: 8K Forward+, ~10M triangles (sort-first split across GPUs)."""
    items = _suzanne_grid(nx, nz, 3.0)
    ext = max(nx, nz) * 3.0 / 2
    lights = make_lights(n_point, n_spot, (-ext, 0.5, -ext), (ext, 3.0, ext), seed=5, range_lo=4.0, range_hi=10.0)
    fp = capi.default_frame_params(shading_model=capi.SHADING_PBR, shadow_enable=0, light_culling=1)
    return SceneData(f"C5_8k_{nx * nz}inst", w, h, 0.1, 1000.0, [load_suzanne()], [], items,
                     (0, ext * 0.7, -ext * 1.5), (0, 0, 0), math.radians(60.0), _SUN_DIR, _SUN_COLOR, 2.2, fp, lights)


def camera_ring(scene: SceneData, n_cameras: int, radius=28.0, height=12.0):
    """64-camera batch of config 5b: cameras on a circle looking at the origin."""
    out = []
    for c in range(n_cameras):
        a = 2.0 * math.pi * c / n_cameras
        out.append(scene.with_camera((radius * math.sin(a), height, -radius * math.cos(a)), (0, 0, 0)))
    return out


def make_sky_faces(size=16, seed=3) -> list:
    """Six small RGBA faces with distinct hues and a gradient (a stand-in for the reference's 2048^2 skybox PNGs)."""
    faces = []
    y, x = np.mgrid[0:size, 0:size]
    for f in range(6):
        t = np.zeros((size, size, 4), np.uint8)
        r = (pseudo_random01((y * size + x + seed * 977 + f * 131).astype(np.uint64)) * 40).astype(np.int32)
        t[..., 0] = np.clip(40 + 35 * f + 4 * x + r, 0, 255)
        t[..., 1] = np.clip(200 - 25 * f + 3 * y - r, 0, 255)
        t[..., 2] = np.clip(90 + 20 * ((f * 3) % 6) + 2 * (x + y), 0, 255)
        t[..., 3] = 255
        faces.append(t)
    return faces


def scene_small(w=160, h=120, shading=capi.SHADING_PBR, n_inst=3, lights=0, tex=False, seed=1, near_clip=False, sky=None, motion=False) -> SceneData:
    """Small parity scene: a few Suzannes + floor; optional texture, lights and a camera that forces frustum clipping."""
    meshes = [load_suzanne(), make_grid_plane(24.0, 8)]
    textures = [make_checker_texture(32)] if tex else []
    sky_desc = None
    if sky == "procedural":
        sky_desc = {"kind": "procedural", "sun_dir": (0.2, -0.35, 0.9)}
    elif sky == "cubemap":
        first = len(textures) + 1
        textures = textures + make_sky_faces()
        sky_desc = {"kind": "cubemap", "faces": list(range(first, first + 6)), "intensity": 1.25}
    items = [{"pos": (0, -1.0, 0), "mesh": 2, "material": dict(_MATERIALS[1], tex=1 if tex else 0), "casts_shadow": False, "object_id": 1000}]
    for k in range(n_inst):
        r = pseudo_random01(np.arange(4, dtype=np.uint64) + np.uint64(seed * 131 + k * 17))
        items.append({"pos": (float(r[0] * 8 - 4), float(r[1] * 1.5), float(r[2] * 8 - 4)), "rot": (0.1 * k, float(r[3] * 6.28), 0.05 * k),
                      "scl": (1.0 + 0.3 * k, 1.0 + 0.3 * k, 1.0 + 0.3 * k), "mesh": 1,
                      "material": (dict(_MATERIALS[k % 4], tex=1) if (tex and k % 2 == 0) else (_MATERIALS[k % 4] if k % 3 else None)),
                      "object_id": 1001 + k})
    lt = make_lights(max(1, lights * 3 // 4), lights - max(1, lights * 3 // 4), (-6, 0.2, -6), (6, 3.0, 6), seed=seed) if lights else None
    fp = capi.default_frame_params(shading_model=shading, shadow_enable=0, light_culling=1 if lights else 0, motion_vectors_enable=1 if motion else 0)
    cam = (0.5, 1.2, -2.2) if near_clip else (0, 4, -8)
    return SceneData(f"small_{w}x{h}", w, h, 0.1, 100.0, meshes, textures, items, cam, (0, 0.5, 0), math.radians(60.0),
                     _SUN_DIR, _SUN_COLOR, 2.2, fp, lt, shadow_size=256, sky=sky_desc)


def scene_ragged(w=150, h=110, shading=capi.SHADING_PBR, shadow=False, nonfinite=True) -> SceneData:
    """Ragged / hostile inputs the reference's triangle loop accepts (sw_render/rasterizer.hpp:196-230, 260-290):
    a NON-indexed mesh; normal / uv arrays shorter than the position array (defaults (0,1,0) / (0,0)); indices beyond
    the position array (triangle counted in tri_input, then skipped); zero-area and repeated-vertex triangles; a vertex
    at infinity and a NaN vertex (non-finite reject); an invisible item; an item without a material; a zero-scale item
    (singular normal matrix: builtin_shaders.hpp:93-96 falls back to mat3(model)); an item that casts no shadow."""
    rng = np.random.default_rng(12)
    # mesh 1: non-indexed soup of 40 triangles around the origin, normals for the first 70 vertices only, uvs for 50
    pos = (rng.random((120, 3), dtype=np.float32) - 0.5) * np.float32(3.0)
    pos[9] = pos[10]                                   # repeated vertex -> zero area
    pos[12:15] = pos[12]                               # a point
    if nonfinite:                                      # (they also poison the scene AABB of PassShadowMap: keep them out of shadow cases)
        pos[30] = (np.float32(np.inf), 0.0, 0.0)       # vertex at infinity
        pos[45] = (np.float32(np.nan), 1.0, 1.0)       # NaN vertex
    nrm = rng.random((70, 3), dtype=np.float32) - 0.5
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    soup = {"positions": pos, "normals": nrm.astype(np.float32), "uvs": rng.random((50, 2), dtype=np.float32), "indices": np.zeros((0,), np.uint32)}
    # mesh 2: indexed quad grid with some indices out of range and a trailing partial triangle
    grid = make_grid_plane(6.0, 4)
    idx = grid["indices"].copy()
    idx[7] = 10_000                                    # beyond positions.size(): triangle skipped after tri_input++
    idx[20] = len(grid["positions"])                   # == size: also out of range
    idx = np.concatenate([idx, np.array([0, 1], np.uint32)])   # 2 dangling indices: indices.size() / 3 truncates
    ragged_grid = dict(grid, indices=idx, uvs=grid["uvs"][:5])
    meshes = [soup, ragged_grid, load_suzanne()]
    textures = [make_checker_texture(16)]
    items = [
        {"pos": (0, -0.8, 0), "mesh": 2, "material": dict(_MATERIALS[1], tex=1), "casts_shadow": False, "object_id": 1},
        {"pos": (-0.5, 0.6, 0.5), "rot": (0.3, 0.8, 0.1), "scl": (1.2, 0.9, 1.1), "mesh": 1, "material": dict(_MATERIALS[0], tex=1), "object_id": 2},
        {"pos": (1.8, 0.4, 0.2), "rot": (0.0, 2.5, 0.0), "mesh": 3, "material": None, "object_id": 3},                     # no material: reference defaults
        {"pos": (-2.0, 0.5, 1.0), "mesh": 3, "material": _MATERIALS[2], "visible": False, "object_id": 4},               # invisible
        {"pos": (0.2, 1.4, -0.4), "scl": (1.0, 0.0, 1.0), "mesh": 3, "material": _MATERIALS[3], "object_id": 5},          # flattened: det(model) = 0
        {"pos": (0.0, 0.2, 1.5), "scl": (0.0, 0.0, 0.0), "mesh": 3, "material": _MATERIALS[3], "object_id": 6},           # collapsed to a point
    ]
    fp = capi.default_frame_params(shading_model=shading, shadow_enable=1 if shadow else 0, light_culling=0, motion_vectors_enable=0)
    return SceneData(f"ragged_{w}x{h}", w, h, 0.1, 60.0, meshes, textures, items, (0.3, 2.2, -4.5), (0, 0.4, 0), math.radians(55.0),
                     _SUN_DIR, _SUN_COLOR, 2.0, fp, None, shadow_size=160)


def scene_adversarial(w=96, h=64, shading=capi.SHADING_PBR) -> SceneData:
    """Geometry chosen to sit ON the decision boundaries of the reference's raster loop (sw_render/rasterizer.hpp:260-361),
    drawn with viewproj = identity (clip = world position, w = 1) so that screen coordinates are under direct control:
      * a fan of triangles whose shared edges and vertices pass exactly through pixel centres (the reference has no fill
        rule: bc >= 0 on both sides, :336-338 -- shared-edge pixels are covered twice and the depth tie goes to the
        earlier triangle, :359);
      * two coplanar quads at the same depth with different materials (tie -> earlier draw);
      * a triangle far larger than the frame (clipped by all four side planes, :232-250);
      * a cloud of sub-pixel triangles, most of which cover no sample;
      * a triangle with one vertex exactly on w = 0 and one behind the camera (perspective matrix item), :111-164;
      * a sliver whose |area| is just above / below the 1e-10 reject (:273-274)."""
    def sx(px):  # NDC x that maps to screen x = px
        return np.float32(2.0 * px / (w - 1) - 1.0)

    def sy(py):
        return np.float32(2.0 * py / (h - 1) - 1.0)

    meshes = []
    # 1. fan around a pixel centre, rim vertices on pixel centres too
    cx, cy = 20.5, 30.5
    rim = [(cx + 8, cy), (cx + 8, cy + 8), (cx, cy + 8), (cx - 8, cy + 8), (cx - 8, cy), (cx - 8, cy - 8), (cx, cy - 8), (cx + 8, cy - 8)]
    pos = [(sx(cx), sy(cy), 0.2)] + [(sx(x), sy(y), 0.2) for x, y in rim]
    idx = []
    for k in range(8):
        idx += [0, 1 + k, 1 + (k + 1) % 8]
    meshes.append({"positions": np.array(pos, np.float32), "normals": np.tile(np.array([[0, 0, -1]], np.float32), (9, 1)),
                   "uvs": np.zeros((9, 2), np.float32), "indices": np.array(idx, np.uint32)})
    # 2. coplanar quads (two meshes, same depth, overlapping)
    def quad(x0, y0, x1, y1, z):
        p = np.array([(sx(x0), sy(y0), z), (sx(x1), sy(y0), z), (sx(x1), sy(y1), z), (sx(x0), sy(y1), z)], np.float32)
        return {"positions": p, "normals": np.tile(np.array([[0, 0, -1]], np.float32), (4, 1)), "uvs": np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32),
                "indices": np.array([0, 1, 2, 0, 2, 3], np.uint32)}
    meshes.append(quad(40.0, 10.0, 70.0, 40.0, 0.4))
    meshes.append(quad(55.5, 25.5, 90.5, 55.5, 0.4))
    # 3. far larger than the frame
    meshes.append({"positions": np.array([(-40.0, -30.0, 0.9), (50.0, -35.0, 0.9), (3.0, 60.0, 0.9)], np.float32),
                   "normals": np.tile(np.array([[0, 0, -1]], np.float32), (3, 1)), "uvs": np.zeros((3, 2), np.float32), "indices": np.array([0, 1, 2], np.uint32)})
    # 4. sub-pixel cloud
    rng = np.random.default_rng(5)
    c = np.stack([rng.uniform(2, w - 3, 600), rng.uniform(2, h - 3, 600)], axis=1)
    tri = np.repeat(c, 3, axis=0) + rng.uniform(-0.6, 0.6, (1800, 2))
    p4 = np.stack([sx(tri[:, 0]), sy(tri[:, 1]), np.full(1800, 0.1, np.float32)], axis=1).astype(np.float32)
    meshes.append({"positions": p4, "normals": np.tile(np.array([[0, 0, -1]], np.float32), (1800, 1)), "uvs": np.zeros((1800, 2), np.float32),
                   "indices": np.arange(1800, dtype=np.uint32)})
    # 6. slivers around the area threshold (screen-space area2 ~ 1e-10 needs NDC offsets ~1e-7)
    base = np.array([(sx(10.0), sy(50.0), 0.3), (sx(30.0), sy(50.0), 0.3)], np.float32)
    sl = []
    for eps in (0.0, 1e-9, 1e-7, 1e-5):
        sl += [base[0], base[1], (np.float32(base[0][0] + 0.3), np.float32(base[0][1] + eps), np.float32(0.3))]
    sl = np.array(sl, np.float32)
    meshes.append({"positions": sl, "normals": np.tile(np.array([[0, 0, -1]], np.float32), (len(sl), 1)), "uvs": np.zeros((len(sl), 2), np.float32),
                   "indices": np.arange(len(sl), dtype=np.uint32)})
    items = [{"pos": (0, 0, 0), "mesh": k + 1, "material": _MATERIALS[k % 4], "object_id": 10 + k} for k in range(len(meshes))]
    fp = capi.default_frame_params(shading_model=shading, shadow_enable=0, light_culling=0, motion_vectors_enable=0, cull_mode=capi.CULL_NONE)
    sd = SceneData(f"adversarial_{w}x{h}", w, h, 0.1, 10.0, meshes, [], items, (0.0, 0.0, -3.0), (0, 0, 0), math.radians(60.0),
                   _SUN_DIR, _SUN_COLOR, 2.0, fp, None, shadow_size=0, viewproj=np.eye(4, dtype=np.float32).reshape(16))
    return sd


def scene_w_zero(w=120, h=80) -> SceneData:
    """A real perspective camera with triangles that have a vertex exactly ON the camera plane (clip.w = 0) and behind it:
    the near-plane clip's t = da / (da - db) and the |da - db| <= 1e-8 skip (rasterizer.hpp:133-138) decide the topology."""
    eye = (0.0, 0.0, -2.0)
    p = np.array([(-1.0, -0.5, 1.0), (1.0, -0.5, 1.0), (0.0, 0.8, -2.0),        # apex exactly on the camera plane (view z = 0)
                  (-1.5, 0.2, 0.5), (-0.5, 0.9, -3.5), (-0.2, -0.1, 0.5),        # apex behind the camera
                  (0.3, 0.1, -1.9), (1.2, 0.2, -1.9), (0.8, 0.9, -1.9),          # whole triangle 0.1 in front of the eye (zn = 0.1: on the near plane)
                  (0.3, -0.9, -1.95), (1.2, -0.8, -1.95), (0.8, -0.2, 3.0)], np.float32)  # crosses the near plane
    mesh = {"positions": p, "normals": np.tile(np.array([[0, 0, -1]], np.float32), (len(p), 1)), "uvs": np.zeros((len(p), 2), np.float32),
            "indices": np.arange(len(p), dtype=np.uint32)}
    items = [{"pos": (0, 0, 0), "mesh": 1, "material": _MATERIALS[2], "object_id": 1}]
    fp = capi.default_frame_params(shading_model=capi.SHADING_BLINN, shadow_enable=0, light_culling=0, motion_vectors_enable=0, cull_mode=capi.CULL_NONE)
    return SceneData(f"w_zero_{w}x{h}", w, h, 0.1, 50.0, [mesh], [], items, eye, (0, 0, 1.0), math.radians(70.0), _SUN_DIR, _SUN_COLOR, 2.0, fp, None)


def scene_mixed_lights(records_u8, w=288, h=180, shading=capi.SHADING_PBR, max_per_tile=128) -> SceneData:
    """The small parity scene lit by a caller-supplied set of CullingLightGPU records (tests/golden/golden_area_lights.npz:
    point lights with all three attenuation models, spot, rect-area and tube-area lights, disabled / zero lights)."""
    sd = scene_small(w=w, h=h, shading=shading, n_inst=4, lights=0, tex=True, seed=8)
    sd.lights = np.ascontiguousarray(records_u8, dtype=np.uint8).reshape(-1, capi.LIGHT_RECORD_BYTES).view(LIGHT_DTYPE).reshape(-1)
    sd.fp.light_culling = 1
    sd.fp.max_lights_per_tile = max_per_tile
    return sd


def scene_odd_texture(w=140, h=100, shading=capi.SHADING_PBR) -> SceneData:
    """Texture sampling corners (shader/builtin_shaders.hpp:25-55): a 5 x 3 and a 1 x 1 texture, uvs far outside [0, 1] and negative
    (repeat wrap via floor), bilinear taps that straddle the wrap seam."""
    rng = np.random.default_rng(21)
    t53 = rng.integers(0, 256, size=(3, 5, 4)).astype(np.uint8)
    t53[..., 3] = 255
    t11 = np.array([[[200, 40, 90, 255]]], np.uint8)
    plane = make_grid_plane(10.0, 6)
    plane["uvs"] = ((plane["uvs"] / 8.0) * np.float32(6.0) - np.float32(2.3)).astype(np.float32)     # uv in [-2.3, 3.7]
    suz = load_suzanne()
    suz = dict(suz, uvs=(suz["uvs"] * np.float32(-3.0) + np.float32(0.5)).astype(np.float32))        # negative, scaled
    items = [{"pos": (0, -1.0, 0), "mesh": 1, "material": dict(_MATERIALS[1], tex=1), "object_id": 1},
             {"pos": (-1.2, 0.2, 0.5), "rot": (0.2, 0.9, 0.0), "mesh": 2, "material": dict(_MATERIALS[0], tex=1), "object_id": 2},
             {"pos": (1.4, 0.1, 0.0), "rot": (0.0, -0.7, 0.1), "mesh": 2, "material": dict(_MATERIALS[3], tex=2), "object_id": 3}]
    fp = capi.default_frame_params(shading_model=shading, shadow_enable=0, light_culling=0, motion_vectors_enable=0)
    return SceneData(f"odd_texture_{w}x{h}", w, h, 0.1, 60.0, [plane, suz], [t53, t11], items, (0.2, 2.0, -4.2), (0, 0.2, 0), math.radians(55.0),
                     _SUN_DIR, _SUN_COLOR, 2.2, fp, None)


def scene_ndc_depth(w=150, h=96) -> SceneData:
    """A depth target whose zf <= zn + 1e-6: the rasteriser falls back from linear view depth to NDC z * 0.5 + 0.5
    (sw_render/rasterizer.hpp:354-357).  The camera keeps a normal near / far; only the RENDER TARGET's zn / zf are equal."""
    sd = scene_small(w=w, h=h, near_clip=True, tex=True, seed=7)
    vp = sd.viewproj
    return SceneData(f"ndc_depth_{w}x{h}", w, h, 5.0, 5.0, sd.meshes, sd.textures, sd.items, sd.cam_pos, sd.cam_target, sd.fovy,
                     sd.sun_dir, sd.sun_color, sd.sun_intensity, sd.fp, None, viewproj=vp)


def scene_shadow_variants(variant: int, w=180, h=110) -> SceneData:
    """PCF corners (lighting/shadow_sample.hpp:65-104): radius 0 (one tap), a fractional step that rounds to 2 or 3, partial strength,
    large biases, a shadow map far smaller / larger than the frame."""
    sd = scene_small(w=w, h=h, tex=True, seed=3 + variant)
    sd.fp.shadow_enable = 1
    r, step, strength, bc, bs, size = [(0, 1.0, 1.0, 0.0008, 0.0015, 64), (1, 2.5, 0.35, 0.004, 0.02, 96), (3, 1.49, 0.8, 0.0, 0.0, 33), (2, 0.2, 1.7, 0.0008, 0.0015, 512)][variant]
    sd.fp.shadow_pcf_radius, sd.fp.shadow_pcf_step, sd.fp.shadow_strength = r, step, strength
    sd.fp.shadow_bias_const, sd.fp.shadow_bias_slope = bc, bs
    sd.shadow_size = size
    return sd


def scene_many_lights(w=128, h=96, n_lights=2600, max_per_tile=8) -> SceneData:
    """More lights than one staging round of the tile kernel holds (1024 candidates) and, with every list saturated, more survivors
    than one pass stages (256): the light loop runs several rounds and passes.  Lights crowd the objects so that they matter."""
    sd = scene_small(w=w, h=h, n_inst=3, lights=0, tex=False, seed=6)
    sd.lights = make_lights(n_lights * 3 // 4, n_lights - n_lights * 3 // 4, (-4, 0.1, -4), (4, 2.5, 4), seed=11, range_lo=1.5, range_hi=4.0)
    sd.fp.light_culling = 1
    sd.fp.max_lights_per_tile = max_per_tile
    return sd


def make_triangle_soup(n_tris: int, seed: int, extent=3.0, indexed=True, zero_normals=True) -> dict:
    """Random triangles for the fuzz scenes: a mix of small, large (frustum-crossing) and sliver triangles with random
    per-vertex normals (not unit length, sometimes zero) and uvs outside [0, 1]."""
    rng = np.random.default_rng(seed)
    centers = rng.uniform(-extent, extent, (n_tris, 1, 3))
    size = rng.choice([0.05, 0.4, 2.5, 12.0], (n_tris, 1, 1), p=[0.2, 0.5, 0.25, 0.05])
    pos = (centers + rng.uniform(-1, 1, (n_tris, 3, 3)) * size).astype(np.float32)
    sl = rng.random(n_tris) < 0.15                                       # slivers: third vertex almost on the first edge
    t = rng.random((n_tris, 1)).astype(np.float32)
    pos[sl, 2] = pos[sl, 0] * (1 - t[sl]) + pos[sl, 1] * t[sl] + rng.normal(0, 1e-4, (int(sl.sum()), 3)).astype(np.float32)
    nrm = rng.normal(0, 1, (n_tris * 3, 3)).astype(np.float32)
    zero = rng.random(n_tris * 3) < 0.03
    if zero_normals:
        nrm[zero] = 0.0                                                  # normalize(0) = NaN colour in the reference: CPU-vs-CPU bit tests only
    uv = rng.uniform(-2.5, 3.5, (n_tris * 3, 2)).astype(np.float32)
    pos = pos.reshape(-1, 3)
    if indexed:
        idx = np.arange(n_tris * 3, dtype=np.uint32)
        rng.shuffle(idx.reshape(-1, 3))                                  # triangle order != vertex order
        return {"positions": pos, "normals": nrm, "uvs": uv, "indices": idx}
    return {"positions": pos, "normals": nrm, "uvs": uv, "indices": np.zeros(0, np.uint32)}


def scene_fuzz(seed: int, lights: bool = False, zero_normals: bool = True) -> SceneData:
    """Seeded random parity scene (tests/test_fuzz_*.py): random target size, camera (often inside or grazing geometry, so the
    near / side planes clip), fov, depth range, shading model, cull mode / winding, shadow parameters, sky model, materials,
    textures of odd sizes, non-uniform and mirrored instance scales, triangle soups.  Everything is a function of `seed`."""
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(17, 230)), int(rng.integers(13, 170))
    meshes = [load_suzanne(), make_grid_plane(float(rng.uniform(6, 30)), int(rng.integers(1, 9))),
              make_triangle_soup(int(rng.integers(8, 120)), seed * 7 + 1, indexed=True, zero_normals=zero_normals),
              make_triangle_soup(int(rng.integers(3, 40)), seed * 7 + 2, extent=1.5, indexed=bool(rng.random() < 0.5), zero_normals=zero_normals)]
    tw, th = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    tex = rng.integers(0, 256, (th, tw, 4), dtype=np.uint8)
    textures = [make_checker_texture(32, seed), tex]
    sky_desc = None
    sk = rng.random()
    if sk < 0.2:
        sky_desc = {"kind": "procedural", "sun_dir": tuple(float(v) for v in rng.normal(0, 1, 3))}
    elif sk < 0.4:
        textures = textures + make_sky_faces(int(rng.integers(1, 12)), seed)
        sky_desc = {"kind": "cubemap", "faces": list(range(3, 9)), "intensity": float(rng.uniform(0.2, 2.0))}
    items = []
    for k in range(int(rng.integers(0, 7))):
        mesh = int(rng.choice([1, 2, 3, 4], p=[0.35, 0.2, 0.3, 0.15]))
        scl = rng.uniform(0.3, 2.0, 3) if rng.random() < 0.5 else np.repeat(rng.uniform(0.3, 2.0), 3)
        if rng.random() < 0.15:
            scl = scl * np.array([-1.0, 1.0, 1.0])                        # mirrored: winding flips on screen
        mat = None
        if rng.random() < 0.8:
            mat = {"base_color": tuple(float(v) for v in rng.uniform(0, 1, 3)), "metallic": float(rng.uniform(-0.2, 1.2)),
                   "roughness": float(rng.uniform(-0.1, 1.3)), "ao": float(rng.uniform(0, 1)), "tex": int(rng.choice([0, 0, 1, 2]))}
        items.append({"pos": tuple(float(v) for v in rng.uniform(-4, 4, 3) * np.array([1, 0.4, 1])), "rot": tuple(float(v) for v in rng.uniform(-3.2, 3.2, 3)),
                      "scl": tuple(float(v) for v in scl), "mesh": mesh, "material": mat, "visible": bool(rng.random() < 0.92),
                      "casts_shadow": bool(rng.random() < 0.8), "object_id": 500 + k})
    shading = capi.SHADING_BLINN if rng.random() < 0.4 else capi.SHADING_PBR
    fp = capi.default_frame_params(shading_model=shading, shadow_enable=1 if rng.random() < 0.4 else 0,
                                   cull_mode=int(rng.choice([capi.CULL_NONE, capi.CULL_BACK, capi.CULL_FRONT], p=[0.3, 0.5, 0.2])),
                                   front_face_ccw=int(rng.random() < 0.7), shadow_pcf_radius=int(rng.integers(0, 4)),
                                   shadow_pcf_step=float(rng.uniform(0.3, 3.0)), shadow_strength=float(rng.uniform(0.0, 1.2)),
                                   shadow_bias_const=float(rng.uniform(0, 0.004)), shadow_bias_slope=float(rng.uniform(0, 0.004)),
                                   exposure=float(rng.uniform(0.2, 3.0)), gamma=float(rng.uniform(1.0, 2.6)),
                                   light_culling=1 if lights else 0, tile_size=16, max_lights_per_tile=int(rng.choice([4, 16, 128])) if lights else 128)
    if rng.random() < 0.1:
        fp.debug_view = int(rng.choice([capi.DEBUG_ALBEDO, capi.DEBUG_NORMAL, capi.DEBUG_DEPTH]))
    mode = rng.random()
    if mode < 0.35:      # orbit camera looking at the cluster
        cam = tuple(float(v) for v in rng.uniform(-9, 9, 3) * np.array([1, 0.5, 1]) + np.array([0, 2.5, 0]))
        tgt = tuple(float(v) for v in rng.uniform(-1, 1, 3))
    elif mode < 0.8:     # inside the cluster: near- and side-plane clipping of large triangles
        cam = tuple(float(v) for v in rng.uniform(-3, 3, 3) * np.array([1, 0.3, 1]))
        tgt = tuple(float(v) for v in np.array(cam) + rng.normal(0, 1, 3))
    else:                # grazing the floor plane
        cam = (float(rng.uniform(-3, 3)), float(rng.uniform(-0.02, 0.05)), float(rng.uniform(-3, 3)))
        tgt = (float(rng.uniform(-3, 3)), float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-3, 3)))
    zn = float(rng.choice([0.01, 0.1, 0.5]))
    zf = float(rng.choice([20.0, 100.0, 1000.0]))
    lt = None
    if lights:
        n = int(rng.integers(1, 90))
        n_spot = int(n * rng.uniform(0, 0.6))
        lt = make_lights(max(1, n - n_spot), n_spot, (-6, -0.5, -6), (6, 3.0, 6), seed=seed, range_lo=0.5, range_hi=6.0)
    return SceneData(f"fuzz{seed}_{w}x{h}", w, h, zn, zf, meshes, textures, items, cam, tgt, math.radians(float(rng.uniform(25, 110))),
                     tuple(_norm(rng.normal(0, 1, 3) * np.array([1, 0.5, 1]) - np.array([0, 0.8, 0]))), tuple(float(v) for v in rng.uniform(0.3, 1.0, 3)),
                     float(rng.uniform(0.0, 4.0)), fp, lt, aspect=None if rng.random() < 0.7 else float(rng.uniform(0.6, 2.2)),
                     shadow_size=int(rng.choice([16, 64, 200])), sky=sky_desc)
