"""ctypes binding of the C-ABI in include/shsb.h (the drop-in boundary).

Only the shared library built from leisure_software_renderer_b200/csrc is loaded here.  There is
no CPU fallback: if the library is missing or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHSB_LIB") or os.path.join(_HERE, "libshsb.so")  # SHSB_LIB: an alternative BUILD of the same library (kernel experiments)

# ---------------------------------------------------------------- enums (include/shsb.h)
OK = 0
CULL_NONE, CULL_BACK, CULL_FRONT = 0, 1, 2
SHADER_PBR_MR, SHADER_BLINN_PHONG, SHADER_DEBUG_ALBEDO, SHADER_DEBUG_NORMAL, SHADER_DEBUG_DEPTH, SHADER_DEPTH_ONLY = range(6)
SHADING_PBR, SHADING_BLINN = 0, 1
SKY_NONE, SKY_PROCEDURAL, SKY_CUBEMAP = 0, 1, 2
DEBUG_FINAL, DEBUG_ALBEDO, DEBUG_NORMAL, DEBUG_DEPTH = 0, 1, 2, 3
RT_COLOR_HDR, RT_COLOR_LDR, RT_DEPTH_MOTION, RT_SHADOW = 1, 2, 3, 4
PLANE_COLOR, PLANE_DEPTH, PLANE_MOTION, PLANE_TRI_ID, PLANE_COVERAGE = 0, 1, 2, 3, 4
TRI_ID_NONE = 0xFFFFFFFF
LIGHT_RECORD_BYTES = 160
LIGHT_CULL_TILED, LIGHT_CULL_TILED_DEPTH01, LIGHT_CULL_TILED_VIEW_DEPTH, LIGHT_CULL_CLUSTERED = 0, 1, 2, 3

F16 = C.c_float * 16
F3 = C.c_float * 3


class Stats(C.Structure):
    _fields_ = [("tri_input", C.c_uint64), ("tri_after_clip", C.c_uint64), ("tri_raster", C.c_uint64),
                ("frag_covered", C.c_uint64), ("frag_shaded", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class RasterCfg(C.Structure):
    _fields_ = [("cull_mode", C.c_int32), ("front_face_ccw", C.c_int32), ("write_aovs", C.c_int32), ("reserved", C.c_int32)]


class Uniforms(C.Structure):
    _fields_ = [("model", F16), ("viewproj", F16), ("light_viewproj", F16),
                ("light_dir_ws", F3), ("light_intensity", C.c_float),
                ("light_color", F3), ("metallic", C.c_float),
                ("camera_pos", F3), ("roughness", C.c_float),
                ("base_color", F3), ("ao", C.c_float),
                ("base_color_tex", C.c_uint32), ("shadow_map", C.c_uint32),
                ("shadow_bias_const", C.c_float), ("shadow_bias_slope", C.c_float),
                ("shadow_pcf_radius", C.c_int32), ("shadow_pcf_step", C.c_float),
                ("shadow_strength", C.c_float), ("enable_motion_vectors", C.c_int32),
                ("prev_model", F16), ("prev_viewproj", F16)]


class MotionBlurParams(C.Structure):
    """ShsbMotionBlurParams; defaults = MotionBlurPassParams (frame/frame_params.hpp:51-59) with the pass enabled."""
    _fields_ = [("enable", C.c_int32), ("samples", C.c_int32), ("strength", C.c_float), ("max_velocity_px", C.c_float),
                ("min_velocity_px", C.c_float), ("depth_reject", C.c_float), ("dt", C.c_float), ("reserved", C.c_int32)]

    def __init__(self, enable=1, samples=10, strength=1.0, max_velocity_px=20.0, min_velocity_px=0.25, depth_reject=0.08, dt=1.0 / 60.0):
        super().__init__(int(enable), int(samples), strength, max_velocity_px, min_velocity_px, depth_reject, dt, 0)


class LightShaftsParams(C.Structure):
    """ShsbLightShaftsParams; defaults = LightShaftsPassParams (frame/frame_params.hpp:35-42)."""
    _fields_ = [("enable", C.c_int32), ("steps", C.c_int32), ("density", C.c_float), ("weight", C.c_float), ("decay", C.c_float),
                ("cam_pos", F3), ("sun_dir_ws", F3), ("reserved", C.c_float), ("cam_viewproj", F16)]

    def __init__(self, cam_viewproj=None, cam_pos=(0, 0, 0), sun_dir_ws=(0, -1, 0), enable=1, steps=48, density=0.8, weight=0.9, decay=0.95):
        super().__init__()
        self.enable, self.steps, self.density, self.weight, self.decay = int(enable), int(steps), density, weight, decay
        set_f(self.cam_pos, cam_pos)
        set_f(self.sun_dir_ws, sun_dir_ws)
        if cam_viewproj is not None:
            set_f(self.cam_viewproj, cam_viewproj)


class LegacyUniforms(C.Structure):
    """ShsbLegacyUniforms: struct Uniforms of the legacy tile-job demo (hello_pipeline_blinn_phong_shading.cpp:35-41) + its job-tile size."""
    _fields_ = [("mvp", F16), ("model", F16), ("light_dir", F3), ("camera_pos", F3), ("color", C.c_uint8 * 4),
                ("job_tile_w", C.c_int32), ("job_tile_h", C.c_int32)]

    def __init__(self, mvp=None, model=None, light_dir=(0, -1, 0), camera_pos=(0, 0, 0), color=(255, 255, 255, 255), job_tile_w=0, job_tile_h=0):
        super().__init__()
        if mvp is not None:
            set_f(self.mvp, mvp)
        if model is not None:
            set_f(self.model, model)
        set_f(self.light_dir, light_dir)
        set_f(self.camera_pos, camera_pos)
        for i, v in enumerate(color):
            self.color[i] = int(v)
        self.job_tile_w, self.job_tile_h = int(job_tile_w), int(job_tile_h)


class Legacy2Uniforms(C.Structure):
    """ShsbLegacy2Uniforms: struct Uniforms of the legacy render-target demos (hello_shadow_mapping_soft.cpp:714-732, hello_pbr.cpp:474-519)
    as plain data + MaterialPBR + the job-tile size."""
    _fields_ = [("mvp", F16), ("prev_mvp", F16), ("model", F16), ("mv", F16), ("normal_mat", C.c_float * 9), ("light_vp", F16),
                ("light_dir_world", F3), ("camera_pos", F3), ("base_color", C.c_uint8 * 4), ("use_texture", C.c_int32), ("albedo", C.c_uint32),
                ("metallic", C.c_float), ("roughness", C.c_float), ("ao", C.c_float),
                ("ibl_diffuse_intensity", C.c_float), ("ibl_specular_intensity", C.c_float), ("ibl_reflection_strength", C.c_float),
                ("job_tile_w", C.c_int32), ("job_tile_h", C.c_int32)]


class FlatDraw(C.Structure):
    """ShsbFlatDraw: one flat-shaded draw of a batch (a DebugMesh, its model matrix and base colour, the object's LightSelection)."""
    _fields_ = [("mesh", C.c_uint32), ("selection_count", C.c_uint32), ("model", F16), ("base_color", F3), ("selection", C.c_uint32 * 8)]


# ShsbLightProperties (128 bytes): shs::LightProperties (lighting/light_runtime.hpp:53-72) + the LightType of the instance's model
LIGHT_PROPS_DTYPE = np.dtype([("color", "<f4", 3), ("intensity", "<f4"), ("position_ws", "<f4", 3), ("range", "<f4"), ("direction_ws", "<f4", 3), ("inner_angle_rad", "<f4"),
                              ("right_ws", "<f4", 3), ("outer_angle_rad", "<f4"), ("up_ws", "<f4", 3), ("tube_half_length", "<f4"), ("rect_half_extents", "<f4", 2),
                              ("tube_radius", "<f4"), ("attenuation_power", "<f4"), ("attenuation_bias", "<f4"), ("attenuation_cutoff", "<f4"), ("attenuation_model", "<u4"),
                              ("flags", "<u4"), ("light_type", "<u4"), ("reserved", "<u4", 3)])
assert LIGHT_PROPS_DTYPE.itemsize == 128
LIGHT_TYPE_POINT, LIGHT_TYPE_SPOT, LIGHT_TYPE_RECT_AREA, LIGHT_TYPE_TUBE_AREA = 1, 2, 3, 4


class LightCullDesc(C.Structure):
    """ShsbLightCullDesc: arguments of the bin builders of lighting/jolt_light_culling.hpp."""
    _fields_ = [("view_proj", F16), ("viewport_w", C.c_uint32), ("viewport_h", C.c_uint32), ("tile_size", C.c_uint32),
                ("max_per_bin", C.c_uint32), ("mode", C.c_int32), ("depth_slices", C.c_uint32), ("z_near", C.c_float), ("z_far", C.c_float)]

    def __init__(self, view_proj=None, w=0, h=0, mode=0, tile_size=16, max_per_bin=128, depth_slices=16, z_near=0.1, z_far=1000.0):
        super().__init__()
        if view_proj is not None:
            set_f(self.view_proj, view_proj)
        self.viewport_w, self.viewport_h, self.tile_size, self.max_per_bin = w, h, tile_size, max_per_bin
        self.mode, self.depth_slices, self.z_near, self.z_far = mode, depth_slices, z_near, z_far

    def tiles(self):
        return ((self.viewport_w + self.tile_size - 1) // self.tile_size) * ((self.viewport_h + self.tile_size - 1) // self.tile_size)

    def bins(self):
        return self.tiles() * (self.depth_slices if self.mode == LIGHT_CULL_CLUSTERED else 1)


class Transform(C.Structure):
    _fields_ = [("pos", F3), ("rot_euler", F3), ("scl", F3)]


class RenderItem(C.Structure):
    _fields_ = [("tr", Transform), ("mesh", C.c_uint32), ("has_material", C.c_uint32),
                ("base_color", F3), ("metallic", C.c_float), ("roughness", C.c_float), ("ao", C.c_float),
                ("base_color_tex", C.c_uint32), ("casts_shadow", C.c_uint32), ("visible", C.c_uint32),
                ("object_id", C.c_uint64)]


class Scene(C.Structure):
    _fields_ = [("cam_viewproj", F16), ("cam_pos", F3), ("sun_intensity", C.c_float),
                ("sun_dir_ws", F3), ("n_items", C.c_uint32),
                ("sun_color", F3), ("reserved", C.c_uint32),
                ("items", C.POINTER(RenderItem)),
                ("cam_prev_viewproj", F16), ("sky_kind", C.c_int32), ("sky_intensity", C.c_float), ("sky_sun_dir_ws", F3),
                ("sky_faces", C.c_uint32 * 6), ("reserved2", C.c_uint32)]


class FrameParams(C.Structure):
    _fields_ = [("shading_model", C.c_int32), ("debug_view", C.c_int32), ("cull_mode", C.c_int32), ("front_face_ccw", C.c_int32),
                ("shadow_enable", C.c_int32), ("shadow_bias_const", C.c_float), ("shadow_bias_slope", C.c_float),
                ("shadow_pcf_radius", C.c_int32), ("shadow_pcf_step", C.c_float), ("shadow_strength", C.c_float),
                ("exposure", C.c_float), ("gamma", C.c_float),
                ("light_culling", C.c_int32), ("tile_size", C.c_uint32), ("max_lights_per_tile", C.c_uint32),
                ("write_aovs", C.c_int32), ("motion_vectors_enable", C.c_int32),
                ("own_row_first", C.c_int32), ("own_row_count", C.c_int32), ("own_row_stride", C.c_int32)]


def default_frame_params(**kw) -> FrameParams:
    """FrameParams defaults of the reference (frame/frame_params.hpp:117-171, :20-31, :73-85)."""
    fp = FrameParams(shading_model=SHADING_PBR, debug_view=DEBUG_FINAL, cull_mode=CULL_BACK, front_face_ccw=1,
                     shadow_enable=1, shadow_bias_const=0.0008, shadow_bias_slope=0.0015, shadow_pcf_radius=2,
                     shadow_pcf_step=1.0, shadow_strength=1.0, exposure=1.0, gamma=2.2,
                     light_culling=0, tile_size=16, max_lights_per_tile=128, write_aovs=0,
                     motion_vectors_enable=0)  # the reference default is true; scenes switch it on explicitly
    for k, v in kw.items():
        setattr(fp, k, v)
    return fp


def fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def u32ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def set_f(arr, values):
    for i, v in enumerate(np.asarray(values, dtype=np.float32).reshape(-1)):
        arr[i] = float(v)


class ShsbError(RuntimeError):
    pass


_lib = None


class GatherExport(C.Structure):
    """ShsbGatherExport: what the root of a sort-first frame assembly hands to the other ranks (184 bytes, plain bytes: ship it
    over any host channel, e.g. torch.distributed.broadcast_object_list(bytes(export)))."""
    _fields_ = [("mem_handle", C.c_ubyte * 64), ("ctl_handle", C.c_ubyte * 64), ("mem_ptr", C.c_uint64), ("ctl_ptr", C.c_uint64),
                ("slot_bytes", C.c_uint64), ("n_ranks", C.c_uint32), ("slots", C.c_uint32), ("root_device", C.c_int32), ("root_pid", C.c_int32)]


def load_library(path: str | None = None):
    """Loads libshsb.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ShsbError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for the raster path)")
    lib = C.CDLL(p)
    P = C.POINTER
    vp = C.c_void_p
    sigs = {
        "shsb_context_create": [C.c_int32, P(vp)],
        "shsb_context_destroy": [vp],
        "shsb_sync": [vp],
        "shsb_stream": [vp, P(vp)],
        "shsb_fence": [vp],
        "shsb_set_tile_streams": [vp, C.c_int32],
        "shsb_launch_count": [vp, P(C.c_uint64)],
        "shsb_last_tile_kernel": [vp, P(C.c_int32)],
        "shsb_software_occlusion": [vp, P(C.c_float), C.c_uint32, P(C.c_uint32), C.c_uint32, P(C.c_uint32), P(C.c_float), P(C.c_uint32), C.c_uint32, P(C.c_float), C.c_uint32,
                                    P(C.c_uint32), C.c_uint32, P(C.c_float), P(C.c_float), C.c_int32, C.c_int32, C.c_float, C.c_int32, P(C.c_uint8), P(C.c_uint32), P(C.c_uint32),
                                    P(C.c_float)],
        "shsb_flat_draw_blinn_phong": [vp, P(FlatDraw), C.c_uint32, P(C.c_float), P(C.c_float), P(C.c_float), C.c_uint32, C.c_uint32],
        "shsb_flat_draw_multi_light": [vp, P(FlatDraw), C.c_uint32, P(C.c_float), P(C.c_float), vp, C.c_uint32, C.c_uint32, C.c_uint32],
        "shsb_gather_create": [vp, C.c_uint32, C.c_uint32, C.c_size_t, P(C.c_uint32), P(GatherExport)],
        "shsb_gather_open": [vp, P(GatherExport), C.c_uint32, P(C.c_uint32)],
        "shsb_gather_destroy": [vp, C.c_uint32],
        "shsb_frame_gather": [vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int32, C.c_size_t, C.c_size_t, C.c_size_t],
        "shsb_gather_commit": [vp, C.c_uint32, C.c_uint64],
        "shsb_gather_wait": [vp, C.c_uint32, C.c_uint64],
        "shsb_gather_release": [vp, C.c_uint32, C.c_uint64],
        "shsb_gather_device_ptr": [vp, C.c_uint32, C.c_uint64, P(vp)],
        "shsb_gather_download": [vp, C.c_uint32, C.c_uint64, C.c_size_t, vp, C.c_size_t],
        "shsb_gather_download_async": [vp, C.c_uint32, C.c_uint64, C.c_size_t, vp, C.c_size_t],
        "shsb_gather_stream": [vp, P(vp)],
        "shsb_mesh_upload": [vp, P(C.c_float), C.c_uint32, P(C.c_float), C.c_uint32, P(C.c_float), C.c_uint32, P(C.c_uint32), C.c_uint32, P(C.c_uint32)],
        "shsb_mesh_destroy": [vp, C.c_uint32],
        "shsb_mesh_load_obj": [vp, C.c_char_p, P(C.c_uint32)],
        "shsb_texture_load_png": [vp, C.c_char_p, C.c_int32, P(C.c_uint32)],
        "shsb_mesh_info": [vp, C.c_uint32, P(C.c_uint32)],
        "shsb_mesh_download": [vp, C.c_uint32, P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_uint32)],
        "shsb_texture_info": [vp, C.c_uint32, P(C.c_int32)],
        "shsb_texture_download": [vp, C.c_uint32, P(C.c_uint8), C.c_size_t],
        "shsb_texture_upload": [vp, P(C.c_uint8), C.c_int32, C.c_int32, P(C.c_uint32)],
        "shsb_texture_destroy": [vp, C.c_uint32],
        "shsb_rt_create": [vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, P(C.c_uint32)],
        "shsb_rt_destroy": [vp, C.c_uint32],
        "shsb_rt_clear": [vp, C.c_uint32, C.c_int32, vp],
        "shsb_rt_upload": [vp, C.c_uint32, C.c_int32, vp, C.c_size_t],
        "shsb_rt_download": [vp, C.c_uint32, C.c_int32, vp, C.c_size_t],
        "shsb_rt_download_async": [vp, C.c_uint32, C.c_int32, vp, C.c_size_t],
        "shsb_rt_device_ptr": [vp, C.c_uint32, C.c_int32, P(vp), P(C.c_size_t)],
        "shsb_model_from_transform": [P(Transform), P(C.c_float)],
        "shsb_camera_viewproj": [P(C.c_float), P(C.c_float), P(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float, P(C.c_float)],
        "shsb_rasterize_mesh": [vp, C.c_uint32, C.c_int32, P(Uniforms), C.c_uint32, C.c_uint32, P(RasterCfg), P(Stats)],
        "shsb_pass_pbr_forward": [vp, P(Scene), P(FrameParams), C.c_uint32, C.c_uint32, C.c_uint32, P(C.c_float), C.c_int32, P(Stats)],
        "shsb_pass_depth_prepass": [vp, P(Scene), P(FrameParams), C.c_uint32, P(Stats)],
        "shsb_pass_shadow_map": [vp, P(Scene), P(FrameParams), C.c_uint32, P(C.c_float)],
        "shsb_history_reset": [vp],
        "shsb_pass_tonemap": [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float],
        "shsb_pass_motion_blur": [vp, P(MotionBlurParams), C.c_uint32, C.c_uint32, C.c_uint32],
        "shsb_pass_light_shafts": [vp, P(LightShaftsParams), C.c_uint32, C.c_uint32, C.c_uint32],
        "shsb_pass_taa": [vp, C.c_uint32],
        "shsb_taa_reset": [vp],
        "shsb_lights_upload": [vp, vp, C.c_uint32],
        "shsb_light_cull": [vp, P(C.c_float), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32],
        "shsb_light_cull_ex": [vp, P(LightCullDesc), P(C.c_float), P(C.c_float)],
        "shsb_cluster_lists_download": [vp, P(C.c_uint32), C.c_size_t, P(C.c_uint32), C.c_size_t],
        "shsb_tile_depth_range": [vp, C.c_uint32, C.c_uint32],
        "shsb_tile_depth_range_ndc01": [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float],
        "shsb_tile_depth_range_download": [vp, P(C.c_float), P(C.c_float), C.c_size_t],
        "shsb_light_lists_download": [vp, P(C.c_uint32), C.c_size_t, P(C.c_uint32), C.c_size_t],
        "shsb_frame_forward_plus": [vp, P(Scene), P(FrameParams), C.c_uint32, C.c_uint32, C.c_uint32, P(Stats)],
        "shsb_last_stage_ms": [vp, P(C.c_float)],
        "shsb_host_submit_us": [vp, P(C.c_double), C.c_int32],
        "shsb_legacy_camera": [P(C.c_float), C.c_float, C.c_float, P(C.c_float), P(C.c_float)],
        "shsb_legacy_world_matrix": [P(C.c_float), P(C.c_float), C.c_float, P(C.c_float)],
        "shsb_legacy_mvp": [P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_float)],
        "shsb_legacy_draw_blinn_phong": [vp, C.c_uint32, P(LegacyUniforms), C.c_uint32, C.c_uint32],
        "shsb_legacy2_shadow_draw": [vp, C.c_uint32, P(C.c_float), P(C.c_float), C.c_int32, C.c_int32, C.c_uint32],
        "shsb_legacy2_draw_softshadow": [vp, C.c_uint32, P(Legacy2Uniforms), C.c_uint32, C.c_uint32, C.c_uint32],
        "shsb_legacy3_ibl_upload": [vp, P(C.c_float), C.c_int32, P(C.c_float), P(C.c_int32), C.c_int32, P(C.c_uint32)],
        "shsb_legacy3_ibl_destroy": [vp, C.c_uint32],
        "shsb_legacy3_draw_pbr": [vp, C.c_uint32, P(Legacy2Uniforms), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32],
        "shsb_cull_objects_frustum": [vp, P(C.c_float), C.c_uint32, P(C.c_float), P(C.c_uint8), P(C.c_uint32), P(C.c_uint32)],
        "shsb_collect_object_lights": [vp, P(C.c_float), C.c_uint32, P(C.c_uint32), C.c_uint32, vp, C.c_uint32, C.c_int32, P(C.c_uint32), P(C.c_uint32), P(C.c_float)],
        "shsb_tile_depth_range_from_scene": [vp, P(C.c_float), C.c_uint32, P(C.c_uint32), C.c_uint32, P(C.c_float), P(C.c_float), C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float],
        "shsb_select_object_lights_from_bins": [vp, P(C.c_float), C.c_uint32, P(C.c_float), P(C.c_float), C.c_int32, C.c_float, C.c_float, vp, C.c_uint32, C.c_int32,
                                                P(C.c_uint32), P(C.c_uint32), P(C.c_float), P(C.c_uint32)],
        "shsb_timing_enable": [vp, C.c_int32],
        "shsb_timing_collect": [vp, P(C.c_float), C.c_size_t, P(C.c_size_t)],
        "shsb_timing_collect_abs": [vp, P(C.c_float), C.c_size_t, P(C.c_size_t)],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export what shsb.h declares
        fn.argtypes = args
        fn.restype = C.c_int32
    lib.shsb_last_error_string.argtypes = [vp]
    lib.shsb_last_error_string.restype = C.c_char_p
    lib.shsb_version.argtypes = []
    lib.shsb_version.restype = C.c_char_p
    lib._shsb_symbols = sorted(list(sigs) + ["shsb_last_error_string", "shsb_version"])
    if path is None:
        _lib = lib
    return lib
