"""Parity cases shared by the CPU (oracle vs reference) and GPU (CUDA vs oracle) tests."""
from leisure_software_renderer_b200 import capi, scenes


def forward_cases():
    """name -> (SceneData factory, kwargs for harness.*_forward)"""
    def shadow_scene():
        sd = scenes.scene_small(w=200, h=120, tex=True)
        sd.fp.shadow_enable = 1
        return sd

    def debug(mode):
        def f():
            sd = scenes.scene_small(w=128, h=96)
            sd.fp.debug_view = mode
            return sd
        return f

    def cull(mode, ccw=1):
        def f():
            sd = scenes.scene_small(w=128, h=96)
            sd.fp.cull_mode = mode
            sd.fp.front_face_ccw = ccw
            return sd
        return f

    return {
        "c1_suzanne_blinn_640x480": (scenes.scene_c1, {}),
        "small_pbr": (lambda: scenes.scene_small(), {}),
        "small_blinn_tex": (lambda: scenes.scene_small(tex=True, shading=capi.SHADING_BLINN), {}),
        "ragged_size_pbr": (lambda: scenes.scene_small(w=173, h=99, n_inst=4, seed=3), {}),
        "nearclip_tex": (lambda: scenes.scene_small(near_clip=True, tex=True), {}),
        "nearclip_blinn": (lambda: scenes.scene_small(w=200, h=150, near_clip=True, shading=capi.SHADING_BLINN, seed=5), {}),
        "painter_no_depth": (lambda: scenes.scene_small(), {"depth": False}),
        "shadow_pcf": (shadow_scene, {"shadow": True}),
        "debug_albedo": (debug(capi.DEBUG_ALBEDO), {}),
        "debug_normal": (debug(capi.DEBUG_NORMAL), {}),
        "debug_depth": (debug(capi.DEBUG_DEPTH), {}),
        "cull_none": (cull(capi.CULL_NONE), {}),
        "cull_front": (cull(capi.CULL_FRONT), {}),
        "front_face_cw": (cull(capi.CULL_BACK, 0), {}),
        "sky_procedural": (lambda: scenes.scene_small(w=192, h=108, sky="procedural"), {}),
        "sky_cubemap_tex": (lambda: scenes.scene_small(w=176, h=100, sky="cubemap", tex=True), {}),
        "sky_cubemap_nearclip_blinn": (lambda: scenes.scene_small(w=150, h=90, sky="cubemap", near_clip=True, shading=capi.SHADING_BLINN), {}),
        "ragged_inputs_pbr": (lambda: scenes.scene_ragged(), {}),
        "ragged_inputs_blinn_shadow": (lambda: scenes.scene_ragged(w=131, h=97, shading=capi.SHADING_BLINN, shadow=True, nonfinite=False), {"shadow": True}),
        "odd_textures_wrap_pbr": (lambda: scenes.scene_odd_texture(), {}),
        "odd_textures_wrap_blinn": (lambda: scenes.scene_odd_texture(w=97, h=71, shading=capi.SHADING_BLINN), {}),
        "ndc_depth_when_zf_equals_zn": (lambda: scenes.scene_ndc_depth(), {}),
        "shadow_pcf_radius0": (lambda: scenes.scene_shadow_variants(0), {"shadow": True}),
        "shadow_pcf_step2_5_strength": (lambda: scenes.scene_shadow_variants(1), {"shadow": True}),
        "shadow_pcf_radius3_tiny_map": (lambda: scenes.scene_shadow_variants(2), {"shadow": True}),
        "shadow_pcf_step_below_1_strength_above_1": (lambda: scenes.scene_shadow_variants(3), {"shadow": True}),
        "adversarial_boundaries": (lambda: scenes.scene_adversarial(), {}),
        "adversarial_boundaries_painter": (lambda: scenes.scene_adversarial(w=81, h=57, shading=capi.SHADING_BLINN), {"depth": False}),
        "vertex_on_camera_plane": (lambda: scenes.scene_w_zero(), {}),
        "empty_scene": (lambda: scenes.scene_small(n_inst=0, w=64, h=48), {}),
        "tiny_target_1x1": (lambda: scenes.scene_small(w=1, h=1), {}),
    }


def forward_plus_cases():
    return {
        "fplus_pbr": (lambda: scenes.scene_small(w=320, h=200, lights=64), {"forward_plus": True}),
        "fplus_blinn_tex": (lambda: scenes.scene_small(w=333, h=207, lights=64, tex=True, shading=capi.SHADING_BLINN), {"forward_plus": True}),
        "fplus_saturated": (lambda: _saturated(), {"forward_plus": True}),
        "fplus_c2_small": (lambda: scenes.scene_c2(w=640, h=360, grid=4, n_point=96, n_spot=32), {"forward_plus": True}),
        "fplus_2600_lights_saturated": (lambda: scenes.scene_many_lights(), {"forward_plus": True}),
        "fplus_1500_lights_long_lists": (lambda: scenes.scene_many_lights(w=96, h=64, n_lights=1500, max_per_tile=2048), {"forward_plus": True}),
        # every light type and attenuation model the reference can pack (records from its own packers), incl. disabled / zero lights
        "fplus_mixed_light_types_pbr": (lambda: _mixed(), {"forward_plus": True}),
        "fplus_mixed_light_types_blinn_saturated": (lambda: _mixed(w=200, h=120, shading=capi.SHADING_BLINN, max_per_tile=6), {"forward_plus": True}),
    }


def _mixed(**kw):
    import os
    import numpy as np
    rec = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_area_lights.npz"))["records"]
    return scenes.scene_mixed_lights(rec, **kw)


def _saturated():
    sd = scenes.scene_small(w=160, h=120, lights=96, seed=9)
    sd.fp.max_lights_per_tile = 8  # forces count >= max -> "walk all lights" fallback (fp_stress_scene.frag:662-668)
    return sd


def motion_cases():
    """name -> (factory of (previous frame, current frame), kwargs).  The current frame's Camera::prev_viewproj is the previous
    frame's viewproj; Context::history holds the previous frame's model matrices (keyed by RenderItem::object_id)."""
    def moving_objects():
        a = scenes.scene_small(w=192, h=120, motion=True, n_inst=4, seed=2)
        return a, a.moved()

    def moving_camera_and_objects():
        a = scenes.scene_small(w=200, h=113, motion=True, tex=True, near_clip=True)
        return a, a.moved(dpos=(0.05, 0.0, 0.1), cam_pos=(0.65, 1.25, -2.3), cam_target=(0.0, 0.55, 0.1))

    def fast_motion_clamped():
        a = scenes.scene_small(w=160, h=100, motion=True, n_inst=3, seed=4)
        return a, a.moved(dpos=(6.0, 0.5, -3.0), drot=(0.3, 1.2, 0.0), cam_pos=(3.0, 4.0, -8.0))   # > 96 px: velocity clamp

    def first_frame():
        a = scenes.scene_small(w=128, h=80, motion=True)
        return None, a                                                                                # no history: zero object motion

    return {"moving_objects": (moving_objects, {}), "moving_camera_and_objects": (moving_camera_and_objects, {}),
            "fast_motion_clamped": (fast_motion_clamped, {}), "first_frame": (first_frame, {})}
