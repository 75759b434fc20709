"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/shsb.h declares.
No compute calls here (no GPU on the build box)."""
import ctypes as C
import os
import re

import pytest

from leisure_software_renderer_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "shsb.h")).read()
    return sorted(set(re.findall(r"SHSB_API\s+[\w\s\*]+?\b(shsb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = C.CDLL(capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    bound = capi.load_library()
    assert set(names) == set(bound._shsb_symbols), set(names) ^ set(bound._shsb_symbols)


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors in capi.py against the C compiler's own view of include/shsb.h: sizes and the offsets of the
    last field and of the fields behind alignment holes."""
    import subprocess
    probes = {"ShsbStats": (capi.Stats, ["frag_shaded"]), "ShsbRasterCfg": (capi.RasterCfg, ["reserved"]),
              "ShsbUniforms": (capi.Uniforms, ["base_color_tex", "enable_motion_vectors", "prev_model", "prev_viewproj"]),
              "ShsbTransform": (capi.Transform, ["scl"]), "ShsbRenderItem": (capi.RenderItem, ["visible", "object_id"]),
              "ShsbScene": (capi.Scene, ["items", "cam_prev_viewproj", "sky_kind", "sky_faces", "reserved2"]),
              "ShsbFrameParams": (capi.FrameParams, ["write_aovs", "motion_vectors_enable", "own_row_first", "own_row_stride"]),
              "ShsbMotionBlurParams": (capi.MotionBlurParams, ["depth_reject", "dt", "reserved"]),
              "ShsbLightShaftsParams": (capi.LightShaftsParams, ["decay", "cam_pos", "sun_dir_ws", "cam_viewproj"]),
              "ShsbLegacyUniforms": (capi.LegacyUniforms, ["model", "camera_pos", "color", "job_tile_w", "job_tile_h"]),
              "ShsbLegacy2Uniforms": (capi.Legacy2Uniforms, ["prev_mvp", "mv", "normal_mat", "light_vp", "light_dir_world", "camera_pos", "base_color", "use_texture",
                                                             "albedo", "metallic", "ao", "ibl_diffuse_intensity", "ibl_reflection_strength", "job_tile_w", "job_tile_h"])}
    lines = ["#include <stdio.h>", "#include <stddef.h>", f'#include "{os.path.join(ROOT, "include", "shsb.h")}"', "int main(void) {"]
    for cname, (_, fields) in probes.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, (cls, fields) in probes.items():
        assert int(got[cname]) == C.sizeof(cls), (cname, got[cname], C.sizeof(cls))
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_no_cpu_fallback_when_no_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = capi.load_library()
    h = C.c_void_p()
    assert lib.shsb_context_create(0, C.byref(h)) == 3  # SHSB_E_NO_DEVICE
    from leisure_software_renderer_b200.renderer import Context
    with pytest.raises(capi.ShsbError):
        Context(0)


def test_built_for_sm_100a_only():
    out = os.popen(f"cuobjdump -lelf {capi.LIB_PATH} 2>/dev/null").read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_legacy2_wrappers_marshal_their_arguments_without_a_device():
    """The numpy-facing wrappers of the legacy render-target demo entries (rows L2 / L3) called on a NULL context: every argument
    must pass ctypes' conversion and the library must answer SHSB_E_INVALID_ARGUMENT (1) -- not crash, not compute.  Catches wrapper /
    signature drift on the CPU box, where no context can be created."""
    import ctypes as C

    import numpy as np

    from leisure_software_renderer_b200 import capi
    from leisure_software_renderer_b200.renderer import Context

    ctx = Context.__new__(Context)            # no device: skip __init__, keep the methods
    ctx.lib, ctx.h, ctx._rt_shape = capi.load_library(), C.c_void_p(), {}
    u = capi.Legacy2Uniforms()
    u.mvp[:] = list(np.eye(4, dtype=np.float32).ravel())
    u.base_color[:] = [1, 2, 3, 255]
    u.albedo, u.use_texture, u.job_tile_w, u.job_tile_h = 3, 1, 160, 160
    u.metallic, u.roughness, u.ao = 0.5, 0.25, 1.0
    eye = np.eye(4)
    irr = np.zeros((6, 2, 2, 3), np.float32)
    calls = [lambda: ctx.legacy2_shadow_draw(1, eye, eye.astype(np.float32).ravel(), 2, 160, 160),
             lambda: ctx.legacy2_draw_softshadow(1, u, 0, 2, 3),
             lambda: ctx.legacy3_ibl_upload(irr, [np.zeros((6, 4, 4, 3), np.float32), np.zeros((6, 2, 2, 3), np.float32)]),
             lambda: ctx.legacy3_ibl_destroy(1),
             lambda: ctx.legacy3_draw_pbr(1, u, 2, 1, 3, 4),
             lambda: ctx.cull_objects_frustum(np.zeros((3, 10), np.float32), eye),
             lambda: ctx.tile_depth_range_from_scene(np.zeros((3, 6), np.float32), np.arange(3, dtype=np.uint32), eye, eye, 64, 48, 16, 0.1, 100.0),
             lambda: ctx.select_object_lights_from_bins(np.zeros((3, 6), np.float32), eye, eye, True, 0.1, 100.0, np.zeros((4, 160), np.uint8), 2),
             lambda: ctx.collect_object_lights(np.zeros((3, 6), np.float32), np.arange(4, dtype=np.uint32), np.zeros((4, 160), np.uint8), 1)]
    for call in calls:
        with pytest.raises(capi.ShsbError, match="status 1"):
            call()
    ctx.h = None                              # keep __del__ from touching the library
    assert C.sizeof(capi.Legacy2Uniforms) == 4 * (16 * 5 + 9 + 3 + 3) + 4 + 4 + 4 + 4 * 6 + 8
