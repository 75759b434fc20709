"""The C++ drop-in binding (leisure_software_renderer_b200/host/shs_b200/drop_in.hpp) compiled against the
reference's own headers: CPU box -> it compiles and refuses to run without a device; GPU box -> reference CPU passes
vs B200 passes through the reference's types."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "drop_in_test")
PLUGIN_BIN = os.path.join(ROOT, "tests", "cpp", "_build", "plugin_test")
LEGACY_BIN = os.path.join(ROOT, "tests", "cpp", "_build", "legacy_drop_in_test")


def _build():
    from leisure_software_renderer_b200 import build
    build.build()
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], check=True, capture_output=True)


def test_drop_in_compiles_and_has_no_cpu_fallback():
    import torch
    if not os.path.isdir("/root/reference") and not os.path.exists(BIN):
        pytest.skip("reference tree absent and no prebuilt binary")
    _build()
    assert os.path.exists(BIN)
    if not torch.cuda.is_available():
        r = subprocess.run([BIN], capture_output=True, text=True)
        assert r.returncode == 77 and "SKIP" in r.stdout


@pytest.mark.gpu
def test_drop_in_parity_on_gpu():
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/_build/drop_in_test was not built (needs /root/reference at build time)")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr


def test_plugin_registry_profiles_and_planner():
    """host/shs_b200/plugin.hpp: the B200 passes behind the reference's IRenderPass / PassFactoryRegistry interface, assembled by
    the reference's own PluggablePipeline and accepted by its planner (part 1 of tests/cpp/plugin_test.cpp; no device needed)."""
    import torch
    if not os.path.isdir("/root/reference") and not os.path.exists(PLUGIN_BIN):
        pytest.skip("reference tree absent and no prebuilt binary")
    _build()
    assert os.path.exists(PLUGIN_BIN)
    if not torch.cuda.is_available():
        r = subprocess.run([PLUGIN_BIN], capture_output=True, text=True)
        print(r.stdout, r.stderr)
        assert r.returncode == 77 and "part 1 (registry / profiles / planner): OK" in r.stdout and "SKIP part 2" in r.stdout


@pytest.mark.gpu
def test_plugin_pipeline_parity_on_gpu():
    """Frames rendered through the reference's PipelineRuntimeExecutor with the B200 registry equal the reference's CPU passes
    (Forward, Forward+ with and without the queue emulation -- quirk Q1 --, clustered), payload counts, opt-in local lights."""
    if not os.path.exists(PLUGIN_BIN):
        pytest.skip("tests/cpp/_build/plugin_test was not built (needs /root/reference at build time)")
    r = subprocess.run([PLUGIN_BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr


def test_legacy_drop_in_compiles_and_has_no_cpu_fallback():
    """host/shs_b200/legacy_drop_in.hpp compiled INTO the reference's config-1 demo source (hello_pipeline_blinn_phong_shading.cpp,
    main renamed): without a device every call of the binding refuses and nothing is rendered on the CPU."""
    import torch
    if not os.path.isdir("/root/reference") and not os.path.exists(LEGACY_BIN):
        pytest.skip("reference tree absent and no prebuilt binary")
    _build()
    assert os.path.exists(LEGACY_BIN)
    if not torch.cuda.is_available():
        r = subprocess.run([LEGACY_BIN], capture_output=True, text=True)
        print(r.stdout, r.stderr)
        assert r.returncode == 77 and "every call refused" in r.stdout


@pytest.mark.gpu
def test_legacy_drop_in_parity_on_gpu():
    """The demo's own RendererSystem::draw_triangle_tile loops on the CPU vs shs::b200::legacy::Renderer over the same shs::Canvas /
    shs::ZBuffer / geometry: z-buffer bit-equal, canvas within 1 LSB."""
    if not os.path.exists(LEGACY_BIN):
        pytest.skip("tests/cpp/_build/legacy_drop_in_test was not built (needs /root/reference at build time)")
    r = subprocess.run([LEGACY_BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("name", ["legacy2_drop_in_test", "legacy3_drop_in_test"])
def test_legacy2_drop_in_compiles_and_has_no_cpu_fallback(name):
    """host/shs_b200/legacy2_drop_in.hpp compiled INTO the reference's render-target demo sources (hello_shadow_mapping_soft.cpp /
    hello_pbr.cpp, main renamed): the reference side renders a non-trivial frame with the demo's own loops; without a device every
    call of the binding refuses and nothing is rendered on the CPU.  (GPU half: tests/test_zz_gpu_legacy2.py.)"""
    import torch
    path = os.path.join(ROOT, "tests", "cpp", "_build", name)
    if not os.path.isdir("/root/reference") and not os.path.exists(path):
        pytest.skip("reference tree absent and no prebuilt binary")
    _build()
    assert os.path.exists(path)
    if not torch.cuda.is_available():
        r = subprocess.run([path], capture_output=True, text=True)
        print(r.stdout, r.stderr)
        assert r.returncode == 77 and "every call refused" in r.stdout and "reference side: 18598 of 30000 px covered" in r.stdout


def test_scene_cull_drop_in_compiles_and_has_no_cpu_fallback():
    """host/shs_b200/scene_cull_drop_in.hpp compiled against the reference's Jolt-guarded headers (JoltPhysics declaration shim): the
    reference side culls / selects / builds ranges with its own functions; without a device the binding refuses every call."""
    import torch
    path = os.path.join(ROOT, "tests", "cpp", "_build", "scene_cull_drop_in_test")
    if not os.path.isdir("/root/reference") and not os.path.exists(path):
        pytest.skip("reference tree absent and no prebuilt binary")
    _build()
    assert os.path.exists(path)
    if not torch.cuda.is_available():
        r = subprocess.run([path], capture_output=True, text=True)
        print(r.stdout, r.stderr)
        assert r.returncode == 77 and "every call refused" in r.stdout and "417 of 600 objects visible" in r.stdout


def test_flat_draw_drop_in_compiles_and_has_no_cpu_fallback():
    """host/shs_b200/flat_draw_drop_in.hpp compiled against the reference's debug_draw.hpp / light_runtime.hpp (JoltPhysics declaration
    shim) and the demo's own draw function: the reference side draws a frame with its per-object calls; without a device the binding
    refuses every call and draws nothing."""
    import torch
    path = os.path.join(ROOT, "tests", "cpp", "_build", "flat_draw_drop_in_test")
    if not os.path.isdir("/root/reference") and not os.path.exists(path):
        pytest.skip("reference tree absent and no prebuilt binary")
    _build()
    assert os.path.exists(path)
    if not torch.cuda.is_available():
        r = subprocess.run([path], capture_output=True, text=True)
        print(r.stdout, r.stderr)
        assert r.returncode == 77 and "every call refused" in r.stdout and "26721 of 120000 texels covered" in r.stdout


def test_gather_test_compiles_against_the_c_abi_alone_and_refuses_without_a_device():
    """tests/cpp/gather_test.cpp includes nothing but include/shsb.h; without a device both of its processes stop at
    shsb_context_create (SHSB_E_NO_DEVICE -> exit code 77)."""
    import torch
    from leisure_software_renderer_b200 import build
    build.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "gather_test")
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "_build/gather_test"], check=True, capture_output=True)
    assert os.path.exists(exe)
    src = open(os.path.join(ROOT, "tests", "cpp", "gather_test.cpp")).read()
    assert '#include "shsb.h"' in src and "torch" not in src.split("#include", 1)[1] and "nccl" not in src.lower().split("#include", 1)[1]
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert r.returncode == 77 and "SKIP" in r.stdout
