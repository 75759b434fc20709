"""Row A9 (the Forward+ per-fragment local-light loop) PINNED: the reference has it only as GLSL
(shaders/vulkan/fp_stress_scene.frag:132-165, 421-523, 644-678; common/light_math.glsl:44-78).  oracle/extract_glsl_a9.py lifts that
text into C++ with lexical rewrites only, oracle/ref_glsl_a9_harness.cpp compiles it against oracle/glsl_shim (GLSL built-ins
evaluated in binary32 with the formulas of the GLSL specification), and these tests hold oracle.cpp's restatement to it:
per-light radiance over every light type / attenuation model / technique, the attenuation function on its own, and the list walk
of main() (tile lookup from gl_FragCoord, capped lists, saturated lists falling back to every light, out-of-range indices).

Bit-equality is demanded: both sides evaluate the same expressions in the same order with the same libm."""
import numpy as np
import pytest

from leisure_software_renderer_b200 import scenes
from oracle import bindings

needs_glsl = pytest.mark.skipif(not bindings.LocalLightEvaluator.available(), reason="oracle/_ref/libshs_glsl_a9_ref.so not built and /root/reference absent")


@pytest.fixture(scope="module")
def glsl():
    return bindings.LocalLightEvaluator("glsl")


@pytest.fixture(scope="module")
def port():
    return bindings.LocalLightEvaluator("port")


def _golden_lights():
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_area_lights.npz")
    return np.load(p)["records"].view(scenes.LIGHT_DTYPE).reshape(-1)


def _surface(rng, lights, k):
    """A surface point near light k (so that a good share of the samples is inside its range and cone), random frame."""
    pos = np.array(lights["position_range"][k][:3], dtype=np.float32)
    reach = float(lights["position_range"][k][3])
    d = rng.normal(size=3).astype(np.float32)
    d /= np.linalg.norm(d)
    P = (pos + d * np.float32(rng.uniform(0.0, 1.3) * max(reach, 0.05))).astype(np.float32)
    N = rng.normal(size=3).astype(np.float32); N /= np.linalg.norm(N)
    V = rng.normal(size=3).astype(np.float32); V /= np.linalg.norm(V)
    albedo = rng.uniform(0.0, 1.0, size=3).astype(np.float32)
    return P, N.astype(np.float32), V.astype(np.float32), albedo, float(np.float32(rng.uniform(0, 1))), float(np.float32(rng.uniform(0.04, 1)))


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


@needs_glsl
def test_attenuation_equals_the_glsl(glsl, port):
    rng = np.random.default_rng(9)
    n = 0
    for _ in range(4000):
        dist, rg = float(np.float32(rng.uniform(0, 12))), float(np.float32(rng.uniform(0.0, 10)))
        model = int(rng.integers(0, 4))
        power = float(np.float32(rng.choice([1.0, 1.25, 0.5, 2.0, 0.0, 3.7])))
        bias = float(np.float32(rng.choice([0.0, 1e-6, 0.01, 0.5])))
        cutoff = float(np.float32(rng.choice([0.0, -1.0, 0.002, 0.05])))
        a, b = glsl.attenuation(dist, rg, model, power, bias, cutoff), port.attenuation(dist, rg, model, power, bias, cutoff)
        assert _bits(a) == _bits(b), (dist, rg, model, power, bias, cutoff, a, b)
        n += int(a > 0)
    assert n > 800


@needs_glsl
@pytest.mark.parametrize("technique", [0, 1])
def test_per_light_radiance_equals_the_glsl_on_the_reference_packed_records(glsl, port, technique):
    """golden_area_lights.npz: 60 records written by the reference's own packers (point lights with all three attenuation models,
    spot, rect-area, tube-area, disabled, zero-intensity and zero-range lights)."""
    lights = _golden_lights()
    rng = np.random.default_rng(100 + technique)
    lit = np.zeros(len(lights), dtype=int)
    for k in range(len(lights)):
        for _ in range(120):
            P, N, V, albedo, metallic, roughness = _surface(rng, lights, k)
            a = glsl.eval_local_light(lights, k, P, N, V, albedo, metallic, roughness, technique)
            b = port.eval_local_light(lights, k, P, N, V, albedo, metallic, roughness, technique)
            assert np.array_equal(_bits(a), _bits(b)), f"light {k} (type {lights['type_shape_flags'][k]}): glsl {a} != restatement {b}"
            lit[k] += int(np.any(a != 0))
    types = lights["type_shape_flags"][:, 0]
    for t in (1, 2, 3, 4):
        assert lit[types == t].sum() > 20, f"light type {t} never lit a sample: the comparison would be vacuous"


@needs_glsl
def test_per_light_radiance_equals_the_glsl_on_fuzzed_point_and_spot_lights(glsl, port):
    rng = np.random.default_rng(7)
    lit = 0
    for seed in range(12):
        lights = scenes.make_lights(24, 24, (-6, 0.2, -6), (6, 3.0, 6), seed=seed)
        for k in range(len(lights)):
            for _ in range(12):
                P, N, V, albedo, metallic, roughness = _surface(rng, lights, k)
                tech = int(rng.integers(0, 2))
                a = glsl.eval_local_light(lights, k, P, N, V, albedo, metallic, roughness, tech)
                b = port.eval_local_light(lights, k, P, N, V, albedo, metallic, roughness, tech)
                assert np.array_equal(_bits(a), _bits(b)), f"seed {seed} light {k}: glsl {a} != restatement {b}"
                lit += int(np.any(a != 0))
    assert lit > 500


@needs_glsl
def test_list_walk_equals_the_glsl(glsl, port):
    """main():644-678 -- tile from gl_FragCoord (the harness maps the bottom-up framebuffer row to Vulkan's top-down window
    coordinate), count = min(tile_counts, max_per_tile), saturated lists walk EVERY light, indices >= light_count are skipped,
    radiance accumulated in list order."""
    rng = np.random.default_rng(21)
    W, H, ts = 100, 70, 16            # 7 x 5 tiles, the last column / row partial
    tiles_x, tiles_y = (W + ts - 1) // ts, (H + ts - 1) // ts
    lights = scenes.make_lights(20, 12, (-3, 0, -3), (3, 3, 3), seed=3)
    n = len(lights)
    for max_per_tile in (4, 16, 64):
        counts = rng.integers(0, max_per_tile + 6, size=tiles_x * tiles_y).astype(np.uint32)      # some above the cap: saturated
        indices = rng.integers(0, n + 3, size=(tiles_x * tiles_y, max_per_tile)).astype(np.uint32)  # some >= light_count: skipped
        for _ in range(150):
            px, py = int(rng.integers(0, W)), int(rng.integers(0, H))
            k = int(rng.integers(0, n))
            P, N, V, albedo, metallic, roughness = _surface(rng, lights, k)
            tech = int(rng.integers(0, 2))
            a = glsl.light_loop(lights, counts, indices, tiles_x, tiles_y, max_per_tile, ts, px, py, H, P, N, V, albedo, metallic, roughness, tech)
            b = port.light_loop(lights, counts, indices, tiles_x, tiles_y, max_per_tile, ts, px, py, H, P, N, V, albedo, metallic, roughness, tech)
            assert np.array_equal(_bits(a), _bits(b)), f"pixel ({px}, {py}) cap {max_per_tile}: glsl {a} != restatement {b}"
    # culling_mode 0 of the shader (no lists) == a saturated list
    P, N, V, albedo, metallic, roughness = _surface(rng, lights, 0)
    every = glsl.light_loop(lights, np.zeros(tiles_x * tiles_y, np.uint32), np.zeros((tiles_x * tiles_y, 4), np.uint32), tiles_x, tiles_y, 4, ts, 5, 5, H,
                            P, N, V, albedo, metallic, roughness, 0, culling_mode=0)
    sat = port.light_loop(lights, np.full(tiles_x * tiles_y, 4, np.uint32), np.zeros((tiles_x * tiles_y, 4), np.uint32), tiles_x, tiles_y, 4, ts, 5, 5, H,
                          P, N, V, albedo, metallic, roughness, 0)
    assert np.array_equal(_bits(every), _bits(sat))


def test_restatement_equals_the_committed_glsl_fixture():
    """tests/golden/golden_a9_glsl.npz (written by tests/golden/make_golden_a9.py from the compiled GLSL): runs wherever the
    restatement runs, also where neither /root/reference nor oracle/_ref exists."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_a9_glsl.npz"))
    port = bindings.LocalLightEvaluator("port")
    recs, rows, want = g["records"], g["inputs"], g["radiance"]
    for r, w in zip(rows, want):
        f = r.astype(np.float32)
        got = port.eval_local_light(recs, int(r[0]), f[1:4], f[4:7], f[7:10], f[10:13], float(f[13]), float(f[14]), int(r[15]))
        assert np.array_equal(_bits(got), _bits(w)), (r, got, w)
    assert np.count_nonzero(np.any(want != 0, axis=1)) > 200
