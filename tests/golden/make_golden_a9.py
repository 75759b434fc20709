"""Writes tests/golden/golden_a9_glsl.npz: outputs of the reference's Forward+ fragment-shader text compiled as C++
(oracle/_ref/libshs_glsl_a9_ref.so, built from /root/reference by `make -C oracle ref`) on seeded inputs -- per-light radiance
for the 60 reference-packed light records of golden_area_lights.npz and for fuzzed point / spot lights.  Run in the container
that has /root/reference; the fixture then pins oracle.cpp's row A9 wherever the tests run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from leisure_software_renderer_b200 import scenes  # noqa: E402
from oracle import bindings  # noqa: E402
from test_a9_pinned_cpu import _golden_lights, _surface  # noqa: E402


def main():
    glsl = bindings.LocalLightEvaluator("glsl")
    rng = np.random.default_rng(2024)
    sets = [_golden_lights()] + [scenes.make_lights(16, 16, (-6, 0.2, -6), (6, 3.0, 6), seed=100 + s) for s in range(4)]
    recs, rows, outs = [], [], []
    base = 0
    for lights in sets:
        recs.append(np.ascontiguousarray(lights).view(np.uint8).reshape(-1, 160))
        for k in range(len(lights)):
            for _ in range(10):
                P, N, V, albedo, metallic, roughness = _surface(rng, lights, k)
                tech = int(rng.integers(0, 2))
                out = glsl.eval_local_light(lights, k, P, N, V, albedo, metallic, roughness, tech)
                rows.append(np.concatenate([[base + k], P, N, V, albedo, [metallic, roughness, tech]]).astype(np.float64))
                outs.append(out)
        base += len(lights)
    out_path = os.path.join(ROOT, "tests", "golden", "golden_a9_glsl.npz")
    np.savez_compressed(out_path, records=np.concatenate(recs), inputs=np.array(rows), radiance=np.array(outs, dtype=np.float32))
    print("wrote", out_path, len(rows), "samples,", int(np.count_nonzero(np.any(np.array(outs) != 0, axis=1))), "lit")


if __name__ == "__main__":
    main()
