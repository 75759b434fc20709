"""Generates tests/golden/golden_legacy.npz by running THE REFERENCE's OWN legacy demo code (oracle/_ref/libshs_legacy_ref.so =
hello_pipeline_blinn_phong_shading.cpp + shs_renderer.hpp compiled where they lie under /root/reference) on BASELINE configs[0] as
the demo sets it up and on two fuzz scenes of tests/fuzz_cases.py.  Run in the container that has /root/reference; the committed
fixture lets the GPU box (which has no reference tree) check the oracle and the CUDA path against reference output."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import fuzz_cases  # noqa: E402
from oracle.bindings import LegacyOracle  # noqa: E402
from test_legacy_cpu import c1_inputs, render  # noqa: E402

CASES = {"c1_640x480": c1_inputs, "fuzz7": lambda: fuzz_cases.legacy_draws(7), "fuzz30": lambda: fuzz_cases.legacy_draws(30)}


def main():
    ref = LegacyOracle("reference")
    out = {}
    for name, make in CASES.items():
        canvas, z = render(ref, *make())
        out[name + "_canvas"], out[name + "_z"] = canvas, z
        print(name, canvas.shape, int((z < np.finfo(np.float32).max).sum()), "px covered")
    np.savez_compressed(os.path.join(HERE, "golden_legacy.npz"), **out)


if __name__ == "__main__":
    main()
