"""CPU property test of the sort-first draw culling (leisure_software_renderer_b200/csrc/host_math.hpp: bounds_touch_owned_rows,
sphere_may_touch_owned_rows): a draw they drop can never reach an owned tile row.  Pure host C++ (g++), no CUDA, no reference tree."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropped_draws_never_reach_owned_rows(tmp_path):
    exe = tmp_path / "host_cull_test"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "leisure_software_renderer_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "host_cull_test.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "violations 0" in r.stdout
