"""One-off wide fuzz of the rows added / restructured in round 2 (beyond the seeds the suite holds): flat draws, legacy L1-L3, light
selection from bins.  python tools/gpu_r2_widefuzz.py -> prints counts; exit code 1 on the first mismatch."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import fuzz_cases
import test_zz_gpu_flat_draw as zf
import test_zz_gpu_legacy2 as zl
import test_legacy2_cpu as t2
import test_legacy3_cpu as t3
from leisure_software_renderer_b200.renderer import Context
from oracle.bindings import FlatDraw, Legacy2Oracle, Legacy3Oracle

gpu = Context(0)
off = 0
for seed in range(40, 240):
    sc = fuzz_cases.flat_draw_scene(seed, dangling=seed % 2 == 1)
    t = zf.Targets(gpu, sc)
    for mode in (0, 1):
        t.reset()
        t.run(mode)
        off += zf.compare(t.read(), FlatDraw("port").run(sc, mode), f"flat seed {seed} mode {mode}")
    t.close()
print("flat draws: 200 more seeds x 2 modes equal; colour channels off by 1:", off)
for seed in range(24, 84):
    sc = fuzz_cases.legacy2_scene(seed)
    zl.check(zl.gpu_render(gpu, sc, with_shadow=seed % 6 != 5), t2.render(Legacy2Oracle("port"), sc, with_shadow=seed % 6 != 5), f"L2 seed {seed}")
    sc = fuzz_cases.legacy3_scene(seed)
    zl.check(zl.gpu_render(gpu, sc, with_shadow=seed % 6 != 5, pbr=True), t3.render(Legacy3Oracle("port"), sc, with_shadow=seed % 6 != 5), f"L3 seed {seed}")
print("legacy L2 / L3: 60 more seeds each within the gates")
gpu.close()
