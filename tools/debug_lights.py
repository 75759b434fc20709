"""Bisects a Forward+ colour mismatch between the CUDA path and the oracle by light type / attenuation model."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness
from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context
from oracle.bindings import Oracle

rec = np.load(os.path.join(ROOT, "tests", "golden", "golden_area_lights.npz"))["records"]
v = rec.view(scenes.LIGHT_DTYPE).reshape(-1)
gpu, port = Context(0), Oracle("port")
def run(sel, label):
    if not sel.any():
        return
    sd = scenes.scene_mixed_lights(rec[sel])
    g = harness.gpu_forward(gpu, sd, forward_plus=True)
    c = harness.cpu_forward(port, sd, forward_plus=True)
    d = np.abs(g.ldr.astype(np.int32) - c.ldr.astype(np.int32))
    dh = np.abs(g.hdr - c.hdr)[..., :3].max(axis=2)
    print(f"{label:40s} lights {int(sel.sum()):3d}  max LSB {int(d.max()):3d}  px>1LSB {int((d.max(axis=2) > 1).sum()):5d}  max |dHDR| {float(dh.max()):.4g}  lists equal {bool(np.array_equal(g.counts, c.counts))}")
    return d
t, model, flags = v["type_shape_flags"][:, 0], v["type_shape_flags"][:, 3], v["type_shape_flags"][:, 2]
run(np.ones(len(v), bool), "all")
for ty in (1, 2, 3, 4):
    run(t == ty, f"type {ty}")
    for m in (0, 1, 2):
        run((t == ty) & (model == m), f"type {ty} attenuation model {m}")
for i in range(len(v)):
    d = run(np.arange(len(v)) == i, f"light {i} type {t[i]} model {model[i]} flags {flags[i]} cutoff {v['shape_attenuation'][i, 3]:.3f} range {v['position_range'][i, 3]:.2f}")
