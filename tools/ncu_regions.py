"""Aggregates executed warp-instructions and stall samples of tile_raster.cu by code region."""
import csv, io, subprocess, sys, re

rep = sys.argv[1]
src_path = "/root/repo/leisure_software_renderer_b200/csrc/tile_raster.cu"
lines = open(src_path).read().split("\n")
# region boundaries by marker text -> name
markers = [("struct Surface", "helpers"), ("eval_brdf_lit(const Surface", "brdf"), ("attenuation_quadratic(float", "attenuation(generic)"),
           ("accumulate_point_spot(const Surface", "point_spot(range,N.d,atten)"), ("eval_light_record(const Surface", "light_record"), ("fake_ibl(V3", "fake_ibl"),
           ("sample_texture(const", "texture"), ("shadow_visibility(const", "shadow"), ("tonemap_pixel(float", "tonemap"), ("resolve_uncovered(const", "resolve_uncovered"),
           ("tile_kernel(const FrameConst", "prologue"), ("// ---------------- empty tiles", "empty_tiles"), ("uint32_t ord = blockIdx.x, cls = 0;", "prologue"),
           ("for (uint32_t base = off0", "raster_loop"), ("// ---------------- resolve: depth", "resolve_depth_counters"),
           ("// ---------------- phase A", "phaseA_shade"), ("// the only barrier every non-empty tile passes", "barrier_stats"), ("// ---------------- phase B", "phaseB_stage_lights"),
           ("const SmLight* lt = s_light;", "phaseB_light_loop"), ("// generic light-tile size", "generic_light_tiles"),
           ("// ---------------- phase C", "phaseC_resolve"), ("light_prep_kernel(", "other")]
bounds = []
for text, name in markers:
    for i, l in enumerate(lines):
        if text.strip() in l:
            bounds.append((i + 1, name)); break
bounds.sort()
def region(ln):
    r = "top"
    for b, n in bounds:
        if ln >= b: r = n
    return r
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = {}; cur = "?"; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples"); continue
    if hdr is None or len(r) < 8 or r[2] != "-": continue
    try: ln = int(r[0]); inst = int(r[ie] or 0); s = int(r[ss] or 0)
    except ValueError: continue
    key = region(ln) if cur == "tile_raster.cu" else cur
    a = agg.setdefault(key, [0, 0]); a[0] += inst; a[1] += s
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} inst {i/ti*100:5.1f}%  samples {s/ts*100:5.1f}%")
