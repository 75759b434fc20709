"""Builds leisure_software_renderer_b200/libshsb.so (sm_100a only) with nvcc, in-tree.

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the .so travels to the GPU box.
Per-file flags:
  geometry.cu, light_cull.cu   --fmad=false   (every expression decides coverage / depth / list bits)
  post_passes.cu               --fmad=false   (RGBA8 outputs of the post passes are bit-exact)
  scene_cull.cu                --fmad=false   (object classes / light selections are bit-exact)
  flat_draw.cu                 --fmad=false   (flat-shaded draws: depth bit-exact, colours the reference's expressions)
  legacy.cu, legacy2.cu        --fmad=false   (the legacy demo variants: every expression is the reference's, unfused)
  tile_raster.cu, binning.cu   FMA allowed; exact expressions use __fmul_rn/__fadd_rn/__fdiv_rn explicitly
  api.cu                       host code with -ffp-contract=off (host float math must equal the reference's)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libshsb.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "--expt-relaxed-constexpr"]
SOURCES = {
    "geometry.cu": ["--fmad=false"],
    "light_cull.cu": ["--fmad=false"],
    "binning.cu": [],
    "post_passes.cu": ["--fmad=false"],
    "legacy.cu": ["--fmad=false"],
    "legacy2.cu": ["--fmad=false"],
    "scene_cull.cu": ["--fmad=false"],
    "flat_draw.cu": ["--fmad=false"],
    "tile_raster.cu": [],
    "gather.cu": [],
    "api.cu": [],
}


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "shsb.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [cc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", o]
        if ptxas_info:
            cmd += ["-Xptxas", "-v"]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src}\n{out}\n")
        elif (verbose or ptxas_info) and out.strip():
            print(f"--- {src}\n{out}")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [cc, *ARCH, "-shared", "-o", OUT, *objs, "-Xcompiler", "-fPIC", "-lz"]  # zlib: PNG fixtures (asset_loaders.hpp)
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


def build_phase_clocks() -> str:
    """Debug variant libshsb_clk.so: tile_raster.cu with -DSHSB_PHASE_CLOCKS (tools/phase_clocks.py, via SHSB_LIB)."""
    build()
    cc = nvcc()
    o = os.path.join(OBJ, "tile_raster_clk.o")
    subprocess.run([cc, *ARCH, *COMMON, "-DSHSB_PHASE_CLOCKS", "-c", os.path.join(CSRC, "tile_raster.cu"), "-o", o], check=True)
    o2 = os.path.join(OBJ, "legacy2_clk.o")  # per-CTA clocks of the legacy render-target demos' raster kernel (tools/l2_clocks.py)
    subprocess.run([cc, *ARCH, *COMMON, *SOURCES["legacy2.cu"], "-DSHSB_PHASE_CLOCKS", "-c", os.path.join(CSRC, "legacy2.cu"), "-o", o2], check=True)
    out = os.path.join(HERE, "libshsb_clk.so")
    objs = [os.path.join(OBJ, f.replace(".cu", ".o")) for f in SOURCES if f not in ("tile_raster.cu", "legacy2.cu")] + [o, o2]
    subprocess.run([cc, *ARCH, "-shared", "-o", out, *objs, "-Xcompiler", "-fPIC", "-lz"], check=True)
    return out


def build_no_stores() -> str:
    """Debug variant libshsb_nostore.so: tile_raster.cu with -DSHSB_NO_STORES (tools/gpu_store_bound.sh, via SHSB_LIB)."""
    build()
    cc = nvcc()
    o = os.path.join(OBJ, "tile_raster_nostore.o")
    subprocess.run([cc, *ARCH, *COMMON, "-DSHSB_NO_STORES", "-c", os.path.join(CSRC, "tile_raster.cu"), "-o", o], check=True)
    out = os.path.join(HERE, "libshsb_nostore.so")
    objs = [os.path.join(OBJ, f.replace(".cu", ".o")) for f in SOURCES if f != "tile_raster.cu"] + [o]
    subprocess.run([cc, *ARCH, "-shared", "-o", out, *objs, "-Xcompiler", "-fPIC", "-lz"], check=True)
    return out


def build_variant(tag: str, defines: list) -> str:
    """Experiment variant libshsb_<tag>.so: EVERY translation unit recompiled with extra -D flags (tools/gpu_ab_lib.sh, via SHSB_LIB)."""
    cc = nvcc()
    obj = os.path.join(OBJ, tag)
    os.makedirs(obj, exist_ok=True)
    procs, objs = [], []
    for src, extra in SOURCES.items():
        o = os.path.join(obj, src.replace(".cu", ".o"))
        procs.append((src, subprocess.Popen([cc, *ARCH, *COMMON, *extra, *defines, "-c", os.path.join(CSRC, src), "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    out = os.path.join(HERE, f"libshsb_{tag}.so")
    subprocess.run([cc, *ARCH, "-shared", "-o", out, *objs, "-Xcompiler", "-fPIC", "-lz"], check=True)
    return out


if __name__ == "__main__":
    if "--tile-h8" in sys.argv:
        print(build_variant("h8", ["-DSHSB_TILE_H=8"]))
        sys.exit(0)
    if "--tile-h4" in sys.argv:
        print(build_variant("h4", ["-DSHSB_TILE_H=4"]))
        sys.exit(0)
    for n in (10, 12):
        if f"--nolight-ctas{n}" in sys.argv:
            print(build_variant(f"nl{n}", [f"-DSHSB_NOLIGHT_CTAS={n}"]))
            sys.exit(0)
    for n in (6, 7, 9, 10, 11, 12, 14):
        if f"--tile-ctas{n}" in sys.argv:
            print(build_variant(f"ctas{n}", [f"-DSHSB_FPLUS_CTAS={n}"]))
            sys.exit(0)
    for n in (4, 6, 12, 16):
        if f"--macro{n}" in sys.argv:
            print(build_variant(f"macro{n}", [f"-DSHSB_MACRO={n}"]))
            sys.exit(0)
    for n in (8, 10, 12):
        if f"--geom-ctas{n}" in sys.argv:
            print(build_variant(f"geom{n}", [f"-DSHSB_GEOM_CTAS={n}"]))
            sys.exit(0)
    if "--cand2" in sys.argv:
        print(build_variant("cand2", ["-DSHSB_CAND_PER_THREAD=2"]))
        sys.exit(0)
    if "--cand8" in sys.argv:
        print(build_variant("cand8", ["-DSHSB_CAND_PER_THREAD=8"]))
        sys.exit(0)
    if "--tile-h16" in sys.argv:
        print(build_variant("h16", ["-DSHSB_TILE_H=16"]))
        sys.exit(0)
    if "--no-stores" in sys.argv:
        print(build_no_stores())
        sys.exit(0)
    if "--phase-clocks" in sys.argv:
        print(build_phase_clocks())
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv))
