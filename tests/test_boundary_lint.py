"""Boundary lint, in the spirit of the reference's own tools/check_vop_boundaries.sh (grep rules on what a layer may include):
* include/shsb.h is a plain-C ABI: C headers only, extern "C", no C++ / CUDA / torch types in any signature;
* the product (package sources, kernels, the C++ bindings) never includes, imports, links or loads anything under oracle/ --
  the oracle is test infrastructure -- and has no CPU rendering fallback to route through;
* the reference-side bindings use only the reference's headers and the C-ABI (no CUDA headers, no dynamic_cast policy branching);
* libshsb.so does not depend on the oracle libraries."""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "leisure_software_renderer_b200")


def _read(p):
    with open(p, encoding="utf-8", errors="replace") as f:
        return f.read()


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def test_c_abi_header_is_plain_c():
    text = _read(os.path.join(ROOT, "include", "shsb.h"))
    code = _strip_comments(text)
    assert set(re.findall(r"#include\s+[<\"]([^>\"]+)[>\"]", code)) == {"stddef.h", "stdint.h"}
    assert 'extern "C"' in code and "#ifdef __cplusplus" in code
    for banned in ("std::", "torch", "at::Tensor", "cudaStream_t", "cuda_runtime", "template", "class ", "namespace", "&)"):
        assert banned not in code, f"include/shsb.h mentions {banned!r}"
    protos = [p for p in re.findall(r"SHSB_API\s+[^;#]+;", code) if "(" in p]
    assert len(protos) >= 40
    for p in protos:   # every entry point returns a status code (or a const char* for the two string getters)
        assert re.match(r"SHSB_API\s+(int32_t|const char\*)\s+shsb_", p), p


def test_product_never_touches_the_oracle():
    files = glob.glob(os.path.join(PKG, "*.py")) + glob.glob(os.path.join(PKG, "csrc", "*")) + glob.glob(os.path.join(PKG, "host", "shs_b200", "*"))
    assert len(files) >= 15
    for p in files:
        text = _read(p)
        code = _strip_comments(text) if not p.endswith(".py") else re.sub(r"#[^\n]*", "", text)
        assert not re.search(r"^\s*(from|import)\s+oracle\b", code, flags=re.M), f"{p} imports the oracle"
        assert not re.search(r"#include\s+[<\"][^>\"]*oracle", code), f"{p} includes an oracle header"
        assert "liboracle" not in code and "libshs_ref" not in code and "libshs_legacy_ref" not in code, f"{p} names an oracle library"
    build_py = _read(os.path.join(PKG, "build.py"))
    assert "oracle" not in build_py, "the product build must not compile or link oracle sources"


def test_bindings_use_only_reference_headers_and_the_c_abi():
    for p in glob.glob(os.path.join(PKG, "host", "shs_b200", "*.hpp")):
        code = _strip_comments(_read(p))
        for inc in re.findall(r"#include\s+[<\"]([^>\"]+)[>\"]", code):
            ok = inc.startswith("shs/") or inc.startswith("shs_b200/") or inc in ("shsb.h", "shs_renderer.hpp") or "/" not in inc and "." not in inc
            assert ok, f"{p} includes {inc}"
        assert not re.search(r"\bcuda[A-Z_]\w*\s*\(|<cuda|cuda_runtime|__global__|<<<", code), f"{p} reaches past the C-ABI into CUDA"
        assert not re.search(r"dynamic_cast\s*<", code), f"{p}: policy branching by dynamic_cast (the reference's own boundary rule)"


def test_library_does_not_link_the_oracle():
    lib = os.path.join(PKG, "libshsb.so")
    if not os.path.exists(lib):
        from leisure_software_renderer_b200 import build
        build.build()
    out = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
    assert "oracle" not in out and "shs_ref" not in out, out
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in syms.splitlines() if " T " in l]
    assert exported and all(s.startswith("shsb_") for s in exported), [s for s in exported if not s.startswith("shsb_")][:10]
