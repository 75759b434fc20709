"""The on-disk fixture readers (csrc/asset_loaders.hpp) on the CPU box, through a g++ build of the same header:
  * Wavefront OBJ: the reference's Suzanne fixture (cpp-folders/src/assets/obj/monkey/monkey.rawobj, when /root/reference is there) gives
    exactly leisure_software_renderer_b200/assets/suzanne.npz -- the arrays every parity test renders; a committed fixture covers
    quads / n-gons, missing vt / vn, relative indices and vertex re-use;
  * PNG: every colour type / bit depth the reader accepts decodes to the same RGBA8 as Pillow, with and without the vertical flip
    load_texture2d_sdl_image applies; interlaced and 16-bit files are refused."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "leisure_software_renderer_b200", "csrc")
MONKEY = "/root/reference/cpp-folders/src/assets/obj/monkey/monkey.rawobj"


@pytest.fixture(scope="module")
def lib():
    out, src, hdr = os.path.join(HERE, "cpp", "_build", "libasset_loaders_emul.so"), os.path.join(HERE, "cpp", "asset_loaders_emul.cpp"), os.path.join(CSRC, "asset_loaders.hpp")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + CSRC, src, "-o", out, "-lz"], check=True)
    lb = C.CDLL(out)
    lb.shsld_error.restype = C.c_char_p
    return lb


def load_obj(lib, path):
    n = (C.c_uint32 * 2)()
    rc = lib.shsld_load_obj(os.fsencode(path), n)
    if rc:
        raise RuntimeError(lib.shsld_error().decode())
    pos, nrm, uv, idx = np.zeros((n[0], 3), np.float32), np.zeros((n[0], 3), np.float32), np.zeros((n[0], 2), np.float32), np.zeros(n[1], np.uint32)
    lib.shsld_mesh_copy(pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p))
    return pos, nrm, uv, idx


def load_png(lib, path, flip):
    wh = (C.c_int32 * 2)()
    rc = lib.shsld_load_png(os.fsencode(path), int(flip), wh)
    if rc:
        raise RuntimeError(lib.shsld_error().decode())
    out = np.zeros((wh[1], wh[0], 4), np.uint8)
    lib.shsld_png_copy(out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.skipif(not os.path.exists(MONKEY), reason="reference fixture not present")
def test_suzanne_from_the_reference_fixture_equals_the_committed_arrays(lib):
    from leisure_software_renderer_b200 import scenes
    want = scenes.load_suzanne()
    pos, nrm, uv, idx = load_obj(lib, MONKEY)
    assert len(idx) == 967 * 3
    for got, key in ((pos, "positions"), (nrm, "normals"), (uv, "uvs")):
        assert np.array_equal(got.view(np.uint32), want[key].view(np.uint32)), key
    assert np.array_equal(idx, want["indices"])


def test_obj_fixture_quads_ngons_missing_attributes_relative_indices(lib):
    pos, nrm, uv, idx = load_obj(lib, os.path.join(HERE, "golden", "assets", "mixed.obj"))
    v = np.array([[0, 0, 0], [1.5, 0, 0.25], [1.5, 1, -0.125], [0, 1, 0.0625], [2.75, 0.5, 0.1], [2.25, 1.5, 3.0]], np.float32)
    # vertices in order of first use: quad 1/1/1 2/2/1 3/3/1 4/4/1 | pentagon 2/2/2 5//2 6//2 3/3/2 (3/3/1 = re-use) | 1 2 5 | 1/1 2/2 3/3
    want_pos = v[[0, 1, 2, 3, 1, 4, 5, 2, 0, 1, 4, 0, 1, 2]]
    assert np.array_equal(pos, want_pos)
    assert np.array_equal(idx, [0, 1, 2, 0, 2, 3, 4, 5, 6, 4, 6, 7, 4, 7, 2, 8, 9, 10, 11, 12, 13])
    assert np.array_equal(nrm[0], [0, 0, -1]) and np.array_equal(nrm[4], np.array([0.6, 0, -0.8], np.float32)) and np.array_equal(nrm[8], [0, 1, 0]) and np.array_equal(nrm[11], [0, 1, 0])
    assert np.array_equal(uv[5], [0, 0]) and np.array_equal(uv[7], [1, 1]) and np.array_equal(uv[8], [0, 0]) and np.array_equal(uv[12], [1, 0])
    with pytest.raises(RuntimeError, match="cannot read"):
        load_obj(lib, "/nonexistent/file.obj")


def test_png_every_supported_layout_equals_pillow(lib, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(3)
    w, h = 37, 23
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    cases = {"rgba": Image.fromarray(rgba, "RGBA"), "rgb": Image.fromarray(rgba[..., :3].copy(), "RGB"), "grey": Image.fromarray(rgba[..., 0].copy(), "L"),
             "grey_alpha": Image.fromarray(rgba[..., :2].copy(), "LA")}
    pal = Image.fromarray(rng.integers(0, 200, (h, w), dtype=np.uint8), "P")
    pal.putpalette([int(x) for x in rng.integers(0, 256, 768)])
    cases["palette"] = pal
    cases["bilevel"] = Image.fromarray((rgba[..., 1] > 127), "1")
    pal16 = Image.fromarray(rng.integers(0, 16, (h, w), dtype=np.uint8), "P")
    pal16.putpalette([int(x) for x in rng.integers(0, 256, 48)] + [0] * 720)
    cases["palette_4bit"] = pal16
    for name, im in cases.items():
        path = str(tmp_path / f"{name}.png")
        kw = {"transparency": bytes(range(255, 55, -1))} if name == "palette" else {}
        if name == "palette_4bit":
            kw["bits"] = 4
        im.save(path, **kw)
        want = np.asarray(Image.open(path).convert("RGBA"))
        assert np.array_equal(load_png(lib, path, False), want), name
        assert np.array_equal(load_png(lib, path, True), want[::-1]), name + " (flipped)"
    path = str(tmp_path / "interlaced.png")
    # Pillow cannot write interlaced files: flip the IHDR interlace byte of a valid file (the CRC is not checked before the refusal)
    raw = bytearray(open(str(tmp_path / "rgb.png"), "rb").read())
    raw[28] = 1
    open(path, "wb").write(bytes(raw))
    with pytest.raises(RuntimeError, match="interlaced"):
        load_png(lib, path, False)
    Image.fromarray(rng.integers(0, 65536, (h, w), dtype=np.uint16)).save(str(tmp_path / "grey16.png"))
    with pytest.raises(RuntimeError, match="unsupported"):
        load_png(lib, str(tmp_path / "grey16.png"), False)
    with pytest.raises(RuntimeError, match="not a PNG"):
        load_png(lib, os.path.join(HERE, "golden", "assets", "mixed.obj"), False)
