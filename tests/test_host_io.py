"""Host-side file formats either side of the hot path: the product's own
Radiance .hdr reader and BMP writer (cpuperformanceraytracer_b200/host/image_io.cpp) against the
reference's loader/writer (asset_loading.cpp via stb, run through oracle/_ref/ref_asset_tool) and
against self-made files (always available)."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "cpuperformanceraytracer_b200", "libdemofox_b200.so")
TEX_DIR = "/root/reference/Textures"


@pytest.fixture(scope="module")
def io():
    if not os.path.exists(HOSTLIB):
        from cpuperformanceraytracer_b200 import build
        build.build()
    L = ctypes.CDLL(HOSTLIB)
    L.b200pt_io_load_hdr.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_float)),
                                     ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.b200pt_io_load_cubemap.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.POINTER(ctypes.c_float)),
                                         ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.b200pt_io_write_bmp32.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
    L.b200pt_io_free.argtypes = [ctypes.c_void_p]
    return L


def load_hdr(io, path):
    data, w, h = ctypes.POINTER(ctypes.c_float)(), ctypes.c_int(), ctypes.c_int()
    rc = io.b200pt_io_load_hdr(str(path).encode(), ctypes.byref(data), ctypes.byref(w), ctypes.byref(h))
    if rc != 0:
        return None
    a = np.ctypeslib.as_array(data, shape=(h.value, w.value, 3)).copy()
    io.b200pt_io_free(data)
    return a


def load_cubemap(io, paths):
    """six face files (px nx py ny pz nz) -> the (6*face, face, 3) atlas LoadCubemapTexture builds, or None"""
    arr = (ctypes.c_char_p * 6)(*[str(f).encode() for f in paths])
    data, cw, ch = ctypes.POINTER(ctypes.c_float)(), ctypes.c_int(), ctypes.c_int()
    if io.b200pt_io_load_cubemap(arr, ctypes.byref(data), ctypes.byref(cw), ctypes.byref(ch)) != 0:
        return None
    a = np.ctypeslib.as_array(data, shape=(ch.value, cw.value, 3)).copy()
    io.b200pt_io_free(data)
    return a


def rgbe_encode(img):
    """float (H, W, 3) top-down -> RGBE bytes (H, W, 4), standard Radiance encoding."""
    m = img.max(axis=2)
    e = np.zeros_like(m, dtype=np.int32)
    nz = m > 1e-32
    mant, ex = np.frexp(m[nz])
    scale = np.zeros_like(m)
    scale[nz] = mant * 256.0 / m[nz]
    e[nz] = ex + 128
    out = np.zeros(img.shape[:2] + (4,), dtype=np.uint8)
    out[..., :3] = np.clip(img * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(nz, e, 0).astype(np.uint8)
    return out


def write_hdr(path, rgbe, rle):
    h, w, _ = rgbe.shape
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\n# made by tests\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(("-Y %d +X %d\n" % (h, w)).encode())
        if not rle:
            f.write(rgbe.tobytes())
            return
        for j in range(h):
            f.write(bytes([2, 2, w >> 8, w & 0xFF]))
            for k in range(4):
                ch = rgbe[j, :, k]
                i = 0
                while i < w:
                    run = 1
                    while i + run < w and run < 127 and ch[i + run] == ch[i]:
                        run += 1
                    if run >= 4:
                        f.write(bytes([128 + run, int(ch[i])]))
                        i += run
                    else:
                        lit = min(w - i, 64)
                        f.write(bytes([lit]) + ch[i:i + lit].tobytes())
                        i += lit


def decode_rgbe(rgbe):
    f = np.ldexp(np.float32(1.0), rgbe[..., 3].astype(np.int32) - 136).astype(np.float32)
    out = rgbe[..., :3].astype(np.float32) * f[..., None]
    out[rgbe[..., 3] == 0] = 0
    return out.astype(np.float32)


@pytest.mark.parametrize("w,h,rle", [(16, 5, True), (64, 9, True), (7, 4, False), (40, 6, False)])
def test_hdr_reader_self_made(io, tmp_path, w, h, rle):
    rng = np.random.default_rng(w * h)
    img = (rng.random((h, w, 3)) ** 4 * 100).astype(np.float32)
    img[0, :3] = 0.0  # exponent byte 0 -> black
    img[1, 2:12] = img[1, 2]  # a run
    rgbe = rgbe_encode(img)
    path = tmp_path / "t.hdr"
    write_hdr(path, rgbe, rle)
    got = load_hdr(io, path)
    assert got is not None and got.shape == (h, w, 3)
    # vertical flip on load: row 0 = bottom (asset_loading.cpp:12)
    assert np.array_equal(got, decode_rgbe(rgbe)[::-1])


def test_hdr_reader_rejects_garbage(io, tmp_path):
    p = tmp_path / "bad.hdr"
    p.write_bytes(b"P6\n1 1\n255\n\0\0\0")
    assert load_hdr(io, p) is None
    p.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_xyze\n\n-Y 1 +X 1\n\0\0\0\0")
    assert load_hdr(io, p) is None
    assert load_hdr(io, tmp_path / "missing.hdr") is None
    # a hostile header: 16M x 16M pixels with 4 bytes of data -> rejected before any allocation
    p.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 16777216 +X 16777216\n\0\0\0\0")
    assert load_hdr(io, p) is None
    # larger than the renderer's 2^24-float texel indexing
    p.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 4096 +X 4096\n" + b"\0" * 64)
    assert load_hdr(io, p) is None
    # truncated pixel data: flat (fewer than 4 bytes per pixel) and RLE (a scanline cut short) fail instead of zero-filling
    p.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 4\n" + b"\x10\x20\x30\x80" * 7)
    assert load_hdr(io, p) is None
    good = tmp_path / "good.hdr"
    write_hdr(good, rgbe_encode(np.full((4, 16, 3), 0.5, dtype=np.float32)), rle=True)
    b = good.read_bytes()
    assert load_hdr(io, good) is not None
    p.write_bytes(b[:-9])
    assert load_hdr(io, p) is None


def test_bmp_writer_layout(io, tmp_path):
    w, h = 5, 3
    px = (np.arange(w * h, dtype=np.uint32) * np.uint32(0x01030507)) | np.uint32(0xFF000000)
    p = tmp_path / "o.bmp"
    assert io.b200pt_io_write_bmp32(str(p).encode(), w, h, px.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))) == 0
    b = p.read_bytes()
    pad = (-w * 3) & 3
    assert b[:2] == b"BM" and len(b) == 14 + 40 + (w * 3 + pad) * h
    size, _, _, off = struct.unpack_from("<IHHI", b, 2)
    assert (size, off) == (len(b), 54)
    hs, bw, bh, planes, bpp, comp = struct.unpack_from("<IiiHHI", b, 14)
    assert (hs, bw, bh, planes, bpp, comp) == (40, w, h, 1, 24, 0)
    # first stored row = bottom image row, bytes B, G, R (opaque alpha: no compositing change)
    first = int(px.reshape(h, w)[h - 1, 0])
    assert tuple(b[54:57]) == ((first >> 16) & 0xFF, (first >> 8) & 0xFF, first & 0xFF)


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_asset_tool")) or
                               not os.path.isdir(TEX_DIR), reason="needs oracle/_ref/ref_asset_tool and the reference textures")


@needs_ref
def test_hdr_reader_matches_reference_loader(io, tmp_path):
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_asset_tool")
    for name in ("HDR_040_Field_Env.hdr", "px.hdr", "nz.hdr"):
        src = os.path.join(TEX_DIR, name)
        out = tmp_path / "ref.f32"
        w, h, c = map(int, subprocess.run([tool, "equirect", src, str(out)], check=True, capture_output=True, text=True).stdout.split())
        ref = np.fromfile(out, dtype=np.float32).reshape(h, w, 3)
        got = load_hdr(io, src)
        assert got is not None and np.array_equal(got, ref), name


@needs_ref
def test_cubemap_atlas_matches_reference_loader(io, tmp_path):
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_asset_tool")
    faces = [os.path.join(TEX_DIR, f + ".hdr") for f in ("px", "nx", "py", "ny", "pz", "nz")]
    out = tmp_path / "cube.f32"
    w, h, c = map(int, subprocess.run([tool, "cubemap"] + faces + [str(out)], check=True, capture_output=True, text=True).stdout.split())
    ref = np.fromfile(out, dtype=np.float32).reshape(h, w, 3)
    arr = (ctypes.c_char_p * 6)(*[f.encode() for f in faces])
    data, cw, ch = ctypes.POINTER(ctypes.c_float)(), ctypes.c_int(), ctypes.c_int()
    assert io.b200pt_io_load_cubemap(arr, ctypes.byref(data), ctypes.byref(cw), ctypes.byref(ch)) == 0
    got = np.ctypeslib.as_array(data, shape=(ch.value, cw.value, 3)).copy()
    io.b200pt_io_free(data)
    assert (cw.value, ch.value) == (w, h) and np.array_equal(got, ref)


@needs_ref
def test_bmp_writer_matches_reference_writer(io, tmp_path):
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_asset_tool")
    w, h = 37, 11
    rng = np.random.default_rng(5)
    px = rng.integers(0, 2 ** 32, size=w * h, dtype=np.uint64).astype(np.uint32)  # arbitrary alpha too
    raw = tmp_path / "in.rgba"
    px.tofile(raw)
    ref = tmp_path / "ref.bmp"
    subprocess.run([tool, "writebmp", str(raw), str(w), str(h), str(ref)], check=True)
    mine = tmp_path / "mine.bmp"
    assert io.b200pt_io_write_bmp32(str(mine).encode(), w, h, px.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))) == 0
    assert mine.read_bytes() == ref.read_bytes()
