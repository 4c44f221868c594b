"""CPU: swegl_b200_decode_image (host/image_decode.cpp) against the texels libpng / libjpeg-turbo produce.

The reference reads embedded images with libpng / libjpeg (src/misc/image.cpp:93-258); the decoder here has to give the
same texels from the same bytes.  Golden: the digests stored in assets/*.scenepack and tests/golden/images/MANIFEST.json
were computed from PIL's decode (libpng / libjpeg-turbo) when those files were made (tools/make_scenepacks.py,
tools/make_image_fixtures.py); where PIL is importable the texels are also compared directly."""
import ctypes as C
import json
import os
import zipfile

import numpy as np
import pytest

from swegl_b200 import _abi, configs
from swegl_b200.scene import Scene, decode_image, texel_digest

IMAGES = os.path.join(os.path.dirname(__file__), "golden", "images")
MANIFEST = json.load(open(os.path.join(IMAGES, "MANIFEST.json")))["images"]


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_fixture_decodes_to_the_golden_texels(name):
    data = open(os.path.join(IMAGES, name), "rb").read()
    t = decode_image(data)
    want = MANIFEST[name]
    assert t.shape == (want["h"], want["w"]) and t.dtype == np.uint32
    assert texel_digest(t) == want["sha256"]


@pytest.mark.parametrize("pack", ["BoxTextured", "CesiumMilkTruck"])
def test_bundled_model_textures(pack):
    """the PNG of BoxTextured.glb and the 2048x2048 progressive JPEG of CesiumMilkTruck.glb, as embedded in the .glb"""
    with zipfile.ZipFile(os.path.join(configs.ASSETS, pack + ".scenepack")) as z:
        meta = json.loads(z.read("meta.json"))
        for i, entry in enumerate(meta["textures"]):
            assert entry["encoding"] == "image"
            t = decode_image(z.read(f"image_{i}.bin"))
            assert t.shape == (entry["h"], entry["w"])
            assert texel_digest(t) == entry["sha256"]
    s = Scene.load_pack(os.path.join(configs.ASSETS, pack + ".scenepack"))      # load_pack decodes with decode_image
    assert len(s.textures) == len(meta["textures"])


def test_against_pil_directly():
    pil = pytest.importorskip("PIL.Image")
    from swegl_b200.scene import decode_image_bgra
    for name in sorted(MANIFEST):
        if name == "gray_16bit.png":                    # PIL's 16 -> 8 conversion is not libpng's strip_16 (see the fixture script)
            continue
        data = open(os.path.join(IMAGES, name), "rb").read()
        assert (decode_image(data) == decode_image_bgra(data)).all(), name


def test_alpha_channel_and_byte_order():
    t = decode_image(open(os.path.join(IMAGES, "rgb.png"), "rb").read())
    assert ((t >> 24) == 255).all()                     # filler 0xFF (image.cpp:137-139)
    a = decode_image(open(os.path.join(IMAGES, "rgba.png"), "rb").read())
    assert ((a >> 24) != 255).any()
    assert ((a & 0x00FFFFFF) == (t & 0x00FFFFFF)).all() # same colour bytes, b in the low byte (colors.hpp:9-18)
    j = decode_image(open(os.path.join(IMAGES, "baseline_444_q100.jpg"), "rb").read())
    d = np.abs(j.view(np.uint8).astype(int) - t.view(np.uint8).astype(int))
    assert d.max() <= 6                                 # quality 100 JPEG of the same picture: same channel order


@pytest.mark.parametrize("data", [b"", b"GIF89a........", b"\x89PNG\r\n\x1a\n" + b"\0" * 20, b"\xff\xd8\xff\xd9",
                                  b"\xff\xd8\xff\xc3\x00\x0b\x08\x00\x10\x00\x10\x01\x01\x11\x00\xff\xd9"])
def test_malformed_input_is_an_error_not_a_crash(data):
    lib = _abi.load()
    ptr, w, h = C.POINTER(C.c_uint32)(), C.c_int32(), C.c_int32()
    rc = lib.swegl_b200_decode_image(data, len(data), C.byref(ptr), C.byref(w), C.byref(h))
    assert rc == _abi.ERR_UNSUPPORTED and lib.swegl_b200_image_error()
    with pytest.raises(ValueError):
        decode_image(data)


def test_truncated_files_do_not_crash():
    """every prefix of a progressive JPEG and of a PNG either decodes (missing scans leave zeros) or reports an error"""
    for name in ("progressive_420_q80.jpg", "palette_trns.png", "restart_420.jpg"):
        data = open(os.path.join(IMAGES, name), "rb").read()
        for cut in range(0, len(data), 97):
            try:
                decode_image(data[:cut])
            except ValueError:
                pass


def _png(w, h, ctype, depth, samples, interlace, plte=None):
    """a PNG file from `samples` (h, w, channels) written by hand (filter type 0 everywhere), one pass or Adam7 -- PIL cannot
    write interlaced files"""
    import struct
    import zlib

    def chunk(ty, body):
        return struct.pack(">I", len(body)) + ty + body + struct.pack(">I", zlib.crc32(ty + body) & 0xFFFFFFFF)

    def pack_row(row):                       # row: (pw, channels) integer samples
        flat = row.reshape(-1)
        if depth == 8:
            return bytes(flat.astype(np.uint8))
        if depth == 16:
            return flat.astype(">u2").tobytes()
        bits = "".join(format(int(v), f"0{depth}b") for v in flat)
        bits += "0" * (-len(bits) % 8)
        return bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))

    passes = [(0, 0, 1, 1)] if not interlace else [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
    raw = b""
    for x0, y0, dx, dy in passes:
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        for row in sub:
            raw += b"\0" + pack_row(row)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if plte is not None:
        out += chunk(b"PLTE", bytes(plte))
    return out + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (5, 3), (8, 8), (9, 7), (37, 21), (64, 33)])
@pytest.mark.parametrize("ctype,depth", [(6, 8), (2, 8), (2, 16), (0, 1), (0, 4), (4, 8), (3, 2), (3, 8)])
def test_adam7_interlaced_png(w, h, ctype, depth):
    """Adam7 (PNG spec 8.2): the seven reduced images, each with its own scanlines and filter bytes, decode to the same
    texels as the one-pass file of the same samples (libpng's png_read_image de-interlaces transparently, image.cpp:93-170);
    sizes below 8 leave some passes empty"""
    rng = np.random.default_rng(w * 1000 + h * 10 + ctype + depth)
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    plte = list(rng.integers(0, 256, 3 * (1 << min(depth, 8)))) if ctype == 3 else None
    samples = rng.integers(0, 1 << depth, (h, w, channels))
    a = decode_image(_png(w, h, ctype, depth, samples, False, plte))
    b = decode_image(_png(w, h, ctype, depth, samples, True, plte))
    assert a.shape == (h, w) and (a == b).all()
    pil = pytest.importorskip("PIL.Image")
    if depth != 16:                                   # (PIL's 16 -> 8 conversion is not libpng's strip_16)
        from swegl_b200.scene import decode_image_bgra
        assert (b == decode_image_bgra(_png(w, h, ctype, depth, samples, True, plte))).all()


def test_png_header_that_promises_more_than_the_data_holds_is_refused_early():
    """a few-byte file with a 32768 x 32768 RGBA16 header must fail before the decoder allocates the 8.6 GB it asks for"""
    import resource
    import struct
    import zlib

    def chunk(ty, body):
        return struct.pack(">I", len(body)) + ty + body + struct.pack(">I", zlib.crc32(ty + body) & 0xFFFFFFFF)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 32768, 32768, 16, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    with pytest.raises(ValueError):
        decode_image(data)
    assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - before < 512 * 1024      # KiB: nowhere near the declared size
