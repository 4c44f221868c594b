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
