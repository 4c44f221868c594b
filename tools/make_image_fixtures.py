"""Small PNG / JPEG fixtures for tests/test_image_decode.py, written with PIL (libpng / libjpeg-turbo) together with the
digest of the texels PIL decodes them to (tests/golden/images/MANIFEST.json).  Run in the build container:
    python tools/make_image_fixtures.py
"""
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swegl_b200.scene import decode_image_bgra, texel_digest  # noqa: E402


def main():
    from PIL import Image
    out = os.path.join(ROOT, "tests", "golden", "images")
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(3)
    h, w = 45, 61                                       # not multiples of 8 or 16: partial MCUs on both edges
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 5 + yy) % 256, (yy * 7) % 256, (xx * yy // 3) % 256], -1).astype(np.uint8)
    noise = rng.integers(0, 256, base.shape, dtype=np.uint8)
    rgb = ((base.astype(int) * 3 + noise) // 4).astype(np.uint8)
    img = Image.fromarray(rgb, "RGB")
    gray = img.convert("L")
    files = {}

    def jpg(name, im, **kw):
        b = io.BytesIO(); im.save(b, "JPEG", **kw); files[name + ".jpg"] = b.getvalue()

    def png(name, im, **kw):
        b = io.BytesIO(); im.save(b, "PNG", **kw); files[name + ".png"] = b.getvalue()
    jpg("baseline_444_q90", img, quality=90, subsampling=0)
    jpg("baseline_422_q85", img, quality=85, subsampling=1)
    jpg("baseline_420_q75", img, quality=75, subsampling=2)
    jpg("baseline_420_q30_optimized", img, quality=30, subsampling=2, optimize=True)
    jpg("baseline_444_q100", img, quality=100, subsampling=0)
    jpg("progressive_444_q95", img, quality=95, subsampling=0, progressive=True)
    jpg("progressive_422_q70", img, quality=70, subsampling=1, progressive=True)
    jpg("progressive_420_q80", img, quality=80, subsampling=2, progressive=True)
    jpg("gray_q80", gray, quality=80)
    jpg("gray_progressive_q80", gray, quality=80, progressive=True)
    jpg("restart_420", img, quality=80, subsampling=2, restart_marker_blocks=3)
    jpg("restart_progressive_420", img, quality=80, subsampling=2, progressive=True, restart_marker_rows=1)
    jpg("one_pixel_wide", img.crop((0, 0, 1, 9)), quality=80, subsampling=0)
    png("rgb", img)
    png("rgba", Image.fromarray(np.dstack([rgb, noise[..., 0]]), "RGBA"))
    png("gray", gray)
    png("gray_alpha", Image.fromarray(np.dstack([np.asarray(gray), noise[..., 1]]), "LA"))
    pal = img.convert("P", palette=Image.ADAPTIVE, colors=37)
    png("palette", pal)
    png("palette_trns", pal, transparency=5)
    png("palette_4bit", img.convert("P", palette=Image.ADAPTIVE, colors=13), bits=4)
    png("bilevel", img.convert("1"))
    png("rgb_trns", img, transparency=tuple(int(v) for v in rgb[5, 5]))
    manifest = {}
    for name, data in sorted(files.items()):
        with open(os.path.join(out, name), "wb") as f:
            f.write(data)
        t = decode_image_bgra(data)
        manifest[name] = {"w": int(t.shape[1]), "h": int(t.shape[0]), "sha256": texel_digest(t), "bytes": len(data)}
    # 16-bit PNG: libpng's png_set_strip_16 keeps the high byte (image.cpp:116-117); PIL converts differently, so the
    # expectation is written down explicitly
    g16 = (np.asarray(gray).astype(np.uint16) * 257 + 13).astype(np.uint16)
    b = io.BytesIO(); Image.fromarray(g16).save(b, "PNG")
    with open(os.path.join(out, "gray_16bit.png"), "wb") as f:
        f.write(b.getvalue())
    hi = (g16 >> 8).astype(np.uint32)
    t = hi | hi << 8 | hi << 16 | np.uint32(0xFF000000)
    manifest["gray_16bit.png"] = {"w": w, "h": h, "sha256": texel_digest(t.astype(np.uint32)), "bytes": len(b.getvalue())}
    with open(os.path.join(out, "MANIFEST.json"), "w") as f:
        json.dump({"_comment": "texel digests (swegl_b200.scene.texel_digest) of the fixtures as PIL = libpng / libjpeg-turbo decodes "
                               "them; made by tools/make_image_fixtures.py", "images": manifest}, f, indent=1)
    print(len(manifest), "fixtures,", sum(v["bytes"] for v in manifest.values()), "bytes")


if __name__ == "__main__":
    main()
