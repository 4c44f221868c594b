"""CPU: scene packs (assets/*.scenepack) load, carry the reference loader's quirks, and detect decoder drift."""
import numpy as np
import pytest

from swegl_b200 import configs
from swegl_b200.scene import Scene


def test_packs_load_with_expected_sizes():
    expect = {"BoxTextured": (24, 12, 1), "CesiumMilkTruck": (3995, 2856, 1), "BrainStem": (34159, 61666, 0)}   # SURVEY §8 sizes
    for name, (nv, nt, ntex) in expect.items():
        s = configs.load_scene(name)
        assert (s.n_vertices, s.n_triangles(), len(s.textures)) == (nv, nt, ntex)
        assert int(s.indices.max()) < 65536                      # gltf.cpp:219 reads uint16 indices
    assert configs.load_scene("CesiumMilkTruck").textures[0].shape == (2048, 2048)


def test_pack_roundtrip_and_digest(tmp_path):
    s = configs.make_sphere_scene(6, texture_size=4)
    p = str(tmp_path / "s.scenepack")
    s.save_pack(p)
    b = Scene.load_pack(p)
    for a in Scene.ARRAYS:
        assert (getattr(s, a) == getattr(b, a)).all()
    assert (s.textures[0] == b.textures[0]).all()
    # corrupt the stored digest -> decoder-drift guard fires
    import json, zipfile
    q = str(tmp_path / "bad.scenepack")
    with zipfile.ZipFile(p) as zin, zipfile.ZipFile(q, "w") as zout:
        for item in zin.infolist():
            data = zin.read(item.filename)
            if item.filename == "meta.json":
                m = json.loads(data); m["textures"][0]["sha256"] = "0" * 64; data = json.dumps(m).encode()
            zout.writestr(item, data)
    with pytest.raises(ValueError):
        Scene.load_pack(q)
