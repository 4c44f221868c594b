"""CPU: Scene.load_glb (the reference's glTF loader restated, src/data/gltf.cpp:56-436, with the package's own PNG /
JPEG decoder) against the scene packs, which tools/make_scenepacks.py exported from the reference's OWN loader.  Every
array, texel and animation key must be identical.  Needs the bundled .glb files of the reference checkout (they are
not copied into this repository): skipped where /root/reference is absent."""
import os

import numpy as np
import pytest

from swegl_b200 import configs
from swegl_b200.scene import Scene

RESOURCES = os.environ.get("SWEGL_RESOURCES", "/root/reference/resources")
MODELS = ["BoxTextured", "CesiumMilkTruck", "BrainStem", "BoxAnimated", "box"]


@pytest.mark.parametrize("name", MODELS)
def test_load_glb_equals_the_reference_loader(name):
    path = os.path.join(RESOURCES, name + ".glb")
    if not os.path.exists(path):
        pytest.skip(f"{path} not present")
    got = Scene.load_glb(path)
    want = Scene.load_pack(os.path.join(configs.ASSETS, name + ".scenepack"))
    for a in Scene.ARRAYS + Scene.ANIM_ARRAYS:
        x, y = getattr(got, a), getattr(want, a)
        assert x.shape == y.shape, a
        assert x.dtype == y.dtype, a
        assert (x.view(np.uint32) == y.view(np.uint32)).all() if x.dtype == np.float32 else (x == y).all(), a
    assert len(got.textures) == len(want.textures)
    for t0, t1 in zip(got.textures, want.textures):
        assert (t0 == t1).all()


def test_gltf_with_external_buffer():
    path = os.path.join(RESOURCES, "Box.gltf")
    if not os.path.exists(path):
        pytest.skip(f"{path} not present")
    s = Scene.load_glb(path)
    assert s.n_triangles() == 12 and s.n_vertices == 24
