"""CPU: the host-side math that stays on the host in this design (camera, viewport matrix, node hierarchy,
make_sphere) restated in swegl_b200/scene.py, against the reference's own camera_t / viewport_t / node_t."""
import numpy as np
import pytest

from swegl_b200 import configs
from swegl_b200.scene import Scene, Viewport, normalized3


@pytest.mark.parametrize("rect,pose", [
    ((0, 0, 640, 480), configs.POSE_TEST1),
    ((0, 0, 1920, 1080), [("rotate_z", 0.3), ("translate", 0.1, -2, 7.5), ("rotate_x", 1.1), ("rotate_y", -2.9)]),
    ((100, 50, 300, 700), [("translate", -4, 0.25, 0.125), ("rotate_y", 0.77)]),      # aspect < 1, offset viewport
    ((7, 9, 333, 333), []),
])
def test_camera_and_viewport_matrices(ref, rect, pose):
    vp = Viewport(*rect)
    vp.camera.apply(pose)
    scr = ref.lib.ref_screen_new(rect[0] + rect[2], rect[1] + rect[3])
    rv = ref.make_viewport(scr, vp, pose)
    view, proj, cam, vpm = ref.viewport_matrices(rv)
    d = vp.desc()
    assert (view.view(np.uint32) == vp.camera.view.view(np.uint32)).all()
    assert (proj.view(np.uint32) == vp.camera.proj.view(np.uint32)).all()
    assert (cam.view(np.uint32) == vp.camera.center.view(np.uint32)).all()
    assert (vpm == np.array([d.vp_m00, d.vp_m03, d.vp_m11, d.vp_m13], np.float32)).all()
    ref.lib.ref_viewport_free(rv); ref.lib.ref_screen_free(scr)


@pytest.mark.parametrize("name", ["BoxTextured", "CesiumMilkTruck", "BrainStem", "BoxAnimated"])
def test_node_hierarchy_matrices(ref, name):
    scene = configs.load_scene(name)
    h = ref.import_scene(scene)
    scr = ref.lib.ref_screen_new(64, 64)
    vp = Viewport(0, 0, 64, 64)
    rv = ref.make_viewport(scr, vp, [])
    ref.lib.ref_render(h, rv)                      # fills node_t::original_to_world_matrix
    rw, rn = ref.node_matrices(h, scene.n_nodes)
    w, n = scene.node_matrices()
    assert (rw.view(np.uint32) == w.view(np.uint32)).all()
    assert (rn.view(np.uint32) == n.view(np.uint32)).all()
    ref.lib.ref_viewport_free(rv); ref.lib.ref_screen_free(scr); ref.lib.ref_scene_free(h)


def test_make_sphere_matches_reference(ref):
    s = configs.make_sphere_scene(37, 2.0, texture_size=8)
    h = ref.new_scene()
    ref.lib.ref_scene_add_material(h, 128, 128, 128, 255, 1.0, 1.0, 0, 0)
    t = s.textures[0]
    ref.lib.ref_scene_add_texture(h, t.ctypes.data, 8, 8)
    ref.lib.ref_scene_add_builtin(h, 3, 37, 2.0, 0, None, None, None)
    r = ref.export(h)
    for a in Scene.ARRAYS:
        x, y = getattr(s, a), getattr(r, a)
        assert x.shape == y.shape, a
        assert (np.ascontiguousarray(x).view(np.uint8) == np.ascontiguousarray(y).view(np.uint8)).all(), a
    ref.lib.ref_scene_free(h)


def test_sun_direction_normalisation(ref):
    s = Scene()
    for raw in [(1.0, -2.0, -1.0), (1.0, -1.0, -1.0), (0.0, 0.0, 0.0), (3e-20, 1e-19, -2e-20)]:
        s.set_lights(0.3, raw, 0.7)
        h = ref.new_scene()
        sun = ref.set_lights(h, s)
        assert (sun.view(np.uint32) == normalized3(*raw).view(np.uint32)).all()
        ref.lib.ref_scene_free(h)


def test_no_undefined_names_in_python_sources():
    """bench.py and the tools only run on the GPU box: catch NameErrors (no pyflakes in the image) before they cost a run"""
    import glob
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    for d in ("swegl_b200", "tools", "tests", "oracle"):
        files += glob.glob(os.path.join(root, d, "*.py"))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "check_names.py")] + files, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
