"""CPU: the C restatement against the unmodified reference itself (oracle/_ref/libswegl_ref.so), bit for bit.
Skipped where the reference library is not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

from swegl_b200 import _abi, configs
from swegl_b200.scene import Scene, Viewport


def both(ref, oracle, scene, vp, screen, pose):
    import ctypes as C
    h = ref.import_scene(scene)
    scr = ref.lib.ref_screen_new(*screen)
    rv = ref.make_viewport(scr, vp, pose)
    rpx, rz = ref.render(h, rv, scr, screen[0], screen[1], vp.w, vp.h)
    rvs = ref.vertex_state(h, scene.n_vertices)
    ref.lib.ref_viewport_free(rv); ref.lib.ref_screen_free(scr); ref.lib.ref_scene_free(h)
    o = oracle.render(scene, vp, screen_wh=screen, want_vertices=True)
    return rpx, rz, rvs, o


@pytest.mark.parametrize("name", ["box_640", "box_640_close", "truck_1080", "sphere100_1080", "brainstem_4k"])
def test_configs_bit_identical(ref, oracle, name):
    scene, vps, screen, cfg = configs.build(name)
    rpx, rz, rvs, o = both(ref, oracle, scene, vps[0], screen, vps[0].pose)
    assert (rpx == o["pixels"]).all()
    assert (rz.view(np.uint32) == o["z"].view(np.uint32)).all()
    for k in ("v_world", "v_viewport", "normal_world"):
        assert (rvs[k].view(np.uint32) == o[k].view(np.uint32)).all(), k
    assert (rvs["yes"] == o["yes"]).all()


@pytest.mark.parametrize("light,tex", [(l, t) for l in (0, 1, 2) for t in (0, 1, 2)])
def test_every_shader_combination(ref, oracle, light, tex):
    """pixel_shader_t / lights_flat / lights_phong x plain / nearest / bilinear (pixel_shaders.hpp:15-179)"""
    scene, vps, screen, cfg = configs.build("truck_1080", light_mode=light, tex_mode=tex)
    vp = Viewport(0, 0, 480, 270, light_mode=light, tex_mode=tex, transparency_layers=0)
    vp.camera.apply(vps[0].pose)
    rpx, rz, rvs, o = both(ref, oracle, scene, vp, (480, 270), vps[0].pose)
    assert (rpx == o["pixels"]).all()
    assert (rz.view(np.uint32) == o["z"].view(np.uint32)).all()


@pytest.mark.parametrize("pose", [
    [("translate", 0, 0.5, -0.9), ("rotate_y", 0.4)],                     # camera inside the scene: near-plane clipping
    [("translate", 1, 2, -5), ("rotate_y", -0.2), ("rotate_x", -0.3), ("rotate_z", 0.5)],
    [("translate", -3, 1, 2), ("rotate_y", 2.2)],
    [("translate", 0, 8, 0), ("rotate_x", -1.5)],
])
def test_near_clip_and_odd_poses(ref, oracle, pose):
    scene, vps, screen, cfg = configs.build("truck_1080")
    vp = Viewport(0, 0, 640, 360, transparency_layers=0)
    vp.camera.apply(pose)
    rpx, rz, rvs, o = both(ref, oracle, scene, vp, (640, 360), pose)
    assert (rpx == o["pixels"]).all()
    assert (rz.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert (rvs["yes"] == o["yes"]).all()


def procedural_scene(ref, alpha):
    """test_1.cpp's build_scene() flavour: torus (strips), cube (fans, scaled), sphere, triangles; optional
    transparent materials (the transparency-layer path, renderer.cpp:500-550)."""
    from swegl_b200.scene import lcg_texture
    h = ref.new_scene()
    tex = lcg_texture(64, seed=99)
    ref.lib.ref_scene_add_texture(h, tex.ctypes.data, 64, 64)
    mats = [(128, 128, 128, 255, 0), (128, 128, 255, 255, -1), (255, 128, 255, 255, -1),
            (128, 128, 255, alpha, -1), (128, 255, 128, alpha, -1), (255, 128, 128, alpha, -1)]
    for b, g, r, a, t in mats:
        ref.lib.ref_scene_add_material(h, b, g, r, a, 1.0, 1.0, t, 0)
    import ctypes as C
    f3 = lambda *v: (C.c_float * 3)(*v)
    ref.lib.ref_scene_add_builtin(h, 2, 24, 1.0, 0, None, f3(0, 0, 0.5), f3(0, 0, -2.5))       # tore
    ref.lib.ref_scene_add_builtin(h, 1, 0, 1.0, 0, f3(2, 1, 1), None, f3(0, 0, 0))              # cube, scale.x = 2
    ref.lib.ref_scene_add_builtin(h, 3, 16, 2.0, 2, None, None, f3(3, 0, -1))                   # sphere
    ref.lib.ref_scene_add_builtin(h, 0, 0, 1.0, 3, None, None, f3(1, 0.5, 2.1))                 # tri
    ref.lib.ref_scene_add_builtin(h, 0, 0, 1.0, 4, None, None, f3(1, 0.5, 2.0))
    ref.lib.ref_scene_add_builtin(h, 0, 0, 1.0, 5, None, None, f3(1, 0.5, 2.2))
    s = ref.export(h, "procedural")
    ref.lib.ref_scene_free(h)
    s.set_lights(0.2, (1.0, -1.0, -1.0), 0.3, configs.POINT_LIGHTS)
    return s


@pytest.mark.parametrize("alpha,layers", [(255, 0), (255, 3), (100, 3), (100, 1), (30, 2)])
def test_procedural_scene_and_transparency_layers(ref, oracle, alpha, layers):
    """strips + fans + node scale + (for alpha < 255) the sorted transparency-layer insertion, flatten and blend"""
    scene = procedural_scene(ref, alpha)
    pose = [("translate", 1, 2, 6), ("rotate_y", 3.0), ("rotate_x", -0.3)]
    vp = Viewport(0, 0, 400, 300, transparency_layers=layers)
    vp.camera.apply(pose)
    rpx, rz, rvs, o = both(ref, oracle, scene, vp, (400, 300), pose)
    assert (rz.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert (rpx == o["pixels"]).all()
    assert o["n_covered"] > 1000
