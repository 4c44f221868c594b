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


def procedural_scene(ref, alpha, tex_alpha=False):
    s = ref.procedural_scene()                      # oracle/binding.py: test_1.cpp build_scene() flavour, via the reference's builtins
    return configs.with_transparency(s, alpha, tex_alpha)


@pytest.mark.parametrize("alpha,layers,tex_alpha", [(255, 0, False), (255, 3, False), (100, 3, False), (100, 1, False), (30, 2, False),
                                                    (100, 2, True), (0, 3, True)])
@pytest.mark.parametrize("pose", ["POSE_PROCEDURAL", "POSE_LAYERS", "POSE_LAYERS_CLOSE"])
def test_procedural_scene_and_transparency_layers(ref, oracle, alpha, layers, tex_alpha, pose):
    """strips + fans + node scale + (for alpha < 255) the sorted transparency-layer insertion, flatten and blend;
    tex_alpha: texels decide per fragment whether the fragment is opaque (renderer.cpp:505)"""
    scene = procedural_scene(ref, alpha, tex_alpha)
    pose = getattr(configs, pose)
    vp = Viewport(0, 0, 400, 300, transparency_layers=layers)
    vp.camera.apply(pose)
    rpx, rz, rvs, o = both(ref, oracle, scene, vp, (400, 300), pose)
    assert (rz.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert (rpx == o["pixels"]).all()
    assert o["n_covered"] > 1000
    if pose is not configs.POSE_PROCEDURAL and layers > 0 and (alpha not in (0, 255) or tex_alpha):
        a = rpx >> 24
        assert ((a != 0) & (a != 255)).sum() > 500          # blended pixels are really in the frame
