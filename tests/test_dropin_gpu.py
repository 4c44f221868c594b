"""GPU: the drop-in boundary.  oracle/_ref/libswegl_dropin.so is the UNMODIFIED reference (scene model, camera,
viewport, pixel-shader classes, swegl::render template) with exactly one translation unit swapped:
swegl_b200/host/renderer_b200.cpp defines swegl::_render on top of the C ABI / CUDA path instead of
src/render/renderer.cpp.  The frame that swegl::render() leaves in SDL_Surface::pixels and viewport_t::m_zbuffer
must match the CPU oracle: depth bit-exact, colour within 1 LSB."""
import os

import numpy as np
import pytest

from swegl_b200 import _abi, configs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dropin():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.binding import Ref, DROPIN_LIB
    if not os.path.exists(DROPIN_LIB):
        pytest.skip("oracle/_ref/libswegl_dropin.so not built (needs /root/reference at build time)")
    return Ref(DROPIN_LIB)


@pytest.mark.parametrize("name,light,tex", [("box_640", 2, 2), ("truck_1080", 2, 2), ("truck_1080", 1, 1), ("sphere100_1080", 2, 2)])
def test_swegl_render_through_the_dropin(dropin, oracle, name, light, tex):
    scene, vps, screen, cfg = configs.build(name, light_mode=light, tex_mode=tex)
    vp = vps[0]
    h = dropin.import_scene(scene)
    scr = dropin.lib.ref_screen_new(*screen)
    rv = dropin.make_viewport(scr, vp, vp.pose)
    for _ in range(2):                                   # second frame reuses the uploaded scene
        px, z = dropin.render(h, rv, scr, screen[0], screen[1], vp.w, vp.h)
    o = oracle.render(scene, vp, screen_wh=screen)
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    d = np.abs(px.view(np.uint8).astype(np.int16) - o["pixels"].view(np.uint8).astype(np.int16))
    assert d.max() <= 1
    dropin.lib.ref_viewport_free(rv); dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)


def test_four_viewports_one_call(dropin, oracle):
    """swegl::render(scene, vp1, vp2, vp3, vp4) (renderer.hpp:20-34) on a shared surface"""
    scene, vps, screen, cfg = configs.build("multiview_1080")
    h = dropin.import_scene(scene)
    scr = dropin.lib.ref_screen_new(*screen)
    rvs = [dropin.make_viewport(scr, vp, vp.pose) for vp in vps]
    dropin.lib.ref_render4(h, *rvs)
    import ctypes as C
    px = np.ctypeslib.as_array(C.cast(dropin.lib.ref_screen_pixels(scr), C.POINTER(C.c_uint32)), shape=(screen[1], screen[0])).copy()
    opx = np.zeros_like(px)
    for vp in vps:
        oracle.render(scene, vp, screen_wh=screen, pixels=opx)
    d = np.abs(px.view(np.uint8).astype(np.int16) - opx.view(np.uint8).astype(np.int16))
    assert d.max() <= 1


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5, 8, 60])
def test_fuzz_scene_through_the_dropin(dropin, oracle, seed):
    """a seeded triangle soup (configs.fuzz_scene: node chain with mirrored scales, all index modes, near-plane clipping)
    built in the reference's own scene_t and rendered by swegl::render() with the replacement renderer"""
    scene, vp, screen, pose = configs.fuzz_case(seed)
    vp.post_mode = _abi.POST_NULL
    h = dropin.import_scene(scene)
    scr = dropin.lib.ref_screen_new(*screen)
    rv = dropin.make_viewport(scr, vp, pose)
    px, z = dropin.render(h, rv, scr, screen[0], screen[1], vp.w, vp.h)
    o = oracle.render(scene, vp, screen_wh=screen)
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    d = np.abs(px.view(np.uint8).astype(np.int16) - o["pixels"].view(np.uint8).astype(np.int16))
    assert d.max() <= 1
    dropin.lib.ref_viewport_free(rv); dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)
