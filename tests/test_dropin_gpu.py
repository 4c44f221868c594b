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


def _surface(dropin, scr, screen):
    import ctypes as C
    return np.ctypeslib.as_array(C.cast(dropin.lib.ref_screen_pixels(scr), C.POINTER(C.c_uint32)), shape=(screen[1], screen[0])).copy()


def _close(px, opx):
    return np.abs(px.view(np.uint8).astype(np.int16) - opx.view(np.uint8).astype(np.int16)).max() <= 1


def test_dof_through_the_dropin(dropin, oracle):
    """src/test_1.cpp:355: the viewport carries a post_shader_depth_box; the adapter's dynamic_cast branch selects DoF-R
    (swegl_b200_adapter.hpp describe()), the frame in SDL_Surface::pixels is the oracle's DoF-R frame"""
    from swegl_b200.scene import Viewport
    scene, vps, screen, cfg = configs.build("truck_1080")
    vp = vps[0]
    vp.post_mode, vp.focal_distance, vp.focal_depth = _abi.POST_DOF, 5.0, 5.0
    try:
        h = dropin.import_scene(scene)
        scr = dropin.lib.ref_screen_new(*screen)
        rv = dropin.make_viewport(scr, vp, vp.pose, with_dof=True)
        px, z = dropin.render(h, rv, scr, screen[0], screen[1], vp.w, vp.h)
        o = oracle.render(scene, vp, screen_wh=screen)
        vp.post_mode = _abi.POST_NULL
        plain = oracle.render(scene, vp, screen_wh=screen)
    finally:
        vp.post_mode = _abi.POST_NULL
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert _close(px, o["pixels"])
    assert (o["pixels"] != plain["pixels"]).sum() > 1000          # the post pass really blurred something
    dropin.lib.ref_viewport_free(rv); dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)


@pytest.mark.parametrize("alpha,layers,tex_alpha", [(100, 3, False), (30, 2, True), (200, 1, False)])
def test_transparency_layers_through_the_dropin(dropin, oracle, alpha, layers, tex_alpha):
    """viewport_t(…, transparency_layers) + materials with alpha < 255: m_got_transparency is set by the reference's own
    viewport constructor, the adapter forwards the layer count, the device resolves and flattens (renderer.cpp:500-550,
    viewport.cpp:43-86)"""
    from swegl_b200.scene import Viewport
    scene = configs.procedural(alpha, tex_alpha)
    screen = (640, 480)
    vp = Viewport(0, 0, *screen, transparency_layers=layers)
    vp.camera.apply(configs.POSE_LAYERS)
    h = dropin.import_scene(scene)
    scr = dropin.lib.ref_screen_new(*screen)
    rv = dropin.make_viewport(scr, vp, configs.POSE_LAYERS)
    px, z = dropin.render(h, rv, scr, screen[0], screen[1], vp.w, vp.h)
    o = oracle.render(scene, vp, screen_wh=screen)
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert _close(px, o["pixels"])
    vp0 = Viewport(0, 0, *screen, transparency_layers=0)
    vp0.camera.apply(configs.POSE_LAYERS)
    assert (oracle.render(scene, vp0, screen_wh=screen)["pixels"] != px).sum() > 100      # the layers matter
    dropin.lib.ref_viewport_free(rv); dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)


@pytest.mark.parametrize("depth", [1, 3])
def test_cpp_pipeline_host(dropin, oracle, depth):
    """swegl_b200::pipeline_t (swegl_b200_host.hpp): 5 frames of a turning camera round robin over `depth` contexts; the last
    one, collected into the viewport's surface, is the oracle's frame of the final camera"""
    scene, vps, screen, cfg = configs.build("truck_1080")
    vp = vps[0]
    h = dropin.import_scene(scene)
    scr = dropin.lib.ref_screen_new(*screen)
    rv = dropin.make_viewport(scr, vp, vp.pose)
    frames, dyaw = 5, 0.02
    dropin.host("ref_render_pipelined", h, rv, depth, frames, dyaw)
    px = _surface(dropin, scr, screen)
    from swegl_b200.scene import Viewport
    ovp = Viewport(vp.x, vp.y, vp.w, vp.h)
    ovp.camera.apply(list(vp.pose) + [("rotate_y", dyaw)] * frames)
    o = oracle.render(scene, ovp, screen_wh=screen)
    assert _close(px, o["pixels"])
    dropin.lib.ref_viewport_free(rv); dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)


@pytest.mark.parametrize("name,n_ctx,dof", [("truck_1080", 2, False), ("truck_1080", 3, True), ("sphere100_1080", 4, False)])
def test_cpp_sharded_host_row_bands(dropin, oracle, name, n_ctx, dof):
    """swegl_b200::sharded_renderer_t::render(scene, vp): ONE frame in n_ctx row bands (contexts of device 0 stand in for
    GPUs), assembled in context 0's screen by the frame protocol, three frames in a row; the surface and the depth buffer
    the caller gets back are the oracle's full frame"""
    scene, vps, screen, cfg = configs.build(name)
    vp = vps[0]
    if dof:
        vp.post_mode, vp.focal_distance, vp.focal_depth = _abi.POST_DOF, 5.0, 5.0
    try:
        h = dropin.import_scene(scene)
        scr = dropin.lib.ref_screen_new(*screen)
        rv = dropin.make_viewport(scr, vp, vp.pose, with_dof=dof)
        dropin.host("ref_render_sharded", h, rv, n_ctx, 3)
        px = _surface(dropin, scr, screen)
        import ctypes as C
        z = np.ctypeslib.as_array(C.cast(dropin.lib.ref_viewport_zbuffer(rv), C.POINTER(C.c_float)), shape=(vp.h, vp.w)).copy()
        o = oracle.render(scene, vp, screen_wh=screen)
    finally:
        vp.post_mode = _abi.POST_NULL
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert _close(px, o["pixels"])
    dropin.lib.ref_viewport_free(rv); dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)


@pytest.mark.parametrize("n_ctx", [2, 4])
def test_cpp_sharded_host_viewport_per_context(dropin, oracle, n_ctx):
    """swegl::render(scene, vp1, vp2, vp3, vp4) (renderer.hpp:20-34) with viewport v on context v mod n_ctx"""
    scene, vps, screen, cfg = configs.build("multiview_1080")
    h = dropin.import_scene(scene)
    scr = dropin.lib.ref_screen_new(*screen)
    rvs = [dropin.make_viewport(scr, vp, vp.pose) for vp in vps]
    dropin.host("ref_render_sharded4", h, *rvs, n_ctx)
    px = _surface(dropin, scr, screen)
    opx = np.zeros_like(px)
    for vp in vps:
        oracle.render(scene, vp, screen_wh=screen, pixels=opx)
    assert _close(px, opx)
    for rv in rvs:
        dropin.lib.ref_viewport_free(rv)
    dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)


def test_device_side_animation_through_the_cpp_host(dropin, oracle):
    """swegl_b200::render_animated(scene, t, viewport): the reference's scene_t with its animations flattened by the adapter,
    animate + hierarchy product on the device; equals the oracle's frame of the scene after Scene.animate(t).  The host
    scene_t is never animated here."""
    import os
    from swegl_b200.scene import Scene, Viewport
    scene = Scene.load_pack(os.path.join(configs.ASSETS, "CesiumMilkTruck.scenepack"))
    scene.set_lights(0.3, (1, -2, -1), 0.7, configs.POINT_LIGHTS)
    res = (960, 540)
    vp = Viewport(0, 0, *res)
    vp.camera.apply(configs.POSE_TEST1)
    h = dropin.import_scene(scene)
    scr = dropin.lib.ref_screen_new(*res)
    rv = dropin.make_viewport(scr, vp, configs.POSE_TEST1)
    import ctypes as C
    prev = None
    for t in (0.0, 0.5, 1.1, 7.3):
        dropin.host("ref_render_animated", h, rv, t)
        px = _surface(dropin, scr, res)
        z = np.ctypeslib.as_array(C.cast(dropin.lib.ref_viewport_zbuffer(rv), C.POINTER(C.c_float)), shape=(vp.h, vp.w)).copy()
        scene.animate(t)
        o = oracle.render(scene, vp, screen_wh=res)
        assert (z.view(np.uint32) == o["z"].view(np.uint32)).all(), t
        assert _close(px, o["pixels"]), t
        assert prev is None or (px != prev).any()
        prev = px
    dropin.lib.ref_viewport_free(rv); dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)
