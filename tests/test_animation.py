"""scene_t::animate (swegl/data/model.hpp:146-177) mirrored on the host side of the package (Scene.animate).

CPU: node TRS and the per-frame matrices after Scene.animate(t) equal the unmodified reference's after
scene.animate(t) + render(), bit for bit, on every bundled model that has animations (key-frame search, linear blend in
fp32, quaternion normalisation, matrix44_t::from_quaternion, node hierarchy product).
GPU: the frames of an animated sequence -- the static scene stays resident, only the 100-byte-per-node matrices travel
per frame -- match the oracle, through the package API and through the C++ drop-in, where the
reference's own scene.animate() runs on its own scene_t exactly as in src/test_1.cpp:374-378."""
import os

import numpy as np
import pytest

from swegl_b200 import _abi, configs
from swegl_b200.scene import Scene, Viewport

ANIMATED = ["BoxAnimated", "CesiumMilkTruck", "BrainStem"]
# inside a key-frame interval, exactly on a key, before the first / after the last key of a channel, wrapped by fmod
TIMES = [0.0, 0.01, 0.4, 1.25, 1.2500001, 2.0, 3.70833, 5.5, 17.3, 34.88, 100.0, 1234.567]


def fresh(name):
    """a private copy: animate() edits the node arrays"""
    return Scene.load_pack(os.path.join(configs.ASSETS, name + ".scenepack"))


@pytest.mark.parametrize("name", ANIMATED)
def test_animate_matches_reference(ref, name):
    scene = fresh(name)
    assert scene.n_animations >= 1
    h = ref.import_scene(scene)
    scr = ref.lib.ref_screen_new(32, 32)
    rv = ref.make_viewport(scr, Viewport(0, 0, 32, 32), [])
    for t in TIMES:
        ref.animate(h, t)
        ref.lib.ref_render(h, rv)                   # fills node_t::original_to_world_matrix
        scene.animate(t)
        rs, rr, rt = ref.node_trs(h, scene.n_nodes)
        assert (rs.view(np.uint32) == scene.node_scale.view(np.uint32)).all(), (t, "scale")
        assert (rr.view(np.uint32) == scene.node_rotation.view(np.uint32)).all(), (t, "rotation")
        assert (rt.view(np.uint32) == scene.node_translation.view(np.uint32)).all(), (t, "translation")
        rw, rn = ref.node_matrices(h, scene.n_nodes)
        w, n = scene.node_matrices()
        assert (rw.view(np.uint32) == w.view(np.uint32)).all(), t
        assert (rn.view(np.uint32) == n.view(np.uint32)).all(), t
    ref.lib.ref_viewport_free(rv); ref.lib.ref_screen_free(scr); ref.lib.ref_scene_free(h)


def test_animated_frames_oracle_equals_reference(ref, oracle):
    """whole frames of the animated truck (wheels turn): reference after its animate() == oracle after Scene.animate()"""
    scene = fresh("CesiumMilkTruck")
    scene.set_lights(0.3, (1, -2, -1), 0.7, configs.POINT_LIGHTS)
    vp = Viewport(0, 0, 480, 270)
    vp.camera.apply(configs.POSE_TEST1)
    h = ref.import_scene(scene)
    scr = ref.lib.ref_screen_new(480, 270)
    rv = ref.make_viewport(scr, vp, configs.POSE_TEST1)
    frames = []
    for t in (0.0, 0.3, 0.9):
        ref.animate(h, t)
        rpx, rz = ref.render(h, rv, scr, 480, 270, 480, 270)
        scene.animate(t)
        o = oracle.render(scene, vp, screen_wh=(480, 270))
        assert (rpx == o["pixels"]).all() and (rz.view(np.uint32) == o["z"].view(np.uint32)).all()
        frames.append(rpx)
    assert (frames[0] != frames[1]).any() and (frames[1] != frames[2]).any()      # it moves
    ref.lib.ref_viewport_free(rv); ref.lib.ref_screen_free(scr); ref.lib.ref_scene_free(h)


def test_scene_pack_without_animations_loads_empty():
    s = fresh("BoxTextured")
    assert s.n_animations == 0 and len(s.chan_node) == 0
    s.animate(1.0)                                   # a no-op


@pytest.mark.gpu
@pytest.mark.parametrize("name,res", [("CesiumMilkTruck", (960, 540)), ("BrainStem", (640, 360)), ("BoxAnimated", (320, 240))])
def test_animated_sequence_on_the_gpu(renderer, oracle, name, res):
    """upload once, then per frame: Scene.animate(t) -> begin_frame (node matrices + lights only) -> render"""
    scene = fresh(name)
    scene.set_lights(0.3, (1, -2, -1), 0.7, configs.POINT_LIGHTS)
    vp = Viewport(0, 0, *res)
    vp.camera.apply(configs.POSE_TEST1)
    renderer.upload_scene(scene)
    renderer.set_screen(*res)
    prev = None
    for t in (0.0, 0.21, 0.7, 1.9, 40.0):
        scene.animate(t)
        renderer.begin_frame(scene)
        px = np.zeros((res[1], res[0]), np.uint32)
        z = np.empty((res[1], res[0]), np.float32)
        renderer.render(vp, px, z)
        o = oracle.render(scene, vp, screen_wh=res)
        assert (z.view(np.uint32) == o["z"].view(np.uint32)).all(), t
        d = np.abs(px.view(np.uint8).astype(np.int16) - o["pixels"].view(np.uint8).astype(np.int16))
        assert d.max() <= 1, t
        if prev is not None and name == "CesiumMilkTruck":      # (BrainStem animates skin joints, which swegl ignores: gltf.cpp)
            assert (px != prev).any()
        prev = px


@pytest.mark.gpu
def test_animated_sequence_through_the_dropin(oracle):
    """src/test_1.cpp:374-378: swegl::render(scene, viewport); scene.animate(t) -- the reference's own animate() on its
    own scene_t, the replacement renderer underneath"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.binding import Ref, DROPIN_LIB
    if not os.path.exists(DROPIN_LIB):
        pytest.skip("oracle/_ref/libswegl_dropin.so not built")
    dropin = Ref(DROPIN_LIB)
    scene = fresh("CesiumMilkTruck")
    scene.set_lights(0.3, (1, -2, -1), 0.7, configs.POINT_LIGHTS)
    res = (960, 540)
    vp = Viewport(0, 0, *res)
    vp.camera.apply(configs.POSE_TEST1)
    h = dropin.import_scene(scene)
    scr = dropin.lib.ref_screen_new(*res)
    rv = dropin.make_viewport(scr, vp, configs.POSE_TEST1)
    for t in (0.0, 0.5, 1.1):
        dropin.animate(h, t)
        px, z = dropin.render(h, rv, scr, res[0], res[1], vp.w, vp.h)
        scene.animate(t)
        o = oracle.render(scene, vp, screen_wh=res)
        assert (z.view(np.uint32) == o["z"].view(np.uint32)).all(), t
        d = np.abs(px.view(np.uint8).astype(np.int16) - o["pixels"].view(np.uint8).astype(np.int16))
        assert d.max() <= 1, t
    dropin.lib.ref_viewport_free(rv); dropin.lib.ref_screen_free(scr); dropin.lib.ref_scene_free(h)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ANIMATED)
def test_device_side_animate_matrices_bit_exact(renderer, name):
    """swegl_b200_set_animation + begin_frame_animated(t): k_animate's node_world / node_normal == Scene.animate(t) +
    node_matrices() bit for bit (which test_animate_matches_reference pins against the unmodified reference), at every
    probe time, in shuffled order (the device state is a function of t alone), interleaved with host-side frames"""
    scene = fresh(name)
    scene.set_lights(0.3, (1, -2, -1), 0.7, configs.POINT_LIGHTS)
    vp = Viewport(0, 0, 64, 48)
    vp.camera.apply(configs.POSE_TEST1)
    renderer.upload_scene(scene)
    renderer.set_screen(64, 48)
    renderer.set_animation(scene)                    # base TRS = the scene as loaded
    host = fresh(name)
    order = TIMES[::2] + TIMES[1::2][::-1]
    for k, t in enumerate(order):
        renderer.begin_frame_animated(t, scene)
        renderer.render_device(vp, stats=(k % 2 == 0))          # direct launches and the captured-graph path
        w, n = renderer.read_node_matrices()
        host.animate(t)
        hw, hn = host.node_matrices()
        assert (w.view(np.uint32) == hw.view(np.uint32)).all(), (t, "node_world")
        assert (n.view(np.uint32) == hn.view(np.uint32)).all(), (t, "node_normal")
        if k == 3:                                   # a host-side frame in between does not disturb the device tables
            renderer.begin_frame(scene)
            renderer.render_device(vp, stats=True)
            w0, _ = renderer.read_node_matrices()
            assert (w0.view(np.uint32) == scene.node_matrices()[0].view(np.uint32)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name,res", [("CesiumMilkTruck", (960, 540)), ("BrainStem", (640, 360)), ("BoxAnimated", (320, 240))])
def test_device_side_animated_frames_match_the_oracle(renderer, oracle, name, res):
    """whole frames: begin_frame_animated(t) + render (blocking, and pipelined through render_async) == oracle on the scene
    after Scene.animate(t)"""
    scene = fresh(name)
    scene.set_lights(0.3, (1, -2, -1), 0.7, configs.POINT_LIGHTS)
    vp = Viewport(0, 0, *res)
    vp.camera.apply(configs.POSE_TEST1)
    renderer.upload_scene(scene)
    renderer.set_screen(*res)
    renderer.set_animation(scene)
    host = fresh(name)
    host.set_lights(0.3, (1, -2, -1), 0.7, configs.POINT_LIGHTS)
    images = [renderer.alloc_host((res[1], res[0]), np.uint32) for _ in range(2)]
    depths = [renderer.alloc_host((res[1], res[0]), np.float32) for _ in range(2)]
    times = (0.0, 0.21, 0.7, 1.9, 40.0)
    want = []
    for t in times:
        host.animate(t)
        want.append(oracle.render(host, vp, screen_wh=res))
    for t, o in zip(times, want):                    # blocking
        renderer.begin_frame_animated(t, scene)
        px = np.zeros((res[1], res[0]), np.uint32)
        z = np.empty((res[1], res[0]), np.float32)
        renderer.render(vp, px, z)
        assert (z.view(np.uint32) == o["z"].view(np.uint32)).all(), t
        assert np.abs(px.view(np.uint8).astype(np.int16) - o["pixels"].view(np.uint8).astype(np.int16)).max() <= 1, t
    tickets = []
    for k, t in enumerate(times):                    # two frames in flight
        if k >= 2:
            renderer.wait(tickets[k - 2])
            o = want[k - 2]
            assert (depths[k & 1].view(np.uint32) == o["z"].view(np.uint32)).all(), times[k - 2]
            assert np.abs(images[k & 1].view(np.uint8).astype(np.int16) - o["pixels"].view(np.uint8).astype(np.int16)).max() <= 1
        renderer.begin_frame_animated(t, scene)
        tickets.append(renderer.render_async(vp, images[k & 1], depths[k & 1]))
    for k in (len(times) - 2, len(times) - 1):
        renderer.wait(tickets[k])
        o = want[k]
        assert (depths[k & 1].view(np.uint32) == o["z"].view(np.uint32)).all()
        assert np.abs(images[k & 1].view(np.uint8).astype(np.int16) - o["pixels"].view(np.uint8).astype(np.int16)).max() <= 1


@pytest.mark.gpu
def test_device_side_animation_argument_checks(renderer):
    from swegl_b200.renderer import SweglB200Error
    scene = fresh("BoxAnimated")
    renderer.upload_scene(scene)
    renderer.set_screen(32, 32)
    with pytest.raises(SweglB200Error) as e:          # begin_frame_animated before set_animation (upload_scene dropped the tables)
        renderer.begin_frame_animated(0.5, scene)
    assert e.value.status == _abi.ERR_STATE
    bad = fresh("BoxAnimated")
    bad.node_parent = bad.node_parent.copy()
    bad.node_parent[0] = 0                            # a node that is its own parent
    with pytest.raises(SweglB200Error) as e:
        renderer.set_animation(bad)
    assert e.value.status == _abi.ERR_ARG
    bad = fresh("BoxAnimated")
    bad.node_parent = bad.node_parent.copy()
    bad.node_parent[0] = 10 ** 6                      # a parent index beyond the node array
    with pytest.raises(SweglB200Error) as e:
        renderer.set_animation(bad)
    assert e.value.status == _abi.ERR_ARG
    bad = fresh("BoxAnimated")
    bad.chan_n_steps = bad.chan_n_steps.copy()
    bad.chan_n_steps[0] = len(bad.step_time) + 5      # key frames beyond the arrays
    with pytest.raises(SweglB200Error) as e:
        renderer.set_animation(bad)
    assert e.value.status == _abi.ERR_ARG
