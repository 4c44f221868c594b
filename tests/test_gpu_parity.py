"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars: vertex stage, yes flags, depth buffer, coverage, DoF-R: bit / pixel identical.
Shaded colour: within +-1 LSB per 8-bit channel, alpha exact (tolerance for pow()/sqrt() double-vs-float
paths) -- and we additionally report how many pixels are not bit-identical.
"""
import json
import os

import numpy as np
import pytest

from swegl_b200 import _abi, configs

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "MANIFEST.json")


def channel_diff(a, b):
    a8 = a.view(np.uint8).reshape(a.shape + (4,)).astype(np.int16)
    b8 = b.view(np.uint8).reshape(b.shape + (4,)).astype(np.int16)
    return np.abs(a8 - b8)


def render_gpu(renderer, scene, vps, screen, want_z=True):
    renderer.upload_scene(scene)
    renderer.set_screen(*screen)
    renderer.begin_frame(scene)
    px = np.zeros((screen[1], screen[0]), np.uint32)
    zs, stats = [], []
    for vp in vps:
        z = np.empty((vp.h, vp.w), np.float32)
        stats.append(renderer.render(vp, px, z))
        zs.append(z)
    return px, zs, stats


def render_oracle(oracle, scene, vps, screen, want_vertices=False):
    px = np.zeros((screen[1], screen[0]), np.uint32)
    outs = [oracle.render(scene, vp, screen_wh=screen, pixels=px, want_vertices=want_vertices) for vp in vps]
    return px, outs


def check_frame(name, gpx, gzs, opx, outs, vps):
    for vp, gz, o in zip(vps, gzs, outs):
        assert (gz.view(np.uint32) == o["z"].view(np.uint32)).all(), f"{name}: depth buffer differs"
    d = channel_diff(gpx, opx)
    assert d[..., 3].max() == 0, f"{name}: alpha differs"
    assert d.max() <= 1, f"{name}: colour differs by more than 1 LSB (max {d.max()})"
    return int((gpx != opx).sum())


SMALL = ["box_640", "box_640_close", "truck_1080_sun", "truck_1080", "sphere100_1080", "multiview_1080", "layers_640", "layers_texalpha_640"]
LARGE = ["brainstem_4k", "truck_4k", "multiview_4k", "sphere1000_8k"]       # sphere1000_8k = BASELINE.json config 5 at its stated size
_ORACLE_FRAMES = {}


def oracle_frame(oracle, name):
    """the oracle's frame of a named workload, rendered once per session (the 8K sphere takes seconds)"""
    if name not in _ORACLE_FRAMES:
        scene, vps, screen, cfg = configs.build(name)
        _ORACLE_FRAMES[name] = render_oracle(oracle, scene, vps, screen, want_vertices=True)
    return _ORACLE_FRAMES[name]


@pytest.fixture
def exact_renderer(renderer):
    """the session renderer switched to bit-exact Phong lighting for one test"""
    renderer.set_shading(_abi.SHADING_EXACT)
    yield renderer
    renderer.set_shading(_abi.SHADING_FAST)


@pytest.mark.parametrize("shading", ["exact", "fast"])
@pytest.mark.parametrize("name", SMALL + LARGE)
def test_frame_matches_oracle(renderer, oracle, name, shading):
    """exact: colour bit-identical to the oracle (hence the golden FNV, which is pinned against the reference);
    fast (the default): within +-1 LSB per channel, alpha exact.  Depth, coverage and vertex state: identical in both."""
    scene, vps, screen, cfg = configs.build(name)
    renderer.set_shading(_abi.SHADING_EXACT if shading == "exact" else _abi.SHADING_FAST)
    try:
        gpx, gzs, stats = render_gpu(renderer, scene, vps, screen)
    finally:
        renderer.set_shading(_abi.SHADING_FAST)
    opx, outs = oracle_frame(oracle, name)
    inexact = check_frame(name, gpx, gzs, opx, outs, vps)
    covered = sum(o["n_covered"] for o in outs)
    assert sum(s.n_covered for s in stats) == covered
    print(f"{name} [{shading}]: covered={covered} pixels_not_bit_identical={inexact}")
    if shading == "exact":
        assert inexact == 0, f"{name}: exact shading differs from the oracle in {inexact} pixels"
    # vertex state of the last viewport, as the reference leaves it in mesh_vertex_t
    gv = renderer.read_vertices()
    for k in ("v_world", "v_viewport", "normal_world"):
        assert (gv[k].view(np.uint32) == outs[-1][k].view(np.uint32)).all(), f"{name}: {k} differs"
    assert (gv["yes"] == outs[-1]["yes"]).all(), f"{name}: yes flags differ"
    if os.path.exists(GOLDEN):
        gold = json.load(open(GOLDEN)).get(name)
        if gold:
            from swegl_b200.renderer import frame_hash
            assert "%016x" % frame_hash(gzs[-1]) == gold["depth_fnv1a64"][-1]
            if inexact == 0:
                assert "%016x" % frame_hash(gpx) == gold["frame_fnv1a64"] == "%016x" % oracle.fnv(gpx)


def test_sphere_8k_in_culled_bands_matches_golden(culling_renderer, oracle):
    """BASELINE.json config 5 the way 8 GPUs render it: the 7680x4320 frame of make_sphere(1000) (model.hpp:393-464) in 8
    row bands with band culling on, reassembled -- depth and (exact shading) colour hashes against tests/golden/MANIFEST.json,
    which tools/make_golden.py pinned against the unmodified reference (729fb9ef5c41fa64)"""
    from swegl_b200.renderer import frame_hash
    from swegl_b200 import sharding
    r = culling_renderer
    name = "sphere1000_8k"
    scene, vps, screen, cfg = configs.build(name)
    vp = vps[0]
    gold = json.load(open(GOLDEN))[name]
    r.set_shading(_abi.SHADING_EXACT)
    try:
        r.upload_scene(scene); r.set_screen(*screen); r.begin_frame(scene)
        px = np.zeros((screen[1], screen[0]), np.uint32)
        z = np.empty((vp.h, vp.w), np.float32)
        skipped = 0
        for k in range(8):
            vp.band = sharding.band_rows(vp.h, 8, k)
            r.render(vp, px, z)
            c = r.cull_counts()
            assert c["culled"]
            skipped += c["clusters"] - c["marked"]
        vp.band = (0, 0)
    finally:
        r.set_shading(_abi.SHADING_FAST)
    assert skipped > 0
    assert "%016x" % frame_hash(z) == gold["depth_fnv1a64"][0]
    assert "%016x" % frame_hash(px) == gold["frame_fnv1a64"] == "729fb9ef5c41fa64"
    # and the default (fast) shading of the same banded frame stays within 1 LSB of it
    fast = np.zeros_like(px)
    for k in range(8):
        vp.band = sharding.band_rows(vp.h, 8, k)
        r.render(vp, fast)
    vp.band = (0, 0)
    d = channel_diff(fast, px)
    assert d.max() <= 1 and d[..., 3].max() == 0
    print(f"sphere1000_8k fast vs exact: {int((fast != px).sum())} pixels differ by 1 LSB")


@pytest.mark.parametrize("light,tex", [(l, t) for l in (0, 1, 2) for t in (0, 1, 2)])
def test_all_shader_combinations(renderer, oracle, light, tex):
    """every built-in pixel_shader_t combination on the point-lit truck at 960x540"""
    scene, vps, screen, cfg = configs.build("truck_1080", light_mode=light, tex_mode=tex)
    vp = vps[0]
    vp.x, vp.y, vp.w, vp.h = 0, 0, 960, 540
    from swegl_b200.scene import Camera
    vp.camera = Camera(1.0 * 960 / 540).apply(vp.pose)
    gpx, gzs, _ = render_gpu(renderer, scene, [vp], (960, 540))
    opx, outs = render_oracle(oracle, scene, [vp], (960, 540))
    check_frame(f"L{light}T{tex}", gpx, gzs, opx, outs, [vp])


@pytest.mark.parametrize("name", ["truck_4k_dof", "brainstem_4k_dof"])
def test_dof(renderer, oracle, name):
    scene, vps, screen, cfg = configs.build(name)
    gpx, gzs, _ = render_gpu(renderer, scene, vps, screen)
    opx, outs = render_oracle(oracle, scene, vps, screen)
    for gz, o in zip(gzs, outs):
        assert (gz.view(np.uint32) == o["z"].view(np.uint32)).all()
    # DoF-R is integer arithmetic on the shaded image: identical wherever the shaded inputs are identical.
    # Check the DoF kernel exactly by feeding the GPU's own pre-DoF image through the oracle's DoF-R.
    vp = vps[0]
    import copy
    vp0 = copy.copy(vp)
    vp0.post_mode = _abi.POST_NULL
    pre, _, _ = render_gpu(renderer, scene, [vp0], screen)
    expect = oracle.dof_r(pre, gzs[0], vp.focal_distance, vp.focal_depth)
    assert (gpx == expect).all(), f"{name}: DoF-R output differs from the oracle on identical inputs"
    d = channel_diff(gpx, opx)
    assert d.max() <= 1


def layered_frame(renderer, oracle, scene, vp, screen):
    gpx, gzs, stats = render_gpu(renderer, scene, [vp], screen)
    opx, outs = render_oracle(oracle, scene, [vp], screen)
    nid = check_frame("layers", gpx, gzs, opx, outs, [vp])
    assert stats[0].n_covered == outs[0]["n_covered"]
    return nid, gpx, outs[0]


@pytest.mark.parametrize("alpha,layers,tex_alpha", [(100, 3, False), (100, 1, False), (30, 2, False), (100, 2, True), (0, 3, True),
                                                    (200, 8, True)])
@pytest.mark.parametrize("size,pose", [((400, 300), "POSE_LAYERS"), ((1920, 1080), "POSE_LAYERS"), ((640, 480), "POSE_LAYERS_CLOSE")])
def test_transparency_layers(renderer, oracle, alpha, layers, tex_alpha, size, pose):
    """renderer.cpp:500-550 + viewport.cpp:43-86 on the device: up to `layers` transparent fragments per pixel in front
    of the opaque one, blended far -> near; with tex_alpha the texel decides per fragment which kind it is"""
    from swegl_b200.scene import Viewport
    scene = configs.procedural(alpha, tex_alpha)
    pose = getattr(configs, pose)
    vp = Viewport(0, 0, size[0], size[1], transparency_layers=layers)
    vp.camera.apply(pose)
    nid, gpx, o = layered_frame(renderer, oracle, scene, vp, size)
    # the layers must actually matter: the same frame without them is different
    vp0 = Viewport(0, 0, size[0], size[1], transparency_layers=0)
    vp0.camera.apply(pose)
    opx0, _ = render_oracle(oracle, scene, [vp0], size)
    assert (opx0 != gpx).sum() > 100
    print(f"layers alpha={alpha} L={layers} tex_alpha={tex_alpha} {size}: {nid} px not bit-identical")


@pytest.mark.parametrize("light,tex", [(0, 0), (1, 1), (2, 0), (0, 2)])
def test_transparency_layers_other_shaders(renderer, oracle, light, tex):
    from swegl_b200.scene import Viewport
    scene = configs.procedural(120, True)
    vp = Viewport(0, 0, 640, 480, light_mode=light, tex_mode=tex, transparency_layers=2)
    vp.camera.apply(configs.POSE_LAYERS)
    layered_frame(renderer, oracle, scene, vp, (640, 480))


def test_transparency_layers_truck_glass_with_dof(renderer, oracle):
    """the milk truck with every material at alpha 140 and DoF-R behind the flatten (post pass reads layer 0)"""
    scene, vps, screen, cfg = configs.build("truck_1080")
    scene.mat_bgra = scene.mat_bgra.copy()
    keep = scene.mat_bgra.copy()
    scene.mat_bgra[:, 3] = 140
    vp = vps[0]
    vp.transparency_layers, vp.post_mode, vp.focal_distance, vp.focal_depth = 3, _abi.POST_DOF, 5.0, 5.0
    try:
        gpx, gzs, _ = render_gpu(renderer, scene, vps, screen)
        opx, outs = render_oracle(oracle, scene, vps, screen)
    finally:
        scene.mat_bgra = keep
        vp.transparency_layers, vp.post_mode = 3, _abi.POST_NULL
    assert (gzs[0].view(np.uint32) == outs[0]["z"].view(np.uint32)).all()
    assert channel_diff(gpx, opx).max() <= 1
    assert len(np.unique(gpx)) > 1000


def test_transparency_layers_limits(renderer):
    from swegl_b200.renderer import SweglB200Error
    from swegl_b200.scene import Viewport
    scene = configs.procedural(100)
    renderer.upload_scene(scene)
    renderer.set_screen(400, 300)
    renderer.begin_frame(scene)
    px = np.zeros((300, 400), np.uint32)
    with pytest.raises(SweglB200Error):
        renderer.render(Viewport(0, 0, 400, 300, transparency_layers=9), px)        # more than 8 layers
    with pytest.raises(SweglB200Error):
        renderer.render(Viewport(8, 0, 392, 300, transparency_layers=2), px)        # flatten() quirk: origin viewports only


def test_pipelined_readback_matches_blocking_render(renderer):
    """render_async/wait (two frames in flight, two pinned host images) delivers the frames render() delivers"""
    from swegl_b200.scene import Viewport
    scene, vps, screen, cfg = configs.build("truck_1080")
    renderer.upload_scene(scene)
    renderer.set_screen(*screen)
    views = []
    for pose in (configs.POSE_TEST1, configs.POSE_CLOSE):
        v = Viewport(0, 0, screen[0], screen[1], transparency_layers=0)
        v.camera.apply(pose)
        views.append(v)
    want = []
    for v in views:
        px = np.zeros((screen[1], screen[0]), np.uint32); z = np.empty((v.h, v.w), np.float32)
        renderer.begin_frame(scene); renderer.render(v, px, z)
        want.append((px, z))
    assert (want[0][0] != want[1][0]).any()
    images = [renderer.alloc_host((screen[1], screen[0]), np.uint32) for _ in range(2)]
    depths = [renderer.alloc_host((screen[1], screen[0]), np.float32) for _ in range(2)]
    tickets = []
    for i in range(5):
        renderer.begin_frame(scene)
        tickets.append(renderer.render_async(views[i & 1], images[i & 1], depths[i & 1]))
        if i >= 1:
            renderer.wait(tickets[i - 1])
            k = (i - 1) & 1
            assert (images[k] == want[k][0]).all() and (depths[k].view(np.uint32) == want[k][1].view(np.uint32)).all(), f"frame {i - 1}"
            images[k][:] = 0                                    # the next frame through this image must really arrive
    renderer.wait(tickets[-1])
    assert (images[0] == want[0][0]).all()


def test_async_pool_overflow_with_two_frames_in_flight():
    """The overflow contract of render_viewport_async (include/swegl_b200.h): a fresh context's fragment pool holds 2^24
    entries; a frame that fills a 6400x3600 screen needs 23 M.  With two frames in flight the frame after the overflowing
    one has already run with the same pools: BOTH tickets report ERR_CAPACITY, the pools are enlarged, and resubmitting
    means begin_frame with THAT frame's node matrices again -- the device block holds the later frame's by then."""
    from swegl_b200 import Renderer
    from swegl_b200.renderer import SweglB200Error
    from swegl_b200.scene import Viewport
    scene, _, _, _ = configs.build("sphere100_1080")
    screen = (6400, 3600)
    r = Renderer(0)
    try:
        r.upload_scene(scene)
        r.set_screen(*screen)
        far, near = Viewport(0, 0, *screen, transparency_layers=0), Viewport(0, 0, *screen, transparency_layers=0)
        far.camera.apply([("translate", 0, 0, -40.0)])          # a dot: fits any pool
        near.camera.apply([("translate", 0, 0, -3.0)])          # the sphere fills the screen: 23 M fragments
        w0, n0 = scene.node_matrices()
        mats = []
        for k in range(3):                                      # every frame has node matrices of its own (the sphere turns)
            w = w0.copy()
            c, s_ = np.float32(np.cos(0.3 * k)), np.float32(np.sin(0.3 * k))
            rot = np.array([[c, 0, s_, 0], [0, 1, 0, 0], [-s_, 0, c, 0], [0, 0, 0, 1]], np.float32)
            w[0] = (rot @ w0[0]).astype(np.float32)
            mats.append((w, n0))
        images = [r.alloc_host((screen[1], screen[0]), np.uint32) for _ in range(2)]
        r.begin_frame(scene, mats[0]); t_a = r.render_async(far, images[0])
        r.begin_frame(scene, mats[1]); t_b = r.render_async(near, images[1])
        r.wait(t_a)                                             # the small frame is fine
        r.begin_frame(scene, mats[2]); t_c = r.render_async(near, images[0])
        failed = []
        for t in (t_b, t_c):
            try:
                r.wait(t)
            except SweglB200Error as e:
                assert e.status == _abi.ERR_CAPACITY
                failed.append(t)
        assert failed == [t_b, t_c]
        # resubmission: begin_frame with frame b's data, then the view again; now it fits
        r.begin_frame(scene, mats[1]); t_b2 = r.render_async(near, images[1])
        r.wait(t_b2)
        r.begin_frame(scene, mats[2]); t_c2 = r.render_async(near, images[0])
        r.wait(t_c2)
        want = np.zeros((screen[1], screen[0]), np.uint32)
        for k, img in ((1, images[1]), (2, images[0])):
            r.begin_frame(scene, mats[k]); st = r.render(near, want)
            assert st.n_covered == screen[0] * screen[1]
            assert (img == want).all(), f"resubmitted frame {k}"
        assert (images[0] != images[1]).any()                   # the two frames really differ (different node matrices)
    finally:
        r.close()


@pytest.mark.parametrize("name", ["truck_1080", "truck_4k_dof"])
def test_color_target_assembles_bands_in_another_screen(renderer, name):
    """multi-GPU output path on one GPU: a second context renders its band straight into the first context's screen
    (set_color_target), as a peer GPU does through a CUDA IPC mapping"""
    from swegl_b200 import Renderer
    scene, vps, screen, cfg = configs.build(name)
    vp = vps[0]
    want, _, _ = render_gpu(renderer, scene, vps, screen, want_z=False)
    a, b = Renderer(0), Renderer(0)
    for r in (a, b):
        r.upload_scene(scene); r.set_screen(*screen); r.begin_frame(scene)
    screen_a, _ = a.device_buffers()
    b.set_color_target(screen_a)
    cut = vp.h // 3
    try:
        vp.band = (0, cut); a.render_device(vp, stats=True)
        vp.band = (cut, vp.h); b.render_device(vp, stats=True)
        a.synchronize(); b.synchronize()
        got = a.read_screen()
        assert (got == want).all()
        b.set_color_target(None)                        # back to its own screen: A's frame is not touched any more
        vp.band = (0, cut); b.render_device(vp, stats=True); b.synchronize()
        assert (a.read_screen() == want).all() and (b.read_screen()[:cut] == want[:cut]).all()
    finally:
        vp.band = (0, 0)


@pytest.mark.parametrize("name", ["multiview_1080", "multiview_4k"])
def test_viewport_per_context_assembles_split_screen(renderer, oracle, name):
    """config 4 on one GPU the way 2 GPUs run it (sharding.viewports_for_rank): context A renders viewports 0 and 2 into
    its own screen, context B renders 1 and 3 straight into A's screen (set_color_target = the CUDA IPC mapping of a
    peer).  The assembled split screen equals the one-context frame bit for bit and the oracle within 1 LSB."""
    from swegl_b200 import Renderer, sharding
    scene, vps, screen, cfg = configs.build(name)
    want, _, _ = render_gpu(renderer, scene, vps, screen, want_z=False)
    a, b = Renderer(0), Renderer(0)
    for r in (a, b):
        r.upload_scene(scene); r.set_screen(*screen); r.begin_frame(scene)
    screen_a, _ = a.device_buffers()
    b.set_color_target(screen_a)
    for rank, r in enumerate((a, b)):
        for v in sharding.viewports_for_rank(len(vps), 2, rank):
            r.render_device(vps[v], stats=True)
    a.synchronize(); b.synchronize()
    got = a.read_screen()
    assert (got == want).all()
    opx, _ = oracle_frame(oracle, name)
    assert channel_diff(got, opx).max() <= 1
    assert ((got >> 24) == (opx >> 24)).all()
    a.close(); b.close()


def test_band_scissor_equals_full_frame(renderer, oracle):
    """sort-first row bands: rendering [0,h) as 3 uneven bands gives the full frame"""
    scene, vps, screen, cfg = configs.build("truck_1080")
    full, fz, _ = render_gpu(renderer, scene, vps, screen)
    vp = vps[0]
    px = np.zeros_like(full)
    z = np.empty_like(fz[0])
    for b0, b1 in [(0, 333), (333, 700), (700, 1080)]:
        vp.band = (b0, b1)
        renderer.render(vp, px, z)
    vp.band = (0, 0)
    assert (px == full).all() and (z.view(np.uint32) == fz[0].view(np.uint32)).all()


@pytest.mark.parametrize("rect", [(13, 7, 611, 403), (0, 0, 1003, 701), (500, 299, 503, 402), (37, 41, 33, 31), (37, 41, 5, 3), (1002, 700, 1, 1)])
@pytest.mark.parametrize("post", [_abi.POST_NULL, _abi.POST_DOF])
def test_odd_viewport_rectangles(renderer, oracle, rect, post):
    """rectangles, offsets and a screen pitch that are multiples of nothing: every vector path (128-bit clears, the
    4-pixel fragment-stream stores, the DoF window loads) has to fall back or peel correctly"""
    from swegl_b200.scene import Viewport
    scene, _, _, _ = configs.build("truck_1080")
    screen = (1003, 701)
    x, y, w, h = rect
    vp = Viewport(x, y, w, h, transparency_layers=0, post_mode=post, focal_distance=5.0, focal_depth=5.0)
    vp.camera.apply(configs.POSE_TEST1 if w < 100 else configs.POSE_CLOSE)
    gpx, gzs, stats = render_gpu(renderer, scene, [vp], screen)
    opx, outs = render_oracle(oracle, scene, [vp], screen)
    check_frame(f"rect {rect}", gpx, gzs, opx, outs, [vp])
    assert stats[0].n_covered == outs[0]["n_covered"]
    outside = np.ones((screen[1], screen[0]), bool)
    outside[y:y + h, x:x + w] = False
    assert (gpx[outside] == 0).all()                    # nothing is written outside the rectangle


def test_odd_rectangle_bands_with_dof(renderer, oracle):
    from swegl_b200.scene import Viewport
    scene, _, _, _ = configs.build("truck_1080")
    screen = (1003, 701)
    vp = Viewport(13, 7, 611, 403, transparency_layers=0, post_mode=_abi.POST_DOF, focal_distance=5.0, focal_depth=5.0)
    vp.camera.apply(configs.POSE_CLOSE)
    full, fz, _ = render_gpu(renderer, scene, [vp], screen)
    px = np.zeros_like(full)
    z = np.empty_like(fz[0])
    for b0, b1 in [(0, 101), (101, 102), (102, 333), (333, 403)]:
        vp.band = (b0, b1)
        renderer.render(vp, px, z)
    vp.band = (0, 0)
    assert (px == full).all() and (z.view(np.uint32) == fz[0].view(np.uint32)).all()


def test_nothing_to_draw(renderer, oracle):
    """camera turned away from the scene (every triangle culled), and a scene without primitives: cleared frames"""
    from swegl_b200.scene import Viewport, Scene
    scene, _, _, _ = configs.build("truck_1080")
    vp = Viewport(0, 0, 640, 360, transparency_layers=0)
    vp.camera.apply([("translate", 0, 0, -5), ("rotate_y", 3.14159)])
    gpx, gzs, stats = render_gpu(renderer, scene, [vp], (640, 360))
    opx, outs = render_oracle(oracle, scene, [vp], (640, 360))
    assert outs[0]["n_covered"] == 0 and stats[0].n_covered == 0
    assert (gpx == 0).all() and (gzs[0].view(np.uint32) == 0x7F7F7F7F).all()
    empty = Scene()
    empty.name = "empty"
    empty.node_scale = np.ones((1, 3), np.float32)
    empty.node_rotation = np.eye(4, dtype=np.float32)[None]
    empty.node_translation = np.zeros((1, 3), np.float32)
    empty.node_parent = np.full(1, -1, np.int32)
    empty.set_lights(0.3, (1.0, -2.0, -1.0), 0.7, ())
    gpx, gzs, stats = render_gpu(renderer, empty, [vp], (640, 360))
    assert stats[0].n_covered == 0 and (gpx == 0).all() and (gzs[0].view(np.uint32) == 0x7F7F7F7F).all()
    # and the context recovers: the next scene renders normally
    vp2 = Viewport(0, 0, 640, 360, transparency_layers=0)
    vp2.camera.apply(configs.POSE_TEST1)
    gpx, gzs, _ = render_gpu(renderer, scene, [vp2], (640, 360))
    opx, outs = render_oracle(oracle, scene, [vp2], (640, 360))
    check_frame("after empty", gpx, gzs, opx, outs, [vp2])


def test_errors(renderer):
    from swegl_b200.renderer import SweglB200Error
    scene, vps, screen, cfg = configs.build("box_640")
    renderer.upload_scene(scene)
    renderer.set_screen(320, 240)
    renderer.begin_frame(scene)
    with pytest.raises(SweglB200Error):
        renderer.render(vps[0], np.zeros((240, 320), np.uint32))      # viewport larger than the screen


# ---- sort-first band culling (include/swegl_b200.h: swegl_b200_set_band_culling) ----
@pytest.fixture(scope="module")
def culling_renderer():
    from swegl_b200 import Renderer
    r = Renderer(0)
    r.set_band_culling(1)              # every banded view, also of scenes below the automatic threshold
    yield r
    r.close()


def _bands(h, n):
    """n uneven bands covering [0, h)"""
    cuts = sorted({0, h} | {int(h * (k / n) ** 1.3) for k in range(1, n)})
    return list(zip(cuts[:-1], cuts[1:]))


CULL_CASES = [("sphere100_1080", None, 8), ("truck_1080", None, 5), ("brainstem_4k_dof", None, 8), ("truck_4k_dof", None, 3),
              ("brainstem_4k", None, 16),
              ("truck_1080", [("translate", 0, 0.5, -0.9), ("rotate_y", 0.4)], 6),          # camera inside the scene: near clipping
              ("box_640_close", None, 4)]


@pytest.mark.parametrize("name,pose,n_bands", CULL_CASES)
def test_band_culling_is_invisible(culling_renderer, name, pose, n_bands):
    """a banded view that skips the triangle clusters / vertex blocks that cannot reach it is bit-identical (colour and
    depth) to the full frame rendered without culling, and something is actually skipped"""
    r = culling_renderer
    scene, vps, screen, cfg = configs.build(name)
    vp = vps[0]
    if pose is not None:
        from swegl_b200.scene import Viewport
        vp = Viewport(vp.x, vp.y, vp.w, vp.h, transparency_layers=0, post_mode=vp.post_mode, focal_distance=vp.focal_distance,
                      focal_depth=vp.focal_depth)
        vp.camera.apply(pose)
    full, fz, _ = render_gpu(r, scene, [vp], screen)
    assert not r.cull_counts()["culled"]
    px = np.zeros_like(full)
    z = np.empty_like(fz[0])
    skipped = 0
    for b0, b1 in _bands(vp.h, n_bands):
        vp.band = (b0, b1)
        r.render(vp, px, z)
        c = r.cull_counts()
        assert c["culled"] and c["live"] <= c["marked"] <= c["clusters"]
        skipped += c["clusters"] - c["marked"]
    vp.band = (0, 0)
    assert (px == full).all() and (z.view(np.uint32) == fz[0].view(np.uint32)).all()
    if scene.n_triangles() >= 2000 and pose is None:
        assert skipped > 0, "no cluster was ever skipped"


def test_band_culling_async_graph_path(culling_renderer):
    """the same through the captured-graph path (render_device) with a changing camera: flags are per view, not cached"""
    r = culling_renderer
    scene, vps, screen, cfg = configs.build("sphere100_1080")
    vp = vps[0]
    r.upload_scene(scene); r.set_screen(*screen)
    for step in range(3):
        vp.camera.apply([("rotate_y", 0.15 * step), ("translate", 0.1 * step, 0, 0)])
        r.begin_frame(scene)
        full = np.zeros((screen[1], screen[0]), np.uint32)
        r.render(vp, full)
        got = np.zeros_like(full)
        for b0, b1 in _bands(vp.h, 4):
            vp.band = (b0, b1)
            r.begin_frame(scene)
            r.render_device(vp, stats=False)
            r.synchronize()
            got[b0:b1] = r.read_screen()[b0:b1]
        vp.band = (0, 0)
        assert (got == full).all(), f"step {step}"


def test_frame_pipeline_renders_the_same_frames(renderer):
    """swegl_b200.FramePipeline (several contexts, frames round robin, overlapping on the GPU) produces exactly the frames
    one context renders one at a time -- with a camera that moves from frame to frame"""
    from swegl_b200.pipeline import FramePipeline
    scene, vps, screen, cfg = configs.build("truck_1080")
    vp = vps[0]
    renderer.upload_scene(scene); renderer.set_screen(*screen)
    pipe = FramePipeline(0, 3)
    try:
        pipe.upload_scene(scene); pipe.set_screen(*screen)
        pipe.size_pools(scene, vps)
        want, slots = [], []
        for i in range(6):
            vp.camera.apply([("rotate_y", 0.07), ("translate", 0.05, 0, 0)])
            renderer.begin_frame(scene)
            px = np.zeros((screen[1], screen[0]), np.uint32)
            renderer.render(vp, px)
            want.append(px)
            slots.append(pipe.submit(scene, [vp.desc()]))
            if i >= 2:                                  # frame i-2's context is about to be reused: read it first
                pass
            if len(slots) == 3:
                pipe.synchronize()
                for k, w in zip(slots, want):
                    assert (pipe.read_screen(k) == w).all()
                want, slots = [], []
    finally:
        pipe.close()


@pytest.mark.parametrize("name", ["sphere100_1080", "truck_1080", "truck_4k_dof"])
def test_frame_sync_protocol_two_contexts(name):
    """the multi-GPU frame protocol (swegl_b200_set_frame_sync) exercised inside one process: context A is "rank 0" and
    assembles the frame, context B is "rank 1", renders the lower band with A's screen as its colour target and skips
    the background.  Flags over (here: same-device) memory only; the camera moves, so stale tiles must disappear."""
    from swegl_b200 import Renderer
    scene, vps, screen, cfg = configs.build(name)
    vp = vps[0]
    a, b, ref = Renderer(0), Renderer(0), Renderer(0)
    try:
        for r in (a, b, ref):
            r.set_band_culling(1)
            r.upload_scene(scene); r.set_screen(*screen)
        cut = int(vp.h * 0.45)
        for r, band in ((a, (0, cut)), (b, (cut, vp.h))):       # size the pools with ordinary frames first
            vp.band = band
            r.begin_frame(scene); r.render_device(vp, stats=True)
        vp.band = (0, 0)
        b.set_color_target(a.device_buffers()[0])
        a.set_frame_sync(0, 2); b.set_frame_sync(1, 2)
        for step in range(4):
            vp.camera.apply([("rotate_y", 0.2), ("translate", 0.15, 0.05, 0)])
            want = np.zeros((screen[1], screen[0]), np.uint32)
            ref.begin_frame(scene); ref.render(vp, want)
            first, second = ((b, (cut, vp.h)), (a, (0, cut))) if step % 2 else ((a, (0, cut)), (b, (cut, vp.h)))
            for r, band in (first, second):                     # either submission order works
                vp.band = band
                r.begin_frame(scene); r.render_device(vp, stats=False)
            vp.band = (0, 0)
            a.synchronize()                                      # rank 0's stream passing the view = frame complete
            got = a.read_screen()
            b.synchronize()
            assert a.frame_sync_errors() == 0 and b.frame_sync_errors() == 0
            assert (got == want).all(), f"{name}: step {step}"
    finally:
        for r in (a, b, ref):
            r.close()


def test_shared_divisor_division_is_correctly_rounded(renderer):
    """div_by() (csrc/common.cuh: quotients by one divisor share the refined reciprocal) == __fdiv_rn, bit for bit, over
    2^30 generated operand pairs: raw bit patterns (every exponent, zeros, denormals, infinities, NaNs), in-range pairs,
    all-ones and power-of-two divisors, short numerators"""
    bad, fast = renderer.selftest_division(1 << 30, seed=20261017)
    assert bad == 0
    assert fast > (1 << 29)                             # the fast path really is what was compared


def test_fast_bilinear_filter_equals_the_exact_one(renderer):
    """tex_filter_bilinear_fast (the fast kernels' filter: no per-operation range checks, rounding through the adder, alpha
    shortcut) == tex_filter<BILINEAR>, bit for bit, over 2^28 generated samples below the 2^21 coordinate guard"""
    assert renderer.selftest_filter(1 << 28, seed=20261018) == 0


# ---- partial read-back of asynchronous frames (include/swegl_b200.h: swegl_b200_render_viewport_async) ----
@pytest.mark.parametrize("name", ["truck_1080", "truck_4k_dof"])
def test_partial_readback_equals_full_copy(renderer, name):
    """render_async copies only the union of the previous and the current bounding box of what was drawn into each host image;
    over a moving camera (the object wanders, leaves the screen, comes back) every image equals the blocking render()'s"""
    from swegl_b200.scene import Viewport
    scene, vps, screen, cfg = configs.build(name)
    base = vps[0]
    renderer.upload_scene(scene)
    renderer.set_screen(*screen)
    poses = [configs.POSE_TEST1,
             [("translate", 2.5, 2, -5), ("rotate_y", -0.2), ("rotate_x", -0.3)],
             [("translate", -1.5, 3, -6), ("rotate_y", 0.3), ("rotate_x", -0.4)],
             [("translate", 0, 0, -5), ("rotate_y", 3.14159)],                     # looks away: nothing drawn
             [("translate", 0, 0, -5), ("rotate_y", 3.14159)],
             configs.POSE_CLOSE,
             [("translate", 1, 2, -9), ("rotate_y", -0.2), ("rotate_x", -0.3)],
             configs.POSE_TEST1]
    views = []
    for pose in poses:
        v = Viewport(0, 0, screen[0], screen[1], transparency_layers=0, post_mode=base.post_mode, focal_distance=base.focal_distance,
                     focal_depth=base.focal_depth)
        v.camera.apply(pose)
        views.append(v)
    want = []
    for v in views:
        px = np.zeros((screen[1], screen[0]), np.uint32); z = np.empty((v.h, v.w), np.float32)
        renderer.begin_frame(scene); renderer.render(v, px, z)
        want.append((px, z))
    images = [renderer.alloc_host((screen[1], screen[0]), np.uint32) for _ in range(2)]
    depths = [renderer.alloc_host((screen[1], screen[0]), np.float32) for _ in range(2)]
    for im in images:
        im[:] = 0xDEADBEEF                                  # the first frame through an image must be a full copy
    renderer.readback_stats(reset=True)
    tickets = []
    for i, v in enumerate(views):
        renderer.begin_frame(scene)
        tickets.append(renderer.render_async(v, images[i & 1], depths[i & 1]))
        if i >= 1:
            renderer.wait(tickets[i - 1])
            k = (i - 1) & 1
            assert (images[k] == want[i - 1][0]).all(), f"frame {i - 1}: colour"
            assert (depths[k].view(np.uint32) == want[i - 1][1].view(np.uint32)).all(), f"frame {i - 1}: depth"
    renderer.wait(tickets[-1])
    assert (images[(len(views) - 1) & 1] == want[-1][0]).all()
    nbytes, nframes = renderer.readback_stats()
    full = screen[0] * screen[1] * 8 * len(views)
    assert nframes == len(views) and nbytes < 0.7 * full, "nothing was saved"
    # an image the caller drew into has to be declared; the next frame through it is then copied in full
    images[0][:] = 0x12345678
    renderer.invalidate_host_image(images[0])
    renderer.begin_frame(scene)
    renderer.wait(renderer.render_async(views[0], images[0], depths[0]))
    assert (images[0] == want[0][0]).all()
    # switched off: every frame is a full copy
    renderer.set_partial_readback(False)
    try:
        images[1][:] = 0x55555555
        renderer.begin_frame(scene)
        renderer.wait(renderer.render_async(views[3], images[1], depths[1]))
        assert (images[1] == want[3][0]).all()
    finally:
        renderer.set_partial_readback(True)


@pytest.mark.parametrize("partial", [False, True])
@pytest.mark.parametrize("name", ["truck_1080", "truck_4k_dof"])
def test_three_frames_in_flight_staged_boxes(renderer, name, partial):
    """render_async keeps three device staging images and copies into one only the union of the box of the frame it holds
    and the new frame's (k_stage_rect): with three frames in flight over a moving camera -- the object wanders, leaves the
    screen, comes back, so every staging image sees growing, shrinking and empty boxes -- every host image equals the
    blocking render()'s.  With partial read-back off the WHOLE staging image crosses PCIe, which checks that it stayed a
    complete frame."""
    from swegl_b200.scene import Viewport
    scene, vps, screen, cfg = configs.build(name)
    base = vps[0]
    renderer.upload_scene(scene)
    renderer.set_screen(*screen)
    poses = [configs.POSE_TEST1,
             [("translate", 2.5, 2, -5), ("rotate_y", -0.2), ("rotate_x", -0.3)],
             [("translate", -1.5, 3, -6), ("rotate_y", 0.3), ("rotate_x", -0.4)],
             [("translate", 0, 0, -5), ("rotate_y", 3.14159)],                     # looks away: nothing drawn
             configs.POSE_CLOSE,
             [("translate", 0, 0, -5), ("rotate_y", 3.14159)],
             [("translate", 0, 0, -5), ("rotate_y", 3.14159)],
             [("translate", 1, 2, -9), ("rotate_y", -0.2), ("rotate_x", -0.3)],
             [("translate", -2.5, 1, -7), ("rotate_y", 0.25)],
             configs.POSE_TEST1, configs.POSE_CLOSE]
    views = []
    for pose in poses:
        v = Viewport(0, 0, screen[0], screen[1], transparency_layers=0, post_mode=base.post_mode, focal_distance=base.focal_distance,
                     focal_depth=base.focal_depth)
        v.camera.apply(pose)
        views.append(v)
    want = []
    for v in views:
        px = np.zeros((screen[1], screen[0]), np.uint32)
        renderer.begin_frame(scene); renderer.render(v, px)
        want.append(px)
    images = [renderer.alloc_host((screen[1], screen[0]), np.uint32) for _ in range(3)]
    for im in images:
        im[:] = 0xDEADBEEF
    renderer.set_partial_readback(partial)
    try:
        renderer.readback_stats(reset=True)
        tickets = []
        for i, v in enumerate(views):
            renderer.begin_frame(scene)
            tickets.append(renderer.render_async(v, images[i % 3]))
            if i >= 2:
                renderer.wait(tickets[i - 2])
                assert (images[(i - 2) % 3] == want[i - 2]).all(), f"frame {i - 2}"
                if not partial:
                    images[(i - 2) % 3][:] = 0xDEADBEEF         # every frame has to arrive whole
        for i in (len(views) - 2, len(views) - 1):
            renderer.wait(tickets[i])
            assert (images[i % 3] == want[i]).all(), f"frame {i}"
        nbytes, nframes = renderer.readback_stats()
        full = screen[0] * screen[1] * 4 * len(views)
        assert nframes == len(views)
        assert nbytes < 0.7 * full if partial else nbytes == full
    finally:
        renderer.set_partial_readback(True)
        renderer.invalidate_host_image(None)


@pytest.mark.parametrize("name", ["truck_1080", "box_640"])
def test_shared_gpu_mode_renders_the_same_frame(renderer, name):
    """swegl_b200_set_shared_gpu selects kernel shapes (the 128-thread span kernel), not results: colour and depth of a
    frame are identical with and without it"""
    scene, vps, screen, cfg = configs.build(name)
    renderer.upload_scene(scene)
    renderer.set_screen(*screen)
    got = []
    try:
        for shared in (False, True, False):
            renderer.set_shared_gpu(shared)
            px = np.zeros((screen[1], screen[0]), np.uint32); z = np.empty((vps[0].h, vps[0].w), np.float32)
            renderer.begin_frame(scene)
            st = renderer.render(vps[0], px, z)
            got.append((px, z, st.n_covered))
    finally:
        renderer.set_shared_gpu(False)
    assert got[0][2] > 0
    for px, z, n in got[1:]:
        assert n == got[0][2] and (px == got[0][0]).all() and (z.view(np.uint32) == got[0][1].view(np.uint32)).all()


def test_host_readback_refused_while_a_colour_target_is_set(renderer):
    from swegl_b200 import Renderer
    from swegl_b200.renderer import SweglB200Error
    scene, vps, screen, cfg = configs.build("box_640")
    a, b = Renderer(0), Renderer(0)
    try:
        for r in (a, b):
            r.upload_scene(scene); r.set_screen(*screen); r.begin_frame(scene)
        b.set_color_target(a.device_buffers()[0])
        px = np.zeros((screen[1], screen[0]), np.uint32)
        with pytest.raises(SweglB200Error) as e:
            b.render(vps[0], px)
        assert e.value.status == _abi.ERR_STATE
        with pytest.raises(SweglB200Error):
            b.render_async(vps[0], px)
        b.set_color_target(None)
        b.render(vps[0], px)
        assert px.any()
    finally:
        a.close(); b.close()


def test_full_view_after_a_culled_band_of_an_animated_frame(culling_renderer, oracle):
    """a band-culled view transforms only the vertex blocks it needs; a full view of the SAME frame after it must not
    trust v_world: moved nodes, band first, then the whole viewport -- equal to the oracle's frame"""
    r = culling_renderer
    scene, vps, screen, cfg = configs.build("truck_1080")
    vp = vps[0]
    r.upload_scene(scene); r.set_screen(*screen)
    px = np.zeros((screen[1], screen[0]), np.uint32)
    r.begin_frame(scene); r.render(vp, px)                          # frame 0: rest pose everywhere in v_world
    world, normal = scene.node_matrices()
    world = world.copy()
    world[:, 0, 3] += 0.75; world[:, 1, 3] -= 0.25                    # frame 1: every node moved
    r.begin_frame(scene, (world, normal))
    vp.band = (500, 560)
    r.render_device(vp, stats=True)
    assert r.cull_counts()["culled"]
    vp.band = (0, 0)
    z = np.empty((vp.h, vp.w), np.float32)
    r.render(vp, px, z)
    o = oracle.render(scene, vp, screen_wh=screen, node_mats=(world, normal))
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert channel_diff(px, o["pixels"]).max() <= 1


@pytest.mark.parametrize("rect", [(12, 7, 612, 400), (0, 0, 1000, 700), (388, 299, 64, 32), (100, 41, 8, 3)])
def test_dof_tma_window_staging_on_odd_offsets(renderer, oracle, rect):
    """viewport widths that are multiples of 4 (the DoF windows arrive by TMA) at offsets and on a screen pitch that are not"""
    from swegl_b200.scene import Viewport
    scene, _, _, _ = configs.build("truck_1080")
    screen = (1003, 701)
    x, y, w, h = rect
    vp = Viewport(x, y, w, h, transparency_layers=0, post_mode=_abi.POST_DOF, focal_distance=5.0, focal_depth=5.0)
    vp.camera.apply(configs.POSE_TEST1 if w < 100 else configs.POSE_CLOSE)
    gpx, gzs, stats = render_gpu(renderer, scene, [vp], screen)
    opx, outs = render_oracle(oracle, scene, [vp], screen)
    check_frame(f"rect {rect}", gpx, gzs, opx, outs, [vp])
    vp0 = Viewport(x, y, w, h, transparency_layers=0)
    vp0.camera = vp.camera
    pre, _, _ = render_gpu(renderer, scene, [vp0], screen)
    expect = oracle.dof_r(pre[y:y + h, x:x + w].copy(), gzs[0], 5.0, 5.0)
    assert (gpx[y:y + h, x:x + w] == expect).all()
