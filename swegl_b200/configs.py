"""The workloads of BASELINE.json / SURVEY.md §8(d), made concrete once for tests, bench and golden
generation: scene + lights + viewport(s) + camera pose(s).

Camera poses are lists of camera_t calls replayed on swegl_b200.scene.Camera (and, in the pinning tests,
on the reference's own camera_t).
"""
import os

import numpy as np

from . import _abi
from .scene import Scene, Viewport, f32, identity44, rotate_y, rotate_z, lcg_texture, from_quaternion

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")

POSE_TEST1 = [("translate", 1, 2, -5), ("rotate_y", -0.2), ("rotate_x", -0.3)]        # src/test_1.cpp:359-361
POSE_CLOSE = [("translate", 1, 2, -1.5), ("rotate_y", -0.2), ("rotate_x", -0.3)]
POSE_BRAIN_CLOSE = [("translate", 1, 2, -1), ("rotate_y", -0.2), ("rotate_x", -0.3)]
POSE_SPHERE = [("translate", 0, 0, -4.5)]                                             # SURVEY §8c synthetic
POINT_LIGHTS = [(0.0, 3.0, 0.0, 0.6), (0.5, 2.0, 0.0, 100.0)]                         # src/test_1.cpp:74-75


def make_sphere_scene(precision, radius=2.0, texture_size=1024):
    """swegl::make_sphere(precision, radius, material 0) (swegl/data/model.hpp:393-464) as a flattened
    Scene with the seeded LCG texture; arithmetic order follows matrix44_t::rotate_* and transform()."""
    P = precision
    angle = f32(f32(2 * f32(3.141592653589)) / f32(P))
    half = f32(angle / f32(2))
    # `small` after k rotate_z(angle/2) calls, `big` after k rotate_y(angle) calls
    smalls = np.zeros((P + 1, 4, 4), np.float32)
    m = identity44()
    for k in range(P + 1):
        smalls[k] = m
        rotate_z(m, half)
    bigs = np.zeros((P + 2, 4, 4), np.float32)
    m = identity44()
    for k in range(P + 2):
        bigs[k] = m
        rotate_y(m, angle)

    def xform(M, v):            # transform(vertex_t, matrix44_t), points.cpp:8-13, vectorised in fp32
        x, y, z = v[..., 0], v[..., 1], v[..., 2]
        out = np.empty(np.broadcast(M[..., 0, 0], x).shape + (3,), np.float32)
        for r in range(3):
            out[..., r] = ((M[..., r, 0] * x + M[..., r, 1] * y) + M[..., r, 2] * z) + M[..., r, 3]
        return out

    base = xform(smalls, np.array([0.0, radius, 0.0], np.float32)[None, :])          # (P+1, 3)
    rows = xform(bigs[:, None], base[None, :, :])                                    # (P+2, P+1, 3)
    l = np.sqrt((rows[..., 0] * rows[..., 0] + rows[..., 1] * rows[..., 1]) + rows[..., 2] * rows[..., 2])
    safe = np.where(l != 0, l, f32(1))
    nrm = np.where((l != 0)[..., None], rows / safe[..., None], rows).astype(np.float32)
    u = (np.arange(P + 1) / P).astype(np.float32)                                     # 1.0*sm/precision
    vrow = (np.arange(P + 2) / P).astype(np.float32)                                  # 1.0*bg/precision

    nprim, nvp = P + 1, 2 * (P + 1)
    s = Scene()
    s.name = f"sphere{P}"
    s.node_scale = np.ones((1, 3), np.float32)
    s.node_rotation = identity44()[None]
    s.node_translation = np.zeros((1, 3), np.float32)
    s.node_parent = np.array([-1], np.int32)
    s.positions = np.concatenate([rows[:-1], rows[1:]], axis=1).reshape(-1, 3).astype(np.float32)
    s.normals = np.concatenate([nrm[:-1], nrm[1:]], axis=1).reshape(-1, 3).astype(np.float32)
    tc = np.empty((nprim, nvp, 2), np.float32)
    tc[:, :P + 1, 0] = u[None]; tc[:, P + 1:, 0] = u[None]
    tc[:, :P + 1, 1] = vrow[:-1, None]; tc[:, P + 1:, 1] = vrow[1:, None]
    s.texcoords = tc.reshape(-1, 2)
    idx = np.empty((P + 1, 2), np.uint32)
    idx[:, 0] = P + 1 + np.arange(P + 1); idx[:, 1] = np.arange(P + 1)
    s.indices = np.tile(idx.reshape(-1), nprim).astype(np.uint32)
    s.prim_node = np.zeros(nprim, np.int32)
    s.prim_mode = np.full(nprim, _abi.MODE_TRIANGLE_STRIP, np.int32)
    s.prim_material = np.zeros(nprim, np.int32)
    s.prim_first_vertex = (np.arange(nprim) * nvp).astype(np.uint32)
    s.prim_n_vertices = np.full(nprim, nvp, np.uint32)
    s.prim_first_index = (np.arange(nprim) * nvp).astype(np.uint32)
    s.prim_n_indices = np.full(nprim, nvp, np.uint32)
    s.mat_bgra = np.array([[128, 128, 128, 255]], np.uint8)
    s.mat_metal_rough = np.ones((1, 2), np.float32)
    s.mat_tex_ds = np.array([[0, 0]], np.int32)
    s.textures = [lcg_texture(texture_size)]
    return s


_SCENE_CACHE = {}


def load_scene(name):
    if name not in _SCENE_CACHE:
        if name.startswith("sphere"):
            _SCENE_CACHE[name] = make_sphere_scene(int(name[len("sphere"):]))
        else:
            _SCENE_CACHE[name] = Scene.load_pack(os.path.join(ASSETS, name + ".scenepack"))
    return _SCENE_CACHE[name]


def _lights_test1(s, points=False):
    return s.set_lights(0.3, (1.0, -2.0, -1.0), 0.7, POINT_LIGHTS if points else ())   # test_1.cpp:334-336


def _lights_synth(s):
    return s.set_lights(0.2, (1.0, -1.0, -1.0), 0.3, POINT_LIGHTS)                     # SURVEY §8c synthetic


POSE_PROCEDURAL = [("translate", 1, 2, 6), ("rotate_y", 3.0), ("rotate_x", -0.3)]      # torus / cube / sphere, strips and fans
# the three stacked (transparent) triangles in front of the cube: 1, 2 and 3 layers deep per pixel
POSE_LAYERS = [("translate", 1.3, 1.2, -4), ("rotate_y", 0.1), ("rotate_x", -0.1)]
POSE_LAYERS_CLOSE = [("translate", 1.5, 1, -3)]


# name -> (scene, screen (w,h), lights fn, [viewport kwargs + pose], description)
CONFIGS = {
    "box_640": dict(scene="BoxTextured", screen=(640, 480), lights="test1", views=[dict(rect=(0, 0, 640, 480), pose=POSE_TEST1, layers=3)],
                    desc="config 1: BoxTextured 640x480, Phong+bilinear, sun, 3 layers, test_1 pose"),
    "box_640_close": dict(scene="BoxTextured", screen=(640, 480), lights="test1", views=[dict(rect=(0, 0, 640, 480), pose=POSE_CLOSE, layers=3)],
                          desc="config 1 close pose (8.3% coverage)"),
    "truck_1080": dict(scene="CesiumMilkTruck", screen=(1920, 1080), lights="test1+points", views=[dict(rect=(0, 0, 1920, 1080), pose=POSE_TEST1, layers=3)],
                       desc="config 2: CesiumMilkTruck 1920x1080, Phong+bilinear, sun + 2 point lights (no bump map: none in the reference)"),
    "truck_1080_sun": dict(scene="CesiumMilkTruck", screen=(1920, 1080), lights="test1", views=[dict(rect=(0, 0, 1920, 1080), pose=POSE_TEST1, layers=3)],
                           desc="CesiumMilkTruck 1920x1080, sun only (survey hash 0418ee20dd2b64e1)"),
    "truck_4k_dof": dict(scene="CesiumMilkTruck", screen=(3840, 2160), lights="test1+points",
                         views=[dict(rect=(0, 0, 3840, 2160), pose=POSE_TEST1, layers=0, post=_abi.POST_DOF)],
                         desc="north-star target: CesiumMilkTruck 3840x2160, Phong+bilinear, sun + 2 point lights, DoF-R(5,5)"),
    "truck_4k": dict(scene="CesiumMilkTruck", screen=(3840, 2160), lights="test1", views=[dict(rect=(0, 0, 3840, 2160), pose=POSE_TEST1, layers=3)],
                     desc="CesiumMilkTruck 3840x2160, sun only (survey hash a8310584d693ad8c)"),
    "brainstem_4k": dict(scene="BrainStem", screen=(3840, 2160), lights="test1", views=[dict(rect=(0, 0, 3840, 2160), pose=POSE_TEST1, layers=3)],
                         desc="BrainStem 3840x2160, test_1 pose (survey hash 8cea197b1c69d6f1)"),
    "brainstem_4k_dof": dict(scene="BrainStem", screen=(3840, 2160), lights="test1",
                             views=[dict(rect=(0, 0, 3840, 2160), pose=POSE_BRAIN_CLOSE, layers=0, post=_abi.POST_DOF)],
                             desc="config 3: BrainStem 3840x2160 close pose, Phong+bilinear+DoF-R"),
    "multiview_1080": dict(scene="CesiumMilkTruck", screen=(1920, 1080), lights="test1",
                           views=[dict(rect=(0, 0, 960, 540), pose=POSE_TEST1, layers=0),
                                  dict(rect=(960, 0, 960, 540), pose=[("translate", -1, 2, -5), ("rotate_y", 0.2), ("rotate_x", -0.3)], layers=0),
                                  dict(rect=(0, 540, 960, 540), pose=[("translate", 0, 4, -4), ("rotate_x", -0.7)], layers=0),
                                  dict(rect=(960, 540, 960, 540), pose=[("translate", 3, 1, -3), ("rotate_y", -0.7), ("rotate_x", -0.1)], layers=0)],
                           desc="config 4: 2x2 split screen, 4 cameras (CesiumMilkTruck substituted for the missing BarramundiFish.glb)"),
    "multiview_4k": dict(scene="CesiumMilkTruck", screen=(3840, 2160), lights="test1+points",
                         views=[dict(rect=(0, 0, 1920, 1080), pose=POSE_TEST1, layers=0),
                                dict(rect=(1920, 0, 1920, 1080), pose=[("translate", -1, 2, -5), ("rotate_y", 0.2), ("rotate_x", -0.3)], layers=0),
                                dict(rect=(0, 1080, 1920, 1080), pose=[("translate", 0, 4, -4), ("rotate_x", -0.7)], layers=0),
                                dict(rect=(1920, 1080, 1920, 1080), pose=[("translate", 3, 1, -3), ("rotate_y", -0.7), ("rotate_x", -0.1)], layers=0)],
                         desc="config 4 at 4K: 2x2 split screen of 1920x1080 viewports, 4 cameras, sun + 2 point lights (CesiumMilkTruck substituted for the missing BarramundiFish.glb)"),
    "layers_640": dict(scene="procedural", alpha=100, screen=(640, 480), lights="procedural",
                       views=[dict(rect=(0, 0, 640, 480), pose=POSE_LAYERS, layers=3)],
                       desc="SURVEY 8f N1: procedural scene (strips, fans, scaled node), three stacked alpha-100 triangles, 3 transparency layers"),
    "layers_texalpha_640": dict(scene="procedural", alpha=100, tex_alpha=True, screen=(640, 480), lights="procedural",
                                views=[dict(rect=(0, 0, 640, 480), pose=POSE_LAYERS_CLOSE, layers=2)],
                                desc="N1: same scene, texels alternate alpha 255/90 (per-fragment opaque/transparent), 2 layers, close pose"),
    "sphere100_1080": dict(scene="sphere100", screen=(1920, 1080), lights="synth", views=[dict(rect=(0, 0, 1920, 1080), pose=POSE_SPHERE, layers=0)],
                           desc="make_sphere(100) 20 200 triangles, LCG texture, 1920x1080 (survey hash 72b4fc66d972867b)"),
    "sphere1000_8k": dict(scene="sphere1000", screen=(7680, 4320), lights="synth", views=[dict(rect=(0, 0, 7680, 4320), pose=POSE_SPHERE, layers=0)],
                          desc="config 5: make_sphere(1000) 2 002 000 triangles, LCG texture, 7680x4320 (survey hash 729fb9ef5c41fa64)"),
}


def build(name, light_mode=_abi.LIGHT_PHONG, tex_mode=_abi.TEX_BILINEAR):
    """-> (scene, [Viewport...], (screen_w, screen_h), config dict)"""
    cfg = CONFIGS[name]
    if cfg["scene"] == "procedural":
        s = procedural(cfg.get("alpha", 255), cfg.get("tex_alpha", False))      # fresh copy, lights set
    else:
        s = load_scene(cfg["scene"])
        {"test1": lambda: _lights_test1(s), "test1+points": lambda: _lights_test1(s, True), "synth": lambda: _lights_synth(s)}[cfg["lights"]]()
    vps = []
    for v in cfg["views"]:
        x, y, w, h = v["rect"]
        vp = Viewport(x, y, w, h, light_mode=light_mode, tex_mode=tex_mode, transparency_layers=v.get("layers", 0),
                      post_mode=v.get("post", _abi.POST_NULL))
        vp.camera.apply(v["pose"])
        vp.pose = v["pose"]
        vps.append(vp)
    return s, vps, cfg["screen"], cfg


def procedural(alpha=255, tex_alpha=False):
    """a fresh (uncached) copy of assets/procedural.scenepack with transparency applied"""
    return with_transparency(Scene.load_pack(os.path.join(ASSETS, "procedural.scenepack")), alpha, tex_alpha)


def with_transparency(scene, alpha, tex_alpha=False):
    """the `procedural` scene with its three stacked triangles' materials (3..5) at `alpha`, and optionally a
    texture whose texels alternate alpha 255 / 90 in 8x8 blocks (per-fragment opaque/transparent decision)."""
    scene.mat_bgra = scene.mat_bgra.copy()
    scene.mat_bgra[3:6, 3] = alpha
    if tex_alpha:
        t = scene.textures[0]
        yy, xx = np.mgrid[0:t.shape[0], 0:t.shape[1]]
        a = np.where(((xx // 8 + yy // 8) & 1) == 1, 255, 90).astype(np.uint32)
        scene.textures[0] = ((t & np.uint32(0x00FFFFFF)) | (a << np.uint32(24))).astype(np.uint32)
        scene.mat_bgra[0, 3] = 200                  # the textured material itself is not what decides: the texel is
    scene.set_lights(0.2, (1.0, -1.0, -1.0), 0.3, POINT_LIGHTS)
    return scene


# ---- seeded random scenes for the parity fuzz tests (tests/test_fuzz_*.py) ----
POSE_FUZZ = [("translate", 0.3, 0.2, -4.0), ("rotate_y", 0.15), ("rotate_x", -0.1)]


def fuzz_scene(seed, n_prims=7, verts_per_prim=36, transparent=False):
    """A triangle soup that pokes at the branches of the reference's frame a well-behaved model never takes: all
    three index modes with random (repeated -> degenerate) indices, vertices behind the camera and across the near
    plane (renderer.cpp:281-356: 1 or 2 vertices clipped), collinear and sub-pixel triangles, screen-filling ones,
    vertices snapped to a grid (exactly shared edges, exactly equal depths: the z-test tie goes to the earlier draw,
    renderer.cpp:491), coplanar duplicates under another material, a node chain with non-uniform and mirrored scales
    (facing flips), unnormalised normals, non-power-of-two textures, 0..3 point lights.  Stays inside the reference's
    defined behaviour: materials >= 0, texture coordinates well above 0, no exact-zero depths."""
    rng = np.random.default_rng(seed)
    u = lambda lo, hi, *shape: rng.uniform(lo, hi, size=shape).astype(np.float32)
    s = Scene()
    s.name = f"fuzz{seed}"
    # nodes: 0 root (identity), 1 child of 0, 2 child of 1 (mirrored), 3 another root
    nn = 4
    s.node_parent = np.array([-1, 0, 1, -1], np.int32)
    s.node_scale = np.ones((nn, 3), np.float32)
    s.node_rotation = np.stack([identity44() for _ in range(nn)]).astype(np.float32)
    s.node_translation = np.zeros((nn, 3), np.float32)
    for i in (1, 2, 3):
        q = rng.normal(size=4).astype(np.float32)
        q = (q / np.float32(np.sqrt(np.float32((q * q).sum())))).astype(np.float32)
        s.node_rotation[i] = from_quaternion(*q)
        s.node_scale[i] = u(0.5, 1.6, 3)
        s.node_translation[i] = u(-0.8, 0.8, 3)
    s.node_scale[2, int(rng.integers(0, 3))] *= np.float32(-1.0)                  # mirrored: winding flips
    # textures + materials
    sizes = [(64, 64), (37, 53), (16, 128), (5, 3)]
    for (tw, th) in sizes:
        t = rng.integers(0, 1 << 24, size=(th, tw), dtype=np.uint32)
        a = rng.integers(40, 256, size=(th, tw), dtype=np.uint32) if transparent else np.full((th, tw), 255, np.uint32)
        s.textures.append((t | (a << 24)).astype(np.uint32))
    nm = 6
    s.mat_bgra = rng.integers(0, 256, size=(nm, 4), dtype=np.uint8)
    s.mat_bgra[:, 3] = rng.integers(60, 256, size=nm, dtype=np.uint8) if transparent else 255
    s.mat_metal_rough = np.ones((nm, 2), np.float32)
    s.mat_tex_ds = np.stack([np.array([0, 1, 2, 3, -1, -1], np.int32), rng.integers(0, 2, size=nm).astype(np.int32)], axis=1)
    pos, nrm, uv, idx = [], [], [], []
    pn, pm, pmat, pfv, pnv, pfi, pni = [], [], [], [], [], [], []
    modes = [_abi.MODE_TRIANGLES, _abi.MODE_TRIANGLE_STRIP, _abi.MODE_TRIANGLE_FAN]
    prim_nodes = sorted(int(k) for k in rng.integers(0, nn, size=n_prims))      # the reference draws node by node (renderer.cpp:86-97)
    for p in range(n_prims):
        nv = verts_per_prim
        v = np.empty((nv, 3), np.float32)
        v[:, 0] = u(-2.5, 2.5, nv); v[:, 1] = u(-2.0, 2.0, nv)
        v[:, 2] = u(-6.5, 3.0, nv)                                                # camera at z = -4: some behind it, some across the near plane
        snap = rng.random(nv) < 0.35
        v[snap] = (np.round(v[snap] * 2.0) / 2.0).astype(np.float32)              # grid: shared positions, equal depths
        kind = p % 5
        if kind == 1:                                                             # a few very large triangles
            v[:6] *= np.float32(6.0)
        elif kind == 2:                                                           # a cloud of sub-pixel triangles
            c = u(-1, 1, 3)
            v[: nv // 2] = c + u(-0.004, 0.004, nv // 2, 3)
        elif kind == 3:                                                           # collinear triples
            d = u(-1, 1, 3)
            for k in range(0, 9, 3):
                v[k + 1] = v[k] + d; v[k + 2] = v[k] + np.float32(2.0) * d
        n = u(-1, 1, nv, 3)
        n[rng.random(nv) < 0.5] *= np.float32(3.0)                                # not unit length (stored as is, gltf.cpp:192-194)
        t = u(1.5, 4.0, nv, 2)                                                    # >= 1.5: slivers extrapolate a little, and a negative texel index is UB in the reference
        mode = modes[p % 3]
        ni = int(rng.integers(12, 3 * nv)) if mode == _abi.MODE_TRIANGLES else int(rng.integers(5, nv))
        if mode == _abi.MODE_TRIANGLES:
            ni -= ni % 3
        ix = rng.integers(0, nv, size=ni).astype(np.uint32)                       # repeats -> degenerate triangles
        pfv.append(sum(pnv)); pnv.append(nv); pfi.append(sum(pni)); pni.append(ni)
        pn.append(prim_nodes[p]); pm.append(mode); pmat.append(int(rng.integers(0, nm)))
        pos.append(v); nrm.append(n); uv.append(t); idx.append(ix)
    # coplanar duplicate of the last primitive under another material, drawn after it: every fragment ties in depth
    pfv.append(sum(pnv)); pnv.append(pnv[-1]); pfi.append(sum(pni)); pni.append(pni[-1])
    pn.append(pn[-1]); pm.append(pm[-1]); pmat.append((pmat[-1] + 1) % nm)
    pos.append(pos[-1].copy()); nrm.append(nrm[-1].copy()); uv.append(uv[-1].copy()); idx.append(idx[-1].copy())
    s.positions = np.concatenate(pos).astype(np.float32)
    s.normals = np.concatenate(nrm).astype(np.float32)
    s.texcoords = np.concatenate(uv).astype(np.float32)
    s.indices = np.concatenate(idx).astype(np.uint32)
    s.prim_node, s.prim_mode, s.prim_material = (np.array(a, np.int32) for a in (pn, pm, pmat))
    s.prim_first_vertex, s.prim_n_vertices, s.prim_first_index, s.prim_n_indices = (np.array(a, np.uint32) for a in (pfv, pnv, pfi, pni))
    nl = int(rng.integers(0, 4))
    lights = [(float(u(-3, 3)), float(u(-1, 4)), float(u(-4, 2)), float(rng.choice([0.3, 2.0, 40.0, 300.0]))) for _ in range(nl)]
    s.set_lights(float(u(0.05, 0.5)), (float(u(-1, 1)), float(u(-2, -0.2)), float(u(-1, 1))), float(u(0.2, 0.9)), lights)
    return s


def fuzz_case(seed, screen=(416, 312)):
    """-> (scene, viewport, screen, pose): shader combination, rectangle, post pass and camera drawn from the seed"""
    rng = np.random.default_rng(1000 + seed)
    scene = fuzz_scene(seed)
    light = int(rng.integers(0, 3)); tex = int(rng.integers(0, 3))
    if seed % 3 == 0:
        light, tex = _abi.LIGHT_PHONG, _abi.TEX_BILINEAR
    x, y = int(rng.integers(0, 24)), int(rng.integers(0, 24))
    w, h = screen[0] - x - int(rng.integers(0, 24)), screen[1] - y - int(rng.integers(0, 24))
    post = _abi.POST_DOF if seed % 4 == 1 else _abi.POST_NULL
    vp = Viewport(x, y, w, h, light_mode=light, tex_mode=tex, transparency_layers=0, post_mode=post,
                  focal_distance=float(rng.uniform(2, 6)), focal_depth=float(rng.uniform(1.5, 6)))
    pose = [("translate", float(rng.uniform(-0.8, 0.8)), float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-5.0, -2.5))),
            ("rotate_y", float(rng.uniform(-0.5, 0.5))), ("rotate_x", float(rng.uniform(-0.4, 0.4))),
            ("rotate_z", float(rng.uniform(-0.3, 0.3)))]
    vp.camera.apply(pose)
    vp.pose = pose
    return scene, vp, screen, pose


def fuzz_layers_case(seed, screen=(416, 312)):
    """the transparent flavour of fuzz_case: material and texel alpha random, 1..3 transparency layers, viewport at the
    screen origin (the reference's flatten ignores the viewport offset, viewport.cpp:61,74)"""
    _, vp0, _, pose = fuzz_case(seed, screen)
    scene = fuzz_scene(seed, transparent=True)
    vp = Viewport(0, 0, screen[0], screen[1], light_mode=vp0.light_mode, tex_mode=vp0.tex_mode, transparency_layers=1 + seed % 3)
    vp.camera.apply(pose)
    vp.pose = pose
    return scene, vp, screen, pose
