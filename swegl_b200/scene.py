"""Host-side scene containers for the Python harness (tests, bench).

The C++ drop-in (swegl_b200/host/swegl_b200_adapter.hpp) reads swegl's own scene_t/viewport_t;
this module is the same thing for Python callers: a flattened scene in the layout
include/swegl_b200.h wants, plus bit-exact restatements of the tiny pieces of host math the
reference does per frame on the CPU and that stay on the host in this design:

  * camera_t                     swegl/projection/camera.hpp:10-24, src/projection/camera.cpp:8-58
  * matrix44_t::rotate_{x,y,z}   src/projection/matrix44.cpp:11-66
  * viewport matrix              src/render/viewport.cpp:30-35
  * node_t::get_local_world_matrix + hierarchy product   swegl/data/model.hpp:56-61,
                                 swegl/render/vertex_shaders.hpp:16-33 (freon::operator*, see
                                 oracle/shims/freon/Matrix.hpp for the product order used)

All arithmetic is done on numpy float32 scalars so every operation rounds like the C++ code.
"""
import io
import json
import math
import os
import zipfile

import numpy as np

from . import _abi

f32 = np.float32


# matrix44_t::rotate_* call cos(a)/sin(a) on a float: C++ overload resolution picks the float versions, and
# g++ -O3 merges them into glibc's sincosf (src/projection/matrix44.cpp:13-14).  glibc's float functions are
# not correctly rounded (sinf(2.9f) is 1 ulp off), so the same libm entry points are called here.
import ctypes as _C
import ctypes.util as _Cu

_libm = _C.CDLL(_Cu.find_library("m") or "libm.so.6")
_libm.cosf.restype = _libm.sinf.restype = _C.c_float
_libm.cosf.argtypes = _libm.sinf.argtypes = [_C.c_float]


def _cosf(a):
    return f32(_libm.cosf(float(f32(a))))


def _sinf(a):
    return f32(_libm.sinf(float(f32(a))))


def identity44():
    return np.eye(4, dtype=np.float32)


def _rotate_rows(m, a, r0, r1):
    """matrix44_t::rotate_*: new_r0 = cos*old_r0 + sin*old_r1 ; new_r1 = -sin*old_r0 + cos*old_r1."""
    c, s = _cosf(a), _sinf(a)
    old0 = m[r0].copy()
    old1 = m[r1].copy()
    for j in range(4):
        m[r0, j] = f32(c * old0[j]) + f32(s * old1[j])
        m[r1, j] = f32(f32(-s) * old0[j]) + f32(c * old1[j])


def rotate_x(m, a):
    _rotate_rows(m, a, 1, 2)


def rotate_y(m, a):
    _rotate_rows(m, a, 0, 2)


def rotate_z(m, a):
    _rotate_rows(m, a, 0, 1)


def matmul44(a, b):
    """freon::operator* as defined by oracle/shims/freon/Matrix.hpp: s = 0; s += a[i][k]*b[k][j]."""
    out = np.zeros((4, 4), dtype=np.float32)
    for i in range(4):
        for j in range(4):
            s = f32(0)
            for k in range(4):
                s = f32(s + f32(a[i, k] * b[k, j]))
            out[i, j] = s
    return out


def from_quaternion(q0, q1, q2, q3):
    """matrix44_t::from_quaternion, swegl/projection/matrix44.hpp:34-43 (argument order as written)."""
    q0, q1, q2, q3 = f32(q0), f32(q1), f32(q2), f32(q3)
    two, one = f32(2), f32(1)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = two * (q0 * q0 + q1 * q1) - one
    m[0, 1] = two * (q1 * q2 - q0 * q3)
    m[0, 2] = two * (q1 * q3 + q0 * q2)
    m[1, 0] = two * (q1 * q2 + q0 * q3)
    m[1, 1] = two * (q0 * q0 + q2 * q2) - one
    m[1, 2] = two * (q2 * q3 - q0 * q1)
    m[2, 0] = two * (q1 * q3 - q0 * q2)
    m[2, 1] = two * (q2 * q3 + q0 * q1)
    m[2, 2] = two * (q0 * q0 + q3 * q3) - one
    m[3, 3] = one
    return m


class Camera:
    """camera_t: view matrix, projection matrix and position (m_center)."""

    def __init__(self, aspect_ratio):
        aspect = f32(aspect_ratio)
        self.center = np.zeros(3, dtype=np.float32)
        self.view = identity44()
        self.proj = identity44()
        self.view[2, 2] = f32(-1)
        n, f, w, h = f32(0.5), f32(10.0), f32(1.0), f32(1.0)
        two = f32(2)
        if aspect > 1:
            self.proj[0, 0] = f32(two * n) / w
            self.proj[1, 1] = f32(f32(aspect * two) * n) / h
        else:
            self.proj[0, 0] = f32(f32(f32(f32(1.0) / aspect) * two) * n) / w
            self.proj[1, 1] = f32(two * n) / h
        self.proj[2, 2] = f / f32(f - n)
        self.proj[2, 3] = f32(f32(-f) * n) / f32(f - n)
        self.proj[3, 2] = f32(1)

    def rotate_x(self, a):
        rotate_x(self.view, -f32(a))

    def rotate_y(self, a):
        rotate_y(self.view, -f32(a))

    def rotate_z(self, a):
        rotate_z(self.view, -f32(a))

    def translate(self, x, y, z):
        x, y, z = f32(x), f32(y), f32(z)
        m = self.view
        m[0, 3] = f32(m[0, 3] + f32(-x))
        m[1, 3] = f32(m[1, 3] + f32(-y))
        m[2, 3] = f32(m[2, 3] + f32(-z))
        for c in range(3):
            self.center[c] = f32(self.center[c] + f32(f32(f32(x * m[0, c]) + f32(y * m[1, c])) + f32(z * m[2, c])))

    def apply(self, ops):
        """ops: sequence of ("translate", x, y, z) / ("rotate_x", a) / ("rotate_y", a) / ("rotate_z", a)."""
        for op in ops:
            getattr(self, op[0])(*op[1:])
        return self


class Viewport:
    """viewport_t: rectangle, camera, shader selection and post pass (swegl/render/viewport.hpp:27-62)."""

    def __init__(self, x, y, w, h, light_mode=_abi.LIGHT_PHONG, tex_mode=_abi.TEX_BILINEAR,
                 transparency_layers=0, post_mode=_abi.POST_NULL, focal_distance=5.0, focal_depth=5.0):
        self.x, self.y, self.w, self.h = int(x), int(y), int(w), int(h)
        self.camera = Camera(1.0 * w / h)          # viewport.cpp:25: m_camera(1.0*w/h)
        self.light_mode, self.tex_mode = light_mode, tex_mode
        self.transparency_layers = transparency_layers
        self.post_mode, self.focal_distance, self.focal_depth = post_mode, focal_distance, focal_depth
        self.band = (0, 0)

    def desc(self):
        d = _abi.ViewportDesc()
        d.x, d.y, d.w, d.h = self.x, self.y, self.w, self.h
        d.view[:] = [float(v) for v in self.camera.view.reshape(-1)]
        d.proj[:] = [float(v) for v in self.camera.proj.reshape(-1)]
        d.cam_pos[:] = [float(v) for v in self.camera.center]
        half = f32(2.0)
        d.vp_m00 = float(f32(self.w) / half)                       # viewport.cpp:30-35
        d.vp_m03 = float(f32(f32(self.x) + f32(self.w) / half))
        d.vp_m11 = float(-(f32(self.h) / half))
        d.vp_m13 = float(f32(f32(self.y) + f32(self.h) / half))
        d.light_mode, d.tex_mode, d.post_mode = self.light_mode, self.tex_mode, self.post_mode
        d.focal_distance, d.focal_depth = self.focal_distance, self.focal_depth
        d.transparency_layers = self.transparency_layers
        d.band_y0, d.band_y1 = self.band
        return d


def normalized3(x, y, z):
    """normal_t(x,y,z): vector_t::normalize, swegl/projection/points.hpp:71-90,139-142."""
    x, y, z = f32(x), f32(y), f32(z)
    l = f32(math.sqrt(float(f32(f32(f32(x * x) + f32(y * y)) + f32(z * z)))))
    if l != 0:
        x, y, z = f32(x / l), f32(y / l), f32(z / l)
    return np.array([x, y, z], dtype=np.float32)


class Scene:
    """Flattened scene_t (swegl/data/model.hpp:128-145): static geometry + per-frame node/light state."""

    ARRAYS = ["node_scale", "node_rotation", "node_translation", "node_parent",
              "prim_node", "prim_mode", "prim_material", "prim_first_vertex", "prim_n_vertices",
              "prim_first_index", "prim_n_indices", "positions", "normals", "texcoords", "indices",
              "mat_bgra", "mat_metal_rough", "mat_tex_ds"]
    # scene_t::animations flattened (model.hpp:88-125); optional in a scene pack (packs made before they existed load empty)
    ANIM_ARRAYS = ["anim_end_time", "chan_anim", "chan_node", "chan_path", "chan_first_step", "chan_n_steps",
                   "step_time", "step_value"]
    PATH_SCALE, PATH_ROTATION, PATH_TRANSLATION, PATH_WEIGHTS = 0, 1, 2, 3     # animation_channel_t::path_t

    def __init__(self):
        self.node_scale = np.zeros((0, 3), np.float32)
        self.node_rotation = np.zeros((0, 4, 4), np.float32)
        self.node_translation = np.zeros((0, 3), np.float32)
        self.node_parent = np.zeros(0, np.int32)
        for n in ["prim_node", "prim_mode", "prim_material"]:
            setattr(self, n, np.zeros(0, np.int32))
        for n in ["prim_first_vertex", "prim_n_vertices", "prim_first_index", "prim_n_indices", "indices"]:
            setattr(self, n, np.zeros(0, np.uint32))
        self.positions = np.zeros((0, 3), np.float32)
        self.normals = np.zeros((0, 3), np.float32)
        self.texcoords = np.zeros((0, 2), np.float32)
        self.mat_bgra = np.zeros((0, 4), np.uint8)
        self.mat_metal_rough = np.zeros((0, 2), np.float32)
        self.mat_tex_ds = np.zeros((0, 2), np.int32)
        self.anim_end_time = np.zeros(0, np.float32)
        for n in ["chan_anim", "chan_node", "chan_path"]:
            setattr(self, n, np.zeros(0, np.int32))
        for n in ["chan_first_step", "chan_n_steps"]:
            setattr(self, n, np.zeros(0, np.uint32))
        self.step_time = np.zeros(0, np.float32)
        self.step_value = np.zeros((0, 4), np.float32)
        self.textures = []                          # list of (h, w) uint32 BGRA arrays
        self.default_material = (255, 255, 255, 255, 1.0, 1.0, -1, 0)   # material_t defaults, model.hpp:80-87
        # lights: test_1.cpp:334-336 defaults
        self.ambient = 0.3
        self.sun_raw = (1.0, -2.0, -1.0)
        self.sun_dir = normalized3(*self.sun_raw)
        self.sun_intensity = 0.7
        self.point_lights = np.zeros((0, 4), np.float32)
        self.name = "scene"
        self._keep = []

    # ---- counts ----
    @property
    def n_nodes(self):
        return len(self.node_parent)

    @property
    def n_primitives(self):
        return len(self.prim_node)

    @property
    def n_vertices(self):
        return len(self.positions)

    def n_triangles(self):
        t = 0
        for mode, n in zip(self.prim_mode, self.prim_n_indices):
            n = int(n)
            if mode == _abi.MODE_TRIANGLES:
                t += n // 3
            elif n >= 3:
                t += n - 2
        return t

    def set_lights(self, ambient, sun_xyz, sun_intensity, point_lights=()):
        self.ambient = ambient
        self.sun_raw = tuple(float(v) for v in sun_xyz)
        self.sun_dir = normalized3(*sun_xyz)
        self.sun_intensity = sun_intensity
        self.point_lights = np.array(point_lights, dtype=np.float32).reshape(-1, 4)
        return self

    # ---- per-frame host math ----
    def animate(self, elapsed_seconds):
        """scene_t::animate (swegl/data/model.hpp:146-177): every channel's two key frames around
        fmod(elapsed, end_time) are blended linearly in fp32 and written to the node's rotation (quaternion
        normalised, then matrix44_t::from_quaternion), translation or scale.  The app calls it once per frame
        (src/test_1.cpp:378); node_matrices() / begin_frame pick the new TRS up."""
        t_in = f32(elapsed_seconds)
        for c in range(len(self.chan_node)):
            end = f32(self.anim_end_time[self.chan_anim[c]])
            rel = f32(math.fmod(float(t_in), float(end)))                   # fmod is exact: float and double agree
            s0, n = int(self.chan_first_step[c]), int(self.chan_n_steps[c])
            times = self.step_time[s0:s0 + n]
            it = int(np.searchsorted(times, rel, side="left"))              # std::lower_bound on step.time < time
            if it == n:
                b = a = s0 + n - 1
            elif it == 0:
                b = a = s0
            else:
                b, a = s0 + it - 1, s0 + it
            tb, ta = f32(self.step_time[b]), f32(self.step_time[a])
            vb, va = self.step_value[b].astype(np.float32), self.step_value[a].astype(np.float32)
            if ta == tb:
                frame = vb.copy()
            else:
                wb = f32(f32(ta - rel) / f32(ta - tb))
                wa = f32(f32(rel - tb) / f32(ta - tb))
                frame = ((vb * wb).astype(np.float32) + (va * wa).astype(np.float32)).astype(np.float32)
            node, path = int(self.chan_node[c]), int(self.chan_path[c])
            if path == self.PATH_ROTATION:
                x, y, z, w = (f32(v) for v in frame)
                ln = f32(math.sqrt(float(f32(f32(f32(f32(x * x) + f32(y * y)) + f32(z * z)) + f32(w * w)))))   # vec2f.hpp:67-77
                if ln != 0:
                    x, y, z, w = f32(x / ln), f32(y / ln), f32(z / ln), f32(w / ln)
                self.node_rotation[node] = from_quaternion(x, y, z, w)
            elif path == self.PATH_TRANSLATION:
                self.node_translation[node] = frame[:3]
            elif path == self.PATH_SCALE:
                self.node_scale[node] = frame[:3]
        return self

    @property
    def n_animations(self):
        return len(self.anim_end_time)

    def node_matrices(self):
        """original_to_world_matrix (n,4,4) and the 3x3 of scale(rotation, scale) (n,3,3) per node."""
        n = self.n_nodes
        local = np.zeros((n, 4, 4), np.float32)
        for i in range(n):
            m = self.node_rotation[i].astype(np.float32).copy()
            for c in range(3):                       # scale(): columns 0..2 of all four rows, points.cpp:14-30
                m[:, c] = (m[:, c] * self.node_scale[i, c]).astype(np.float32)
            for r in range(3):                       # translate(): model.hpp:59
                m[r, 3] = f32(m[r, 3] + self.node_translation[i, r])
            local[i] = m
        world = np.zeros((n, 4, 4), np.float32)
        done = np.zeros(n, bool)

        def resolve(i):
            if done[i]:
                return
            p = int(self.node_parent[i])
            if p < 0:
                world[i] = matmul44(identity44(), local[i])
            else:
                resolve(p)
                world[i] = matmul44(world[p], local[i])
            done[i] = True

        for i in range(n):
            resolve(i)
        normal = np.ascontiguousarray(local[:, :3, :3])
        # the translation is not part of the 3x3, so `local` before translate == after for these entries
        return world, normal

    # ---- ABI descriptors (keep numpy buffers alive in self._keep) ----
    def scene_desc(self):
        import ctypes as C
        d = _abi.SceneDesc()
        keep = []
        prims = (_abi.Primitive * max(1, self.n_primitives))()
        for i in range(self.n_primitives):
            prims[i] = _abi.Primitive(int(self.prim_node[i]), int(self.prim_mode[i]), int(self.prim_material[i]),
                                      int(self.prim_first_vertex[i]), int(self.prim_n_vertices[i]),
                                      int(self.prim_first_index[i]), int(self.prim_n_indices[i]))
        mats = (_abi.Material * max(1, len(self.mat_bgra)))()
        for i in range(len(self.mat_bgra)):
            b, g, r, a = (int(v) for v in self.mat_bgra[i])
            mats[i] = _abi.Material(b, g, r, a, float(self.mat_metal_rough[i, 0]), float(self.mat_metal_rough[i, 1]),
                                    int(self.mat_tex_ds[i, 0]), int(self.mat_tex_ds[i, 1]))
        texs = (_abi.Texture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            t = np.ascontiguousarray(t, dtype=np.uint32)
            keep.append(t)
            texs[i] = _abi.Texture(t.ctypes.data_as(C.POINTER(C.c_uint32)), t.shape[1], t.shape[0])

        def fptr(a, dt, ct):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(ct))

        d.n_nodes, d.n_primitives, d.n_vertices = self.n_nodes, self.n_primitives, self.n_vertices
        d.n_indices, d.n_materials, d.n_textures = len(self.indices), len(self.mat_bgra), len(self.textures)
        d.primitives = prims
        d.positions = fptr(self.positions, np.float32, C.c_float)
        d.normals = fptr(self.normals, np.float32, C.c_float)
        d.texcoords = fptr(self.texcoords, np.float32, C.c_float)
        d.indices = fptr(self.indices, np.uint32, C.c_uint32)
        d.materials = mats
        dm = self.default_material
        d.default_material = _abi.Material(dm[0], dm[1], dm[2], dm[3], dm[4], dm[5], dm[6], dm[7])
        d.textures = texs
        keep += [prims, mats, texs]
        self._keep = keep
        return d

    def frame_desc(self, node_world=None, node_normal=None, lights_only=False):
        import ctypes as C
        d = _abi.FrameDesc()
        if not lights_only:
            if node_world is None:
                node_world, node_normal = self.node_matrices()
            self._nw = np.ascontiguousarray(node_world, dtype=np.float32).reshape(-1)
            self._nn = np.ascontiguousarray(node_normal, dtype=np.float32).reshape(-1)
            d.node_world = self._nw.ctypes.data_as(C.POINTER(C.c_float))
            d.node_normal = self._nn.ctypes.data_as(C.POINTER(C.c_float))
        self._pl = np.ascontiguousarray(self.point_lights, dtype=np.float32).reshape(-1)
        d.ambient = float(f32(self.ambient))
        d.sun_dir[:] = [float(v) for v in self.sun_dir]
        d.sun_intensity = float(f32(self.sun_intensity))
        d.n_point_lights = len(self.point_lights)
        d.point_lights = self._pl.ctypes.data_as(C.POINTER(C.c_float))
        return d

    def animation_desc(self):
        """swegl_b200_animation_desc of the scene AS IT IS NOW (take it before the first animate(): the device evaluates
        scene_t::animate from the base TRS, model.hpp:146-177)"""
        import ctypes as C
        d = _abi.AnimationDesc()
        n, nc = self.n_nodes, len(self.chan_node)
        keep = [np.ascontiguousarray(self.node_parent, dtype=np.int32), np.ascontiguousarray(self.node_rotation, dtype=np.float32),
                np.ascontiguousarray(self.node_translation, dtype=np.float32), np.ascontiguousarray(self.node_scale, dtype=np.float32),
                np.ascontiguousarray(self.anim_end_time, dtype=np.float32), np.ascontiguousarray(self.step_time, dtype=np.float32),
                np.ascontiguousarray(self.step_value, dtype=np.float32)]
        chans = (_abi.AnimChannel * max(1, nc))()
        for c in range(nc):
            chans[c] = _abi.AnimChannel(int(self.chan_anim[c]), int(self.chan_node[c]), int(self.chan_path[c]),
                                        int(self.chan_first_step[c]), int(self.chan_n_steps[c]))
        d.n_nodes = n
        d.node_parent = keep[0].ctypes.data_as(C.POINTER(C.c_int32))
        d.node_rotation = keep[1].ctypes.data_as(C.POINTER(C.c_float))
        d.node_translation = keep[2].ctypes.data_as(C.POINTER(C.c_float))
        d.node_scale = keep[3].ctypes.data_as(C.POINTER(C.c_float))
        d.n_animations = len(self.anim_end_time)
        d.end_time = keep[4].ctypes.data_as(C.POINTER(C.c_float))
        d.n_channels = nc
        d.channels = chans
        d.n_steps = len(self.step_time)
        d.step_time = keep[5].ctypes.data_as(C.POINTER(C.c_float))
        d.step_value = keep[6].ctypes.data_as(C.POINTER(C.c_float))
        assert keep[6].size == 4 * len(self.step_time)
        self._keep_anim = keep + [chans]
        return d

    # ---- scene packs (assets/*.scenepack): npz-style zip of the arrays + encoded images ----
    def save_pack(self, path, encoded_images=None):
        """encoded_images: list of original PNG/JPEG byte strings (kept small); else raw texels are stored."""
        with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
            meta = {"format": "swegl_b200.scenepack.v1", "name": self.name,
                    "default_material": list(self.default_material),
                    "textures": []}
            for name in self.ARRAYS + (self.ANIM_ARRAYS if self.n_animations else []):
                buf = io.BytesIO()
                np.save(buf, getattr(self, name))
                z.writestr(name + ".npy", buf.getvalue())
            for i, t in enumerate(self.textures):
                entry = {"w": int(t.shape[1]), "h": int(t.shape[0]), "sha256": texel_digest(t)}
                if encoded_images is not None and encoded_images[i] is not None:
                    z.writestr(f"image_{i}.bin", encoded_images[i])
                    entry["encoding"] = "image"
                else:
                    buf = io.BytesIO()
                    np.save(buf, np.ascontiguousarray(t, dtype=np.uint32))
                    z.writestr(f"image_{i}.npy", buf.getvalue())
                    entry["encoding"] = "raw"
                meta["textures"].append(entry)
            z.writestr("meta.json", json.dumps(meta, indent=1))

    @staticmethod
    def load_glb(path):
        """see load_glb() below"""
        return load_glb(path)

    @classmethod
    def load_pack(cls, path):
        s = cls()
        with zipfile.ZipFile(path) as z:
            meta = json.loads(z.read("meta.json"))
            if meta.get("format") != "swegl_b200.scenepack.v1":
                raise ValueError(f"{path}: not a swegl_b200 scene pack")
            s.name = meta["name"]
            s.default_material = tuple(meta["default_material"])
            for name in cls.ARRAYS:
                setattr(s, name, np.load(io.BytesIO(z.read(name + ".npy"))))
            if "anim_end_time.npy" in z.namelist():
                for name in cls.ANIM_ARRAYS:
                    setattr(s, name, np.load(io.BytesIO(z.read(name + ".npy"))))
            for i, entry in enumerate(meta["textures"]):
                if entry["encoding"] == "raw":
                    t = np.load(io.BytesIO(z.read(f"image_{i}.npy")))
                else:
                    t = decode_image(z.read(f"image_{i}.bin"))
                if texel_digest(t) != entry["sha256"]:
                    raise ValueError(f"{path}: image {i} decodes to different texels than when the pack was made "
                                     "(image decoder drift); golden frames would not match")
                s.textures.append(t)
        return s


def load_glb(path):
    """A .glb / .gltf file -> Scene, the way the reference's loader reads it (src/data/gltf.cpp:56-436) -- quirks
    included, because they shape what gets drawn: indices are read as uint16 whatever their
    component type (:219), TEXCOORD_0 lands swapped (v in tex_coords.x, u in .y, :205-206), baseColorFactor becomes
    (uchar)(255 * f) in b,g,r,a order (:130-134), a node's glTF quaternion [x,y,z,w] is passed as
    from_quaternion(q0..q3) (:234-238), `matrix` nodes are decomposed with row lengths but column division (:279-306),
    a mesh referenced by several nodes is moved into the first one only (:308-309), skins / morph targets / cameras /
    samplers / scenes are ignored.  Embedded PNG / JPEG images are decoded by
    decode_image (the package's own decoder).  Returns the same arrays tools/make_scenepacks.py exports from the
    reference's own loader (tests/test_load_glb.py compares them bit for bit)."""
    import struct
    raw = open(path, "rb").read()
    root = os.path.dirname(os.path.abspath(path))
    if raw[:4] == b"glTF":
        chunks, pos = [], 12
        total = struct.unpack_from("<I", raw, 8)[0]
        while pos + 8 <= min(total, len(raw)):
            ln, _ty = struct.unpack_from("<II", raw, pos)
            chunks.append((pos + 8, ln))
            pos += 8 + ln
        j = json.loads(raw[chunks[0][0]:chunks[0][0] + chunks[0][1]])
        glb_bin = chunks[1] if len(chunks) > 1 else (0, 0)
    else:
        j, glb_bin = json.loads(raw), None
    buffers = []
    for jb in j.get("buffers", []):
        if "uri" in jb:
            buffers.append(open(os.path.join(root, jb["uri"]), "rb").read()[: jb["byteLength"]])
        else:
            buffers.append(raw[glb_bin[0]:glb_bin[0] + jb["byteLength"]])
    views = []
    for bv in j.get("bufferViews", []):
        b = buffers[bv["buffer"]]
        o = bv.get("byteOffset", 0)
        views.append((b[o:o + bv["byteLength"]], bv.get("byteStride", 0)))
    comp_size = {5120: 1, 5121: 1, 5122: 2, 5123: 2, 5125: 4, 5126: 4}
    n_comp = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}

    def accessor(i):                                    # accessor_t, gltf.cpp:13-51 -> (bytes, stride, count)
        a = j["accessors"][i]
        data, stride = views[a["bufferView"]]
        if stride == 0:
            stride = comp_size[a["componentType"]] * n_comp[a["type"]]
        o = a.get("byteOffset", 0)
        return data[o:o + stride * a["count"]], stride, a["count"]

    def floats(acc, n, first=0):                        # n consecutive floats at byte `first` of every element
        data, stride, count = acc
        out = np.zeros((count, n), np.float32)
        for k in range(n):
            out[:, k] = np.frombuffer(b"".join(data[i * stride + first + 4 * k: i * stride + first + 4 * k + 4] for i in range(count)), "<f4")
        return out
    s = Scene()
    s.name = os.path.splitext(os.path.basename(path))[0]
    for im in j.get("images", []):
        if "uri" in im:
            s.textures.append(decode_image(open(os.path.join(root, im["uri"]), "rb").read()))
        else:
            s.textures.append(decode_image(views[im["bufferView"]][0]))
    mats = j.get("materials", [])
    s.mat_bgra = np.full((len(mats), 4), 255, np.uint8)
    s.mat_metal_rough = np.ones((len(mats), 2), np.float32)
    s.mat_tex_ds = np.zeros((len(mats), 2), np.int32)
    s.mat_tex_ds[:, 0] = -1
    for m, mat in enumerate(mats):
        pbr = mat.get("pbrMetallicRoughness")
        if pbr is None:
            continue                                    # material_t{color, metallic, roughness, img_idx}: double_sided stays false (:124-128)
        s.mat_tex_ds[m, 1] = 1 if mat.get("doubleSided", False) else 0
        if "baseColorFactor" in pbr:
            f = [f32(v) for v in pbr["baseColorFactor"]]
            s.mat_bgra[m] = [int(f32(f32(255) * f[2])) & 255, int(f32(f32(255) * f[1])) & 255, int(f32(f32(255) * f[0])) & 255,
                             int(f32(f32(255) * f[3])) & 255]
        if "baseColorTexture" in pbr:
            s.mat_tex_ds[m, 0] = j["textures"][pbr["baseColorTexture"]["index"]]["source"]
        s.mat_metal_rough[m] = [f32(pbr.get("metallicFactor", 1.0)), f32(pbr.get("roughnessFactor", 1.0))]
    meshes = []                                         # temp_meshes: only meshes that have primitives get a slot (:155-161)
    for mesh in j.get("meshes", []):
        if "primitives" not in mesh:
            continue
        prims = []
        for pr in mesh["primitives"]:
            at = pr["attributes"]
            pos = floats(accessor(at["POSITION"]), 3)
            nv = len(pos)                               # (the two spare vertices of :175 are the CPU clipper's scratch space: not part of a Scene)
            P, N, T = np.zeros((nv, 3), np.float32), np.zeros((nv, 3), np.float32), np.zeros((nv, 2), np.float32)
            P[:len(pos)] = pos
            if "NORMAL" in at:
                nr = floats(accessor(at["NORMAL"]), 3)
                N[:len(nr)] = nr
            if "TEXCOORD_0" in at:
                acc = accessor(at["TEXCOORD_0"])
                T[:acc[2], 0] = floats(acc, 1, first=4)[:, 0]
                T[:acc[2], 1] = floats(acc, 1, first=0)[:, 0]
            idx = np.zeros(0, np.uint32)
            if "indices" in pr:
                data, stride, count = accessor(pr["indices"])
                idx = np.array([struct.unpack_from("<H", data, i * stride)[0] for i in range(count)], np.uint32)
            prims.append(dict(material=pr.get("material", -1), mode=pr.get("mode", 4), P=P, N=N, T=T, idx=idx))
        meshes.append(prims)
    nodes = j.get("nodes", [])
    n = len(nodes)
    s.node_scale = np.ones((n, 3), np.float32)
    s.node_rotation = np.stack([identity44() for _ in range(n)]).astype(np.float32) if n else np.zeros((0, 4, 4), np.float32)
    s.node_translation = np.zeros((n, 3), np.float32)
    s.node_parent = np.full(n, -1, np.int32)
    pn, pm, pmat, pfv, pnv, pfi, pni, Ps, Ns, Ts, Is = [], [], [], [], [], [], [], [], [], [], []
    for i, nd in enumerate(nodes):
        if "rotation" in nd:
            s.node_rotation[i] = from_quaternion(*[f32(v) for v in nd["rotation"]])
        if "translation" in nd:
            s.node_translation[i] = [f32(v) for v in nd["translation"]]
        if "scale" in nd:
            s.node_scale[i] = [f32(v) for v in nd["scale"]]
        if "matrix" in nd:
            m = np.array([f32(v) for v in nd["matrix"]], np.float32).reshape(4, 4).T.copy()     # column-major file, row-major swegl
            s.node_translation[i] = m[:3, 3]
            sc = [f32(math.sqrt(float(f32(f32(f32(m[r, 0] * m[r, 0]) + f32(m[r, 1] * m[r, 1])) + f32(m[r, 2] * m[r, 2]))))) for r in range(3)]
            s.node_scale[i] = sc
            m[:3, 3] = 0
            for c in range(3):                          # rows measured, COLUMNS divided (:279-306)
                if sc[c] != 0:
                    for r in range(3):
                        m[r, c] = f32(m[r, c] / sc[c])
            s.node_rotation[i] = m
        if "mesh" in nd:
            for pr in meshes[nd["mesh"]]:
                pn.append(i); pm.append(pr["mode"]); pmat.append(pr["material"])
                pfv.append(sum(pnv)); pnv.append(len(pr["P"])); pfi.append(sum(pni)); pni.append(len(pr["idx"]))
                Ps.append(pr["P"]); Ns.append(pr["N"]); Ts.append(pr["T"]); Is.append(pr["idx"])
            meshes[nd["mesh"]] = []                     # std::move: the next node naming this mesh gets nothing (:308-309)
        for ch in nd.get("children", []):
            s.node_parent[ch] = i
    s.prim_node, s.prim_mode, s.prim_material = (np.array(a, np.int32) for a in (pn, pm, pmat))
    s.prim_first_vertex, s.prim_n_vertices, s.prim_first_index, s.prim_n_indices = (np.array(a, np.uint32) for a in (pfv, pnv, pfi, pni))
    s.positions = np.concatenate(Ps).astype(np.float32) if Ps else np.zeros((0, 3), np.float32)
    s.normals = np.concatenate(Ns).astype(np.float32) if Ns else np.zeros((0, 3), np.float32)
    s.texcoords = np.concatenate(Ts).astype(np.float32) if Ts else np.zeros((0, 2), np.float32)
    s.indices = np.concatenate(Is).astype(np.uint32) if Is else np.zeros(0, np.uint32)
    # animations, gltf.cpp:331-405
    ends, ca, cn, cp, cf, cc, st, sv = [], [], [], [], [], [], [], []
    paths = {"scale": Scene.PATH_SCALE, "rotation": Scene.PATH_ROTATION, "translation": Scene.PATH_TRANSLATION, "weights": Scene.PATH_WEIGHTS}
    for a, an in enumerate(j.get("animations", [])):
        end = f32(0)
        for ch in an["channels"]:
            smp = an["samplers"][ch["sampler"]]
            path = paths.get(ch["target"]["path"], 3)
            times = floats(accessor(smp["input"]), 1)[:, 0]
            vals = np.zeros((len(times), 4), np.float32)
            vals[:, 3] = 1
            width = {Scene.PATH_SCALE: 3, Scene.PATH_ROTATION: 4, Scene.PATH_TRANSLATION: 3}.get(path, 0)
            if width:
                vals[:, :width] = floats(accessor(smp["output"]), width)[:len(times)]
            order = np.argsort(times, kind="stable")
            times, vals = times[order], vals[order]
            ca.append(a); cn.append(ch["target"]["node"]); cp.append(path); cf.append(len(st)); cc.append(len(times))
            st += list(times); sv += list(vals)
            end = max(end, f32(times[-1]))
        ends.append(end)
    if ends:
        s.anim_end_time = np.array(ends, np.float32)
        s.chan_anim, s.chan_node, s.chan_path = (np.array(a, np.int32) for a in (ca, cn, cp))
        s.chan_first_step, s.chan_n_steps = np.array(cf, np.uint32), np.array(cc, np.uint32)
        s.step_time, s.step_value = np.array(st, np.float32), np.array(sv, np.float32).reshape(-1, 4)
    return s


def decode_image(data):
    """PNG/JPEG bytes -> (h, w) uint32 texels with bytes b,g,r,a (alpha 255 when absent): the layout and the values the
    reference's libpng / libjpeg readers produce (src/misc/image.cpp:93-258), decoded by the package's own C++ decoder
    (swegl_b200_decode_image, host/image_decode.cpp) -- no imaging library involved."""
    import ctypes as C
    if os.environ.get("SWEGL_B200_IMAGE_DECODER") == "pil":
        # bench.py --impl reference: the reference arm must not load the product library at all; the scene packs' texel
        # digests (texel_digest) make sure both decoders give the same texels
        return decode_image_bgra(data)
    lib = _abi.load()
    ptr, w, h = C.POINTER(C.c_uint32)(), C.c_int32(), C.c_int32()
    data = bytes(data)
    rc = lib.swegl_b200_decode_image(data, len(data), C.byref(ptr), C.byref(w), C.byref(h))
    if rc != _abi.OK:
        raise ValueError(f"swegl_b200_decode_image: status {rc}: {lib.swegl_b200_image_error().decode()}")
    try:
        return np.ctypeslib.as_array(ptr, shape=(h.value, w.value)).copy()
    finally:
        lib.swegl_b200_image_free(ptr)


def decode_image_bgra(data):
    """The same through PIL (libpng / libjpeg-turbo): used where the scene packs and the image fixtures are MADE
    (tools/), and as the cross-check of decode_image in tests/test_image_decode.py."""
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    has_alpha = im.mode in ("RGBA", "LA") or "transparency" in im.info
    rgba = np.asarray(im.convert("RGBA"), dtype=np.uint8)
    out = np.empty(rgba.shape, np.uint8)
    out[..., 0], out[..., 1], out[..., 2] = rgba[..., 2], rgba[..., 1], rgba[..., 0]
    out[..., 3] = rgba[..., 3] if has_alpha else 255
    return np.ascontiguousarray(out).view(np.uint32).reshape(rgba.shape[0], rgba.shape[1])


def texel_digest(a):
    """sha256 of the texel bytes: guards scene packs against image-decoder drift."""
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint32).tobytes()).hexdigest()


def lcg_texture(size=1024, seed=12345):
    """The seeded synthetic texture of SURVEY §8c: s = s*1664525 + 1013904223 (mod 2^32);
    texel = 0xFF000000 | (s >> 8)."""
    s, a, c, m32 = seed, 1664525, 1013904223, (1 << 32) - 1
    vals = []
    for _ in range(size * size):
        s = (s * a + c) & m32
        vals.append(0xFF000000 | (s >> 8))
    return np.array(vals, dtype=np.uint64).astype(np.uint32).reshape(size, size)
