"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the two CPU checkers:

  Oracle : oracle/liboracle.so        (C restatement, swegl_oracle.c)
  Ref    : oracle/_ref/libswegl_ref.so (the unmodified reference behind ref_driver.cpp)

Imported only by tests/, tools/ (fixture generation), __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by swegl_b200/.
"""
import ctypes as C
import io
import os
import subprocess

import numpy as np

from swegl_b200 import _abi
from swegl_b200.scene import Scene, decode_image_bgra

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libswegl_ref.so")
# the same driver + unmodified reference sources, with swegl::_render supplied by swegl_b200/host/renderer_b200.cpp
DROPIN_LIB = os.path.join(HERE, "_ref", "libswegl_dropin.so")
# the reference with post_shader_depth_box repaired by oracle/dof_r.patch (the pin of the oracle's DoF-R)
REF_DOFR_LIB = os.path.join(HERE, "_ref", "libswegl_ref_dofr.so")


def build(ref=True):
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class OrcDump(C.Structure):
    _fields_ = [("v_world", C.c_void_p), ("v_viewport", C.c_void_p), ("normal_world", C.c_void_p), ("yes", C.c_void_p),
                ("n_fill_triangle", C.c_uint64), ("n_setup_triangles", C.c_uint64), ("n_spans", C.c_uint64),
                ("n_fragments", C.c_uint64), ("n_covered", C.c_uint64), ("n_texel_guard", C.c_uint64)]


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_LIB):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_LIB)
        self.lib.orc_render.restype = C.c_int
        self.lib.orc_render.argtypes = [C.POINTER(_abi.SceneDesc), C.POINTER(_abi.FrameDesc), C.POINTER(_abi.ViewportDesc),
                                        C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(OrcDump)]
        self.lib.orc_fnv1a64_words.restype = C.c_uint64
        self.lib.orc_fnv1a64_words.argtypes = [C.c_void_p, C.c_size_t]
        self.lib.orc_dof_r.restype = None
        self.lib.orc_dof_r.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]

    def fnv(self, words):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        return int(self.lib.orc_fnv1a64_words(a.ctypes.data, a.size))

    def render(self, scene, viewport, screen_wh=None, pixels=None, node_mats=None, want_vertices=False):
        """One frame of one viewport. Returns dict(pixels (H,W) u32, z (h,w) f32, counters, [vertex state])."""
        sw, sh = screen_wh or (viewport.x + viewport.w, viewport.y + viewport.h)
        if pixels is None:
            pixels = np.zeros((sh, sw), dtype=np.uint32)
        z = np.empty((viewport.h, viewport.w), dtype=np.float32)
        sd = scene.scene_desc()
        fd = scene.frame_desc(*(node_mats or (None, None)))
        vd = viewport.desc()
        dump = OrcDump()
        out = {}
        if want_vertices:
            nv = scene.n_vertices
            out["v_world"] = np.zeros((nv, 3), np.float32)
            out["v_viewport"] = np.zeros((nv, 3), np.float32)
            out["normal_world"] = np.zeros((nv, 3), np.float32)
            out["yes"] = np.zeros(nv, np.uint8)
            dump.v_world, dump.v_viewport = out["v_world"].ctypes.data, out["v_viewport"].ctypes.data
            dump.normal_world, dump.yes = out["normal_world"].ctypes.data, out["yes"].ctypes.data
        rc = self.lib.orc_render(C.byref(sd), C.byref(fd), C.byref(vd), pixels.ctypes.data, sw * 4, sw, sh,
                                 z.ctypes.data, C.byref(dump))
        if rc != 0:
            raise RuntimeError(f"orc_render failed: {rc}")
        out.update(pixels=pixels, z=z, n_fill_triangle=dump.n_fill_triangle, n_setup_triangles=dump.n_setup_triangles,
                   n_spans=dump.n_spans, n_fragments=dump.n_fragments, n_covered=dump.n_covered, n_texel_guard=dump.n_texel_guard)
        return out

    def dof_r(self, src, depth, focal_distance=5.0, focal_depth=5.0):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        dst = np.empty_like(src)
        self.lib.orc_dof_r(src.ctypes.data, depth.ctypes.data, dst.ctypes.data, src.shape[1], src.shape[0],
                           focal_distance, focal_depth)
        return dst


_DECODE_CB = C.CFUNCTYPE(C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint32))


class Ref:
    """The unmodified reference renderer (swegl::render) behind oracle/ref_driver.cpp."""

    def __init__(self, lib_path=REF_LIB):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(lib_path)
        L = self.lib = C.CDLL(lib_path)
        vp = C.c_void_p
        fp, ip, up = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint)
        sig = {
            "ref_set_image_decoder": (None, [_DECODE_CB]),
            "ref_scene_load": (vp, [C.c_char_p]),
            "ref_scene_new": (vp, []),
            "ref_scene_free": (None, [vp]),
            "ref_scene_add_material": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]),
            "ref_scene_add_texture": (C.c_int, [vp, vp, C.c_int, C.c_int]),
            "ref_scene_add_builtin": (C.c_int, [vp, C.c_int, C.c_uint, C.c_float, C.c_int, fp, fp, fp]),
            "ref_scene_import": (vp, [C.c_uint, vp, vp, vp, vp, C.c_uint, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "ref_scene_set_lights": (None, [vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint, vp]),
            "ref_scene_get_sun": (None, [vp, vp]),
            "ref_scene_counts": (None, [vp, vp]),
            "ref_scene_export": (None, [vp] + [vp] * 19),
            "ref_scene_export_texture": (None, [vp, C.c_int, vp]),
            "ref_scene_animate": (None, [vp, C.c_float]),
            "ref_scene_animation_counts": (None, [vp, vp]),
            "ref_scene_export_animations": (None, [vp] + [vp] * 8),
            "ref_scene_import_animations": (None, [vp, C.c_uint, vp, C.c_uint, vp, vp, vp, vp, vp, vp, vp]),
            "ref_scene_node_trs": (None, [vp, vp, vp, vp]),
            "ref_scene_node_matrices": (None, [vp, vp, vp]),
            "ref_scene_vertex_state": (None, [vp, vp, vp, vp, vp]),
            "ref_screen_new": (vp, [C.c_int, C.c_int]),
            "ref_screen_free": (None, [vp]),
            "ref_screen_pixels": (vp, [vp]),
            "ref_viewport_new": (vp, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
            "ref_viewport_free": (None, [vp]),
            "ref_viewport_set_dof": (None, [vp, C.c_float, C.c_float]),
            "ref_viewport_camera": (None, [vp, C.c_int, C.c_float, C.c_float, C.c_float]),
            "ref_viewport_get": (None, [vp, vp, vp, vp, vp]),
            "ref_viewport_zbuffer": (vp, [vp]),
            "ref_render": (None, [vp, vp]),
            "ref_render2": (None, [vp, vp, vp]),
            "ref_render4": (None, [vp, vp, vp, vp, vp]),
            "ref_time_render": (C.c_double, [vp, vp, C.c_int, C.c_int, vp]),
        }
        # only in libswegl_dropin.so: the C++ multi-context hosts and the device-side animation (ref_driver.cpp, SWEGL_B200_DROPIN)
        dropin_only = {
            "ref_render_pipelined": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_float]),
            "ref_render_sharded": (C.c_int, [vp, vp, C.c_int, C.c_int]),
            "ref_render_sharded4": (C.c_int, [vp, vp, vp, vp, vp, C.c_int]),
            "ref_render_animated": (C.c_int, [vp, vp, C.c_float]),
            "ref_dropin_last_error": (C.c_char_p, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        self.is_dropin = hasattr(L, "ref_render_pipelined")
        if self.is_dropin:
            for name, (res, args) in dropin_only.items():
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        del fp, ip, up
        self.encoded_images = []     # filled by the decode callback: the original bytes of each image
        self._cb = _DECODE_CB(self._decode)
        L.ref_set_image_decoder(self._cb)
        self._decoded = {}

    # called by the reference loader for every image (gltf.cpp:77-111)
    def _decode(self, filename, offset, w, h, out):
        key = (filename, offset)
        if key not in self._decoded:
            with open(filename, "rb") as f:
                f.seek(offset)
                data = f.read()
            t = decode_image_bgra(data)
            # keep only the image's own bytes (PIL stops at the end marker; re-encode length unknown,
            # so find it by re-parsing: PNG ends with IEND chunk, JPEG with FFD9)
            self._decoded[key] = (t, _trim_image_bytes(data))
        t, _ = self._decoded[key]
        w[0], h[0] = t.shape[1], t.shape[0]
        if out:
            C.memmove(out, t.ctypes.data, t.nbytes)
            self.encoded_images.append(self._decoded[key][1])
        return 0

    def host(self, name, *args):
        """call one of the C++-host entry points of libswegl_dropin.so; a C++ exception becomes a RuntimeError"""
        if getattr(self.lib, name)(*args) != 0:
            raise RuntimeError(f"{name}: {self.lib.ref_dropin_last_error().decode()}")

    # ---- scenes ----
    def load(self, path):
        self.encoded_images = []
        return self.lib.ref_scene_load(os.fsencode(path))

    def new_scene(self):
        return self.lib.ref_scene_new()

    def import_scene(self, s: Scene):
        a = lambda x, dt: np.ascontiguousarray(x, dtype=dt)
        arrs = [a(s.node_scale, np.float32), a(s.node_rotation, np.float32), a(s.node_translation, np.float32),
                a(s.node_parent, np.int32)]
        parrs = [a(s.prim_node, np.int32), a(s.prim_mode, np.int32), a(s.prim_material, np.int32),
                 a(s.prim_first_vertex, np.uint32), a(s.prim_n_vertices, np.uint32), a(s.prim_first_index, np.uint32),
                 a(s.prim_n_indices, np.uint32), a(s.positions, np.float32), a(s.normals, np.float32),
                 a(s.texcoords, np.float32), a(s.indices, np.uint32)]
        h = self.lib.ref_scene_import(s.n_nodes, *[x.ctypes.data for x in arrs], s.n_primitives,
                                      *[x.ctypes.data for x in parrs])
        for i in range(len(s.mat_bgra)):
            b, g, r, al = (int(v) for v in s.mat_bgra[i])
            self.lib.ref_scene_add_material(h, b, g, r, al, float(s.mat_metal_rough[i, 0]), float(s.mat_metal_rough[i, 1]),
                                            int(s.mat_tex_ds[i, 0]), int(s.mat_tex_ds[i, 1]))
        for t in s.textures:
            t = np.ascontiguousarray(t, dtype=np.uint32)
            self.lib.ref_scene_add_texture(h, t.ctypes.data, t.shape[1], t.shape[0])
        self.set_lights(h, s)
        if s.n_animations:
            a = [np.ascontiguousarray(s.anim_end_time, np.float32), np.ascontiguousarray(s.chan_anim, np.int32),
                 np.ascontiguousarray(s.chan_node, np.int32), np.ascontiguousarray(s.chan_path, np.int32),
                 np.ascontiguousarray(s.chan_first_step, np.uint32), np.ascontiguousarray(s.chan_n_steps, np.uint32),
                 np.ascontiguousarray(s.step_time, np.float32), np.ascontiguousarray(s.step_value, np.float32)]
            self.lib.ref_scene_import_animations(h, len(a[0]), a[0].ctypes.data, len(a[1]), *[x.ctypes.data for x in a[1:]])
        return h

    def set_lights(self, h, s: Scene):
        pl = np.ascontiguousarray(s.point_lights, dtype=np.float32)
        # the driver builds normal_t(x,y,z) from the raw direction exactly like src/test_1.cpp:335
        self.lib.ref_scene_set_lights(h, s.ambient, float(s.sun_raw[0]), float(s.sun_raw[1]), float(s.sun_raw[2]),
                                      s.sun_intensity, len(pl), pl.ctypes.data)
        sun = np.zeros(3, np.float32)
        self.lib.ref_scene_get_sun(h, sun.ctypes.data)
        return sun

    def export(self, h, name="scene") -> Scene:
        cnt = np.zeros(6, np.uint32)
        self.lib.ref_scene_counts(h, cnt.ctypes.data)
        nn, npr, nv, ni, nm, nt = (int(c) for c in cnt)
        s = Scene()
        s.name = name
        s.node_scale = np.zeros((nn, 3), np.float32)
        s.node_rotation = np.zeros((nn, 4, 4), np.float32)
        s.node_translation = np.zeros((nn, 3), np.float32)
        s.node_parent = np.zeros(nn, np.int32)
        s.prim_node, s.prim_mode, s.prim_material = (np.zeros(npr, np.int32) for _ in range(3))
        s.prim_first_vertex, s.prim_n_vertices, s.prim_first_index, s.prim_n_indices = (np.zeros(npr, np.uint32) for _ in range(4))
        s.positions, s.normals = np.zeros((nv, 3), np.float32), np.zeros((nv, 3), np.float32)
        s.texcoords = np.zeros((nv, 2), np.float32)
        s.indices = np.zeros(ni, np.uint32)
        s.mat_bgra = np.zeros((nm, 4), np.uint8)
        s.mat_metal_rough = np.zeros((nm, 2), np.float32)
        s.mat_tex_ds = np.zeros((nm, 2), np.int32)
        tex_wh = np.zeros((max(nt, 1), 2), np.int32)
        arrs = [s.node_scale, s.node_rotation, s.node_translation, s.node_parent, s.prim_node, s.prim_mode, s.prim_material,
                s.prim_first_vertex, s.prim_n_vertices, s.prim_first_index, s.prim_n_indices, s.positions, s.normals,
                s.texcoords, s.indices, s.mat_bgra, s.mat_metal_rough, s.mat_tex_ds, tex_wh]
        self.lib.ref_scene_export(h, *[x.ctypes.data for x in arrs])
        for t in range(nt):
            tex = np.zeros((int(tex_wh[t, 1]), int(tex_wh[t, 0])), np.uint32)
            self.lib.ref_scene_export_texture(h, t, tex.ctypes.data)
            s.textures.append(tex)
        ac = np.zeros(3, np.uint32)
        self.lib.ref_scene_animation_counts(h, ac.ctypes.data)
        na, nc, ns = (int(v) for v in ac)
        if na:
            s.anim_end_time = np.zeros(na, np.float32)
            s.chan_anim, s.chan_node, s.chan_path = (np.zeros(nc, np.int32) for _ in range(3))
            s.chan_first_step, s.chan_n_steps = (np.zeros(nc, np.uint32) for _ in range(2))
            s.step_time, s.step_value = np.zeros(ns, np.float32), np.zeros((ns, 4), np.float32)
            self.lib.ref_scene_export_animations(h, *[getattr(s, n).ctypes.data for n in Scene.ANIM_ARRAYS])
        return s

    def animate(self, h, seconds):
        """scene_t::animate(seconds) on the reference scene (model.hpp:146-177)"""
        self.lib.ref_scene_animate(h, float(seconds))

    def node_trs(self, h, n_nodes):
        sc, ro, tr = np.zeros((n_nodes, 3), np.float32), np.zeros((n_nodes, 4, 4), np.float32), np.zeros((n_nodes, 3), np.float32)
        self.lib.ref_scene_node_trs(h, sc.ctypes.data, ro.ctypes.data, tr.ctypes.data)
        return sc, ro, tr

    def procedural_scene(self) -> Scene:
        """test_1.cpp's build_scene() flavour from the reference's own mesh builders (src/data/model.cpp):
        torus (triangle strips), cube (fans, node scale.x = 2), textured sphere, three stacked triangles whose
        materials tests make transparent.  All materials opaque here; see configs.with_transparency."""
        from swegl_b200.scene import lcg_texture
        h = self.new_scene()
        tex = lcg_texture(64, seed=99)
        self.lib.ref_scene_add_texture(h, tex.ctypes.data, 64, 64)
        mats = [(128, 128, 128, 255, 0), (128, 128, 255, 255, -1), (255, 128, 255, 255, -1),
                (128, 128, 255, 255, -1), (128, 255, 128, 255, -1), (255, 128, 128, 255, -1)]
        for b, g, r, a, t in mats:
            self.lib.ref_scene_add_material(h, b, g, r, a, 1.0, 1.0, t, 0)
        f3 = lambda *v: (C.c_float * 3)(*v)
        self.lib.ref_scene_add_builtin(h, 2, 24, 1.0, 0, None, f3(0, 0, 0.5), f3(0, 0, -2.5))      # tore
        self.lib.ref_scene_add_builtin(h, 1, 0, 1.0, 0, f3(2, 1, 1), None, f3(0, 0, 0))             # cube, scale.x = 2
        self.lib.ref_scene_add_builtin(h, 3, 16, 2.0, 2, None, None, f3(3, 0, -1))                  # sphere
        self.lib.ref_scene_add_builtin(h, 0, 0, 1.0, 3, None, None, f3(1, 0.5, 2.1))                # tri
        self.lib.ref_scene_add_builtin(h, 0, 0, 1.0, 4, None, None, f3(1, 0.5, 2.0))
        self.lib.ref_scene_add_builtin(h, 0, 0, 1.0, 5, None, None, f3(1, 0.5, 2.2))
        s = self.export(h, "procedural")
        self.lib.ref_scene_free(h)
        return s

    def node_matrices(self, h, n_nodes):
        w = np.zeros((n_nodes, 4, 4), np.float32)
        n = np.zeros((n_nodes, 3, 3), np.float32)
        self.lib.ref_scene_node_matrices(h, w.ctypes.data, n.ctypes.data)
        return w, n

    def vertex_state(self, h, n_vertices):
        vw, vv, nw = (np.zeros((n_vertices, 3), np.float32) for _ in range(3))
        yes = np.zeros(n_vertices, np.uint8)
        self.lib.ref_scene_vertex_state(h, vw.ctypes.data, vv.ctypes.data, nw.ctypes.data, yes.ctypes.data)
        return dict(v_world=vw, v_viewport=vv, normal_world=nw, yes=yes)

    # ---- viewports ----
    def make_viewport(self, screen, viewport, camera_ops, with_dof=False):
        """viewport: swegl_b200.Viewport (rectangle + shader modes); camera_ops replayed on the reference camera."""
        v = self.lib.ref_viewport_new(screen, viewport.x, viewport.y, viewport.w, viewport.h,
                                      viewport.light_mode, viewport.tex_mode, viewport.transparency_layers)
        opcode = {"translate": 0, "rotate_x": 1, "rotate_y": 2, "rotate_z": 3}
        for op in camera_ops:
            args = list(op[1:]) + [0.0] * (3 - len(op[1:]))
            self.lib.ref_viewport_camera(v, opcode[op[0]], *args)
        if with_dof and viewport.post_mode == 1:            # post_shader_depth_box (src/test_1.cpp:355)
            self.lib.ref_viewport_set_dof(v, float(viewport.focal_distance), float(viewport.focal_depth))
        return v

    def viewport_matrices(self, v):
        view, proj = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
        cam, vpm = np.zeros(3, np.float32), np.zeros(4, np.float32)
        self.lib.ref_viewport_get(v, view.ctypes.data, proj.ctypes.data, cam.ctypes.data, vpm.ctypes.data)
        return view, proj, cam, vpm

    def render(self, h, v, screen, sw, sh, vw, vh):
        self.lib.ref_render(h, v)
        px = np.ctypeslib.as_array(C.cast(self.lib.ref_screen_pixels(screen), C.POINTER(C.c_uint32)), shape=(sh, sw)).copy()
        z = np.ctypeslib.as_array(C.cast(self.lib.ref_viewport_zbuffer(v), C.POINTER(C.c_float)), shape=(vh, vw)).copy()
        return px, z

    def time_render(self, h, v, warmup, frames):
        ms = np.zeros(frames, np.float64)
        self.lib.ref_time_render(h, v, warmup, frames, ms.ctypes.data)
        return ms


def _trim_image_bytes(data):
    """Cut `data` (image bytes followed by the rest of a .glb) at the end of the PNG/JPEG stream."""
    if data[:8] == b"\x89PNG\r\n\x1a\n":
        pos = 8
        while pos + 8 <= len(data):
            ln = int.from_bytes(data[pos:pos + 4], "big")
            typ = data[pos + 4:pos + 8]
            pos += 12 + ln
            if typ == b"IEND":
                return data[:pos]
        return data
    if data[:2] == b"\xff\xd8":
        # walk JPEG segments up to SOS, then scan entropy-coded data for EOI
        pos = 2
        while pos + 4 <= len(data):
            if data[pos] != 0xFF:
                break
            marker = data[pos + 1]
            if marker == 0xDA:
                end = data.find(b"\xff\xd9", pos)
                while end != -1:
                    # FFD9 cannot appear inside entropy data unescaped (FF is stuffed as FF00)
                    return data[:end + 2]
                break
            ln = int.from_bytes(data[pos + 2:pos + 4], "big")
            pos += 2 + ln
        return data
    return data
