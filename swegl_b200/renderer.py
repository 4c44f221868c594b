"""Python front end of the C ABI (include/swegl_b200.h), mirroring the reference's call sequence

    swegl::render(scene, viewport...)          swegl/render/renderer.hpp:27-34
      -> original_to_world(scene)              Renderer.begin_frame(scene)
      -> _render(scene, viewport) each         Renderer.render(viewport, pixels)

There is no CPU path behind these calls: they fail loudly when the CUDA library or device is missing.
"""
import ctypes as C

import numpy as np

from . import _abi


class SweglB200Error(RuntimeError):
    def __init__(self, msg, status=None):
        super().__init__(msg)
        self.status = status


def frame_hash(words):
    """FNV-1a-64 over the 32-bit words of a frame (tests/golden/MANIFEST.json's fingerprint), computed by the library"""
    a = np.ascontiguousarray(words).view(np.uint32)
    return int(_abi.load().swegl_b200_frame_hash(a.ctypes.data, a.size))


class Renderer:
    def __init__(self, device=0, stream=None):
        self.lib = _abi.load()
        if self.lib.swegl_b200_abi_version() != _abi.ABI_VERSION:
            raise SweglB200Error("ABI version mismatch between swegl_b200/_abi.py and libswegl_b200.so")
        self.ctx = C.c_void_p()
        rc = self.lib.swegl_b200_create(int(device), C.byref(self.ctx))
        if rc != _abi.OK:
            raise SweglB200Error(f"swegl_b200_create(device={device}) failed with status {rc} "
                                 "(no CUDA device? swegl_b200 has no CPU fallback)")
        if stream is not None:
            self._check(self.lib.swegl_b200_set_stream(self.ctx, C.c_void_p(int(stream))))
        self.scene = None
        self.screen_wh = None

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.swegl_b200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _abi.OK:
            msg = self.lib.swegl_b200_last_error(self.ctx)
            raise SweglB200Error(f"swegl_b200 status {rc}: {msg.decode() if msg else ''}", rc)

    def synchronize(self):
        self._check(self.lib.swegl_b200_synchronize(self.ctx))

    def alloc_host(self, shape, dtype):
        """numpy array backed by page-locked memory (for `pixels` / `zbuffer` of render())."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        if self.lib.swegl_b200_alloc_host(n, C.byref(p)) != _abi.OK:
            raise SweglB200Error("swegl_b200_alloc_host failed")
        buf = (C.c_uint8 * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        self._pinned = getattr(self, "_pinned", []) + [p]
        return arr

    def set_timing(self, enabled=True):
        self._check(self.lib.swegl_b200_set_timing(self.ctx, int(enabled)))

    def upload_scene(self, scene):
        self.scene = scene
        sd = scene.scene_desc()
        self._check(self.lib.swegl_b200_upload_scene(self.ctx, C.byref(sd)))

    def set_screen(self, w, h):
        self.screen_wh = (int(w), int(h))
        self._check(self.lib.swegl_b200_set_screen(self.ctx, int(w), int(h)))

    def begin_frame(self, scene=None, node_mats=None):
        scene = scene or self.scene
        fd = scene.frame_desc(*(node_mats or (None, None)))
        self._check(self.lib.swegl_b200_begin_frame(self.ctx, C.byref(fd)))

    def set_animation(self, scene=None):
        """upload the scene's key frames, base TRS and hierarchy (swegl_b200_set_animation); call after upload_scene and
        BEFORE the first Scene.animate() on `scene` (the base TRS is the scene as loaded)"""
        scene = scene or self.scene
        ad = scene.animation_desc()
        self._check(self.lib.swegl_b200_set_animation(self.ctx, C.byref(ad)))

    def begin_frame_animated(self, elapsed_seconds, scene=None):
        """scene_t::animate(t) + the node-hierarchy product on the device: only the time stamp and the lights travel"""
        scene = scene or self.scene
        fd = scene.frame_desc(lights_only=True)
        self._check(self.lib.swegl_b200_begin_frame_animated(self.ctx, C.c_float(elapsed_seconds), C.byref(fd)))

    def read_node_matrices(self):
        """-> (node_world (n,4,4), node_normal (n,3,3)) as the device holds them after the last rendered frame"""
        n = self.scene.n_nodes
        w, nn = np.zeros((n, 4, 4), np.float32), np.zeros((n, 3, 3), np.float32)
        self._check(self.lib.swegl_b200_read_node_matrices(self.ctx, w.ctypes.data, nn.ctypes.data))
        return w, nn

    def render_device(self, viewport, stats=True):
        """_render() with the result left in HBM."""
        vd = viewport.desc() if not isinstance(viewport, _abi.ViewportDesc) else viewport
        st = _abi.Stats()
        self._check(self.lib.swegl_b200_render_viewport_device(self.ctx, C.byref(vd), C.byref(st) if stats else None))
        return st

    def render(self, viewport, pixels, zbuffer=None, stats=True):
        """_render() into host memory: `pixels` is the (screen_h, screen_w) uint32 surface."""
        vd = viewport.desc() if not isinstance(viewport, _abi.ViewportDesc) else viewport
        st = _abi.Stats()
        zptr = zbuffer.ctypes.data if zbuffer is not None else None
        self._check(self.lib.swegl_b200_render_viewport(self.ctx, C.byref(vd), pixels.ctypes.data, pixels.strides[0],
                                                        zptr, C.byref(st) if stats else None))
        return st

    def render_async(self, viewport, pixels, zbuffer=None):
        """Pipelined render(): queues the frame and its copy into `pixels` (page-locked: alloc_host) and returns a
        ticket; the image is complete after wait(ticket).  Up to three frames may be in flight: rotate over two or three
        host images and wait for a frame before its image is used again."""
        vd = viewport.desc() if not isinstance(viewport, _abi.ViewportDesc) else viewport
        zptr = zbuffer.ctypes.data if zbuffer is not None else None
        ticket = C.c_uint64(0)
        self._check(self.lib.swegl_b200_render_viewport_async(self.ctx, C.byref(vd), pixels.ctypes.data, pixels.strides[0],
                                                              zptr, C.byref(ticket)))
        return ticket.value

    def wait(self, ticket):
        """Blocks until the frame of `ticket` is in host memory.  Raises SweglB200Error(ERR_CAPACITY) if that frame must
        be submitted again (its pools were too small and have been enlarged)."""
        self._check(self.lib.swegl_b200_wait(self.ctx, C.c_uint64(ticket)))

    def set_shading(self, mode):
        """_abi.SHADING_EXACT (bit-exact Phong lighting) or _abi.SHADING_FAST (within +-1 LSB per colour channel, default)"""
        self._check(self.lib.swegl_b200_set_shading(self.ctx, int(mode)))

    def set_shared_gpu(self, shared=True):
        """other contexts render on this GPU at the same time (FramePipeline): prefer kernels that hold fewer SM resources"""
        self._check(self.lib.swegl_b200_set_shared_gpu(self.ctx, int(bool(shared))))

    def set_partial_readback(self, enabled=True):
        self._check(self.lib.swegl_b200_set_partial_readback(self.ctx, int(bool(enabled))))

    def invalidate_host_image(self, pixels=None):
        """tell the library that `pixels` (None: every host image) was modified by the caller since the last frame through it"""
        self._check(self.lib.swegl_b200_invalidate_host_image(self.ctx, C.c_void_p(pixels.ctypes.data) if pixels is not None else None))

    def readback_stats(self, reset=False):
        """-> (bytes copied device->host by render_async so far, frames)"""
        out = (C.c_uint64 * 2)()
        self._check(self.lib.swegl_b200_readback_stats(self.ctx, out, int(bool(reset))))
        return int(out[0]), int(out[1])

    def export_screen(self):
        """64-byte CUDA IPC handle of this context's device screen (ship it to the other ranks)"""
        buf = (C.c_uint8 * 64)()
        self._check(self.lib.swegl_b200_export_screen(self.ctx, buf))
        return bytes(buf)

    def import_screen(self, handle):
        """map another process's exported screen; returns its device pointer in this process"""
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self.lib.swegl_b200_import_screen(self.ctx, buf, C.byref(p)))
        return p.value

    def set_color_target(self, device_ptr):
        """finished colour goes to `device_ptr` (another context's / GPU's screen) instead of the own screen; None resets"""
        self._check(self.lib.swegl_b200_set_color_target(self.ctx, C.c_void_p(device_ptr) if device_ptr else None))

    def read_screen(self, y0=0, y1=None):
        w, h = self.screen_wh
        y1 = h if y1 is None else y1
        out = np.empty((y1 - y0, w), dtype=np.uint32)
        self._check(self.lib.swegl_b200_read_screen(self.ctx, y0, y1, out.ctypes.data, w * 4))
        return out

    def read_depth(self, w, h):
        out = np.empty((h, w), dtype=np.float32)
        self._check(self.lib.swegl_b200_read_depth(self.ctx, out.ctypes.data))
        return out

    def read_vertices(self):
        nv = self.scene.n_vertices
        vw, vv, nw = (np.zeros((nv, 3), np.float32) for _ in range(3))
        yes = np.zeros(nv, np.uint8)
        self._check(self.lib.swegl_b200_read_vertices(self.ctx, vw.ctypes.data, vv.ctypes.data, nw.ctypes.data, yes.ctypes.data))
        return dict(v_world=vw, v_viewport=vv, normal_world=nw, yes=yes)

    def set_frame_sync(self, rank, world=0):
        """frame protocol of the band-sharded single frame over peer memory (include/swegl_b200.h); rank < 0: off"""
        self._check(self.lib.swegl_b200_set_frame_sync(self.ctx, int(rank), int(world)))

    def frame_sync_status(self):
        """-> (timed-out waits of this context, rank 0's own share of the last frame in ms)"""
        n, ms = C.c_uint32(0), C.c_float(0)
        self._check(self.lib.swegl_b200_frame_sync_status(self.ctx, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def frame_sync_errors(self):
        return self.frame_sync_status()[0]

    def set_band_culling(self, policy):
        """-1 automatic, 0 off, 1 on for every banded view; call before upload_scene (include/swegl_b200.h)"""
        self._check(self.lib.swegl_b200_set_band_culling(self.ctx, int(policy)))

    def cull_counts(self):
        c = (C.c_uint32 * 6)()
        self._check(self.lib.swegl_b200_cull_counts(self.ctx, c))
        return dict(clusters=c[0], vertex_blocks=c[1], live=c[2], marked=c[3], vertex_blocks_needed=c[4], culled=bool(c[5]))

    def selftest_division(self, n_pairs, seed=1):
        """-> (quotients that differ from __fdiv_rn, pairs that took the shared-reciprocal fast path)"""
        out = (C.c_uint64 * 2)()
        self._check(self.lib.swegl_b200_selftest_division(self.ctx, C.c_uint64(int(n_pairs)), C.c_uint32(int(seed)), out))
        return int(out[0]), int(out[1])

    def selftest_filter(self, n_samples, seed=1):
        """-> samples on which the fast kernels' bilinear filter differs from the exact one (must be 0)"""
        out = (C.c_uint64 * 1)()
        self._check(self.lib.swegl_b200_selftest_filter(self.ctx, C.c_uint64(int(n_samples)), C.c_uint32(int(seed)), out))
        return int(out[0])

    def device_buffers(self):
        s, d = C.c_void_p(), C.c_void_p()
        self._check(self.lib.swegl_b200_device_buffers(self.ctx, C.byref(s), C.byref(d)))
        return s.value, d.value
