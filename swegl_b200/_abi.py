"""ctypes mirror of include/swegl_b200.h and loader of the in-tree CUDA library.

The product path has no CPU fallback: `load()` raises if libswegl_b200.so is missing, and every
entry point of the library itself fails with SWEGL_B200_ERR_CUDA without an sm_100 device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWEGL_B200_LIB") or os.path.join(_HERE, "libswegl_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_CAPACITY, ERR_STATE = range(6)
MODE_TRIANGLES, MODE_TRIANGLE_STRIP, MODE_TRIANGLE_FAN = 4, 5, 6
LIGHT_NONE, LIGHT_FLAT, LIGHT_PHONG = 0, 1, 2
TEX_PLAIN, TEX_NEAREST, TEX_BILINEAR = 0, 1, 2
POST_NULL, POST_DOF = 0, 1
SHADING_EXACT, SHADING_FAST = 0, 1
ABI_VERSION = 4


class Primitive(C.Structure):
    _fields_ = [("node", C.c_int32), ("mode", C.c_int32), ("material_id", C.c_int32),
                ("first_vertex", C.c_uint32), ("n_vertices", C.c_uint32),
                ("first_index", C.c_uint32), ("n_indices", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("b", C.c_uint8), ("g", C.c_uint8), ("r", C.c_uint8), ("a", C.c_uint8),
                ("metallic", C.c_float), ("roughness", C.c_float),
                ("texture_idx", C.c_int32), ("double_sided", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("texels", C.POINTER(C.c_uint32)), ("width", C.c_int32), ("height", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_primitives", C.c_uint32), ("n_vertices", C.c_uint32),
                ("n_indices", C.c_uint32), ("n_materials", C.c_uint32), ("n_textures", C.c_uint32),
                ("primitives", C.POINTER(Primitive)),
                ("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
                ("texcoords", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_uint32)),
                ("materials", C.POINTER(Material)), ("default_material", Material),
                ("textures", C.POINTER(Texture))]


class FrameDesc(C.Structure):
    _fields_ = [("node_world", C.POINTER(C.c_float)), ("node_normal", C.POINTER(C.c_float)),
                ("ambient", C.c_float), ("sun_dir", C.c_float * 3), ("sun_intensity", C.c_float),
                ("n_point_lights", C.c_uint32), ("point_lights", C.POINTER(C.c_float))]


class AnimChannel(C.Structure):
    _fields_ = [("animation", C.c_int32), ("node", C.c_int32), ("path", C.c_int32), ("first_step", C.c_uint32), ("n_steps", C.c_uint32)]


class AnimationDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("node_parent", C.POINTER(C.c_int32)), ("node_rotation", C.POINTER(C.c_float)),
                ("node_translation", C.POINTER(C.c_float)), ("node_scale", C.POINTER(C.c_float)),
                ("n_animations", C.c_uint32), ("end_time", C.POINTER(C.c_float)),
                ("n_channels", C.c_uint32), ("channels", C.POINTER(AnimChannel)),
                ("n_steps", C.c_uint32), ("step_time", C.POINTER(C.c_float)), ("step_value", C.POINTER(C.c_float))]


class ViewportDesc(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("view", C.c_float * 16), ("proj", C.c_float * 16), ("cam_pos", C.c_float * 3),
                ("vp_m00", C.c_float), ("vp_m03", C.c_float), ("vp_m11", C.c_float), ("vp_m13", C.c_float),
                ("light_mode", C.c_int32), ("tex_mode", C.c_int32), ("post_mode", C.c_int32),
                ("focal_distance", C.c_float), ("focal_depth", C.c_float),
                ("transparency_layers", C.c_int32), ("band_y0", C.c_int32), ("band_y1", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_setup_triangles", C.c_uint32), ("n_spans", C.c_uint32), ("n_chunks", C.c_uint32),
                ("n_covered", C.c_uint32), ("n_launches", C.c_uint32), ("pool_grows", C.c_uint32),
                ("ms_vertex", C.c_float), ("ms_setup", C.c_float), ("ms_raster", C.c_float),
                ("ms_fragment", C.c_float), ("ms_post", C.c_float), ("ms_total", C.c_float), ("n_busy_tiles", C.c_uint32)]


# every symbol include/swegl_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("swegl_b200_abi_version", C.c_int, []),
    ("swegl_b200_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("swegl_b200_destroy", None, [C.c_void_p]),
    ("swegl_b200_last_error", C.c_char_p, [C.c_void_p]),
    ("swegl_b200_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swegl_b200_synchronize", C.c_int, [C.c_void_p]),
    ("swegl_b200_alloc_host", C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    ("swegl_b200_free_host", C.c_int, [C.c_void_p]),
    ("swegl_b200_set_timing", C.c_int, [C.c_void_p, C.c_int]),
    ("swegl_b200_upload_scene", C.c_int, [C.c_void_p, C.POINTER(SceneDesc)]),
    ("swegl_b200_set_screen", C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    ("swegl_b200_begin_frame", C.c_int, [C.c_void_p, C.POINTER(FrameDesc)]),
    ("swegl_b200_set_animation", C.c_int, [C.c_void_p, C.POINTER(AnimationDesc)]),
    ("swegl_b200_begin_frame_animated", C.c_int, [C.c_void_p, C.c_float, C.POINTER(FrameDesc)]),
    ("swegl_b200_read_node_matrices", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swegl_b200_render_viewport_device", C.c_int, [C.c_void_p, C.POINTER(ViewportDesc), C.POINTER(Stats)]),
    ("swegl_b200_render_viewport", C.c_int, [C.c_void_p, C.POINTER(ViewportDesc), C.c_void_p, C.c_int32,
                                             C.c_void_p, C.POINTER(Stats)]),
    ("swegl_b200_render_viewport_async", C.c_int, [C.c_void_p, C.POINTER(ViewportDesc), C.c_void_p, C.c_int32,
                                                   C.c_void_p, C.POINTER(C.c_uint64)]),
    ("swegl_b200_wait", C.c_int, [C.c_void_p, C.c_uint64]),
    ("swegl_b200_set_partial_readback", C.c_int, [C.c_void_p, C.c_int]),
    ("swegl_b200_invalidate_host_image", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swegl_b200_readback_stats", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("swegl_b200_set_shading", C.c_int, [C.c_void_p, C.c_int]),
    ("swegl_b200_set_shared_gpu", C.c_int, [C.c_void_p, C.c_int]),
    ("swegl_b200_frame_hash", C.c_uint64, [C.c_void_p, C.c_size_t]),
    ("swegl_b200_export_screen", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swegl_b200_import_screen", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    ("swegl_b200_set_color_target", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swegl_b200_device_buffers", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    ("swegl_b200_read_screen", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    ("swegl_b200_read_depth", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swegl_b200_read_rect", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    ("swegl_b200_read_depth_rows", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    ("swegl_b200_enable_peer", C.c_int, [C.c_void_p, C.c_int]),
    ("swegl_b200_device_of", C.c_int, [C.c_void_p]),
    ("swegl_b200_scene_opaque", C.c_int, [C.c_void_p]),
    ("swegl_b200_read_vertices", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swegl_b200_set_frame_sync", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("swegl_b200_frame_sync_status", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swegl_b200_set_band_culling", C.c_int, [C.c_void_p, C.c_int]),
    ("swegl_b200_cull_counts", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swegl_b200_selftest_division", C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    ("swegl_b200_selftest_filter", C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    ("swegl_b200_decode_image", C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("swegl_b200_image_free", None, [C.POINTER(C.c_uint32)]),
    ("swegl_b200_image_error", C.c_char_p, []),
]

_lib = None


def load():
    """Load libswegl_b200.so and bind every declared symbol. Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(swegl_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the header and the library drift apart
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
