/*
 * swegl_b200.h — C ABI of the B200-native replacement for swegl's per-frame
 * rendering hot path.
 *
 * The reference (gbizzotto/swegl) has no FFI; its boundary is one C++ call,
 *     swegl::render(scene_t&, viewport_t&...)          swegl/render/renderer.hpp:27-34
 * which runs vertex_shader_t::original_to_world once (vertex_shaders.hpp:16-33)
 * and then swegl::_render(scene, viewport) per viewport (src/render/renderer.cpp:77-235).
 * This header is what a maintainer binds from the body of those two functions
 * (see INTEGRATION.md and swegl_b200/host/swegl_b200_adapter.hpp): plain
 * pointers and sizes only, `int` status returns (0 = ok), no C++ or torch
 * types.  There is no CPU fallback: every entry point fails with
 * SWEGL_B200_ERR_CUDA when no sm_100 device is usable.
 *
 * Data layout conventions (all little-endian, tightly packed):
 *   - matrices are row-major float[16] exactly like swegl::matrix44_t
 *     (freon::Matrix<float,4,4>, m[row][col]),
 *   - colours are 32-bit words with bytes b,g,r,a (swegl/render/colors.hpp:9-18),
 *   - vertex attributes are SoA copies of mesh_vertex_t::{v, normal, tex_coords}
 *     (swegl/data/model.hpp:20-29) flattened over nodes -> primitives in
 *     scene.nodes order, which is also the draw order of renderer.cpp:83-231.
 */
#ifndef SWEGL_B200_H
#define SWEGL_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWEGL_B200_ABI_VERSION 4

/* ---- status codes (reference has none: void + assert, renderer.cpp:98-100) ---- */
enum {
    SWEGL_B200_OK              = 0,
    SWEGL_B200_ERR_ARG         = 1,  /* null pointer / bad size / bad enum               */
    SWEGL_B200_ERR_CUDA        = 2,  /* CUDA runtime error, see swegl_b200_last_error()   */
    SWEGL_B200_ERR_UNSUPPORTED = 3,  /* shader / transparency combination not on device   */
    SWEGL_B200_ERR_CAPACITY    = 4,  /* internal span/chunk pool exhausted even after grow*/
    SWEGL_B200_ERR_STATE       = 5   /* call order (render before upload, ...)            */
};

/* primitive_t::index_mode_t, swegl/data/model.hpp:33-42 (glTF mode codes) */
enum {
    SWEGL_B200_MODE_TRIANGLES      = 4,
    SWEGL_B200_MODE_TRIANGLE_STRIP = 5,
    SWEGL_B200_MODE_TRIANGLE_FAN   = 6
};

/* which pixel_shader_t the viewport carries (swegl/render/pixel_shaders.hpp:15-179).
 * COMBINED = pixel_shader_light_and_texture<L,T> (:121-179): texture colour scaled by light. */
enum { SWEGL_B200_LIGHT_NONE = 0, SWEGL_B200_LIGHT_FLAT = 1, SWEGL_B200_LIGHT_PHONG = 2 };
enum { SWEGL_B200_TEX_PLAIN = 0, SWEGL_B200_TEX_NEAREST = 1, SWEGL_B200_TEX_BILINEAR = 2 };
/* Phong lighting arithmetic (swegl_b200_set_shading).  EXACT reproduces the reference's unfused fp32 / fp64 evaluation
 * bit for bit; FAST (the default) stays within +-1 LSB per 8-bit colour channel of it (alpha, coverage and depth are
 * always exact) at about a third of the instructions.  Flat lighting and the texture filters are always exact. */
enum { SWEGL_B200_SHADING_EXACT = 0, SWEGL_B200_SHADING_FAST = 1 };
/* post_shader_t (null / copy, post_shaders.hpp:15-48) or post_shader_depth_box (:51-132, as
 * repaired: "DoF-R") */
enum { SWEGL_B200_POST_NULL = 0, SWEGL_B200_POST_DOF = 1 };

typedef struct swegl_b200_primitive {
    int32_t  node;          /* index of the owning node_t in scene.nodes                 */
    int32_t  mode;          /* SWEGL_B200_MODE_*                                         */
    int32_t  material_id;   /* primitive_t::material_id, -1 = scene.default_material     */
    uint32_t first_vertex;  /* offset into the SoA vertex arrays                         */
    uint32_t n_vertices;    /* vertices of the primitive as the caller counts them; indices must stay below it.
                               (swegl's loader appends 2 spare slots for the near clipper, gltf.cpp:175: the
                               adapter uploads them along -- harmless, nothing indexes them; the device clipper
                               keeps its vertices in registers) */
    uint32_t first_index;   /* offset into indices[]                                     */
    uint32_t n_indices;     /* indices are relative to first_vertex, as in primitive_t   */
} swegl_b200_primitive;

/* material_t, swegl/data/model.hpp:80-87 */
typedef struct swegl_b200_material {
    uint8_t  b, g, r, a;    /* material_t::color                                         */
    float    metallic, roughness;
    int32_t  texture_idx;   /* -1 = none: a 1x1 bitmap of `color` is sampled instead     */
    int32_t  double_sided;
} swegl_b200_material;

/* texture_t level 0 only (pixel_shaders.cpp:298-300), row-major BGRA words */
typedef struct swegl_b200_texture {
    const uint32_t *texels;
    int32_t width, height;
} swegl_b200_texture;

/* everything that is static after load_scene() */
typedef struct swegl_b200_scene_desc {
    uint32_t n_nodes, n_primitives, n_vertices, n_indices, n_materials, n_textures;
    const swegl_b200_primitive *primitives;   /* in draw order                             */
    const float    *positions;                /* 3*n_vertices, mesh_vertex_t::v            */
    const float    *normals;                  /* 3*n_vertices, mesh_vertex_t::normal       */
    const float    *texcoords;                /* 2*n_vertices, tex_coords.x(), .y()        */
    const uint32_t *indices;                  /* n_indices                                  */
    const swegl_b200_material *materials;     /* n_materials                                */
    swegl_b200_material default_material;     /* scene_t::default_material                  */
    const swegl_b200_texture *textures;       /* n_textures, scene_t::images                */
} swegl_b200_scene_desc;

/* what may change every frame: node transforms and lights (test_1.cpp:288-303,378).
 * The node-hierarchy product (freon::operator*, vertex_shaders.hpp:18) stays on the
 * host; the device only does matrix x vertex. */
typedef struct swegl_b200_frame_desc {
    const float *node_world;   /* n_nodes*16: node_t::original_to_world_matrix            */
    const float *node_normal;  /* n_nodes*9 : upper 3x3 of scale(node.rotation,node.scale) */
    float    ambient;          /* scene_t::ambient_light_intensity                         */
    float    sun_dir[3];       /* scene_t::sun_direction (already normalised)              */
    float    sun_intensity;
    uint32_t n_point_lights;
    const float *point_lights; /* n*4: position xyz, intensity (point_source_light)        */
} swegl_b200_frame_desc;

/* ---- device-side animation: scene_t::animate, swegl/data/model.hpp:146-177, and the node-hierarchy product of
 * vertex_shader_t::original_to_world, swegl/render/vertex_shaders.hpp:16-33, evaluated on the device ----
 * animation_channel_t (model.hpp:94-122) flattened: key frames [first_step, first_step + n_steps) of step_time / step_value */
typedef struct swegl_b200_anim_channel {
    int32_t  animation;        /* index of the owning animation_t (its end_time wraps the clock, model.hpp:150)   */
    int32_t  node;             /* animation_channel_t::node_idx                                                     */
    int32_t  path;             /* animation_channel_t::path_t: 0 scale, 1 rotation, 2 translation, 3 weights/none   */
    uint32_t first_step, n_steps;   /* n_steps >= 1                                                                 */
} swegl_b200_anim_channel;
typedef struct swegl_b200_animation_desc {
    uint32_t n_nodes;               /* must equal the uploaded scene's                                               */
    const int32_t *node_parent;     /* n_nodes: the node whose children_idx holds this node, -1 for scene_t::root_nodes */
    const float *node_rotation;     /* n_nodes*16: node_t::rotation (matrix44_t) as loaded                           */
    const float *node_translation;  /* n_nodes*3 : node_t::translation                                               */
    const float *node_scale;        /* n_nodes*3 : node_t::scale                                                     */
    uint32_t n_animations;
    const float *end_time;          /* n_animations: animation_t::end_time                                           */
    uint32_t n_channels;
    const swegl_b200_anim_channel *channels;   /* in scene order: animation by animation, channel by channel         */
    uint32_t n_steps;
    const float *step_time;         /* n_steps  : animation_step_t::time                                             */
    const float *step_value;        /* n_steps*4: animation_step_t::value (x, y, z, w)                               */
} swegl_b200_animation_desc;

/* viewport_t + camera_t + shader selection, swegl/render/viewport.hpp:27-62 */
typedef struct swegl_b200_viewport_desc {
    int32_t x, y, w, h;            /* m_x, m_y, m_w, m_h                                   */
    float   view[16];              /* camera_t::m_viewmatrix                               */
    float   proj[16];              /* camera_t::m_projectionmatrix                         */
    float   cam_pos[3];            /* camera_t::m_center                                   */
    float   vp_m00, vp_m03;        /* m_viewportmatrix[0][0], [0][3]  (viewport.cpp:30-35) */
    float   vp_m11, vp_m13;        /* m_viewportmatrix[1][1], [1][3]                       */
    int32_t light_mode;            /* SWEGL_B200_LIGHT_*                                   */
    int32_t tex_mode;              /* SWEGL_B200_TEX_*                                     */
    int32_t post_mode;             /* SWEGL_B200_POST_*                                    */
    float   focal_distance;        /* post_shader_depth_box::focal_distance                */
    float   focal_depth;           /* post_shader_depth_box::focal_depth                   */
    int32_t transparency_layers;   /* viewport_t ctor argument (viewport.hpp:19-25): up to 8 on the device, and only on a
                                      viewport at the screen origin; ignored when every material and
                                      texel has alpha 255 (the layer logic is then the identity)                     */
    int32_t band_y0, band_y1;      /* sort-first scissor, viewport-relative rows
                                      [band_y0, band_y1); (0,0) = whole viewport           */
} swegl_b200_viewport_desc;

/* per-frame counters, filled by the render calls when non-null */
typedef struct swegl_b200_stats {
    uint32_t n_setup_triangles;    /* triangles that reached fill_triangle_2               */
    uint32_t n_spans;              /* scanline spans walked                                */
    uint32_t n_chunks;             /* <=32-pixel span pieces binned                        */
    uint32_t n_covered;            /* pixels whose depth != 0x7F7F7F7F after the frame     */
    uint32_t n_launches;           /* kernels launched by the call                         */
    uint32_t pool_grows;           /* times the span/chunk pools had to be enlarged        */
    float    ms_vertex, ms_setup, ms_raster, ms_fragment, ms_post, ms_total;
                                   /* CUDA-event times, only when timing is enabled        */
    uint32_t n_busy_tiles;         /* screen tiles of 8 rows x 128 pixels something was drawn into: with a colour target
                                      (multi-GPU output over peer memory) and the frame protocol, 4 KiB of finished
                                      colour per busy tile is what crosses NVLink                                      */
} swegl_b200_stats;

typedef struct swegl_b200_ctx swegl_b200_ctx;

/* ---- lifetime ---- */
int  swegl_b200_abi_version(void);
int  swegl_b200_create(int device, swegl_b200_ctx **out);
void swegl_b200_destroy(swegl_b200_ctx *ctx);
const char *swegl_b200_last_error(const swegl_b200_ctx *ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the ctx's own */
int  swegl_b200_set_stream(swegl_b200_ctx *ctx, void *cuda_stream);
/* block until everything queued on the context's stream has finished.  Returns SWEGL_B200_ERR_CAPACITY when
 * the last render_viewport_device() frame ran out of internal pool space (the pools are enlarged by this call;
 * re-issue the frame).  render_viewport() and calls with stats != NULL handle that case themselves. */
int  swegl_b200_synchronize(swegl_b200_ctx *ctx);
/* page-locked host memory for `pixels` / `zbuffer` (what SDL_Surface::pixels should live in for
 * full-speed read-back); plain malloc'ed memory works too, just slower */
int  swegl_b200_alloc_host(size_t bytes, void **out);
int  swegl_b200_free_host(void *p);
/* enable CUDA-event stage timing into swegl_b200_stats (adds synchronisation) */
int  swegl_b200_set_timing(swegl_b200_ctx *ctx, int enabled);

/* ---- scene ---- */
/* replaces nothing in the reference (its scene lives in host vectors); uploads the static part */
int  swegl_b200_upload_scene(swegl_b200_ctx *ctx, const swegl_b200_scene_desc *scene);
/* the screen (SDL_Surface w,h) the viewports draw into; device copy is cleared to 0 */
int  swegl_b200_set_screen(swegl_b200_ctx *ctx, int32_t screen_w, int32_t screen_h);

/* ---- frame ---- */
/* replaces vertex_shader_t::original_to_world's per-vertex loop (vertex_shaders.hpp:20-24):
 * uploads node matrices + lights and computes v_world for every vertex. Once per frame. */
int  swegl_b200_begin_frame(swegl_b200_ctx *ctx, const swegl_b200_frame_desc *frame);

/* Device-side animation.  set_animation uploads the key frames, the nodes' TRS as loaded and the hierarchy once (after
 * upload_scene; a new upload_scene drops them).  begin_frame_animated(t, frame) then replaces the application's
 *     scene.animate(t);  (test_1.cpp:378)   +   the node loop of original_to_world (vertex_shaders.hpp:16-33)
 * for that frame: `frame` supplies the lights only (node_world / node_normal are ignored and may be NULL), 4 bytes of time
 * stamp travel instead of 100 bytes per node, and the first kernel of the frame computes node_t::rotation / translation /
 * scale from the key frames and from them the world matrices, with the reference's fp32 operations in the reference's
 * order -- bit-identical to animate() + original_to_world on the host.  The state is a function of t alone (every call of
 * animate() rewrites the same node paths), so frames may be submitted in any order of t.  Plain begin_frame and
 * begin_frame_animated can be mixed freely.  read_node_matrices returns the matrices the device holds after the last
 * rendered frame (either pointer may be NULL). */
int  swegl_b200_set_animation(swegl_b200_ctx *ctx, const swegl_b200_animation_desc *animation);
int  swegl_b200_begin_frame_animated(swegl_b200_ctx *ctx, float elapsed_seconds, const swegl_b200_frame_desc *frame);
int  swegl_b200_read_node_matrices(swegl_b200_ctx *ctx, float *node_world, float *node_normal);

/* replaces swegl::_render(scene, viewport) (renderer.cpp:77-235), result left in HBM.  Asynchronous when
 * stats == NULL (see swegl_b200_synchronize for the pool-overflow contract); with stats it synchronises,
 * grows the pools if needed and redoes the frame. */
int  swegl_b200_render_viewport_device(swegl_b200_ctx *ctx, const swegl_b200_viewport_desc *vp,
                                       swegl_b200_stats *stats);
/* same + copies the viewport rectangle into host `pixels` (SDL_Surface::pixels, rows
 * `pitch_bytes` apart, absolute screen coordinates) and, if non-null, the viewport's
 * depth buffer into host `zbuffer` (w*h floats, viewport_t::m_zbuffer). Synchronous. */
int  swegl_b200_render_viewport(swegl_b200_ctx *ctx, const swegl_b200_viewport_desc *vp,
                                void *pixels, int32_t pitch_bytes, float *zbuffer,
                                swegl_b200_stats *stats);

/* Pipelined variant of render_viewport ("async readback / double-buffered frames"): queues the
 * frame, a device-side copy of the finished rectangle into one of three staging images (only the part that differs
 * from the frame the staging image already holds), and the copy of that image to host `pixels` / `zbuffer` on a
 * second stream -- then returns.  The next frames render while this one crosses PCIe.  `*ticket` identifies the
 * frame; its host data is complete after swegl_b200_wait(ticket).
 * At most three frames are in flight (a fourth submit waits, on the device, for the oldest one's copy), so callers
 * rotate over two or three host images and wait for a frame before reusing its image; the images should be
 * page-locked (swegl_b200_alloc_host), otherwise the copy blocks the submit.
 * swegl_b200_wait returns SWEGL_B200_ERR_CAPACITY when that frame ran out of pool space (the pools are then
 * enlarged: submit it again), like swegl_b200_synchronize.
 *
 * Overflow with several frames in flight: by the time wait(ticket i) reports ERR_CAPACITY, frame i+1 has already replaced the
 * per-frame data on the device (node matrices, lights) and ran with the same undersized pools, so expect its ticket to
 * fail as well.  Resubmitting frame i therefore means swegl_b200_begin_frame with THAT frame's data again, then
 * render_viewport_async; render_viewport() (blocking) redoes an overflowed frame itself.
 *
 * Partial read-back (default on): outside the bounding box of what was drawn (grown by the blur radius with DoF-R) a frame
 * is one constant.  The library remembers, per host image (`pixels` pointer), what it last left there, and copies only
 * the union of the previous and the current box -- the background of a 4K frame does not cross PCIe every frame.  The
 * image must therefore not be modified by the caller between two frames that go through it; after drawing into it
 * (an overlay), call swegl_b200_invalidate_host_image (pixels == NULL: all images), or switch the feature off.  The first
 * frame through an image, a change of viewport / pitch / band / post pass, and views with transparency layers are
 * copied in full.  render_viewport (blocking, the drop-in's path) always copies the whole rectangle.
 * Both host-read-back entry points fail with SWEGL_B200_ERR_STATE while a colour target is set (set_color_target): the
 * finished pixels are then in the target's screen, not in this context's. */
int  swegl_b200_render_viewport_async(swegl_b200_ctx *ctx, const swegl_b200_viewport_desc *vp,
                                      void *pixels, int32_t pitch_bytes, float *zbuffer, uint64_t *ticket);
int  swegl_b200_wait(swegl_b200_ctx *ctx, uint64_t ticket);
int  swegl_b200_set_partial_readback(swegl_b200_ctx *ctx, int enabled);
int  swegl_b200_invalidate_host_image(swegl_b200_ctx *ctx, const void *pixels);
/* out[0] = bytes copied device->host by render_viewport_async so far, out[1] = frames; reset != 0 zeroes both */
int  swegl_b200_readback_stats(swegl_b200_ctx *ctx, uint64_t out[2], int reset);

/* SWEGL_B200_SHADING_EXACT / _FAST for the frames submitted from now on (environment override at create:
 * SWEGL_B200_SHADING=exact|fast).  No reference counterpart: the reference has one arithmetic. */
int  swegl_b200_set_shading(swegl_b200_ctx *ctx, int mode);

/* shared != 0: other contexts render on this context's GPU at the same time (several frames in flight, one context each:
 * swegl_b200::pipeline_t, swegl_b200/pipeline.py).  The context then prefers kernel shapes that hold fewer SM resources
 * while they wait over the ones with the shortest critical path (today: the span kernel at 128 instead of 256 threads
 * per 32 scanlines -- a single frame takes ~5 us longer, four contexts together render ~6 % more frames per second).
 * Same pixels either way.  Default off (environment override at create: SWEGL_B200_SHARED_GPU=0|1).  No reference
 * counterpart. */
int  swegl_b200_set_shared_gpu(swegl_b200_ctx *ctx, int shared);

/* FNV-1a-64 over 32-bit words (offset basis 1469598103934665603, prime 1099511628211, one multiply per word): the frame
 * fingerprint of SURVEY 8c / tests/golden/MANIFEST.json.  Host side, no device involved. */
uint64_t swegl_b200_frame_hash(const uint32_t *words, size_t n_words);

/* ---- multi-GPU single-frame output over peer memory ----
 * Sort-first sharding gives every GPU a row band of the viewport (viewport_desc.band_y0/y1).  Instead of rendering
 * into its own screen and sending the band afterwards, a GPU can write its finished pixels straight into the screen
 * of the GPU that assembles the frame: that GPU exports its screen (a 64-byte CUDA IPC handle, to be shipped to the
 * other processes by any means, e.g. a torch.distributed broadcast), the others import it and select it as their
 * colour target.  The stores of the last kernel of the frame (k_fragments, or k_dof with the DoF pass) then travel
 * over NVLink while the kernel is still computing; what is left of the "gather" is a barrier.
 * set_color_target(NULL) restores the context's own screen; the target must have the context's screen size.
 * Within one process a target can simply be another context's screen pointer (swegl_b200_device_buffers). */
int  swegl_b200_export_screen(swegl_b200_ctx *ctx, void *handle64);
int  swegl_b200_import_screen(swegl_b200_ctx *ctx, const void *handle64, void **peer_screen);
int  swegl_b200_set_color_target(swegl_b200_ctx *ctx, void *device_screen);

/* Frame protocol for the band-sharded single frame, without a collective.  After every rank has selected rank 0's
 * screen as its colour target, swegl_b200_set_frame_sync(ctx, rank, world) makes the banded, opaque views submitted
 * through swegl_b200_render_viewport_device(stats == NULL) follow this protocol, entirely on the GPUs:
 *   rank 0   first kernel: clears the other ranks' rows of its screen (local HBM writes) and publishes "ready(frame)";
 *            last kernel: waits until every other rank's "done(frame)" flag has arrived -- when the stream reaches the
 *            end of the view, the whole frame is in rank 0's screen;
 *   rank r   waits for "ready(frame)" (a load over NVLink) in front of its last kernel, stores ONLY the tiles it drew
 *            into (the background is already there), then writes "done(frame)" into rank 0's memory.
 * The flags live behind the pixels of the exported screen, so import_screen is all the set-up it needs.  Every rank
 * must submit the same sequence of such views; call it on all ranks (rank 0 first, then a host barrier) to (re)start
 * the sequence, e.g. after a frame reported SWEGL_B200_ERR_CAPACITY.  rank < 0 switches the protocol off.
 * Waits give up after about 2 s (a lost peer must not hang the GPU); frame_sync_status counts those and, on rank 0,
 * reports how long rank 0's own share of the last frame took (clear + own band, without the wait for the others:
 * the input of a time-based band balance).  Either pointer may be null. */
int  swegl_b200_set_frame_sync(swegl_b200_ctx *ctx, int rank, int world);
int  swegl_b200_frame_sync_status(swegl_b200_ctx *ctx, uint32_t *timeouts, float *rank0_own_ms);

/* ---- read-back of device state (parity tests, multi-GPU gather) ---- */
/* device pointers of the screen (screen_w*screen_h words) and of the last viewport's depth */
int  swegl_b200_device_buffers(swegl_b200_ctx *ctx, void **screen_dev, void **depth_dev);
/* copy rows [y0,y1) of the device screen to host */
int  swegl_b200_read_screen(swegl_b200_ctx *ctx, int32_t y0, int32_t y1, void *pixels, int32_t pitch_bytes);
int  swegl_b200_read_depth(swegl_b200_ctx *ctx, float *zbuffer);
/* a rectangle of the device screen into its place in a host surface (`pixels` = the surface's origin), and rows
 * [row0, row1) of the last viewport's depth into their place in a host depth buffer: what a multi-context host uses to
 * assemble viewport_t::m_screen / m_zbuffer (swegl_b200/host/swegl_b200_host.hpp) */
int  swegl_b200_read_rect(swegl_b200_ctx *ctx, int32_t x, int32_t y, int32_t w, int32_t h, void *pixels, int32_t pitch_bytes);
int  swegl_b200_read_depth_rows(swegl_b200_ctx *ctx, int32_t row0, int32_t row1, float *zbuffer);
/* Several contexts of ONE process, one per GPU: lets this context's device store into `peer_device`'s memory (NVLink peer
 * access), so that another context's screen pointer (swegl_b200_device_buffers) can be this one's colour target without the
 * CUDA IPC detour of export / import_screen.  SWEGL_B200_ERR_UNSUPPORTED when the devices cannot reach each other. */
int  swegl_b200_enable_peer(swegl_b200_ctx *ctx, int peer_device);
int  swegl_b200_device_of(const swegl_b200_ctx *ctx);
/* 1 when every material and texel of the uploaded scene has alpha 255 (viewport_desc.transparency_layers is then ignored:
 * the layer logic of renderer.cpp:500-550 is the identity), 0 when not, -1 before upload_scene */
int  swegl_b200_scene_opaque(const swegl_b200_ctx *ctx);
/* post-render vertex state of the last viewport, as the reference leaves it in
 * mesh_vertex_t: any pointer may be null. v_viewport holds pixel coordinates
 * for yes-vertices and NDC for the others (vertex_shaders.hpp:72-84).  After a band-culled view only the vertex
 * blocks that view needed are current (the others keep older values); render an un-culled view to read them all. */
int  swegl_b200_read_vertices(swegl_b200_ctx *ctx, float *v_world, float *v_viewport,
                              float *normal_world, uint8_t *yes);

/* ---- sort-first band culling (multi-GPU row bands) ----
 * A banded view (viewport_desc.band_y0/y1) only has to transform, mark and set up what can reach its rows.  The
 * library keeps a static table of triangle clusters (64 consecutive triangles of the draw order, with the box of
 * their vertices) and skips, per banded view, the clusters whose projected box misses the band -- dilated over the
 * "shares a vertex" adjacency so that the `yes` flags of every vertex that is looked at (renderer.cpp:86-185,
 * 248-253) are exactly the reference's.  The frame is bit-identical with and without it.
 * policy: -1 automatic (banded views of scenes with >= 16384 triangles), 0 never, 1 every banded view.
 * Call before upload_scene (the tables are built there); environment override: SWEGL_B200_CULL=0|1. */
int  swegl_b200_set_band_culling(swegl_b200_ctx *ctx, int policy);
/* diagnostics of the last view: counts[0..5] = clusters, vertex blocks, live clusters, clusters marked,
 * vertex blocks transformed, 1 if that view was culled at all (else the three counts equal the totals) */
int  swegl_b200_cull_counts(swegl_b200_ctx *ctx, uint32_t counts[6]);


/* Host side, no device involved: decode a PNG or JPEG image as embedded in a .glb into the texel layout
 * the reference's loader produces with libpng / libjpeg -- row-major uint32, bytes b,g,r,a, alpha 255 where the file has
 * none (src/misc/image.cpp:93-258: read_png_file, read_jpeg_file; called from src/data/gltf.cpp:77-111).  The texels are
 * the ones libpng / libjpeg(-turbo) produce, bit for bit (JPEG: islow inverse DCT, fancy upsampling, jdcolor tables),
 * so a texture decoded here and one decoded by the reference sample identically.  PNG: all colour types and bit depths,
 * tRNS, Adam7 interlacing.  JPEG: baseline / extended / progressive Huffman, 8 bit, 1 or 3 components, sampling 1x1 2x1 2x2,
 * restart intervals.  *texels_bgra is malloc'ed; release it with swegl_b200_image_free.  Errors: SWEGL_B200_ERR_ARG,
 * SWEGL_B200_ERR_UNSUPPORTED (malformed or unsupported file; swegl_b200_image_error() says which, per thread). */
int  swegl_b200_decode_image(const void *data, size_t size, uint32_t **texels_bgra, int32_t *width, int32_t *height);
void swegl_b200_image_free(uint32_t *texels_bgra);
const char *swegl_b200_image_error(void);

/* Self-test of the device's shared-divisor division (csrc/common.cuh div_by: several quotients by one divisor reuse the
 * refined reciprocal of div.rn.f32's own expansion) against __fdiv_rn over `n_pairs` generated operand pairs.
 * out[0] = quotients whose bits differ (must be 0), out[1] = pairs that took the fast path.  No reference counterpart:
 * it guards the bit-exactness of normalize() (points.hpp:71-90) and of the light sums (pixel_shaders.cpp:173-203). */
int  swegl_b200_selftest_division(swegl_b200_ctx *ctx, uint64_t n_pairs, uint32_t seed, uint64_t out[2]);
/* Self-test of the FAST kernels' bilinear filter (same arithmetic as pixel_shader_texture_bilinear::shade,
 * pixel_shaders.cpp:348-384, without the per-operation range checks; csrc/fragment.cu tex_filter_bilinear_fast) against the
 * exact kernels' filter over `n_samples` generated (texels, texture coordinate) samples below the 2^21 guard.
 * out[0] = samples whose filtered colour or texel index differs (must be 0). */
int  swegl_b200_selftest_filter(swegl_b200_ctx *ctx, uint64_t n_samples, uint32_t seed, uint64_t out[1]);

#ifdef __cplusplus
}
#endif
#endif /* SWEGL_B200_H */
