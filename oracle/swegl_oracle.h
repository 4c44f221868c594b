/*
 * swegl_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, serial) of the swegl per-frame hot path, used as the
 * parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (swegl_b200/) never links, imports or executes it.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so
 * this restatement is pinned against the reference ITSELF: oracle/Makefile builds
 * the unmodified reference sources into oracle/_ref/libswegl_ref.so and
 * tests/test_oracle_vs_ref.py requires bit-identical frames, depth buffers and
 * per-vertex state on every scene/pose in tests/golden/MANIFEST.json; the hashes
 * those runs produced are committed there and re-checked without the reference.
 * Two stated deviations, both where the reference has no defined behaviour:
 *   - DoF is the repaired "DoF-R" semantics (post_shaders.hpp:51-132 is
 *     out-of-bounds at HEAD)  -> DoF parity is oracle-vs-CUDA only ("unpinned");
 *   - material_id == -1 samples scene.default_material's colour instead of
 *     materials[-1] (pixel_shaders.cpp:288-294 is UB there), and negative wrapped
 *     texel indices are wrapped instead of read out of bounds.
 */
#ifndef SWEGL_ORACLE_H
#define SWEGL_ORACLE_H

#include "../include/swegl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_dump {
    /* optional outputs, n_vertices long (x3 / x3 / x3 / x1); null = skip */
    float   *v_world;
    float   *v_viewport;
    float   *normal_world;
    uint8_t *yes;
    /* counters */
    uint64_t n_fill_triangle;   /* fill_triangle calls that passed the yes test     */
    uint64_t n_setup_triangles; /* fill_triangle_2 calls                            */
    uint64_t n_spans;           /* scanlines with x1 < x2                           */
    uint64_t n_fragments;       /* fragments that passed the z test (were shaded)   */
    uint64_t n_covered;         /* pixels with depth != 0x7F7F7F7F at the end       */
    uint64_t n_texel_guard;     /* bilinear fetches whose row / column was negative: the reference reads out of bounds there */
} orc_dump;

/* One swegl::render(scene, viewport) for a single viewport.
 * pixels: the whole screen (screen_h rows of pitch_bytes), like SDL_Surface::pixels;
 * zbuffer: vp->w * vp->h floats (viewport_t::m_zbuffer). Returns 0 on success. */
int orc_render(const swegl_b200_scene_desc *scene, const swegl_b200_frame_desc *frame,
               const swegl_b200_viewport_desc *vp,
               uint32_t *pixels, int32_t pitch_bytes, int32_t screen_w, int32_t screen_h,
               float *zbuffer, orc_dump *dump);

/* DoF-R alone: src/dst are w*h colour words, depth w*h floats. */
void orc_dof_r(const uint32_t *src, const float *depth, uint32_t *dst, int w, int h,
               float focal_distance, float focal_depth);

/* FNV-1a-64 over 32-bit words (offset 1469598103934665603, prime 1099511628211,
 * one multiply per word). */
uint64_t orc_fnv1a64_words(const uint32_t *words, size_t n);

#ifdef __cplusplus
}
#endif
#endif
