// geometry.cu — vertex stage, cull/mark, near clip + triangle setup, edge walk, span setup.
//
// Stage map against the reference (paths relative to the swegl checkout):
//   k_vertex_world  vertex_shader_t::original_to_world            vertex_shaders.hpp:16-33
//   k_vertex_view   world_to_camera_or_frustum / camera_to_frustum vertex_shaders.hpp:35-52,61-71
//   k_mark          the mark pass of _render                      renderer.cpp:86-185
//   k_setup         fill_triangle + the head of fill_triangle_2   renderer.cpp:240-394
//   k_edgewalk      fill_triangle_2 + the y loop of fill_half_triangle   renderer.cpp:396-460,467,553-556
//   k_spans         the per-scanline part of fill_half_triangle   renderer.cpp:469-480
// The fp32 recurrences (edge x += ratio, topalpha += topstep ...) are replayed step by step so
// coverage and depth are bit-identical to the CPU renderer; everything that is not a recurrence
// runs one thread per vertex / triangle / scanline.
#include "common.cuh"

namespace sb {

static constexpr int TPB = 256;

// ----------------------------------------------------------------------------------------
// vertex stage
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_vertex_world(DeviceScene s)
{
    uint32_t i = blockIdx.x * TPB + threadIdx.x;
    if (i >= s.n_vertices) return;
    const float *M = s.node_world + 16 * s.vert_node[i];
    V3 w = xform(M, v3(s.pos[3 * i], s.pos[3 * i + 1], s.pos[3 * i + 2]));
    s.v_world[3 * i] = w.x; s.v_world[3 * i + 1] = w.y; s.v_world[3 * i + 2] = w.z;
}

__global__ void __launch_bounds__(TPB) k_vertex_view(DeviceScene s, const __grid_constant__ ViewParams vp)
{
    uint32_t i = blockIdx.x * TPB + threadIdx.x;
    if (i >= s.n_vertices) return;
    V3 w = v3(s.v_world[3 * i], s.v_world[3 * i + 1], s.v_world[3 * i + 2]);
    V3 p = project(vp.proj, xform(vp.view, w));
    V3 n = normal_to_world(s.node_normal + 9 * s.vert_node[i], v3(s.nrm[3 * i], s.nrm[3 * i + 1], s.nrm[3 * i + 2]));
    s.v_ndc[3 * i] = p.x; s.v_ndc[3 * i + 1] = p.y; s.v_ndc[3 * i + 2] = p.z;
    s.n_world[3 * i] = n.x; s.n_world[3 * i + 1] = n.y; s.n_world[3 * i + 2] = n.z;
    s.yes[i] = 0;
}

// ----------------------------------------------------------------------------------------
// mark pass: a triangle that is inside the frustum and front facing (or double sided) marks its
// three vertices; fill_triangle later draws every triangle whose vertices are all marked.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_mark(DeviceScene s)
{
    uint32_t t = blockIdx.x * TPB + threadIdx.x;
    if (t >= s.n_tris) return;
    Tri tr = s.tris[t];
    V3 a = v3(s.v_ndc[3 * tr.i0], s.v_ndc[3 * tr.i0 + 1], s.v_ndc[3 * tr.i0 + 2]);
    V3 b = v3(s.v_ndc[3 * tr.i1], s.v_ndc[3 * tr.i1 + 1], s.v_ndc[3 * tr.i1 + 2]);
    V3 c = v3(s.v_ndc[3 * tr.i2], s.v_ndc[3 * tr.i2 + 1], s.v_ndc[3 * tr.i2 + 2]);
    // inside_camera_frustum, renderer.cpp:58-70 (symmetric in the three vertices)
    bool inside = (a.x >= -1.f || b.x >= -1.f || c.x >= -1.f)
               && (a.y >= -1.f || b.y >= -1.f || c.y >= -1.f)
               && (a.x < 1.f || b.x < 1.f || c.x < 1.f)
               && (a.y < 1.f || b.y < 1.f || c.y < 1.f)
               && (a.z >= NEAR_Z || b.z >= NEAR_Z || c.z >= NEAR_Z)
               && (a.x != b.x || a.x != c.x)
               && (a.y != b.y || a.y != c.y);
    if (!inside) return;
    if (!s.prims[tr.prim].double_sided) {
        // front_face_visible, renderer.cpp:72-75: sign of the NORMALISED cross product's z
        if (!(cross_n(sub(b, a), sub(c, a)).z > 0.f)) return;
    }
    s.yes[tr.i0] = 1; s.yes[tr.i1] = 1; s.yes[tr.i2] = 1;
}

// ----------------------------------------------------------------------------------------
// setup
// ----------------------------------------------------------------------------------------
struct SV {             // one vertex as fill_triangle sees it
    V3 s;               // v_viewport (pixel x,y + depth)
    V3 w;               // v_world
    V3 n;               // normal_world
    float tu, tv;       // tex_coords
    uint32_t idx;       // index in the scene arrays (for the object-space normal when clipping)
};

SB_DEV void swap_sv(SV &a, SV &b) { SV t = a; a = b; b = t; }

SB_DEV SV load_sv(const DeviceScene &s, const ViewParams &vp, uint32_t i)
{
    SV v;
    v.s = v3(s.v_ndc[3 * i], s.v_ndc[3 * i + 1], s.v_ndc[3 * i + 2]);
    to_viewport(vp, v.s);                                   // frustum_to_viewport (only yes vertices get here)
    v.w = v3(s.v_world[3 * i], s.v_world[3 * i + 1], s.v_world[3 * i + 2]);
    v.n = v3(s.n_world[3 * i], s.n_world[3 * i + 1], s.n_world[3 * i + 2]);
    v.tu = s.uv[2 * i]; v.tv = s.uv[2 * i + 1];
    v.idx = i;
    return v;
}

// the vertex the near clipper creates on edge from->to (renderer.cpp:289-297) and pushes through
// vertex_shader_t::world_to_viewport (vertex_shaders.hpp:54-59)
SB_DEV SV make_clip_vertex(const DeviceScene &s, const ViewParams &vp, const float *m9, const SV &from, const SV &to, float cut)
{
    SV r;
    r.w = add(from.w, mul(sub(to.w, from.w), cut));
    r.tu = fadd(from.tu, fmul(fsub(to.tu, from.tu), cut));
    r.tv = fadd(from.tv, fmul(fsub(to.tv, from.tv), cut));
    V3 nf = v3(s.nrm[3 * from.idx], s.nrm[3 * from.idx + 1], s.nrm[3 * from.idx + 2]);
    V3 nt = v3(s.nrm[3 * to.idx], s.nrm[3 * to.idx + 1], s.nrm[3 * to.idx + 2]);
    V3 n = nf;
    if (!(nt.x == nf.x && nt.y == nf.y && nt.z == nf.z))
        n = normalize(add(nf, mul(sub(nt, nf), cut)));      // normal_t::operator=(vector_t) normalises
    r.n = normal_to_world(m9, n);
    r.s = project(vp.proj, xform(vp.view, r.w));
    to_viewport(vp, r.s);
    r.idx = 0xFFFFFFFFu;
    return r;
}

// fill_triangle_2 up to the row count: y sort, ceil limits, and the records the later stages need
SB_DEV void emit_slot(const DeviceScene &s, const ViewParams &vp, const FrameParams &fp, const Pools &pl,
                      uint32_t slot, uint32_t prim_id, const Prim &pr, SV a, SV b, SV c, bool front_face_visible)
{
    bool inverted = !front_face_visible;
    if (b.s.y < a.s.y) swap_sv(a, b);                       // renderer.cpp:375-387
    if (c.s.y < b.s.y) swap_sv(b, c);
    if (b.s.y < a.s.y) swap_sv(a, b);
    int y0 = ceil_i(a.s.y), y1 = ceil_i(b.s.y), y2 = ceil_i(c.s.y);
    if (y0 == y2) return;                                   // renderer.cpp:394

    // scanlines this slot will walk inside the band (upper half :416-421, lower half :439-444)
    int n = 0;
    if (y1 >= vp.vy) {
        int ya = max(max(y0, vp.vy), vp.band0), yb = min(min(y1, vp.vy + vp.vh), vp.band1);
        n += max(0, yb - ya);
    }
    if (y1 < vp.vy + vp.vh) {
        int ya = max(max(y1, vp.vy), vp.band0), yb = min(min(y2, vp.vy + vp.vh), vp.band1);
        n += max(0, yb - ya);
    }
    if (n == 0) return;
    uint32_t base = atomicAdd(&pl.counters->n_rows, (uint32_t)n);
    if (base + (uint32_t)n > pl.rows_cap) { atomicOr(&pl.counters->overflow, 1u); return; }

    SlotEdge e;
    e.x0 = a.s.x; e.y0 = a.s.y; e.z0 = a.s.z;
    e.x1 = b.s.x; e.y1 = b.s.y; e.z1 = b.s.z;
    e.x2 = c.s.x; e.y2 = c.s.y; e.z2 = c.s.z;
    e.span_base = (int32_t)base; e.pad0 = 0; e.pad1 = 0;
    pl.edges[slot] = e;

    SlotShade sh;
    sh.w0[0] = a.w.x; sh.w0[1] = a.w.y; sh.w0[2] = a.w.z;
    sh.w1[0] = b.w.x; sh.w1[1] = b.w.y; sh.w1[2] = b.w.z;
    sh.w2[0] = c.w.x; sh.w2[1] = c.w.y; sh.w2[2] = c.w.z;
    float sg = inverted ? -1.0f : 1.0f;                     // pixel_shaders.cpp:88-105 (sign flip is exact)
    sh.n0[0] = sg * a.n.x; sh.n0[1] = sg * a.n.y; sh.n0[2] = sg * a.n.z;
    sh.n1[0] = sg * b.n.x; sh.n1[1] = sg * b.n.y; sh.n1[2] = sg * b.n.z;
    sh.n2[0] = sg * c.n.x; sh.n2[1] = sg * c.n.y; sh.n2[2] = sg * c.n.z;
    float tw = (float)pr.tw, th = (float)pr.th;             // pixel_shaders.cpp:304-318
    sh.t0[0] = fmul(a.tu, tw); sh.t0[1] = fmul(a.tv, th);
    sh.t1[0] = fmul(b.tu, tw); sh.t1[1] = fmul(b.tv, th);
    sh.t2[0] = fmul(c.tu, tw); sh.t2[1] = fmul(c.tv, th);
    sh.flat_light = 0.0f;
    if (vp.light_mode == SWEGL_B200_LIGHT_FLAT) {
        // pixel_shader_lights_flat::prepare_for_triangle, pixel_shaders.cpp:33-84
        V3 nw = cross_n(sub(b.w, a.w), sub(c.w, a.w));
        if (inverted) nw = normalize(neg(nw));              // operator-(normal_t) re-normalises
        float sun = -dot(nw, v3(fp.sun[0], fp.sun[1], fp.sun[2]));
        if (sun < 0.0f) sun = 0.0f; else sun = fmul(sun, fp.sun_intensity);
        V3 center = add(add(a.w, b.w), c.w);
        center = v3(fdiv(center.x, 3.0f), fdiv(center.y, 3.0f), fdiv(center.z, 3.0f));
        V3 camv = normalize(sub(v3(vp.cam[0], vp.cam[1], vp.cam[2]), center));
        float dyn = point_lights_sum(fp, center, nw, camv);
        sh.flat_light = fmul(fadd(fadd(fp.ambient, sun), dyn), 65536.0f);
    }
    sh.prim = prim_id;
    #pragma unroll
    for (int k = 0; k < 6; k++) sh.pad[k] = 0;
    pl.shades[slot] = sh;

    pl.live[atomicAdd(&pl.counters->n_live, 1u)] = slot;
}

__global__ void __launch_bounds__(128) k_setup(DeviceScene s, const __grid_constant__ ViewParams vp,
                                               const __grid_constant__ FrameParams fp, Pools pl)
{
    uint32_t t = blockIdx.x * 128 + threadIdx.x;
    if (t >= s.n_tris) return;
    Tri tr = s.tris[t];
    if (!s.yes[tr.i0] || !s.yes[tr.i1] || !s.yes[tr.i2]) return;        // renderer.cpp:248-253
    Prim pr = s.prims[tr.prim];
    SV a = load_sv(s, vp, tr.i0), b = load_sv(s, vp, tr.i1), c = load_sv(s, vp, tr.i2);

    bool ffv = cross_n(sub(b.s, a.s), sub(c.s, a.s)).z < 0.0f;          // renderer.cpp:258
    bool inverted_order = false;
    if (b.s.z > a.s.z) { swap_sv(a, b); inverted_order = !inverted_order; }   // sort by z DESC, :264-279
    if (c.s.z > b.s.z) { swap_sv(b, c); inverted_order = !inverted_order; }
    if (b.s.z > a.s.z) { swap_sv(a, b); inverted_order = !inverted_order; }

    const float *m9 = s.node_normal + 9 * pr.node;
    if (c.s.z >= NEAR_Z) {
        emit_slot(s, vp, fp, pl, 2 * t, tr.prim, pr, a, b, c, ffv);
    } else if (b.s.z < NEAR_Z) {
        // only a in front of the camera, renderer.cpp:286-317
        float cut_1 = fdiv(fsub(a.s.z, 0.001f), fsub(a.s.z, b.s.z));
        SV n1 = make_clip_vertex(s, vp, m9, a, b, cut_1);
        float cut_2 = fdiv(fsub(a.s.z, 0.001f), fsub(a.s.z, c.s.z));
        SV n2 = make_clip_vertex(s, vp, m9, a, c, cut_2);
        ffv = cross_n(sub(n1.s, a.s), sub(n2.s, a.s)).z < 0.0f;
        if (inverted_order) ffv = !ffv;
        emit_slot(s, vp, fp, pl, 2 * t, tr.prim, pr, a, n1, n2, ffv);
    } else if (c.s.z < NEAR_Z) {
        // only c behind the camera: two triangles, renderer.cpp:318-356
        float cut_0 = fdiv(fsub(a.s.z, 0.001f), fsub(a.s.z, c.s.z));
        SV n1 = make_clip_vertex(s, vp, m9, a, c, cut_0);
        float cut_1 = fdiv(fsub(b.s.z, 0.001f), fsub(b.s.z, c.s.z));
        SV n2 = make_clip_vertex(s, vp, m9, b, c, cut_1);
        ffv = cross_n(sub(b.s, a.s), sub(n2.s, a.s)).z < 0.0f;
        if (inverted_order) ffv = !ffv;
        emit_slot(s, vp, fp, pl, 2 * t, tr.prim, pr, a, b, n2, ffv);
        ffv = cross_n(sub(n2.s, a.s), sub(n1.s, a.s)).z < 0.0f;
        if (inverted_order) ffv = !ffv;
        emit_slot(s, vp, fp, pl, 2 * t + 1, tr.prim, pr, a, n2, n1, ffv);
    }
}

// ----------------------------------------------------------------------------------------
// edge walk: one thread per live slot replays the per-scanline recurrences of both edges
// ----------------------------------------------------------------------------------------
struct Side { Interp ip; float ratio, x; };

SB_DEV void walk_half(const ViewParams &vp, Row *rows, uint32_t &k, uint32_t slot, int lower,
                      int y, int y_end, bool lor, Side &lng, Side &sht)
{
    uint32_t flags = (slot << 2) | ((uint32_t)lower << 1) | (lor ? 1u : 0u);
    for (; y < y_end; y++) {
        if (y >= vp.band0 && y < vp.band1) {
            Row r;
            const Side &L = lor ? sht : lng;
            const Side &R = lor ? lng : sht;
            r.lx = L.x; r.rx = R.x;
            r.ltop = L.ip.top; r.lbot = L.ip.bottom; r.rtop = R.ip.top; r.rbot = R.ip.bottom;
            r.slot_flags = flags; r.y = y;
            rows[k++] = r;
        }
        lng.x = fadd(lng.x, lng.ratio);                     // renderer.cpp:553-556
        sht.x = fadd(sht.x, sht.ratio);
        interp_step(lng.ip);
        interp_step(sht.ip);
    }
}

__global__ void __launch_bounds__(128) k_edgewalk(const __grid_constant__ ViewParams vp, Pools pl)
{
    uint32_t n_live = pl.counters->n_live;
    for (uint32_t li = blockIdx.x * 128 + threadIdx.x; li < n_live; li += gridDim.x * 128) {
        uint32_t slot = pl.live[li];
        SlotEdge e = pl.edges[slot];
        int y0 = ceil_i(e.y0), y1 = ceil_i(e.y1), y2 = ceil_i(e.y2);
        uint32_t k = (uint32_t)e.span_base;

        Side lng, sht;
        lng.ratio = fdiv(fsub(e.x2, e.x0), fsub(e.y2, e.y0));              // renderer.cpp:400-409
        interp_init_self(lng.ip, fsub(e.y2, e.y0), e.z0, e.z2);
        {
            float move = (y0 < vp.vy) ? fsub((float)vp.vy, e.y0) : fsub((float)y0, e.y0);
            interp_displace(lng.ip, move);
            lng.x = fadd(e.x0, fmul(lng.ratio, move));
        }
        if (y1 >= vp.vy) {                                                  // upper half, :416-436
            sht.ratio = fdiv(fsub(e.x1, e.x0), fsub(e.y1, e.y0));
            interp_init_self(sht.ip, fsub(e.y1, e.y0), e.z0, e.z1);
            int y = max(y0, vp.vy), y_end = min(y1, vp.vy + vp.vh);
            float move = fsub((float)y, e.y0);
            interp_displace(sht.ip, move);
            sht.x = fadd(e.x0, fmul(sht.ratio, move));
            bool lor = lng.ratio > sht.ratio;
            walk_half(vp, pl.rows, k, slot, 0, y, y_end, lor, lng, sht);
        }
        if (y1 < vp.vy + vp.vh) {                                           // lower half, :439-459
            sht.ratio = fdiv(fsub(e.x2, e.x1), fsub(e.y2, e.y1));
            interp_init_self(sht.ip, fsub(e.y2, e.y1), e.z1, e.z2);
            int y = max(y1, vp.vy), y_end = min(y2, vp.vy + vp.vh);
            float move = fsub((float)y, e.y1);
            interp_displace(sht.ip, move);
            sht.x = fadd(e.x1, fmul(sht.ratio, move));
            bool lor = lng.ratio < sht.ratio;
            walk_half(vp, pl.rows, k, slot, 1, y, y_end, lor, lng, sht);
        }
    }
}

// ----------------------------------------------------------------------------------------
// spans: one thread per scanline record turns the edge state into the per-pixel interpolator,
// replays it along x and drops a checkpoint ("chunk") at every 32-column bin it crosses
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_spans(const __grid_constant__ ViewParams vp, Pools pl)
{
    if (pl.counters->overflow & 1u) return;      // some rows were never written: the host grows the pool and redoes the frame
    uint32_t n_rows = min(pl.counters->n_rows, pl.rows_cap);
    Span *spans = reinterpret_cast<Span *>(pl.rows);
    for (uint32_t i = blockIdx.x * TPB + threadIdx.x; i < n_rows; i += gridDim.x * TPB) {
        Row r = pl.rows[i];
        Span sp;
        sp.slot_flags = r.slot_flags;
        int x1 = max(ceil_i(r.lx), vp.vx);                                  // renderer.cpp:469-470
        int x2 = min(ceil_i(r.rx), vp.vx + vp.vw);
        if (!(x1 < x2)) {
            sp.topstep = sp.bottomstep = sp.v0 = sp.v1 = sp.pl = sp.pr = 0.f; sp.x1x2 = 0;
            spans[i] = sp;
            continue;
        }
        uint32_t slot = r.slot_flags >> 2;
        bool lower = (r.slot_flags >> 1) & 1u, lor = r.slot_flags & 1u;
        const SlotEdge &e = pl.edges[slot];
        float z0 = e.z0, z1 = e.z1, z2 = e.z2;
        // edge interpolators' v[0]: long (z0, z2-z0); short (z0, z1-z0) upper / (z1, z2-z1) lower
        float la = lor ? (lower ? z1 : z0) : z0;
        float lb = lor ? (lower ? z2 : z1) : z2;
        float ra = lor ? z0 : (lower ? z1 : z0);
        float rb = lor ? z2 : (lower ? z2 : z1);
        sp.pl = fdiv(r.ltop, r.lbot);                                       // interpolator progress()
        sp.pr = fdiv(r.rtop, r.rbot);
        float zl = fadd(la, fmul(fsub(lb, la), sp.pl));                     // value(0), interpolator.hpp:103
        float zr = fadd(ra, fmul(fsub(rb, ra), sp.pr));
        Interp q;
        interp_init_self(q, fsub(r.rx, r.lx), zl, zr);                      // renderer.cpp:476-480
        interp_displace(q, fsub((float)x1, r.lx));
        sp.topstep = q.topstep; sp.bottomstep = q.bottomstep; sp.v0 = q.v0; sp.v1 = q.v1;
        sp.x1x2 = (uint32_t)x1 | ((uint32_t)x2 << 16);
        spans[i] = sp;

        int b0 = (x1 - vp.vx) >> 5, b1 = (x2 - 1 - vp.vx) >> 5;
        uint32_t nchunks = (uint32_t)(b1 - b0 + 1);
        uint32_t cbase = atomicAdd(&pl.counters->n_chunks, nchunks);
        if (cbase + nchunks > pl.chunks_cap) { atomicOr(&pl.counters->overflow, 2u); continue; }
        int32_t *heads = pl.bin_head + (size_t)(r.y - vp.vy) * vp.nbx;
        int b = b0;
        for (int x = x1;;) {
            Chunk ch;
            ch.top = q.top; ch.bottom = q.bottom; ch.span = i;
            ch.next = atomicExch(&heads[b], (int32_t)cbase);
            pl.chunks[cbase] = ch;
            cbase++; b++;
            int xn = min(vp.vx + (b << 5), x2);                             // first column of the next bin
            if (xn >= x2) break;
            for (; x < xn; x++) interp_step(q);                             // renderer.cpp:486 (Step per pixel)
        }
    }
}

// ----------------------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------------------
static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

void launch_vertex_world(const DeviceScene &s, cudaStream_t st)
{
    if (s.n_vertices) k_vertex_world<<<cdiv(s.n_vertices, TPB), TPB, 0, st>>>(s);
}
void launch_vertex_view(const DeviceScene &s, const ViewParams &vp, cudaStream_t st)
{
    if (s.n_vertices) k_vertex_view<<<cdiv(s.n_vertices, TPB), TPB, 0, st>>>(s, vp);
}
void launch_mark(const DeviceScene &s, cudaStream_t st)
{
    if (s.n_tris) k_mark<<<cdiv(s.n_tris, TPB), TPB, 0, st>>>(s);
}
void launch_setup(const DeviceScene &s, const ViewParams &vp, const FrameParams &fp, const Pools &p, cudaStream_t st)
{
    if (s.n_tris) k_setup<<<cdiv(s.n_tris, 128), 128, 0, st>>>(s, vp, fp, p);
}
void launch_edgewalk(const ViewParams &vp, const Pools &p, uint32_t max_live, cudaStream_t st)
{
    if (!max_live) return;
    unsigned blocks = min(cdiv(max_live, 128u), 148u * 16u);
    k_edgewalk<<<blocks, 128, 0, st>>>(vp, p);
}
void launch_spans(const ViewParams &vp, const Pools &p, cudaStream_t st)
{
    k_spans<<<148 * 8, TPB, 0, st>>>(vp, p);
}

} // namespace sb
