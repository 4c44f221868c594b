// geometry.cu — vertex stage, cull/mark, near clip + triangle setup, edge walk, span setup.
//
// Stage map against the reference (paths relative to the swegl checkout):
//   k_vertex<W>     vertex_shader_t::original_to_world (W) +      vertex_shaders.hpp:16-33
//                   world_to_camera_or_frustum / camera_to_frustum vertex_shaders.hpp:35-52,61-71
//   k_mark          the mark pass of _render                      renderer.cpp:86-185
//   k_setup         fill_triangle + fill_triangle_2 (edge set-up)  renderer.cpp:240-460
//   k_spans         the y loop and per-scanline part of fill_half_triangle   renderer.cpp:467-480,553-556
// The fp32 recurrences (edge x += ratio, topalpha += topstep ...) must see the same additions as the
// CPU renderer for coverage and depth to be bit-identical; radd() (radd.h) performs k of them in
// O(1), so there is no serial edge walk: one thread per vertex / triangle / scanline everywhere.
#include <cstddef>
#include "common.cuh"

namespace sb {

static constexpr int TPB = 256;
#ifndef SPAN_SEG_V
#define SPAN_SEG_V 2
#endif
static constexpr uint32_t SPAN_SEG = SPAN_SEG_V;      // 32-column bins per span segment (one lane walks one segment).  Measured 1 / 2 / 4 / 8: truck 4K spans 29.7 / 29.6 / 32.6 / 39.4 us, BrainStem 71.0 / 67.8 / 67.4 / 71.7, 8K sphere 252 / 250 / 250 / 309

// ----------------------------------------------------------------------------------------
// band culling (row-band sharding): which triangle clusters / vertex blocks this view has to look at at all
// ----------------------------------------------------------------------------------------
// One thread per cluster.  Every vertex the cluster references lies inside its object-space box, and
// object -> world -> camera -> (y', z') is affine, so in exact arithmetic y'/z' of any such vertex lies between the
// extremes over the 8 box corners as long as z' > 0 on all of them.  The fp32 results the vertex stage really
// produces differ from exact by at most E = 64 ulp-units of the summed magnitudes (12 roundings on the way); the
// interval below carries 2E (corner + vertex).  A box that comes closer than 0.002 to the camera plane is never
// culled (the near clipper, renderer.cpp:286-356, then creates vertices of its own).  Rows walked by an unclipped
// triangle are [ceil(min y), ceil(max y)) of its pixel-space vertices (renderer.cpp:375-394).
// list[atomicAdd(count, #set lanes) + rank among the set lanes] = id, one atomic per warp (every lane of the warp calls it)
SB_DEV void append_ids(uint32_t *list, uint32_t *count, bool set, uint32_t id)
{
    const unsigned m = __ballot_sync(0xFFFFFFFFu, set);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (set) list[base + __popc(m & ((1u << lane) - 1u))] = id;
}

__global__ void __launch_bounds__(128) k_cull_live(DeviceScene s, const ViewParams *__restrict__ vpp, CullTables ct)
{
    pdl_trigger(s.early_trigger);
    const uint32_t c = blockIdx.x * 128 + threadIdx.x;
    const bool valid = c < ct.n_clusters;
    ClusterBox box;
    box.node = -1;
    if (valid) box = ct.boxes[c];
    bool live = valid;
    if (valid && box.node >= 0) {
        const float *M = s.node_world + 16 * box.node;
        const float *V = vpp->view, *P = vpp->proj;
        double aw[3], ac[3];
        const double pm[3] = { fmax(fabs((double)box.lo[0]), fabs((double)box.hi[0])), fmax(fabs((double)box.lo[1]), fabs((double)box.hi[1])),
                               fmax(fabs((double)box.lo[2]), fabs((double)box.hi[2])) };
        for (int r = 0; r < 3; r++)
            aw[r] = fabs((double)M[4 * r]) * pm[0] + fabs((double)M[4 * r + 1]) * pm[1] + fabs((double)M[4 * r + 2]) * pm[2] + fabs((double)M[4 * r + 3]);
        for (int r = 0; r < 3; r++)
            ac[r] = fabs((double)V[4 * r]) * aw[0] + fabs((double)V[4 * r + 1]) * aw[1] + fabs((double)V[4 * r + 2]) * aw[2] + fabs((double)V[4 * r + 3]);
        const double ay = fabs((double)P[4]) * ac[0] + fabs((double)P[5]) * ac[1] + fabs((double)P[6]) * ac[2] + fabs((double)P[7]);
        const double az = fabs((double)P[8]) * ac[0] + fabs((double)P[9]) * ac[1] + fabs((double)P[10]) * ac[2] + fabs((double)P[11]);
        const double E = 2.0 * 64.0 * 5.9604644775390625e-8;            // 2 x 64 x 2^-24
        const double Ey = E * ay, Ez = E * az;
        double rlo = 1e300, rhi = -1e300, zmin = 1e300;
        for (int k = 0; k < 8; k++) {
            const V3 p = v3((k & 1) ? box.hi[0] : box.lo[0], (k & 2) ? box.hi[1] : box.lo[1], (k & 4) ? box.hi[2] : box.lo[2]);
            const V3 q = xform(P, xform(V, xform(M, p)));                   // the vertex stage's own operations
            const double y = (double)q.y, z = (double)q.z;
            zmin = fmin(zmin, z);
            const double zl = z - Ez, zh = z + Ez, yh = y + Ey, yl = y - Ey;
            rhi = fmax(rhi, yh >= 0.0 ? yh / zl : yh / zh);
            rlo = fmin(rlo, yl >= 0.0 ? yl / zh : yl / zl);
        }
        if (zmin - Ez > 0.002) {
            const double m11 = (double)vpp->vp_m11, m13 = (double)vpp->vp_m13;
            const double a = m11 * rlo + m13, b = m11 * rhi + m13;
            const double ylo = fmin(a, b), yhi = fmax(a, b);
            const double margin = 1.0 + 1e-5 * (fabs(m11) * fmax(fabs(rlo), fabs(rhi)) + fabs(m13));
            // NaNs anywhere compare false: the cluster stays live
            if (yhi + margin < (double)vpp->band0 - 1.0 || ylo - margin > (double)vpp->band1 + 1.0) live = false;
        }
    }
    if (valid) ct.cl_live[c] = live ? 1 : 0;
    append_ids(ct.live_list, &ct.counts[0], live, c);
}

// One thread per cluster (mark_need) and per vertex block (vert_need): OR of cl_live over a static adjacency list.
__global__ void __launch_bounds__(128) k_cull_need(CullTables ct)
{
    pdl_trigger(ct.early_trigger);
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    const bool is_cl = i < ct.n_clusters;
    const uint32_t j = is_cl ? i : i - ct.n_clusters;
    const bool valid = is_cl || j < ct.n_vblocks;
    const uint32_t *off = is_cl ? ct.cl_adj_off : ct.vb_adj_off, *adj = is_cl ? ct.cl_adj : ct.vb_adj;
    uint32_t a = 0, b = 0;
    if (valid) { a = off[j]; b = off[j + 1]; }
    pdl_wait();                                                             // k_cull_live's flags
    bool need = false;
    for (uint32_t k = a; k < b && !need; k++) {
        const uint32_t c = adj[k];
        need = c == CULL_ALWAYS || ct.cl_live[c] != 0;
    }
    if (valid) (is_cl ? ct.mark_need : ct.vert_need)[j] = need ? 1 : 0;
    append_ids(ct.mark_list, &ct.counts[1], need && is_cl, j);              // (a warp may straddle the cluster / block boundary)
    append_ids(ct.vert_list, &ct.counts[2], need && valid && !is_cl, j);
}

// ----------------------------------------------------------------------------------------
// vertex stage
// ----------------------------------------------------------------------------------------
// v_world = M_node * v (A1, once per frame) and, per viewport, camera/projection transform, world normal and the
// yes reset (A2).  WORLD selects whether this launch is the first of the frame and also has to produce v_world.
// Block 0 clears the frame counters (no memset node in the graph).
template <bool WORLD>
SB_DEV void vertex_one(const DeviceScene &s, const ViewParams &vp, uint32_t i)
{
    const uint32_t node = s.vert_node[i];
    V3 w;
    if (WORLD) {
        w = xform(s.node_world + 16 * node, v3(s.pos[3 * i], s.pos[3 * i + 1], s.pos[3 * i + 2]));     // vertex_shaders.hpp:20-24
        s.v_world[3 * i] = w.x; s.v_world[3 * i + 1] = w.y; s.v_world[3 * i + 2] = w.z;
    } else {
        w = v3(s.v_world[3 * i], s.v_world[3 * i + 1], s.v_world[3 * i + 2]);
    }
    V3 p = project(vp.proj, xform(vp.view, w));
    V3 n = normal_to_world(s.node_normal + 9 * node, v3(s.nrm[3 * i], s.nrm[3 * i + 1], s.nrm[3 * i + 2]));
    s.v_ndc[3 * i] = p.x; s.v_ndc[3 * i + 1] = p.y; s.v_ndc[3 * i + 2] = p.z;
    s.n_world[3 * i] = n.x; s.n_world[3 * i + 1] = n.y; s.n_world[3 * i + 2] = n.z;
    s.yes[i] = 0;
}

template <bool WORLD>
__global__ void __launch_bounds__(TPB, 8) k_vertex(DeviceScene s, const ViewParams *__restrict__ vpp, Counters *__restrict__ counters)
{
    pdl_trigger(s.early_trigger);
    __shared__ ViewParams vp;
    if (blockIdx.x == 0 && threadIdx.x < sizeof(Counters) / 4)
        reinterpret_cast<uint32_t *>(counters)[threadIdx.x] = (threadIdx.x == offsetof(Counters, bb_x0) / 4 || threadIdx.x == offsetof(Counters, bb_y0) / 4) ? COUNTERS_BB_MIN_INIT : 0u;
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += TPB) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    __syncthreads();
    if (s.vert_list) {
        // culled view: persistent CTAs walk the list of vertex blocks some triangle that matters to the band uses
        pdl_wait();                                                         // k_cull_need's list
        constexpr uint32_t PER = TPB / CULL_CL;                             // vertex blocks per CTA pass
        const uint32_t n_items = s.cull_counts[2];
        for (uint32_t it = blockIdx.x * PER + threadIdx.x / CULL_CL; it < n_items; it += gridDim.x * PER) {
            const uint32_t i = s.vert_list[it] * CULL_CL + threadIdx.x % CULL_CL;
            if (i < s.n_vertices) vertex_one<WORLD>(s, vp, i);
        }
        return;
    }
    uint32_t i = blockIdx.x * TPB + threadIdx.x;
    if (i >= s.n_vertices) return;
    vertex_one<WORLD>(s, vp, i);
}

// ----------------------------------------------------------------------------------------
// mark pass: a triangle that is inside the frustum and front facing (or double sided) marks its
// three vertices; fill_triangle later draws every triangle whose vertices are all marked.
// ----------------------------------------------------------------------------------------
SB_DEV void mark_one(const DeviceScene &s, const Tri &tr)
{
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

__global__ void __launch_bounds__(TPB) k_mark(DeviceScene s)
{
    pdl_trigger(s.early_trigger);
    if (s.mark_list) {
        // culled view: only clusters that share a vertex with a cluster reaching the band can mark a vertex that is looked at
        pdl_wait();                                                         // k_vertex (and, through it, k_cull_need's list)
        constexpr uint32_t PER = TPB / CULL_CL;
        const uint32_t n_items = s.cull_counts[1];
        for (uint32_t it = blockIdx.x * PER + threadIdx.x / CULL_CL; it < n_items; it += gridDim.x * PER) {
            const uint32_t t = s.mark_list[it] * CULL_CL + threadIdx.x % CULL_CL;
            if (t < s.n_tris) mark_one(s, s.tris[t]);
        }
        return;
    }
    uint32_t t = blockIdx.x * TPB + threadIdx.x;
    if (t >= s.n_tris) { pdl_wait(); return; }                              // (every CTA waits: completion stays transitive along the chain)
    Tri tr = s.tris[t];                                                     // static scene data: may be read before the wait
    pdl_wait();                                                             // k_vertex's v_ndc and yes reset
    mark_one(s, tr);
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
struct RowRange { uint32_t base, n; };     // scanline records a slot allocated (n == 0: nothing to draw)

SB_DEV uint32_t warp_incl_scan(uint32_t v, int lane)
{
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= d) v += t; }
    return v;
}

// y sort + the number of scanlines the slot walks inside the band
struct SlotPlan { int y0, y1, y2, n; };

SB_DEV SlotPlan plan_slot(const ViewParams &vp, SV &a, SV &b, SV &c)
{
    SlotPlan p = { 0, 0, 0, 0 };
    if (b.s.y < a.s.y) swap_sv(a, b);                       // renderer.cpp:375-387
    if (c.s.y < b.s.y) swap_sv(b, c);
    if (b.s.y < a.s.y) swap_sv(a, b);
    p.y0 = ceil_i(a.s.y); p.y1 = ceil_i(b.s.y); p.y2 = ceil_i(c.s.y);
    if (p.y0 == p.y2) return p;                             // renderer.cpp:394
    // scanlines this slot will walk inside the band (upper half :416-421, lower half :439-444)
    int n = 0;
    if (p.y1 >= vp.vy) {
        int ya = max(max(p.y0, vp.vy), vp.band0), yb = min(min(p.y1, vp.vy + vp.vh), vp.band1);
        if (yb > ya) n += yb - ya;                          // (not max(0, yb - ya): a limit may be INT_MIN, see f2i)
    }
    if (p.y1 < vp.vy + vp.vh) {
        int ya = max(max(p.y1, vp.vy), vp.band0), yb = min(min(p.y2, vp.vy + vp.vh), vp.band1);
        if (yb > ya) n += yb - ya;                          // y2 = (int)ceil(+huge) is INT_MIN on x86: the reference's loop does not run
    }
    p.n = n;
    return p;
}

// the records of a planned slot whose `plan.n` scanline records start at `base`; a, b, c are y-sorted (plan_slot)
SB_DEV void finish_slot(const DeviceScene &s, const ViewParams &vp, const FrameParams &fp, const Pools &pl,
                        uint32_t slot, uint32_t prim_id, const Prim &pr, const SV &a, const SV &b, const SV &c,
                        bool front_face_visible, const SlotPlan &plan, uint32_t base)
{
    const bool inverted = !front_face_visible;
    const int y0 = plan.y0, y1 = plan.y1, y2 = plan.y2;

    // fill_triangle_2's edge set-up (renderer.cpp:396-459): every side's state at the first scanline it is
    // walked on.  k_spans jumps from here to any scanline with radd().
    SlotEdge e;
    const float x0 = a.s.x, yy0 = a.s.y, z0 = a.s.z, x1 = b.s.x, yy1 = b.s.y, z1 = b.s.z, x2 = c.s.x, yy2 = c.s.y, z2 = c.s.z;
    {
        Interp ip;
        e.lng.ratio = fdiv(fsub(x2, x0), fsub(yy2, yy0));                   // :400-409
        interp_init_self(ip, fsub(yy2, yy0), z0, z2);
        const float move = (y0 < vp.vy) ? fsub((float)vp.vy, yy0) : fsub((float)y0, yy0);
        interp_displace(ip, move);
        e.lng.x = fadd(x0, fmul(e.lng.ratio, move));
        e.lng.top = ip.top; e.lng.topstep = ip.topstep; e.lng.bottom = ip.bottom; e.lng.bottomstep = ip.bottomstep;
    }
    e.y_long = max(y0, vp.vy);
    e.flags = 0;
    e.ya_u = e.yb_u = e.ya_l = e.yb_l = 0;
    e.su = e.lng; e.sl = e.lng;
    if (y1 >= vp.vy) {                                                      // upper half, :416-436
        Interp ip;
        e.su.ratio = fdiv(fsub(x1, x0), fsub(yy1, yy0));
        interp_init_self(ip, fsub(yy1, yy0), z0, z1);
        e.ya_u = max(y0, vp.vy); e.yb_u = min(y1, vp.vy + vp.vh);
        const float move = fsub((float)e.ya_u, yy0);
        interp_displace(ip, move);
        e.su.x = fadd(x0, fmul(e.su.ratio, move));
        e.su.top = ip.top; e.su.topstep = ip.topstep; e.su.bottom = ip.bottom; e.su.bottomstep = ip.bottomstep;
        if (e.lng.ratio > e.su.ratio) e.flags |= 1u;
    }
    if (y1 < vp.vy + vp.vh) {                                               // lower half, :439-459
        Interp ip;
        e.sl.ratio = fdiv(fsub(x2, x1), fsub(yy2, yy1));
        interp_init_self(ip, fsub(yy2, yy1), z1, z2);
        e.ya_l = max(y1, vp.vy); e.yb_l = min(y2, vp.vy + vp.vh);
        const float move = fsub((float)e.ya_l, yy1);
        interp_displace(ip, move);
        e.sl.x = fadd(x1, fmul(e.sl.ratio, move));
        e.sl.top = ip.top; e.sl.topstep = ip.topstep; e.sl.bottom = ip.bottom; e.sl.bottomstep = ip.bottomstep;
        if (e.lng.ratio < e.sl.ratio) e.flags |= 2u;
    }
    e.z0 = z0; e.z1 = z1; e.z2 = z2;
    e.span_base = (int32_t)base; e.pad = 0;
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
    // plain colour shaders never sample the texture: alpha is the material's (pixel_shaders.hpp:28)
    sh.alpha_class = vp.tex_mode == SWEGL_B200_TEX_PLAIN ? ((pr.color >> 24) == 255u ? ALPHA_OPAQUE : ALPHA_UNIFORM) : pr.alpha_class;
    sh.pad0 = 0;
    sh.color = pr.color; sh.tex_off = pr.tex_off; sh.tw = pr.tw; sh.th = pr.th;
    pl.shades[slot] = sh;
}

// plan + allocate + finish by one thread (the near clipper's rare second triangle); the common first slot of a
// triangle goes through the warp-aggregated allocation in k_setup instead
SB_DEV RowRange emit_slot(const DeviceScene &s, const ViewParams &vp, const FrameParams &fp, const Pools &pl,
                          uint32_t slot, uint32_t prim_id, const Prim &pr, SV a, SV b, SV c, bool front_face_visible)
{
    const RowRange none = { 0u, 0u };
    const SlotPlan plan = plan_slot(vp, a, b, c);
    if (plan.n == 0) return none;
    const uint32_t base = atomicAdd(&pl.counters->n_rows, (uint32_t)plan.n);
    if (base + (uint32_t)plan.n > pl.rows_cap) { atomicOr(&pl.counters->overflow, 1u); return none; }
    finish_slot(s, vp, fp, pl, slot, prim_id, pr, a, b, c, front_face_visible, plan, base);
    atomicAdd(&pl.counters->n_slots, 1u);
    const RowRange rr = { base, (uint32_t)plan.n };
    return rr;
}

// row_slot[base .. base+n) = slot (scanline record -> owning slot).  Short ranges are written by their
// own thread; tall triangles are written by the whole warp so no lane runs a thousand-iteration loop.
SB_DEV void fill_row_slots(const Pools &pl, RowRange rr, uint32_t slot)
{
    const int lane = threadIdx.x & 31;
    const bool big = rr.n > 48;
    if (!big) for (uint32_t k = 0; k < rr.n; k++) pl.row_slot[rr.base + k] = slot;
    unsigned m = __ballot_sync(0xFFFFFFFFu, big);
    while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t b = __shfl_sync(0xFFFFFFFFu, rr.base, l), n = __shfl_sync(0xFFFFFFFFu, rr.n, l);
        const uint32_t sl = __shfl_sync(0xFFFFFFFFu, slot, l);
        for (uint32_t k = lane; k < n; k += 32) pl.row_slot[b + k] = sl;
    }
}

#ifndef SETUP_MINB
#define SETUP_MINB 4
#endif
// fill_triangle for triangle t (go == false: an idle lane that only takes part in the warp's collectives)
SB_DEV void setup_one(const DeviceScene &s, const ViewParams &vp, const FrameParams &fp, const Pools &pl, const uint32_t t, bool go)
{
    RowRange r0 = { 0u, 0u }, r1 = { 0u, 0u };
    Tri tr = { 0u, 0u, 0u, 0u };
    if (go) tr = s.tris[t];
    pdl_wait();                                                             // k_mark's yes flags
    if (go) {
        go = s.yes[tr.i0] && s.yes[tr.i1] && s.yes[tr.i2];                  // renderer.cpp:248-253
    }
    if (go && (vp.band0 > vp.vy || vp.band1 < vp.vy + vp.vh)) {
        // a row band of the viewport (sort-first sharding): a triangle that is not near-clipped walks scanlines
        // [ceil(min y), ceil(max y)) only (renderer.cpp:375-394), so one that misses the band is dropped here,
        // before its world positions, normals and texture coordinates are fetched
        const float za = s.v_ndc[3 * tr.i0 + 2], zb = s.v_ndc[3 * tr.i1 + 2], zc = s.v_ndc[3 * tr.i2 + 2];
        if (za >= NEAR_Z && zb >= NEAR_Z && zc >= NEAR_Z) {
            const int ya = ceil_i(fadd(fmul(vp.vp_m11, s.v_ndc[3 * tr.i0 + 1]), vp.vp_m13));   // frustum_to_viewport, as load_sv
            const int yb = ceil_i(fadd(fmul(vp.vp_m11, s.v_ndc[3 * tr.i1 + 1]), vp.vp_m13));
            const int yc = ceil_i(fadd(fmul(vp.vp_m11, s.v_ndc[3 * tr.i2 + 1]), vp.vp_m13));
            if (max(ya, max(yb, yc)) <= vp.band0 || min(ya, min(yb, yc)) >= vp.band1) go = false;
        }
    }
    // the slot most triangles produce (2t) is planned inside the branches and allocated for the whole warp at once
    // below; the near clipper's second triangle (2t+1, rare) allocates on its own
    bool have0 = false, ffv0 = false;
    SV A, B, C;
    Prim pr;
    if (go) {
        pr = s.prims[tr.prim];
        SV a = load_sv(s, vp, tr.i0), b = load_sv(s, vp, tr.i1), c = load_sv(s, vp, tr.i2);

        bool ffv = cross_n(sub(b.s, a.s), sub(c.s, a.s)).z < 0.0f;          // renderer.cpp:258
        bool inverted_order = false;
        if (b.s.z > a.s.z) { swap_sv(a, b); inverted_order = !inverted_order; }   // sort by z DESC, :264-279
        if (c.s.z > b.s.z) { swap_sv(b, c); inverted_order = !inverted_order; }
        if (b.s.z > a.s.z) { swap_sv(a, b); inverted_order = !inverted_order; }

        const float *m9 = s.node_normal + 9 * pr.node;
        if (c.s.z >= NEAR_Z) {
            A = a; B = b; C = c; ffv0 = ffv; have0 = true;
        } else if (b.s.z < NEAR_Z) {
            // only a in front of the camera, renderer.cpp:286-317
            float cut_1 = fdiv(fsub(a.s.z, 0.001f), fsub(a.s.z, b.s.z));
            SV n1 = make_clip_vertex(s, vp, m9, a, b, cut_1);
            float cut_2 = fdiv(fsub(a.s.z, 0.001f), fsub(a.s.z, c.s.z));
            SV n2 = make_clip_vertex(s, vp, m9, a, c, cut_2);
            ffv = cross_n(sub(n1.s, a.s), sub(n2.s, a.s)).z < 0.0f;
            if (inverted_order) ffv = !ffv;
            A = a; B = n1; C = n2; ffv0 = ffv; have0 = true;
        } else if (c.s.z < NEAR_Z) {
            // only c behind the camera: two triangles, renderer.cpp:318-356
            float cut_0 = fdiv(fsub(a.s.z, 0.001f), fsub(a.s.z, c.s.z));
            SV n1 = make_clip_vertex(s, vp, m9, a, c, cut_0);
            float cut_1 = fdiv(fsub(b.s.z, 0.001f), fsub(b.s.z, c.s.z));
            SV n2 = make_clip_vertex(s, vp, m9, b, c, cut_1);
            ffv = cross_n(sub(b.s, a.s), sub(n2.s, a.s)).z < 0.0f;
            if (inverted_order) ffv = !ffv;
            A = a; B = b; C = n2; ffv0 = ffv; have0 = true;
            ffv = cross_n(sub(n2.s, a.s), sub(n1.s, a.s)).z < 0.0f;
            if (inverted_order) ffv = !ffv;
            r1 = emit_slot(s, vp, fp, pl, 2 * t + 1, tr.prim, pr, a, n2, n1, ffv);
        }
    }
    {
        // one n_rows / n_slots atomic per warp instead of one per triangle
        const int lane = threadIdx.x & 31;
        SlotPlan plan = { 0, 0, 0, 0 };
        if (have0) plan = plan_slot(vp, A, B, C);
        const uint32_t n = (uint32_t)plan.n;
        const uint32_t incl = warp_incl_scan(n, lane);
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (tot) {                                                          // warp-uniform
            uint32_t wbase = 0;
            if (lane == 31) wbase = atomicAdd(&pl.counters->n_rows, tot);
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
            const unsigned live = __ballot_sync(0xFFFFFFFFu, n != 0);
            if (wbase + tot > pl.rows_cap) {
                if (lane == 0) atomicOr(&pl.counters->overflow, 1u);
            } else {
                if (lane == 0) atomicAdd(&pl.counters->n_slots, (uint32_t)__popc(live));
                if (n) {
                    const uint32_t base = wbase + incl - n;
                    finish_slot(s, vp, fp, pl, 2 * t, tr.prim, pr, A, B, C, ffv0, plan, base);
                    r0.base = base; r0.n = n;
                }
            }
        }
    }
    fill_row_slots(pl, r0, 2 * t);
    if (__any_sync(0xFFFFFFFFu, r1.n != 0)) fill_row_slots(pl, r1, 2 * t + 1);
}

__global__ void __launch_bounds__(128, SETUP_MINB) k_setup(DeviceScene s, const ViewParams *__restrict__ vpp,
                                               const FrameParams *__restrict__ fpp, Pools pl)
{
    pdl_trigger(s.early_trigger);
    __shared__ ViewParams vp;                   // staged once per CTA: used all over the set-up code
    __shared__ FrameParams fp;
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += 128) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    for (int w = threadIdx.x; w < (int)(sizeof(FrameParams) / 4); w += 128) reinterpret_cast<uint32_t *>(&fp)[w] = reinterpret_cast<const uint32_t *>(fpp)[w];
    __syncthreads();
    // One call site for both cases (set-up is a large body: a second inlined copy costs a small frame ~4 us of
    // instruction fetch): a culled view walks the list of clusters that may hold a triangle reaching the band with
    // persistent CTAs; otherwise the grid covers every cluster once and the loop runs a single pass.
    static_assert(128 % CULL_CL == 0, "whole clusters per CTA pass");
    constexpr uint32_t PER = 128 / CULL_CL;
    const bool listed = s.live_list != nullptr;
    uint32_t n_items = (s.n_tris + CULL_CL - 1) / CULL_CL;
    if (listed) {
        pdl_wait();                                                         // k_mark (and, through the chain, k_cull_live's list)
        n_items = s.cull_counts[0];
    }
    for (uint32_t base = blockIdx.x * PER; base < n_items; base += gridDim.x * PER) {          // CTA-uniform trip count (warp collectives inside)
        const uint32_t it = base + threadIdx.x / CULL_CL;
        uint32_t t = 0;
        bool go = false;
        if (it < n_items) { t = (listed ? s.live_list[it] : it) * CULL_CL + threadIdx.x % CULL_CL; go = t < s.n_tris; }
        setup_one(s, vp, fp, pl, t, go);
    }
}

// ----------------------------------------------------------------------------------------
// spans: one thread per scanline record.  The edges' state on this scanline is the state stored by
// k_setup advanced by (y - first row) steps of `x += ratio; topalpha += topstep; bottomalpha += bottomstep`
// (renderer.cpp:553-556) -- radd() does those steps exactly.  The thread then sets up the per-pixel
// interpolator (renderer.cpp:469-480), replays it along x and drops a checkpoint ("chunk") at every
// 32-column bin the span crosses.
// ----------------------------------------------------------------------------------------

// One SEGMENT (up to SPAN_SEG bins = 64 pixels) of a span per thread.
//  * The thread jumps to its segment's first column with radd() (free for a span's first segment) and replays
//    qpixel.Step() pixel by pixel (renderer.cpp:486); the recurrence is serial in x.  It writes ualpha = topalpha /
//    bottomalpha (progress(), interpolator.hpp:98) of every pixel to the fragment stream, four pixels per 128-bit
//    store: the four divisions are independent (they pipeline), and a warp whose 32 threads write 32 different
//    streams issues a quarter of the memory transactions it would with one store per pixel.  Stream indices are
//    congruent to the column (relative to the viewport) modulo 4, see the allocation in phase B.
//  * One chunk is linked into every 32-column bin crossed.  The atomics that return something (previous bin head,
//    previous tile stamp) are issued side by side, around the pixel walk, so their round trips to L2 overlap.
template <typename SH>
SB_DEV void walk_segment(const Pools &pl, const ViewParams &vp, const SH &sh, int owner, uint32_t sg, uint32_t span_id)
{
    const uint32_t o_nch = sh.nchunks[owner];
    const uint32_t c0 = sg * SPAN_SEG, c1 = min(o_nch, c0 + SPAN_SEG);                          // chunks [c0, c1) of the span
    const int n = (int)(c1 - c0);
    const int o_x1 = sh.x1[owner], o_x2 = sh.x2[owner];
    const int b0 = ((o_x1 - vp.vx) >> 5) + (int)c0;
    int x = c0 == 0 ? o_x1 : vp.vx + (b0 << 5);                             // first column of the segment
    const int xend = min(vp.vx + ((b0 + n) << 5), o_x2);                    // one past its last column
    const int row_rel = sh.y[owner] - vp.vy;
    const size_t bin0 = (size_t)row_rel * vp.nbx + b0;
    const uint32_t cid0 = sh.cbase[owner] + c0;
    int32_t pos[SPAN_SEG];                                                  // arrival order of this piece in its bin
    #pragma unroll
    for (int j = 0; j < (int)SPAN_SEG; j++) pos[j] = j < n ? atomicAdd(&pl.bin_cnt[bin0 + j], 1) : 0;
    const uint32_t steps = (uint32_t)(x - o_x1);
    Interp w;
    w.topstep = sh.topstep[owner]; w.bottomstep = sh.bottomstep[owner];
    w.top = radd(sh.top[owner], w.topstep, steps);
    w.bottom = radd(sh.bottom[owner], w.bottomstep, steps);
    const uint32_t o_fb = sh.fbase[owner];
    // ---- the pixel walk ----
    {
        float *fu = pl.frag_u + o_fb - o_x1;                                // fu[x] = ualpha at column x; (&fu[x] - pool) % 4 == (x - vx) % 4
        for (; x < xend && ((x - vp.vx) & 3); x++) { fu[x] = fdiv(w.top, w.bottom); interp_step(w); }
        for (; x + 4 <= xend; x += 4) {
            const float t0 = w.top, d0 = w.bottom; interp_step(w);
            const float t1 = w.top, d1 = w.bottom; interp_step(w);
            const float t2 = w.top, d2 = w.bottom; interp_step(w);
            const float t3 = w.top, d3 = w.bottom; interp_step(w);
            *reinterpret_cast<float4 *>(&fu[x]) = make_float4(fdiv(t0, d0), fdiv(t1, d1), fdiv(t2, d2), fdiv(t3, d3));
        }
        for (; x < xend; x++) { fu[x] = fdiv(w.top, w.bottom); interp_step(w); }
    }
    // ---- chunk records ----
    Chunk ch;
    ch.span = span_id; ch.v0 = sh.v0[owner]; ch.v1 = sh.v1[owner]; ch.slot = sh.slot[owner]; ch.alpha_class = sh.aclass[owner];
    int xc = sg == 0 ? o_x1 : vp.vx + (b0 << 5);
    #pragma unroll
    for (int j = 0; j < (int)SPAN_SEG; j++) {
        if (j >= n) break;
        const int binx0 = vp.vx + ((b0 + j) << 5);
        const int xn = min(binx0 + 32, o_x2);                               // end of this bin's piece of the span
        ch.frag0 = o_fb + (uint32_t)(binx0 - o_x1);                         // wraps for the span's first bin; lanes < xs never read
        ch.xs_xe = (uint32_t)(xc - binx0) | ((uint32_t)(xn - binx0) << 8);
        ch.next = -1;
        if (pos[j] < BIN_SLOTS) {                                           // the bin's own slots: no list
            pl.bin_slots[(bin0 + j) * BIN_SLOTS + pos[j]] = ch;
        } else {                                                            // a crowded bin: the rest is linked
            ch.next = atomicExch(&pl.bin_head[bin0 + j], (int32_t)(cid0 + j));
            pl.chunks[cid0 + j] = ch;
        }
        xc = xn;
    }
    // first chunk of a bin this frame: note the bin's tile once per frame (Pools::busy_list)
    uint32_t tl[SPAN_SEG], old[SPAN_SEG];
    #pragma unroll
    for (int j = 0; j < (int)SPAN_SEG; j++) {
        tl[j] = 0xFFFFFFFFu; old[j] = 0;
        if (j < n && pos[j] == 0) {
            const uint32_t t = (uint32_t)((row_rel - (vp.band0 - vp.vy)) / FRAG_ROWS) * (uint32_t)vp.ntx + (uint32_t)((b0 + j) / FRAG_STRETCH);
            bool dup = false;
            #pragma unroll
            for (int q = 0; q < j; q++) dup = dup || tl[q] == t;
            if (!dup) { tl[j] = t; old[j] = atomicExch(&pl.tile_stamp[t], vp.stamp); }
        }
    }
    // (the list entry is the tile's row and column, tile row << 16 | tile column: its consumer, one warp per tile row, would
    // otherwise divide by the tiles-per-row count)
    #pragma unroll
    for (int j = 0; j < (int)SPAN_SEG; j++)
        if (tl[j] != 0xFFFFFFFFu && old[j] != vp.stamp) {
            const uint32_t c = pl.tile_cost[tl[j]], m = pl.cost_acc[1];      // last measured cost of this tile, mean of that frame
#ifdef FRAG_NO_COST_ORDER       // A/B: one class, first-touch order
            const uint32_t cls = c + m == 0xFFFFFFFFu ? 1u : 0u;
#else
            const uint32_t cls = c * 2u >= m * 3u ? 0u : (c >= m ? 1u : (c * 2u >= m ? 2u : 3u));
#endif
            pl.busy_list[cls * pl.busy_stride + atomicAdd(&pl.counters->n_busy_b[cls], 1u)] =
                ((uint32_t)((row_rel - (vp.band0 - vp.vy)) / FRAG_ROWS) << 16) | (uint32_t)((b0 + j) / FRAG_STRETCH);
            atomicAdd(&pl.counters->n_busy, 1u);
        }
}

// shared state of one CTA pass over SPAN_ROWS scanline records
#ifndef SPAN_ROWS_V
#define SPAN_ROWS_V 32
#endif
static constexpr int SPAN_ROWS = SPAN_ROWS_V;                 // scanline records per CTA pass (a multiple of 32, <= TPB)
struct SpanCta {
    float val[6][SPAN_ROWS];            // phase A: long x/top/bottom, short x/top/bottom on each scanline
    int x1[SPAN_ROWS], x2[SPAN_ROWS], y[SPAN_ROWS];          // phase B: the spans ...
    float top[SPAN_ROWS], topstep[SPAN_ROWS], bottom[SPAN_ROWS], bottomstep[SPAN_ROWS];   // ... their qpixel at x1 ...
    float v0[SPAN_ROWS], v1[SPAN_ROWS]; uint32_t slot[SPAN_ROWS], aclass[SPAN_ROWS];      // ... depth plane, draw-order key, alpha class ...
    uint32_t nchunks[SPAN_ROWS], cbase[SPAN_ROWS], fbase[SPAN_ROWS], seg_incl[SPAN_ROWS]; // ... and their allocations / segment prefix
    uint32_t warp_segs[(SPAN_ROWS + 31) / 32]; // segments of each 32-row group (phase B warp)
};

#ifndef SPAN_MINB
#define SPAN_MINB 6
#endif
#ifndef SPAN_TPB_V
#define SPAN_TPB_V 256
#endif
// Two widths of the same kernel.  256 threads per 32 scanline records give the shortest critical path (phase A is one
// round, phase C one round) -- but phase B occupies one warp of the eight and phase C a quarter of the threads, and the
// idle ones hold registers and warp slots.  A context that shares its GPU with other contexts' frames
// (swegl_b200_set_shared_gpu) takes the 128-thread width: alone the kernel is slower (truck 4K 31 -> 36 us), four contexts
// together render 6 % more frames per second (profiles/README.md).
#ifndef SPAN_NARROW_TPB
#define SPAN_NARROW_TPB 128
#endif
#ifndef SPAN_NARROW_MINB
#define SPAN_NARROW_MINB 10
#endif
template <int SPAN_TPB, int MINB>
__global__ void __launch_bounds__(SPAN_TPB, MINB) k_spans(const ViewParams *__restrict__ vpp, Pools pl)
{
    __shared__ SpanCta sh;
    __shared__ ViewParams vp;
    pdl_trigger(pl.early_trigger);
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += SPAN_TPB) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    __syncthreads();
    pdl_wait();                                                             // k_setup's records and counters
    if (pl.cull_counts && blockIdx.x == 0 && threadIdx.x < 3) pl.cull_counts[threadIdx.x] = 0;   // the view's cull lists are spent (k_setup was their last reader)
    if (pl.counters->overflow & 1u) return;      // some rows were never allocated: the host grows the pool and redoes the frame
    const uint32_t n_rows = min(pl.counters->n_rows, pl.rows_cap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Span *spans = pl.spans;
    for (uint32_t i0 = blockIdx.x * SPAN_ROWS; i0 < n_rows; i0 += gridDim.x * SPAN_ROWS) {      // CTA-uniform
        // ---- phase A: 6 recurrences x SPAN_ROWS scanlines, one (recurrence, scanline) task per thread and round:
        //      advance it to its row with radd() (renderer.cpp:553-556 applied (y - first walked row) times);
        //      a warp works on one recurrence, so its lanes share the cost profile ----
        for (int task = tid; task < 6 * SPAN_ROWS; task += SPAN_TPB) {
            const int rec = task / SPAN_ROWS, r = task - rec * SPAN_ROWS;
            const uint32_t i = i0 + (uint32_t)r;
            float val = 0.f;
            if (i < n_rows) {
                const uint32_t slot = pl.row_slot[i];
                const SlotEdge &e = pl.edges[slot];
                const int j = (int)i - e.span_base;
                const int ua = max(e.ya_u, vp.band0), ub = min(e.yb_u, vp.band1);
                const int nu = max(0, ub - ua);
                const bool lower = j >= nu;                                 // in-band rows of the upper half come first
                const int y = lower ? max(e.ya_l, vp.band0) + (j - nu) : ua + j;
                const SideRec &S = rec < 3 ? e.lng : (lower ? e.sl : e.su);
                const uint32_t k = rec < 3 ? (uint32_t)(y - e.y_long) : (uint32_t)(y - (lower ? e.ya_l : e.ya_u));
                const int f = rec % 3;
                const float start = f == 0 ? S.x : (f == 1 ? S.top : S.bottom);
                const float step = f == 0 ? S.ratio : (f == 1 ? S.topstep : S.bottomstep);
                val = radd(start, step, k);
            }
            sh.val[rec][r] = val;
        }
        __syncthreads();
        // ---- phase B: warps 0 .. SPAN_ROWS/32-1, lane = scanline: the span (renderer.cpp:469-480), its shading
        //      constants, and one chunk / fragment-stream allocation per warp ----
        if (warp < (SPAN_ROWS + 31) / 32) {
            const int r = warp * 32 + lane;
            const uint32_t i = r < SPAN_ROWS ? i0 + (uint32_t)r : 0xFFFFFFFFu;
            uint32_t nchunks = 0, npix = 0, lead = 0;
            int x1 = 0, x2 = 0, y = 0;
            Interp q; q.top = q.topstep = q.bottom = q.bottomstep = q.v0 = q.v1 = 0.f;
            Span sp; sp.frag_base = 0; sp.pad0 = 0; sp.v0 = sp.v1 = sp.pl = sp.pr = 0.f; sp.x1x2 = 0; sp.slot_flags = 0;
            if (i < n_rows) {
                const uint32_t slot = pl.row_slot[i];
                const SlotEdge &e = pl.edges[slot];
                const int j = (int)i - e.span_base;
                const int ua = max(e.ya_u, vp.band0), ub = min(e.yb_u, vp.band1);
                const int nu = max(0, ub - ua);
                const bool lower = j >= nu;
                y = lower ? max(e.ya_l, vp.band0) + (j - nu) : ua + j;
                const bool lor = lower ? ((e.flags >> 1) & 1u) : (e.flags & 1u);
                const float gx = sh.val[0][r], gtop = sh.val[1][r], gbot = sh.val[2][r];
                const float sx = sh.val[3][r], stop = sh.val[4][r], sbot = sh.val[5][r];
                const float lx = lor ? sx : gx, rx = lor ? gx : sx;
                const float ltop = lor ? stop : gtop, lbot = lor ? sbot : gbot, rtop = lor ? gtop : stop, rbot = lor ? gbot : sbot;

                sp.slot_flags = (slot << 2) | ((uint32_t)lower << 1) | (lor ? 1u : 0u);
                x1 = max(ceil_i(lx), vp.vx);
                x2 = min(ceil_i(rx), vp.vx + vp.vw);
                if (x1 < x2) {
                    const float z0 = e.z0, z1 = e.z1, z2 = e.z2;
                    // edge interpolators' v[0]: long (z0, z2-z0); short (z0, z1-z0) upper / (z1, z2-z1) lower
                    const float la = lor ? (lower ? z1 : z0) : z0;
                    const float lb = lor ? (lower ? z2 : z1) : z2;
                    const float ra = lor ? z0 : (lower ? z1 : z0);
                    const float rb = lor ? z2 : (lower ? z2 : z1);
                    sp.pl = fdiv(ltop, lbot);                               // interpolator progress()
                    sp.pr = fdiv(rtop, rbot);
                    const float zl = fadd(la, fmul(fsub(lb, la), sp.pl));   // value(0), interpolator.hpp:103
                    const float zr = fadd(ra, fmul(fsub(rb, ra), sp.pr));
                    interp_init_self(q, fsub(rx, lx), zl, zr);
                    interp_displace(q, fsub((float)x1, lx));
                    sp.v0 = q.v0; sp.v1 = q.v1;
                    sp.x1x2 = (uint32_t)x1 | ((uint32_t)x2 << 16);
                    nchunks = (uint32_t)(((x2 - 1 - vp.vx) >> 5) - ((x1 - vp.vx) >> 5) + 1);
                    // stream entries: index % 4 == (column - vx) % 4 so that walk_segment can store four at a time
                    lead = (uint32_t)(x1 - vp.vx) & 3u;
                    npix = (lead + (uint32_t)(x2 - x1) + 3u) & ~3u;
                    // prepare_for_{upper,lower}_triangle + prepare_for_scanline of the pixel shaders
                    // (pixel_shaders.cpp:106-158, 320-346), once per span instead of once per pixel.
                    // long side: base v0, dir v2-v0; short side: upper (v0, v1-v0) / lower (v1, v2-v1)
                    const SlotShade &ssh = pl.shades[slot];
                    SpanShade ss;
                    #pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float w0 = ssh.w0[c], w1 = ssh.w1[c], w2 = ssh.w2[c], m0 = ssh.n0[c], m1 = ssh.n1[c], m2 = ssh.n2[c];
                        const float lgd = fsub(w2, w0), shb = lower ? w1 : w0, shd = lower ? fsub(w2, w1) : fsub(w1, w0);
                        const float nlgd = fsub(m2, m0), nshb = lower ? m1 : m0, nshd = lower ? fsub(m2, m1) : fsub(m1, m0);
                        const float vl = lor ? shb : w0, vld = lor ? shd : lgd, vr = lor ? w0 : shb, vrd = lor ? lgd : shd;
                        const float nl = lor ? nshb : m0, nld = lor ? nshd : nlgd, nr = lor ? m0 : nshb, nrd = lor ? nlgd : nshd;
                        ss.v[c] = fadd(vl, fmul(vld, sp.pl));
                        ss.vdir[c] = fsub(fadd(vr, fmul(vrd, sp.pr)), ss.v[c]);
                        ss.n[c] = fadd(nl, fmul(nld, sp.pl));
                        ss.ndir[c] = fsub(fadd(nr, fmul(nrd, sp.pr)), ss.n[c]);
                    }
                    #pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const float a0 = ssh.t0[c], a1 = ssh.t1[c], a2 = ssh.t2[c];
                        const float lgd = fsub(a2, a0), shb = lower ? a1 : a0, shd = lower ? fsub(a2, a1) : fsub(a1, a0);
                        const float tl = lor ? shb : a0, tld = lor ? shd : lgd, tr = lor ? a0 : shb, trd = lor ? lgd : shd;
                        ss.t_left[c] = fadd(tl, fmul(tld, sp.pl));
                        ss.t_dir[c] = fsub(fadd(tr, fmul(trd, sp.pr)), ss.t_left[c]);
                    }
                    pl.span_shades[i] = ss;
                }
            }
            const uint32_t c_incl = warp_incl_scan(nchunks, lane), p_incl = warp_incl_scan(npix, lane);
            const uint32_t c_tot = __shfl_sync(0xFFFFFFFFu, c_incl, 31), p_tot = __shfl_sync(0xFFFFFFFFu, p_incl, 31);
            uint32_t wbase = 0, fbase = 0;
            if (lane == 0 && c_tot) {
                wbase = atomicAdd(&pl.counters->n_chunks, c_tot);
                fbase = atomicAdd(&pl.counters->n_frags, p_tot);
            }
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
            fbase = __shfl_sync(0xFFFFFFFFu, fbase, 0);
            const bool room = wbase + c_tot <= pl.chunks_cap && fbase + p_tot <= pl.frags_cap;
            if (!room && lane == 0)
                atomicOr(&pl.counters->overflow, (wbase + c_tot > pl.chunks_cap ? 2u : 0u) | (fbase + p_tot > pl.frags_cap ? 4u : 0u));
            sp.frag_base = fbase + p_incl - npix + lead;                    // entry of column x1
            if (i < n_rows) spans[i] = sp;
            if (!room) nchunks = 0;                                         // nothing is walked; the host redoes the frame
            const uint32_t nseg = (nchunks + SPAN_SEG - 1) / SPAN_SEG;
            if (r < SPAN_ROWS) {
            sh.x1[r] = x1; sh.x2[r] = x2; sh.y[r] = y;
            sh.top[r] = q.top; sh.topstep[r] = q.topstep; sh.bottom[r] = q.bottom; sh.bottomstep[r] = q.bottomstep;
            sh.v0[r] = sp.v0; sh.v1[r] = sp.v1; sh.slot[r] = sp.slot_flags >> 2;
            sh.aclass[r] = (i < n_rows && nchunks) ? pl.shades[sp.slot_flags >> 2].alpha_class : 0u;
            sh.nchunks[r] = nchunks; sh.cbase[r] = wbase + c_incl - (room ? nchunks : 0u); sh.fbase[r] = sp.frag_base;
            }
            const uint32_t s_incl = warp_incl_scan(nseg, lane);
            if (r < SPAN_ROWS) sh.seg_incl[r] = s_incl;                      // warp-local; made CTA-wide below
            if (lane == 31) sh.warp_segs[warp] = s_incl;
        }
        __syncthreads();
        if (SPAN_ROWS > 32 && tid >= 32 && tid < SPAN_ROWS) {               // add the segment totals of the warps before
            uint32_t before = 0;
            for (int w = 0; w < warp; w++) before += sh.warp_segs[w];
            sh.seg_incl[tid] += before;
        }
        if (SPAN_ROWS > 32) __syncthreads();
        // ---- phase C: thread = SEGMENT (SPAN_SEG bins = 64 pixels) of the group's spans, so a screen-wide span
        //      is walked by many threads (walk_segment) ----
        const uint32_t total = sh.seg_incl[SPAN_ROWS - 1];
        for (uint32_t t = tid; t < total; t += SPAN_TPB) {
            int owner = 0;                                                  // first row whose inclusive prefix exceeds t
            #pragma unroll
            for (int stp = SPAN_ROWS / 2; stp >= 1; stp >>= 1)
                if (sh.seg_incl[owner + stp - 1] <= t) owner += stp;
            const uint32_t sg = t - (sh.seg_incl[owner] - (sh.nchunks[owner] + SPAN_SEG - 1) / SPAN_SEG);   // segment index inside its span
            walk_segment(pl, vp, sh, owner, sg, i0 + (uint32_t)owner);
        }
        __syncthreads();                                                    // sh is reused by the next group
    }
}

// ----------------------------------------------------------------------------------------
// k_spans_dense: the same work with thread = scanline record and 256 records per CTA pass.  Every thread does its
// six radd() jumps itself (cheap when triangles are small: a handful of real additions), the span set-up runs on all
// 256 threads, there is one allocation per CTA pass, and the segment phase again gives long spans to many threads.
// Preferred when scanlines vastly outnumber SMs x warps (small triangles); k_spans (6 warps per 32 rows) keeps the
// critical path short when a few big triangles have long radd() chains.
// ----------------------------------------------------------------------------------------
struct SpanCtaDense {
    int x1[TPB], x2[TPB], y[TPB];
    float top[TPB], topstep[TPB], bottom[TPB], bottomstep[TPB];
    float v0[TPB], v1[TPB]; uint32_t slot[TPB], aclass[TPB];
    uint32_t nchunks[TPB], cbase[TPB], fbase[TPB], seg_incl[TPB];
    uint32_t wsum[3][TPB / 32];          // per-warp totals of the three block scans
    uint32_t base[2];                    // chunk / fragment-stream allocation of this pass
    uint32_t room;
};

__global__ void __launch_bounds__(TPB) k_spans_dense(const ViewParams *__restrict__ vpp, Pools pl)
{
    __shared__ SpanCtaDense sh;
    __shared__ ViewParams vp;
    pdl_trigger(pl.early_trigger);
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += TPB) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    __syncthreads();
    pdl_wait();                                                             // k_setup's records and counters
    if (pl.cull_counts && blockIdx.x == 0 && threadIdx.x < 3) pl.cull_counts[threadIdx.x] = 0;   // the view's cull lists are spent (k_setup was their last reader)
    if (pl.counters->overflow & 1u) return;
    const uint32_t n_rows = min(pl.counters->n_rows, pl.rows_cap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Span *spans = pl.spans;
    for (uint32_t i0 = blockIdx.x * TPB; i0 < n_rows; i0 += gridDim.x * TPB) {                  // CTA-uniform
        const uint32_t i = i0 + tid;
        uint32_t nchunks = 0, npix = 0, lead = 0;
        int x1 = 0, x2 = 0, y = 0;
        Interp q; q.top = q.topstep = q.bottom = q.bottomstep = q.v0 = q.v1 = 0.f;
        Span sp; sp.frag_base = 0; sp.pad0 = 0; sp.v0 = sp.v1 = sp.pl = sp.pr = 0.f; sp.x1x2 = 0; sp.slot_flags = 0;
        if (i < n_rows) {
            const uint32_t slot = pl.row_slot[i];
            const SlotEdge &e = pl.edges[slot];
            const int j = (int)i - e.span_base;
            const int ua = max(e.ya_u, vp.band0), ub = min(e.yb_u, vp.band1);
            const int nu = max(0, ub - ua);
            const bool lower = j >= nu;                                     // in-band rows of the upper half come first
            y = lower ? max(e.ya_l, vp.band0) + (j - nu) : ua + j;
            const bool lor = lower ? ((e.flags >> 1) & 1u) : (e.flags & 1u);
            const SideRec S = lower ? e.sl : e.su;
            const SideRec G = e.lng;
            // renderer.cpp:553-556 applied (y - first walked row) times
            const uint32_t kl = (uint32_t)(y - e.y_long), ks = (uint32_t)(y - (lower ? e.ya_l : e.ya_u));
            const float gx = radd(G.x, G.ratio, kl), gtop = radd(G.top, G.topstep, kl), gbot = radd(G.bottom, G.bottomstep, kl);
            const float sx = radd(S.x, S.ratio, ks), stop = radd(S.top, S.topstep, ks), sbot = radd(S.bottom, S.bottomstep, ks);
            const float lx = lor ? sx : gx, rx = lor ? gx : sx;
            const float ltop = lor ? stop : gtop, lbot = lor ? sbot : gbot, rtop = lor ? gtop : stop, rbot = lor ? gbot : sbot;
            sp.slot_flags = (slot << 2) | ((uint32_t)lower << 1) | (lor ? 1u : 0u);
            x1 = max(ceil_i(lx), vp.vx);                                    // renderer.cpp:469-470
            x2 = min(ceil_i(rx), vp.vx + vp.vw);
            if (x1 < x2) {
                const float z0 = e.z0, z1 = e.z1, z2 = e.z2;
                const float la = lor ? (lower ? z1 : z0) : z0;
                const float lb = lor ? (lower ? z2 : z1) : z2;
                const float ra = lor ? z0 : (lower ? z1 : z0);
                const float rb = lor ? z2 : (lower ? z2 : z1);
                sp.pl = fdiv(ltop, lbot);
                sp.pr = fdiv(rtop, rbot);
                const float zl = fadd(la, fmul(fsub(lb, la), sp.pl));
                const float zr = fadd(ra, fmul(fsub(rb, ra), sp.pr));
                interp_init_self(q, fsub(rx, lx), zl, zr);                  // renderer.cpp:476-480
                interp_displace(q, fsub((float)x1, lx));
                sp.v0 = q.v0; sp.v1 = q.v1;
                sp.x1x2 = (uint32_t)x1 | ((uint32_t)x2 << 16);
                nchunks = (uint32_t)(((x2 - 1 - vp.vx) >> 5) - ((x1 - vp.vx) >> 5) + 1);
                lead = (uint32_t)(x1 - vp.vx) & 3u;                         // stream index % 4 == (column - vx) % 4, as in k_spans
                npix = (lead + (uint32_t)(x2 - x1) + 3u) & ~3u;
                const SlotShade &ssh = pl.shades[slot];                     // prepare_for_scanline, once per span
                SpanShade ss;
                #pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float w0 = ssh.w0[c], w1 = ssh.w1[c], w2 = ssh.w2[c], m0 = ssh.n0[c], m1 = ssh.n1[c], m2 = ssh.n2[c];
                    const float lgd = fsub(w2, w0), shb = lower ? w1 : w0, shd = lower ? fsub(w2, w1) : fsub(w1, w0);
                    const float nlgd = fsub(m2, m0), nshb = lower ? m1 : m0, nshd = lower ? fsub(m2, m1) : fsub(m1, m0);
                    const float vl = lor ? shb : w0, vld = lor ? shd : lgd, vr = lor ? w0 : shb, vrd = lor ? lgd : shd;
                    const float nl = lor ? nshb : m0, nld = lor ? nshd : nlgd, nr = lor ? m0 : nshb, nrd = lor ? nlgd : nshd;
                    ss.v[c] = fadd(vl, fmul(vld, sp.pl));
                    ss.vdir[c] = fsub(fadd(vr, fmul(vrd, sp.pr)), ss.v[c]);
                    ss.n[c] = fadd(nl, fmul(nld, sp.pl));
                    ss.ndir[c] = fsub(fadd(nr, fmul(nrd, sp.pr)), ss.n[c]);
                }
                #pragma unroll
                for (int c = 0; c < 2; c++) {
                    const float a0 = ssh.t0[c], a1 = ssh.t1[c], a2 = ssh.t2[c];
                    const float lgd = fsub(a2, a0), shb = lower ? a1 : a0, shd = lower ? fsub(a2, a1) : fsub(a1, a0);
                    const float tl = lor ? shb : a0, tld = lor ? shd : lgd, tr = lor ? a0 : shb, trd = lor ? lgd : shd;
                    ss.t_left[c] = fadd(tl, fmul(tld, sp.pl));
                    ss.t_dir[c] = fsub(fadd(tr, fmul(trd, sp.pr)), ss.t_left[c]);
                }
                pl.span_shades[i] = ss;
            }
        }
        // ---- block-wide exclusive scans of chunks / pixels / segments; one allocation per pass ----
        const uint32_t nseg = (nchunks + SPAN_SEG - 1) / SPAN_SEG;
        const uint32_t c_incl = warp_incl_scan(nchunks, lane), p_incl = warp_incl_scan(npix, lane), s_incl = warp_incl_scan(nseg, lane);
        if (lane == 31) { sh.wsum[0][warp] = c_incl; sh.wsum[1][warp] = p_incl; sh.wsum[2][warp] = s_incl; }
        __syncthreads();
        uint32_t c_off = 0, p_off = 0, s_off = 0, c_tot = 0, p_tot = 0;
        #pragma unroll
        for (int w = 0; w < TPB / 32; w++) {
            const uint32_t a = sh.wsum[0][w], b = sh.wsum[1][w], c = sh.wsum[2][w];
            if (w < warp) { c_off += a; p_off += b; s_off += c; }
            c_tot += a; p_tot += b;
        }
        if (tid == 0) {
            uint32_t wb = 0, fb = 0;
            if (c_tot) { wb = atomicAdd(&pl.counters->n_chunks, c_tot); fb = atomicAdd(&pl.counters->n_frags, p_tot); }
            const bool room = wb + c_tot <= pl.chunks_cap && fb + p_tot <= pl.frags_cap;
            if (!room) atomicOr(&pl.counters->overflow, (wb + c_tot > pl.chunks_cap ? 2u : 0u) | (fb + p_tot > pl.frags_cap ? 4u : 0u));
            sh.base[0] = wb; sh.base[1] = fb; sh.room = room ? 1u : 0u;
        }
        __syncthreads();
        const bool room = sh.room != 0;
        sp.frag_base = sh.base[1] + p_off + p_incl - npix + lead;
        if (i < n_rows) spans[i] = sp;
        sh.x1[tid] = x1; sh.x2[tid] = x2; sh.y[tid] = y;
        sh.top[tid] = q.top; sh.topstep[tid] = q.topstep; sh.bottom[tid] = q.bottom; sh.bottomstep[tid] = q.bottomstep;
        sh.v0[tid] = sp.v0; sh.v1[tid] = sp.v1; sh.slot[tid] = sp.slot_flags >> 2;
        sh.aclass[tid] = (i < n_rows && nchunks) ? pl.shades[sp.slot_flags >> 2].alpha_class : 0u;
        sh.nchunks[tid] = room ? nchunks : 0u; sh.cbase[tid] = sh.base[0] + c_off + c_incl - nchunks; sh.fbase[tid] = sp.frag_base;
        sh.seg_incl[tid] = room ? s_off + s_incl : 0u;
        __syncthreads();
        // ---- thread = segment of SPAN_SEG bins, as in k_spans ----
        const uint32_t total = sh.seg_incl[TPB - 1];
        for (uint32_t t = tid; t < total; t += TPB) {
            int owner = 0;                                                  // first row whose inclusive prefix exceeds t
            #pragma unroll
            for (int stp = TPB / 2; stp >= 1; stp >>= 1)
                if (sh.seg_incl[owner + stp - 1] <= t) owner += stp;
            const uint32_t sg = t - (sh.seg_incl[owner] - (sh.nchunks[owner] + SPAN_SEG - 1) / SPAN_SEG);
            walk_segment(pl, vp, sh, owner, sg, i0 + (uint32_t)owner);
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------------------
static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

void launch_cull(const DeviceScene &s, const ViewParams *d_vp, const CullTables &ct, cudaStream_t st)
{
    // first kernel of the view: it follows the upload of the parameter block (a copy, not a kernel) -> ordinary launch
    launch_chain(k_cull_live, max(1u, cdiv(ct.n_clusters, 128u)), 128, st, false, s, d_vp, ct);
    launch_chain(k_cull_need, max(1u, cdiv(ct.n_clusters + ct.n_vblocks, 128u)), 128, st, true, ct);
}
void launch_vertex(const DeviceScene &s, const ViewParams *d_vp, Counters *counters, bool with_world, cudaStream_t st)
{
    // culled views walk a list with persistent CTAs (k_vertex): one wave is enough
    const unsigned blocks = s.vert_list ? min(max(1u, cdiv(s.n_vertices, TPB)), 148u * 8u) : max(1u, cdiv(s.n_vertices, TPB));
    // first kernel of the frame unless the view is culled: then it follows the upload of the parameter block (a copy,
    // not a kernel) -> ordinary launch.  A culled view skips vertices, so it can never rely on an earlier view's v_world.
    const bool chained = s.vert_need != nullptr;
    if (with_world || chained) launch_chain(k_vertex<true>, blocks, TPB, st, chained, s, d_vp, counters);
    else launch_chain(k_vertex<false>, blocks, TPB, st, chained, s, d_vp, counters);
}
void launch_mark(const DeviceScene &s, cudaStream_t st)
{
    if (s.n_tris) launch_chain(k_mark, s.mark_list ? min(cdiv(s.n_tris, TPB), 148u * 8u) : cdiv(s.n_tris, TPB), TPB, st, true, s);
}
void launch_setup(const DeviceScene &s, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p, cudaStream_t st)
{
    if (s.n_tris) launch_chain(k_setup, s.live_list ? min(cdiv(s.n_tris, 128u), 148u * SETUP_MINB * 2u) : cdiv(s.n_tris, 128u), 128, st, true, s, d_vp, d_fp, p);
}
void launch_spans(const ViewParams *d_vp, const Pools &p, bool dense, bool narrow, cudaStream_t st)
{
    if (dense) launch_chain(k_spans_dense, 148 * 8, TPB, st, true, d_vp, p);      // persistent CTAs, 256 scanline records per pass
    else if (narrow) launch_chain(k_spans<SPAN_NARROW_TPB, SPAN_NARROW_MINB>, 148 * 16 * (256 / SPAN_NARROW_TPB), SPAN_NARROW_TPB, st, true, d_vp, p);
    else launch_chain(k_spans<SPAN_TPB_V, SPAN_MINB>, 148 * 16 * (256 / SPAN_TPB_V), SPAN_TPB_V, st, true, d_vp, p);   // persistent CTAs, SPAN_ROWS scanline records per pass
}

} // namespace sb
