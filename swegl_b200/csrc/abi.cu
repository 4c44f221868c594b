// abi.cu — the C ABI of include/swegl_b200.h: context, HBM pools, scene upload and the per-frame
// kernel sequence that replaces swegl::render / swegl::_render (renderer.hpp:27-34, renderer.cpp:77-235).
//
// Frame protocol.  Everything that changes per frame or per viewport lives in ONE device block
//     [ FrameParams | ViewParams | node_world 16/node | node_normal 9/node | lights 4/light ]
// mirrored by two pinned staging blocks.  begin_frame() only fills the frame part of a staging block;
// the first render_viewport*() of the frame uploads the whole block with one copy and runs the
// vertex stage with the world transform, later viewports of the same frame upload just their
// ViewParams.  The launch sequence therefore has no per-frame kernel arguments and is replayed as a
// CUDA graph (one per viewport configuration, staging slot and with/without frame upload).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

using namespace sb;

struct swegl_b200_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    bool timing = false;
    std::string err;

    // scene (static)
    bool have_scene = false, have_frame = false;
    DeviceScene ds{};
    uint32_t n_nodes = 0;
    float *d_pos = nullptr, *d_nrm = nullptr, *d_uv = nullptr;
    uint32_t *d_vert_node = nullptr, *d_texels = nullptr;
    Tri *d_tris = nullptr; Prim *d_prims = nullptr;
    float *d_v_world = nullptr, *d_v_ndc = nullptr, *d_n_world = nullptr; uint8_t *d_yes = nullptr;
    bool opaque = true;            // every material and texel has alpha 255
    bool fast_shading = true;      // Phong lighting within +-1 LSB (swegl_b200_set_shading); false: bit-exact
    bool shared_gpu = false;       // other contexts render on this GPU at the same time (swegl_b200_set_shared_gpu): kernels that hold fewer SM resources
    // device-side scene_t::animate (animate.cu): one allocation holding every static table + the scratch matrices
    AnimTables anim{}; uint8_t *d_anim = nullptr; bool have_anim = false;
    bool frame_animated = false;   // the staged frame's node matrices come from k_animate (begin_frame_animated), not from the host
    bool world_complete = false;   // v_world holds this frame's world positions of EVERY vertex (a culled view only fills its blocks)
    // DoF-R source tensor maps (TMA), valid for (dof_tm_w, dof_tm_h) and the current d_tmp_color / d_depth
    CUtensorMap dof_tm_color{}, dof_tm_depth{}; int dof_tm_w = 0, dof_tm_h = 0; bool dof_tm_ok = false; int dof_tma_policy = 1;
    // band culling (common.cuh): static tables + per-view flags; policy -1 automatic (banded views of big scenes), 0 off, 1 on
    CullTables cull{};
    ClusterBox *d_cl_box = nullptr; uint32_t *d_cl_adj_off = nullptr, *d_cl_adj = nullptr, *d_vb_adj_off = nullptr, *d_vb_adj = nullptr;
    uint8_t *d_cull_flags = nullptr;
    uint32_t *d_cull_lists = nullptr;   // live_list (ncl) | mark_list (ncl) | vert_list (nvb) | counts (4)
    int cull_policy = -1;
    uint32_t stamp = 0;            // ViewParams::stamp of the last staged view

    // the frame block
    uint8_t *d_block = nullptr; size_t block_bytes = 0;
    size_t off_vp = 0, off_nw = 0, off_nn = 0, off_lights = 0; uint32_t lights_cap = 0;
    struct Slot {
        uint8_t *block = nullptr; Counters *counters = nullptr;     // pinned
        cudaEvent_t done = nullptr; bool pending = false;           // last launch that read this slot
        uint64_t ticket = 0;                                        // of that launch, when it came through render_viewport_async
    } slots[2];
    std::vector<uint64_t> failed_tickets;                           // async frames found to have overflowed a pool
    int next_slot = 0, frame_slot = 0; bool frame_dirty = false;
    int last_slot = 0;                  // staging slot of the most recently issued view
    // pipelined read-back (render_viewport_async): N_OUT device staging images, filled on `stream`, drained on `copy_stream`.
    // Three, not two: the host learns a frame's bounding box (and so can queue its copy) only when that frame's kernels are
    // done, and the copy of a 4K frame's box takes about as long as the next frame's kernels -- with two images the submit
    // of frame i+2 has to wait for that copy and reaches the GPU just after frame i+1 has ended.
    static constexpr int N_OUT = 3;
    struct OutBuf {
        uint32_t *color = nullptr; float *depth = nullptr; size_t cap = 0;
        cudaEvent_t ready = nullptr, copied = nullptr;
        uint64_t ticket = 0; int slot = 0; bool in_flight = false;
        // the read-back itself is issued once the frame's bounding box is known on the host (issue_readback)
        bool d2h_issued = true, partial_ok = false, dof = false;
        void *pixels = nullptr; int32_t pitch_bytes = 0; float *zbuffer = nullptr; ViewParams vp{};
        // what the staging image holds (k_stage_rect): a complete frame of view `vp` over background `stage_bg`, its box in rec[rec_par]
        StageRect *rec = nullptr; int rec_par = 0; bool stage_valid = false; uint32_t stage_bg = 0;
    } out[N_OUT];
    // partial read-back: what the library last left in each host image (the caller's `pixels`), so that only the
    // rectangle that differs from it crosses PCIe
    struct HostImage {
        void *pixels = nullptr; float *zbuffer = nullptr; int32_t pitch_bytes = 0;
        int32_t vx = 0, vy = 0, vw = 0, vh = 0, band0 = 0, band1 = 0; uint32_t bg = 0;
        int32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;     // viewport-relative rectangle outside of which the image holds `bg` (depth: 0x7F7F7F7F)
        uint64_t age = 0;
    };
    std::vector<HostImage> host_images;
    bool partial_readback = true;
    uint64_t readback_bytes = 0, readback_frames = 0;
    cudaStream_t copy_stream = nullptr;
    uint64_t ticket_seq = 0;
    struct ViewGraph { int32_t key[15]; cudaGraphExec_t exec[2]; };
    std::vector<ViewGraph> view_graphs;
    bool graphs_enabled = true;
    bool dense_spans = false;           // which span kernel the next frames use (see choose_span_kernel)
    int span_policy = -1;               // -1 automatic, 0 always k_spans, 1 always k_spans_dense

    // pools
    Pools pools{};
    uint32_t slots_cap = 0; size_t bins_cap = 0;
    Counters *h_counters = nullptr;     // pinned

    // screen
    int sw = 0, sh = 0;
    uint32_t *d_screen = nullptr; float *d_depth = nullptr; uint32_t *d_tmp_color = nullptr;
    uint32_t *color_target = nullptr;   // where finished colour goes instead of d_screen (another context's / GPU's screen)
    // frame protocol of the band-sharded single frame (FrameSync in common.cuh; swegl_b200_set_frame_sync)
    int sync_rank = -1, sync_world = 0; uint32_t sync_seq = 0;
    FrameSync *own_sync() const { return reinterpret_cast<FrameSync *>(d_screen + (size_t)sw * sh); }
    std::vector<void *> imported;       // cudaIpcOpenMemHandle mappings to close
    std::vector<cudaIpcMemHandle_t> imported_handles;   // the handle of each mapping (a handle is opened once per context)

    ViewParams last_vp{}; bool have_vp = false; bool last_dof = false;
    ViewParams dof_cache{}; float dof_cache_depth = 0.f; bool dof_cache_valid = false;   // DoF thresholds per focal_depth
    cudaEvent_t ev[8]{};

    ViewParams *d_vp() const { return reinterpret_cast<ViewParams *>(d_block + off_vp); }
    FrameParams *d_fp() const { return reinterpret_cast<FrameParams *>(d_block); }
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return SWEGL_B200_ERR_CUDA; } } while (0)

static int fail(swegl_b200_ctx *ctx, int code, const char *msg) { if (ctx) ctx->err = msg; return code; }

// every captured graph bakes in device pointers, grid sizes and the stream: drop them whenever one of those changes
static void drop_graphs(swegl_b200_ctx *ctx)
{
    for (auto &g : ctx->view_graphs) for (auto &e : g.exec) if (e) cudaGraphExecDestroy(e);
    ctx->view_graphs.clear();
}

template <typename T> static cudaError_t dalloc(T *&p, size_t n)
{
    if (p) { cudaFree(p); p = nullptr; }
    return cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T));
}

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// (re)allocate the frame block for the current node count and `lights_cap` point lights
static int layout_block(swegl_b200_ctx *ctx, uint32_t lights_cap)
{
    CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx);
    ctx->lights_cap = lights_cap;
    ctx->off_vp = align16(sizeof(FrameParams));
    ctx->off_nw = ctx->off_vp + align16(sizeof(ViewParams));
    ctx->off_nn = ctx->off_nw + (size_t)64 * ctx->n_nodes;
    ctx->off_lights = align16(ctx->off_nn + (size_t)36 * ctx->n_nodes);
    ctx->block_bytes = align16(ctx->off_lights + (size_t)16 * lights_cap);
    if (ctx->d_block) { cudaFree(ctx->d_block); ctx->d_block = nullptr; }
    CK(cudaMalloc((void **)&ctx->d_block, ctx->block_bytes));
    CK(cudaMemset(ctx->d_block, 0, ctx->block_bytes));
    for (auto &sl : ctx->slots) {
        if (sl.block) cudaFreeHost(sl.block);
        sl.block = nullptr; sl.pending = false;
        CK(cudaMallocHost((void **)&sl.block, ctx->block_bytes));
        memset(sl.block, 0, ctx->block_bytes);
    }
    ctx->ds.node_world = reinterpret_cast<const float *>(ctx->d_block + ctx->off_nw);
    ctx->ds.node_normal = reinterpret_cast<const float *>(ctx->d_block + ctx->off_nn);
    ctx->have_frame = false; ctx->frame_dirty = false;
    return SWEGL_B200_OK;
}

// Static tables of the band culling (common.cuh): clusters of CULL_CL consecutive triangles with the object-space box
// of the vertices they reference, and the two adjacency lists k_cull_need ORs cl_live over:
//   cl_adj(c) = clusters sharing at least one vertex with c (c included)        -> mark_need
//   vb_adj(j) = union of cl_adj(c) over the clusters c referencing block j      -> vert_need
// Lists longer than 32 entries collapse to CULL_ALWAYS (e.g. the hub vertex of a big fan).
static int build_cull_tables(swegl_b200_ctx *ctx, const std::vector<Tri> &tris, const float *pos, const std::vector<uint32_t> &vert_node, uint32_t nv)
{
    const uint32_t nt = (uint32_t)tris.size();
    const uint32_t ncl = (nt + CULL_CL - 1) / CULL_CL, nvb = (nv + CULL_CL - 1) / CULL_CL;
    const size_t LIST_CAP = 32;
    std::vector<ClusterBox> boxes(ncl);
    std::vector<uint64_t> pairs;                    // vertex << 32 | cluster, one per (cluster, referenced vertex)
    pairs.reserve((size_t)ncl * (CULL_CL + 2));
    std::vector<uint32_t> vs;
    for (uint32_t c = 0; c < ncl; c++) {
        vs.clear();
        for (uint32_t t = c * CULL_CL; t < nt && t < (c + 1) * CULL_CL; t++) { vs.push_back(tris[t].i0); vs.push_back(tris[t].i1); vs.push_back(tris[t].i2); }
        std::sort(vs.begin(), vs.end());
        vs.erase(std::unique(vs.begin(), vs.end()), vs.end());
        ClusterBox &b = boxes[c];
        b.node = (int32_t)vert_node[vs[0]]; b.pad = 0;
        for (int k = 0; k < 3; k++) { b.lo[k] = pos[3 * (size_t)vs[0] + k]; b.hi[k] = b.lo[k]; }
        for (uint32_t v : vs) {
            if ((int32_t)vert_node[v] != b.node) b.node = -1;               // more than one node: one box cannot describe it
            for (int k = 0; k < 3; k++) {
                const float x = pos[3 * (size_t)v + k];
                if (!(x >= b.lo[k])) b.lo[k] = x;                           // (a NaN position ends up in the box and keeps the cluster live)
                if (!(x <= b.hi[k])) b.hi[k] = x;
            }
            pairs.push_back(((uint64_t)v << 32) | c);
        }
        if (b.node < 0) b.node = -1;
    }
    std::sort(pairs.begin(), pairs.end());
    std::vector<std::vector<uint32_t>> cl_adj(ncl);
    std::vector<uint8_t> cl_always(ncl, 0);
    for (uint32_t c = 0; c < ncl; c++) cl_adj[c].push_back(c);
    for (size_t a = 0; a < pairs.size();) {
        size_t b = a;
        while (b < pairs.size() && (pairs[b] >> 32) == (pairs[a] >> 32)) b++;
        if (b - a > LIST_CAP) { for (size_t i = a; i < b; i++) cl_always[(uint32_t)pairs[i]] = 1; }
        else if (b - a > 1)
            for (size_t i = a; i < b; i++) for (size_t j = a; j < b; j++) if (i != j) cl_adj[(uint32_t)pairs[i]].push_back((uint32_t)pairs[j]);
        a = b;
    }
    for (uint32_t c = 0; c < ncl; c++) {
        auto &l = cl_adj[c];
        std::sort(l.begin(), l.end()); l.erase(std::unique(l.begin(), l.end()), l.end());
        if (l.size() > LIST_CAP) cl_always[c] = 1;
    }
    std::vector<uint32_t> cl_off(ncl + 1, 0), cl_list, vb_off(nvb + 1, 0), vb_list;
    for (uint32_t c = 0; c < ncl; c++) {
        cl_off[c] = (uint32_t)cl_list.size();
        if (cl_always[c]) cl_list.push_back(CULL_ALWAYS); else cl_list.insert(cl_list.end(), cl_adj[c].begin(), cl_adj[c].end());
    }
    cl_off[ncl] = (uint32_t)cl_list.size();
    {
        size_t a = 0;
        std::vector<uint32_t> l;
        for (uint32_t j = 0; j < nvb; j++) {
            vb_off[j] = (uint32_t)vb_list.size();
            l.clear();
            bool always = false;
            uint32_t last_c = 0xFFFFFFFFu;
            for (; a < pairs.size() && (uint32_t)(pairs[a] >> 32) < (j + 1) * CULL_CL; a++) {
                const uint32_t c = (uint32_t)pairs[a];
                if (c == last_c) continue;
                last_c = c;
                if (cl_always[c]) always = true; else l.insert(l.end(), cl_adj[c].begin(), cl_adj[c].end());
            }
            std::sort(l.begin(), l.end()); l.erase(std::unique(l.begin(), l.end()), l.end());
            if (always || l.size() > LIST_CAP) vb_list.push_back(CULL_ALWAYS); else vb_list.insert(vb_list.end(), l.begin(), l.end());
        }
        vb_off[nvb] = (uint32_t)vb_list.size();
    }
    CK(dalloc(ctx->d_cl_box, (size_t)ncl)); CK(dalloc(ctx->d_cl_adj_off, (size_t)ncl + 1)); CK(dalloc(ctx->d_cl_adj, cl_list.size()));
    CK(dalloc(ctx->d_vb_adj_off, (size_t)nvb + 1)); CK(dalloc(ctx->d_vb_adj, vb_list.size()));
    CK(dalloc(ctx->d_cull_flags, (size_t)2 * ncl + nvb));
    CK(cudaMemcpy(ctx->d_cl_box, boxes.data(), (size_t)ncl * sizeof(ClusterBox), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_cl_adj_off, cl_off.data(), ((size_t)ncl + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_cl_adj, cl_list.data(), cl_list.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_vb_adj_off, vb_off.data(), ((size_t)nvb + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_vb_adj, vb_list.data(), vb_list.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(ctx->d_cull_flags, 1, (size_t)2 * ncl + nvb));
    CK(dalloc(ctx->d_cull_lists, (size_t)2 * ncl + nvb + 4));
    CK(cudaMemset(ctx->d_cull_lists, 0, ((size_t)2 * ncl + nvb + 4) * 4));
    CullTables &ct = ctx->cull;
    ct.n_clusters = ncl; ct.n_vblocks = nvb; ct.boxes = ctx->d_cl_box;
    ct.cl_adj_off = ctx->d_cl_adj_off; ct.cl_adj = ctx->d_cl_adj; ct.vb_adj_off = ctx->d_vb_adj_off; ct.vb_adj = ctx->d_vb_adj;
    ct.cl_live = ctx->d_cull_flags; ct.mark_need = ctx->d_cull_flags + ncl; ct.vert_need = ctx->d_cull_flags + 2 * (size_t)ncl;
    ct.live_list = ctx->d_cull_lists; ct.mark_list = ctx->d_cull_lists + ncl; ct.vert_list = ctx->d_cull_lists + 2 * (size_t)ncl;
    ct.counts = ctx->d_cull_lists + 2 * (size_t)ncl + nvb;
    ctx->pools.cull_counts = ct.counts;
    return SWEGL_B200_OK;
}

// does this view run the band culling?  (a pure function of the view and the context's policy: graph keys stay valid)
static bool view_culled(const swegl_b200_ctx *ctx, const ViewParams &vp)
{
    if (ctx->cull_policy == 0 || ctx->cull.n_clusters == 0) return false;
    const bool banded = vp.band0 > vp.vy || vp.band1 < vp.vy + vp.vh;
    if (!banded) return false;
    return ctx->cull_policy == 1 || ctx->ds.n_tris >= 16384;
}

extern "C" {

int swegl_b200_abi_version(void) { return SWEGL_B200_ABI_VERSION; }

int swegl_b200_create(int device, swegl_b200_ctx **out)
{
    if (!out) return SWEGL_B200_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return SWEGL_B200_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SWEGL_B200_ERR_CUDA;
    if (prop.major != 10) return SWEGL_B200_ERR_CUDA;          // sm_100a code only, no fallback
    if (cudaSetDevice(device) != cudaSuccess) return SWEGL_B200_ERR_CUDA;
    configure_kernels();
    swegl_b200_ctx *ctx = new (std::nothrow) swegl_b200_ctx;
    if (!ctx) return SWEGL_B200_ERR_ARG;
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SWEGL_B200_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    cudaMallocHost((void **)&ctx->h_counters, sizeof(Counters));
    for (auto &sl : ctx->slots) {
        cudaMallocHost((void **)&sl.counters, sizeof(Counters));
        memset(sl.counters, 0, sizeof(Counters));
        cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
    }
    cudaMalloc((void **)&ctx->pools.counters, sizeof(Counters));
    if (const char *e = getenv("SWEGL_B200_SPANS")) {          // "dense" / "coop" pin the span kernel, default: automatic
        ctx->span_policy = !strcmp(e, "dense") ? 1 : (!strcmp(e, "coop") ? 0 : -1);
        ctx->dense_spans = ctx->span_policy == 1;
    }
    if (const char *e = getenv("SWEGL_B200_SHARED_GPU"))      // "1" / "0": the default of swegl_b200_set_shared_gpu
        ctx->shared_gpu = atoi(e) != 0;
    if (const char *e = getenv("SWEGL_B200_SHADING"))         // "exact" pins bit-exact Phong lighting, "fast" the +-1 LSB path (default)
        ctx->fast_shading = strcmp(e, "exact") != 0;
    if (const char *e = getenv("SWEGL_B200_DOF_TMA"))         // "0": stage the DoF windows with plain loads instead of TMA
        ctx->dof_tma_policy = strcmp(e, "0") != 0;
    if (const char *e = getenv("SWEGL_B200_CULL"))            // "0" / "1" pin band culling off / on, default: automatic
        ctx->cull_policy = !strcmp(e, "0") ? 0 : (!strcmp(e, "1") ? 1 : -1);
    *out = ctx;
    return SWEGL_B200_OK;
}

void swegl_b200_destroy(swegl_b200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drop_graphs(ctx);
    void *ptrs[] = { ctx->d_pos, ctx->d_nrm, ctx->d_uv, ctx->d_vert_node, ctx->d_texels, ctx->d_tris, ctx->d_prims,
                     ctx->d_v_world, ctx->d_v_ndc, ctx->d_n_world, ctx->d_yes, ctx->d_block,
                     ctx->pools.edges, ctx->pools.shades, ctx->pools.spans, ctx->pools.span_shades, ctx->pools.frag_u, ctx->pools.row_slot,
                     ctx->pools.chunks, ctx->pools.bin_head, ctx->pools.bin_cnt, ctx->pools.bin_slots, ctx->pools.dof_list, ctx->pools.tile_stamp, ctx->pools.busy_list, ctx->pools.tile_cost,
                     ctx->pools.counters, ctx->d_screen, ctx->d_depth,
                     ctx->d_tmp_color, ctx->d_cl_box, ctx->d_cl_adj_off, ctx->d_cl_adj, ctx->d_vb_adj_off, ctx->d_vb_adj, ctx->d_cull_flags, ctx->d_cull_lists, ctx->d_anim };
    for (void *p : ptrs) if (p) cudaFree(p);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    for (auto &sl : ctx->slots) {
        if (sl.block) cudaFreeHost(sl.block);
        if (sl.counters) cudaFreeHost(sl.counters);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    for (void *m : ctx->imported) cudaIpcCloseMemHandle(m);
    for (auto &ob : ctx->out) {
        if (ob.color) cudaFree(ob.color);
        if (ob.depth) cudaFree(ob.depth);
        if (ob.rec) cudaFree(ob.rec);
        if (ob.ready) cudaEventDestroy(ob.ready);
        if (ob.copied) cudaEventDestroy(ob.copied);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *swegl_b200_last_error(const swegl_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int swegl_b200_set_stream(swegl_b200_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return SWEGL_B200_ERR_ARG;
    cudaStreamSynchronize(ctx->stream);
    drop_graphs(ctx);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SWEGL_B200_OK;
}

int swegl_b200_alloc_host(size_t bytes, void **out)
{
    if (!out) return SWEGL_B200_ERR_ARG;
    return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? SWEGL_B200_OK : SWEGL_B200_ERR_CUDA;
}

int swegl_b200_free_host(void *p)
{
    return cudaFreeHost(p) == cudaSuccess ? SWEGL_B200_OK : SWEGL_B200_ERR_CUDA;
}

int swegl_b200_set_timing(swegl_b200_ctx *ctx, int enabled)
{
    if (!ctx) return SWEGL_B200_ERR_ARG;
    ctx->timing = enabled != 0;
    return SWEGL_B200_OK;
}

static int ensure_pools(swegl_b200_ctx *ctx, uint32_t rows_cap, uint32_t chunks_cap, uint32_t frags_cap)
{
    if (frags_cap > ctx->pools.frags_cap || rows_cap > ctx->pools.rows_cap || chunks_cap > ctx->pools.chunks_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        drop_graphs(ctx);
    }
    if (frags_cap > ctx->pools.frags_cap) {
        CK(dalloc(ctx->pools.frag_u, (size_t)frags_cap));
        // entry 0 is what k_fragments' lanes OUTSIDE a piece load (and ignore) instead of branching around the load; the
        // stream's alignment rule may leave it unwritten (compute-sanitizer initcheck)
        CK(cudaMemset(ctx->pools.frag_u, 0, 64));
        ctx->pools.frags_cap = frags_cap;
    }
    if (rows_cap > ctx->pools.rows_cap) {
        CK(dalloc(ctx->pools.spans, (size_t)rows_cap));
        CK(dalloc(ctx->pools.span_shades, (size_t)rows_cap));
        CK(dalloc(ctx->pools.row_slot, (size_t)rows_cap));
        ctx->pools.rows_cap = rows_cap;
    }
    if (chunks_cap > ctx->pools.chunks_cap) {
        CK(dalloc(ctx->pools.chunks, (size_t)chunks_cap));
        ctx->pools.chunks_cap = chunks_cap;
    }
    return SWEGL_B200_OK;
}

static int grow_pools_for(swegl_b200_ctx *ctx, const Counters &c)
{
    uint64_t want_rows = (uint64_t)c.n_rows + c.n_rows / 4 + 1024, want_chunks = (uint64_t)c.n_chunks + c.n_chunks / 4 + 1024;
    uint64_t want_frags = (uint64_t)c.n_frags + c.n_frags / 4 + 1024;
    // when rows overflowed, k_spans did not run: chunk / fragment demand is unknown, so at least double them
    if (c.overflow & 1u) want_chunks = want_chunks > 2 * want_rows ? want_chunks : 2 * want_rows;
    if (c.overflow & 3u) want_frags = want_frags > 2 * (uint64_t)ctx->pools.frags_cap ? want_frags : 2 * (uint64_t)ctx->pools.frags_cap;
    if (want_rows > 0xFFFFFFF0ull || want_chunks > 0x7FFFFFF0ull || want_frags > 0xFFFFFFF0ull)
        return fail(ctx, SWEGL_B200_ERR_CAPACITY, "frame needs more than 2^32 fragments / 2^31 chunks");
    // bin lists were consumed by k_fragments except for chunks that never got linked: reset them all
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemsetAsync(ctx->pools.bin_head, 0xFF, ctx->bins_cap * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->pools.bin_cnt, 0, ctx->bins_cap * 4, ctx->stream));
    return ensure_pools(ctx, (uint32_t)want_rows, (uint32_t)want_chunks, (uint32_t)want_frags);
}

// k_spans (6 warps x 32 rows) has the shorter critical path when a few big triangles need long radd() chains;
// k_spans_dense (thread = row) has ~10x the throughput when there are very many scanlines.  Decide from the
// scanline count of the most recent frame whose counters are known.
static void choose_span_kernel(swegl_b200_ctx *ctx, const Counters &c)
{
    if (ctx->span_policy >= 0) return;
    const bool dense = c.n_rows >= 150000u;
    if (dense != ctx->dense_spans) ctx->dense_spans = dense;      // (the graph key carries the choice)
}

// an asynchronous frame that ran out of pool space is incomplete: enlarge the pools so re-issuing it succeeds
static int check_slot_overflow(swegl_b200_ctx *ctx, swegl_b200_ctx::Slot &sl)
{
    if (!sl.pending) return SWEGL_B200_OK;
    sl.pending = false;
    choose_span_kernel(ctx, *sl.counters);
    if (!sl.counters->overflow) return SWEGL_B200_OK;
    Counters c = *sl.counters;
    sl.counters->overflow = 0;
    if (sl.ticket) { if (ctx->failed_tickets.size() >= 16) ctx->failed_tickets.erase(ctx->failed_tickets.begin()); ctx->failed_tickets.push_back(sl.ticket); }
    int rc = grow_pools_for(ctx, c);
    if (rc) return rc;
    return fail(ctx, SWEGL_B200_ERR_CAPACITY, "an asynchronous frame overflowed the span/chunk/fragment pools (now enlarged): render it again");
}

int swegl_b200_synchronize(swegl_b200_ctx *ctx)
{
    if (!ctx) return SWEGL_B200_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = SWEGL_B200_OK;
    for (auto &sl : ctx->slots) { int r = check_slot_overflow(ctx, sl); if (r) rc = r; }
    return rc;
}

// wait until the launch that last read this staging slot has finished, and act on its pool overflow
static int acquire_slot(swegl_b200_ctx *ctx, swegl_b200_ctx::Slot &sl)
{
    if (sl.pending) {
        CK(cudaEventSynchronize(sl.done));
        int rc = check_slot_overflow(ctx, sl);
        if (rc == SWEGL_B200_ERR_CAPACITY) ctx->err.clear();     // pools were enlarged; later frames are fine
        else if (rc) return rc;
    }
    return SWEGL_B200_OK;
}

int swegl_b200_upload_scene(swegl_b200_ctx *ctx, const swegl_b200_scene_desc *sc)
{
    if (!ctx || !sc) return SWEGL_B200_ERR_ARG;
    if ((sc->n_vertices && (!sc->positions || !sc->normals || !sc->texcoords)) || (sc->n_primitives && !sc->primitives)
        || (sc->n_indices && !sc->indices) || (sc->n_materials && !sc->materials) || (sc->n_textures && !sc->textures))
        return fail(ctx, SWEGL_B200_ERR_ARG, "upload_scene: null array");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx);

    // texel pool: [textures..., one 1x1 colour per material, the default material's colour]
    std::vector<uint32_t> tex_off(sc->n_textures);
    size_t n_texels = 0;
    bool opaque = true;
    for (uint32_t t = 0; t < sc->n_textures; t++) {
        const auto &tx = sc->textures[t];
        if (!tx.texels || tx.width <= 0 || tx.height <= 0) return fail(ctx, SWEGL_B200_ERR_ARG, "upload_scene: bad texture");
        tex_off[t] = (uint32_t)n_texels;
        n_texels += (size_t)tx.width * tx.height;
    }
    std::vector<uint32_t> texels(n_texels + sc->n_materials + 1);
    std::vector<uint8_t> tex_opaque(sc->n_textures, 1);
    for (uint32_t t = 0; t < sc->n_textures; t++) {
        const auto &tx = sc->textures[t];
        size_t n = (size_t)tx.width * tx.height;
        memcpy(&texels[tex_off[t]], tx.texels, n * 4);
        for (size_t i = 0; i < n; i++) if ((tx.texels[i] >> 24) != 255) { tex_opaque[t] = 0; opaque = false; break; }
    }
    auto mat_color = [](const swegl_b200_material &m) {
        return (uint32_t)m.b | ((uint32_t)m.g << 8) | ((uint32_t)m.r << 16) | ((uint32_t)m.a << 24);
    };
    for (uint32_t m = 0; m < sc->n_materials; m++) {
        texels[n_texels + m] = mat_color(sc->materials[m]);
        if (sc->materials[m].a != 255) opaque = false;
        if (sc->materials[m].texture_idx >= (int32_t)sc->n_textures || sc->materials[m].texture_idx < -1)
            return fail(ctx, SWEGL_B200_ERR_ARG, "upload_scene: texture_idx out of range");
    }
    texels[n_texels + sc->n_materials] = mat_color(sc->default_material);

    // primitives and the expanded triangle list (renderer.cpp:197-229 argument order)
    std::vector<Prim> prims(sc->n_primitives);
    std::vector<uint32_t> vert_node(sc->n_vertices, 0);
    std::vector<Tri> tris;
    for (uint32_t p = 0; p < sc->n_primitives; p++) {
        const auto &sp = sc->primitives[p];
        if (sp.node < 0 || (uint32_t)sp.node >= sc->n_nodes || sp.material_id < -1 || sp.material_id >= (int32_t)sc->n_materials
            || (uint64_t)sp.first_vertex + sp.n_vertices > sc->n_vertices || (uint64_t)sp.first_index + sp.n_indices > sc->n_indices)
            return fail(ctx, SWEGL_B200_ERR_ARG, "upload_scene: primitive out of range");
        Prim &d = prims[p];
        const swegl_b200_material &m = sp.material_id >= 0 ? sc->materials[sp.material_id] : sc->default_material;
        if (sp.material_id == -1 && sc->default_material.a != 255) opaque = false;
        d.color = mat_color(m);
        d.node = sp.node;
        d.double_sided = (sp.material_id != -1 && m.double_sided) ? 1 : 0;          // renderer.cpp:90
        if (sp.material_id == -1 || m.texture_idx == -1) {                           // pixel_shaders.cpp:288-294
            d.tex_off = (uint32_t)(n_texels + (sp.material_id >= 0 ? (uint32_t)sp.material_id : sc->n_materials));
            d.tw = 1; d.th = 1;
            d.alpha_class = m.a == 255 ? ALPHA_OPAQUE : ALPHA_UNIFORM;               // a filtered 1x1 texel keeps alpha 255 / < 255
        } else {
            d.tex_off = tex_off[m.texture_idx];
            d.tw = sc->textures[m.texture_idx].width; d.th = sc->textures[m.texture_idx].height;
            d.alpha_class = tex_opaque[m.texture_idx] ? ALPHA_OPAQUE : ALPHA_PER_FRAGMENT;
        }
        d.tw_mask = (d.tw & (d.tw - 1)) == 0 ? d.tw - 1 : -1;
        d.th_mask = (d.th & (d.th - 1)) == 0 ? d.th - 1 : -1;
        for (uint32_t k = 0; k < sp.n_vertices; k++) vert_node[sp.first_vertex + k] = (uint32_t)sp.node;
        const uint32_t *I = sc->indices + sp.first_index;
        auto idx = [&](uint32_t i) -> uint32_t { return sp.first_vertex + I[i]; };
        for (uint32_t i = 0; i < sp.n_indices; i++)
            if (I[i] >= sp.n_vertices) return fail(ctx, SWEGL_B200_ERR_ARG, "upload_scene: index out of range");
        if (sp.mode == SWEGL_B200_MODE_TRIANGLE_STRIP)
            for (uint32_t i = 2; i < sp.n_indices; i++) tris.push_back(Tri{ idx(i - 2), idx(i - 1 + (i & 1)), idx(i - (i & 1)), p });
        else if (sp.mode == SWEGL_B200_MODE_TRIANGLE_FAN)
            for (uint32_t i = 2; i < sp.n_indices; i++) tris.push_back(Tri{ idx(0), idx(i - 1), idx(i), p });
        else if (sp.mode == SWEGL_B200_MODE_TRIANGLES)
            for (uint32_t i = 2; i < sp.n_indices; i += 3) tris.push_back(Tri{ idx(i - 2), idx(i - 1), idx(i), p });
        // other modes (points, lines) draw nothing in the reference either
    }
    if (tris.size() >= (1u << 29)) return fail(ctx, SWEGL_B200_ERR_UNSUPPORTED, "upload_scene: more than 2^29 triangles");

    const uint32_t nv = sc->n_vertices, nt = (uint32_t)tris.size();
    CK(dalloc(ctx->d_pos, (size_t)3 * nv)); CK(dalloc(ctx->d_nrm, (size_t)3 * nv)); CK(dalloc(ctx->d_uv, (size_t)2 * nv));
    CK(dalloc(ctx->d_vert_node, (size_t)nv)); CK(dalloc(ctx->d_texels, texels.size()));
    CK(dalloc(ctx->d_tris, (size_t)nt)); CK(dalloc(ctx->d_prims, (size_t)sc->n_primitives));
    CK(dalloc(ctx->d_v_world, (size_t)3 * nv)); CK(dalloc(ctx->d_v_ndc, (size_t)3 * nv)); CK(dalloc(ctx->d_n_world, (size_t)3 * nv));
    CK(dalloc(ctx->d_yes, (size_t)nv));
    CK(cudaMemcpy(ctx->d_pos, sc->positions, (size_t)12 * nv, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_nrm, sc->normals, (size_t)12 * nv, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_uv, sc->texcoords, (size_t)8 * nv, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_vert_node, vert_node.data(), (size_t)4 * nv, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_texels, texels.data(), texels.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_tris, tris.data(), (size_t)nt * sizeof(Tri), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_prims, prims.data(), prims.size() * sizeof(Prim), cudaMemcpyHostToDevice));
    CK(cudaMemset(ctx->d_yes, 0, nv ? nv : 1));

    // per-slot records: 2 slots per triangle, addressed by slot id (sparse; only live slots are touched)
    ctx->slots_cap = 2 * nt;
    CK(dalloc(ctx->pools.edges, (size_t)ctx->slots_cap)); CK(dalloc(ctx->pools.shades, (size_t)ctx->slots_cap));
    // initial pool sizes in 64 bit (nt < 2^29 is accepted above), clamped to what the 32-bit record indices can address
    const uint64_t rows64 = std::max<uint64_t>((uint64_t)nt * 8u, 1u << 20);
    const uint32_t rows0 = (uint32_t)std::min<uint64_t>(rows64, 0xFFFFFFF0ull);
    const uint32_t chunks0 = (uint32_t)std::min<uint64_t>(rows64 * 2, 0x7FFFFFF0ull);
    int rc = ensure_pools(ctx, rows0, chunks0, 1u << 24);
    if (rc) return rc;

    DeviceScene &ds = ctx->ds;
    ds.n_vertices = nv; ds.n_tris = nt; ds.n_prims = sc->n_primitives; ds.n_nodes = sc->n_nodes;
    ds.pos = ctx->d_pos; ds.nrm = ctx->d_nrm; ds.uv = ctx->d_uv; ds.vert_node = ctx->d_vert_node;
    ds.tris = ctx->d_tris; ds.prims = ctx->d_prims; ds.texels = ctx->d_texels;
    ds.v_world = ctx->d_v_world; ds.v_ndc = ctx->d_v_ndc; ds.n_world = ctx->d_n_world; ds.yes = ctx->d_yes;
    ds.cl_live = ds.mark_need = ds.vert_need = nullptr;
    ds.live_list = ds.mark_list = ds.vert_list = ds.cull_counts = nullptr;
    ctx->cull = CullTables{};
    if (nt && ctx->cull_policy != 0) { rc = build_cull_tables(ctx, tris, sc->positions, vert_node, nv); if (rc) return rc; }
    ctx->n_nodes = sc->n_nodes;
    ctx->have_anim = false; ctx->frame_animated = false;    // the animation tables belong to the previous scene's nodes
    rc = layout_block(ctx, ctx->lights_cap ? ctx->lights_cap : 8);
    if (rc) return rc;
    ctx->opaque = opaque;
    ctx->have_scene = true; ctx->have_frame = false; ctx->have_vp = false;
    return SWEGL_B200_OK;
}

int swegl_b200_set_screen(swegl_b200_ctx *ctx, int32_t w, int32_t h)
{
    if (!ctx || w <= 0 || h <= 0 || w > 65535 || h > 65535) return fail(ctx, SWEGL_B200_ERR_ARG, "set_screen: bad size");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx);
    size_t n = (size_t)w * h;
    CK(dalloc(ctx->d_screen, n + SYNC_WORDS)); CK(dalloc(ctx->d_depth, n)); CK(dalloc(ctx->d_tmp_color, n));   // + the FrameSync block
    CK(cudaMemset(ctx->d_screen, 0, (n + SYNC_WORDS) * 4));
    ctx->sync_rank = -1;
    CK(cudaMemset(ctx->d_depth, 0x7F, n * 4));
    CK(cudaMemset(ctx->d_tmp_color, 0, n * 4));
    size_t bins = (size_t)((w + 31) / 32 + 1) * h;
    CK(dalloc(ctx->pools.bin_head, bins));
    CK(cudaMemset(ctx->pools.bin_head, 0xFF, bins * 4));
    CK(dalloc(ctx->pools.bin_cnt, bins));
    CK(cudaMemset(ctx->pools.bin_cnt, 0, bins * 4));
    CK(dalloc(ctx->pools.bin_slots, bins * BIN_SLOTS));     // 1 KB per 32-pixel bin: 265 MB at 4K, 1.06 GB at 8K (only the used slots are ever touched)

    CK(dalloc(ctx->pools.dof_list, bins));              // (a DoF tile is more than one bin)
    ctx->dof_tm_w = ctx->dof_tm_h = 0;                  // d_tmp_color / d_depth moved: the tensor maps are stale
    CK(dalloc(ctx->pools.tile_stamp, bins));            // (a tile is at least one bin)
    CK(cudaMemset(ctx->pools.tile_stamp, 0, bins * 4));
    CK(dalloc(ctx->pools.busy_list, 4 * bins));         // four cost classes, each able to hold every tile
    ctx->pools.busy_stride = (uint32_t)bins;
    CK(dalloc(ctx->pools.tile_cost, bins + 2));         // + the running sum and the last frame's mean
    CK(cudaMemset(ctx->pools.tile_cost, 0, (bins + 2) * 4));
    ctx->pools.cost_acc = ctx->pools.tile_cost + bins;
    ctx->bins_cap = bins;
    ctx->sw = w; ctx->sh = h;
    ctx->color_target = nullptr;
    return SWEGL_B200_OK;
}

// stage one frame's data in a pinned slot: lights always; node matrices from the caller, or (animated) only the time stamp
static int stage_frame(swegl_b200_ctx *ctx, const swegl_b200_frame_desc *fr, bool animated, float elapsed_seconds)
{
    if (!ctx->have_scene) return fail(ctx, SWEGL_B200_ERR_STATE, "begin_frame before upload_scene");
    if ((!animated && ctx->n_nodes && (!fr->node_world || !fr->node_normal)) || (fr->n_point_lights && !fr->point_lights))
        return fail(ctx, SWEGL_B200_ERR_ARG, "begin_frame: null array");
    CK(cudaSetDevice(ctx->device));
    if (fr->n_point_lights > ctx->lights_cap) {
        int rc = layout_block(ctx, fr->n_point_lights + 8);
        if (rc) return rc;
    }
    // stage the caller's arrays in a pinned block: no device work here, the caller may reuse its buffers at once,
    // and the first render_viewport of the frame uploads the block and computes v_world
    const int si = ctx->frame_dirty ? ctx->frame_slot : ctx->next_slot;
    auto &sl = ctx->slots[si];
    int rc = acquire_slot(ctx, sl);
    if (rc) return rc;
    FrameParams fp{};
    fp.ambient = fr->ambient;
    fp.sun[0] = fr->sun_dir[0]; fp.sun[1] = fr->sun_dir[1]; fp.sun[2] = fr->sun_dir[2];
    fp.sun_intensity = fr->sun_intensity;
    fp.n_lights = fr->n_point_lights;
    fp.lights = reinterpret_cast<const float4 *>(ctx->d_block + ctx->off_lights);
    fp.anim_time = elapsed_seconds;
    memcpy(sl.block, &fp, sizeof fp);
    if (!animated) {
        memcpy(sl.block + ctx->off_nw, fr->node_world, (size_t)64 * ctx->n_nodes);
        memcpy(sl.block + ctx->off_nn, fr->node_normal, (size_t)36 * ctx->n_nodes);
    }
    if (fr->n_point_lights) memcpy(sl.block + ctx->off_lights, fr->point_lights, (size_t)16 * fr->n_point_lights);
    ctx->frame_slot = si; ctx->frame_dirty = true; ctx->have_frame = true;
    ctx->frame_animated = animated;
    ctx->world_complete = false;
    return SWEGL_B200_OK;
}

int swegl_b200_begin_frame(swegl_b200_ctx *ctx, const swegl_b200_frame_desc *fr)
{
    if (!ctx || !fr) return SWEGL_B200_ERR_ARG;
    return stage_frame(ctx, fr, false, 0.0f);
}

int swegl_b200_begin_frame_animated(swegl_b200_ctx *ctx, float elapsed_seconds, const swegl_b200_frame_desc *fr)
{
    if (!ctx || !fr) return SWEGL_B200_ERR_ARG;
    if (!ctx->have_anim) return fail(ctx, SWEGL_B200_ERR_STATE, "begin_frame_animated before set_animation");
    return stage_frame(ctx, fr, true, elapsed_seconds);
}

// Upload what scene_t::animate and the hierarchy product need (animate.cu).  Everything is validated here so that the
// kernel can index without checks.
int swegl_b200_set_animation(swegl_b200_ctx *ctx, const swegl_b200_animation_desc *an)
{
    if (!ctx || !an) return SWEGL_B200_ERR_ARG;
    if (!ctx->have_scene) return fail(ctx, SWEGL_B200_ERR_STATE, "set_animation before upload_scene");
    const uint32_t n = ctx->n_nodes;
    if (an->n_nodes != n) return fail(ctx, SWEGL_B200_ERR_ARG, "set_animation: n_nodes differs from the uploaded scene's");
    if (n && (!an->node_parent || !an->node_rotation || !an->node_translation || !an->node_scale))
        return fail(ctx, SWEGL_B200_ERR_ARG, "set_animation: null node array");
    if ((an->n_channels && (!an->channels || !an->n_animations)) || (an->n_animations && !an->end_time) || (an->n_steps && (!an->step_time || !an->step_value)))
        return fail(ctx, SWEGL_B200_ERR_ARG, "set_animation: null animation array");
    // hierarchy levels (vertex_shaders.hpp:16-33 recurses from the roots; a node's matrix needs its parent's)
    std::vector<int32_t> level(n, -1);
    uint32_t n_levels = 0;
    for (uint32_t i = 0; i < n; i++)
        if (an->node_parent[i] < -1 || an->node_parent[i] >= (int32_t)n) return fail(ctx, SWEGL_B200_ERR_ARG, "set_animation: node_parent out of range");
    for (uint32_t i = 0; i < n; i++) {
        uint32_t depth = 0; int32_t j = (int32_t)i;
        while (j >= 0 && level[j] < 0) {
            if (++depth > n) return fail(ctx, SWEGL_B200_ERR_ARG, "set_animation: node_parent forms a cycle");
            j = an->node_parent[j];
        }
        // walk down again assigning levels
        int32_t base = j < 0 ? -1 : level[j];
        std::vector<int32_t> chain;
        for (int32_t k = (int32_t)i; k >= 0 && level[k] < 0; k = an->node_parent[k]) chain.push_back(k);
        for (size_t c = chain.size(); c-- > 0;) level[chain[c]] = ++base;
    }
    for (uint32_t i = 0; i < n; i++) n_levels = std::max(n_levels, (uint32_t)level[i] + 1);
    std::vector<uint32_t> level_off(n_levels + 1, 0), order(n);
    for (uint32_t i = 0; i < n; i++) level_off[level[i] + 1]++;
    for (uint32_t l = 0; l < n_levels; l++) level_off[l + 1] += level_off[l];
    { std::vector<uint32_t> fill(level_off.begin(), level_off.end() - 1); for (uint32_t i = 0; i < n; i++) order[fill[level[i]]++] = i; }
    // channels, grouped by node in scene order (animation by animation, channel by channel: model.hpp:148-153)
    std::vector<uint32_t> chan_off(n + 1, 0);
    for (uint32_t c = 0; c < an->n_channels; c++) {
        const auto &ch = an->channels[c];
        if (ch.node < 0 || (uint32_t)ch.node >= n || ch.animation < 0 || (uint32_t)ch.animation >= an->n_animations || ch.n_steps == 0
            || (uint64_t)ch.first_step + ch.n_steps > an->n_steps)
            return fail(ctx, SWEGL_B200_ERR_ARG, "set_animation: channel out of range (node, animation or key frames; a channel needs at least one key frame)");
        chan_off[ch.node + 1]++;
    }
    for (uint32_t i = 0; i < n; i++) chan_off[i + 1] += chan_off[i];
    std::vector<AnimChannel> chans(an->n_channels);
    { std::vector<uint32_t> fill(chan_off.begin(), chan_off.end() - 1);
      for (uint32_t c = 0; c < an->n_channels; c++) {
          const auto &ch = an->channels[c];
          chans[fill[ch.node]++] = AnimChannel{ ch.path, ch.first_step, ch.n_steps, an->end_time[ch.animation] };
      } }
    // one device allocation, 16-byte aligned pieces
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx);
    size_t off = 0;
    auto piece = [&](size_t bytes) { const size_t o = off; off = align16(off + bytes); return o; };
    const size_t o_rot = piece((size_t)64 * n), o_tr = piece((size_t)12 * n), o_sc = piece((size_t)12 * n), o_par = piece((size_t)4 * n),
                 o_ord = piece((size_t)4 * n), o_lvl = piece((size_t)4 * (n_levels + 1)), o_cho = piece((size_t)4 * (n + 1)),
                 o_ch = piece(sizeof(AnimChannel) * chans.size()), o_st = piece((size_t)4 * an->n_steps), o_sv = piece((size_t)16 * an->n_steps),
                 o_loc = piece((size_t)64 * n);
    std::vector<uint8_t> host(off, 0);
    if (n) {
        memcpy(&host[o_rot], an->node_rotation, (size_t)64 * n); memcpy(&host[o_tr], an->node_translation, (size_t)12 * n);
        memcpy(&host[o_sc], an->node_scale, (size_t)12 * n); memcpy(&host[o_par], an->node_parent, (size_t)4 * n);
        memcpy(&host[o_ord], order.data(), (size_t)4 * n);
    }
    memcpy(&host[o_lvl], level_off.data(), 4 * level_off.size());
    memcpy(&host[o_cho], chan_off.data(), 4 * chan_off.size());
    if (!chans.empty()) memcpy(&host[o_ch], chans.data(), sizeof(AnimChannel) * chans.size());
    if (an->n_steps) { memcpy(&host[o_st], an->step_time, (size_t)4 * an->n_steps); memcpy(&host[o_sv], an->step_value, (size_t)16 * an->n_steps); }
    CK(dalloc(ctx->d_anim, off));
    CK(cudaMemcpy(ctx->d_anim, host.data(), off, cudaMemcpyHostToDevice));
    AnimTables &a = ctx->anim;
    a.n_nodes = n; a.n_levels = n_levels;
    a.base_rotation = reinterpret_cast<const float *>(ctx->d_anim + o_rot); a.base_translation = reinterpret_cast<const float *>(ctx->d_anim + o_tr);
    a.base_scale = reinterpret_cast<const float *>(ctx->d_anim + o_sc); a.parent = reinterpret_cast<const int32_t *>(ctx->d_anim + o_par);
    a.order = reinterpret_cast<const uint32_t *>(ctx->d_anim + o_ord); a.level_off = reinterpret_cast<const uint32_t *>(ctx->d_anim + o_lvl);
    a.node_chan_off = reinterpret_cast<const uint32_t *>(ctx->d_anim + o_cho); a.channels = reinterpret_cast<const AnimChannel *>(ctx->d_anim + o_ch);
    a.step_time = reinterpret_cast<const float *>(ctx->d_anim + o_st); a.step_value = reinterpret_cast<const float4 *>(ctx->d_anim + o_sv);
    a.local = reinterpret_cast<float *>(ctx->d_anim + o_loc);
    ctx->have_anim = true;
    return SWEGL_B200_OK;
}

// the node matrices the device currently holds (after an animated frame: k_animate's), for parity checks
int swegl_b200_read_node_matrices(swegl_b200_ctx *ctx, float *node_world, float *node_normal)
{
    if (!ctx) return SWEGL_B200_ERR_ARG;
    if (!ctx->have_scene || !ctx->d_block) return fail(ctx, SWEGL_B200_ERR_STATE, "read_node_matrices before upload_scene");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (node_world) CK(cudaMemcpy(node_world, ctx->d_block + ctx->off_nw, (size_t)64 * ctx->n_nodes, cudaMemcpyDeviceToHost));
    if (node_normal) CK(cudaMemcpy(node_normal, ctx->d_block + ctx->off_nn, (size_t)36 * ctx->n_nodes, cudaMemcpyDeviceToHost));
    return SWEGL_B200_OK;
}

// remap_clipped(1.0f, focal_depth, 0.0f, 5.0f, t) (lerp.hpp:24-43) in IEEE fp32 on the host (this TU is built with
// -ffp-contract=off; volatile keeps every intermediate in fp32)
static float dof_blur_of_t(float t, float focal_depth)
{
    const float a = 1.0f, b = focal_depth;
    volatile float xq;
    if (a == b) xq = 0.5f; else if (t <= a) xq = 0.0f; else if (t >= b) xq = 1.0f;
    else { volatile float num = t - a; volatile float den = b - a; xq = num / den; }
    if (xq <= 0.0f) return 0.0f;
    if (xq >= 5.0f) return 5.0f;
    volatile float m = 5.0f * xq; volatile float r = 0.0f + m;
    return r;
}

// DoF-R's radius (int)blur(t) is a monotone step function of t = |focal_distance - z| >= 0: find, for k = 1..5, the
// smallest float t with radius(t) >= k by bisection over the (monotone) bit patterns of non-negative floats, and the
// largest t whose blur is exactly 0.  The kernel then classifies a pixel with 6 comparisons instead of a division.
static void dof_thresholds(float focal_depth, ViewParams &vp)
{
    auto f = [](uint32_t bits) { float x; memcpy(&x, &bits, 4); return x; };
    vp.dof_const_radius = -1;
    if (focal_depth == 1.0f) { vp.dof_const_radius = (int32_t)dof_blur_of_t(2.0f, focal_depth); }   // a == b: blur 2.5 everywhere
    const uint32_t top = 0x7F800000u;                         // +inf: radius(inf) is the maximum
    for (int k = 1; k <= 5; k++) {
        uint32_t lo = 0, hi = top;                            // invariant: radius(lo) < k (or lo == 0), radius(hi) >= k or hi == top
        if ((int)dof_blur_of_t(f(top), focal_depth) < k) { vp.dof_t[k - 1] = f(0x7FC00000u); continue; }   // never reached: NaN compares false
        if ((int)dof_blur_of_t(0.0f, focal_depth) >= k) { vp.dof_t[k - 1] = 0.0f; continue; }
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if ((int)dof_blur_of_t(f(mid), focal_depth) >= k) hi = mid; else lo = mid;
        }
        vp.dof_t[k - 1] = f(hi);
    }
    // taps count iff blur != 0: largest t with blur == 0 (blur is monotone too)
    {
        uint32_t lo = 0, hi = top;
        if (dof_blur_of_t(f(top), focal_depth) == 0.0f) vp.dof_on = f(top);
        else if (dof_blur_of_t(0.0f, focal_depth) != 0.0f) vp.dof_on = -1.0f;
        else {
            while (hi - lo > 1) {
                const uint32_t mid = lo + (hi - lo) / 2;
                if (dof_blur_of_t(f(mid), focal_depth) != 0.0f) hi = mid; else lo = mid;
            }
            vp.dof_on = f(lo);
        }
    }
}

static int build_view(swegl_b200_ctx *ctx, const swegl_b200_viewport_desc *v, ViewParams &vp)
{
    if (v->w <= 0 || v->h <= 0 || v->x < 0 || v->y < 0 || v->x + v->w > ctx->sw || v->y + v->h > ctx->sh)
        return fail(ctx, SWEGL_B200_ERR_ARG, "viewport rectangle outside the screen (call set_screen first)");
    if (v->light_mode < 0 || v->light_mode > 2 || v->tex_mode < 0 || v->tex_mode > 2 || v->post_mode < 0 || v->post_mode > 1)
        return fail(ctx, SWEGL_B200_ERR_ARG, "bad shader / post mode");
    // transparency layers only change the frame when something can be non-opaque (renderer.cpp:505)
    const int n_layers = (v->transparency_layers > 0 && !ctx->opaque) ? v->transparency_layers : 0;
    if (n_layers > 8)
        return fail(ctx, SWEGL_B200_ERR_UNSUPPORTED, "more than 8 transparency layers");
    if (n_layers > 0 && (v->x != 0 || v->y != 0))
        return fail(ctx, SWEGL_B200_ERR_UNSUPPORTED,
                    "transparency layers on a viewport that is not at the screen origin: viewport_t::flatten (viewport.cpp:62) "
                    "blends against screen rows/columns counted from 0, i.e. against another viewport's pixels");
    memset(&vp, 0, sizeof vp);
    vp.n_layers = n_layers;
    memcpy(vp.view, v->view, sizeof vp.view);
    memcpy(vp.proj, v->proj, sizeof vp.proj);
    memcpy(vp.cam, v->cam_pos, sizeof vp.cam);
    vp.vp_m00 = v->vp_m00; vp.vp_m03 = v->vp_m03; vp.vp_m11 = v->vp_m11; vp.vp_m13 = v->vp_m13;
    vp.vx = v->x; vp.vy = v->y; vp.vw = v->w; vp.vh = v->h;
    vp.band0 = v->y; vp.band1 = v->y + v->h;
    if (v->band_y0 != 0 || v->band_y1 != 0) {
        if (v->band_y0 < 0 || v->band_y1 > v->h || v->band_y0 >= v->band_y1) return fail(ctx, SWEGL_B200_ERR_ARG, "bad band");
        vp.band0 = v->y + v->band_y0; vp.band1 = v->y + v->band_y1;
    }
    vp.nbx = (v->w + 31) / 32;
    vp.ntx = (vp.nbx + FRAG_STRETCH - 1) / FRAG_STRETCH;
    vp.screen_w = ctx->sw;
    vp.light_mode = v->light_mode; vp.tex_mode = v->tex_mode;
    vp.focal_distance = v->focal_distance; vp.focal_depth = v->focal_depth;
    vp.dof_const_radius = -1;
    if (v->post_mode == SWEGL_B200_POST_DOF) {
        if (!(v->focal_depth == v->focal_depth) || !(v->focal_distance == v->focal_distance))
            return fail(ctx, SWEGL_B200_ERR_ARG, "DoF focal parameters are NaN");
        if (v->focal_depth != ctx->dof_cache_depth || !ctx->dof_cache_valid) {
            dof_thresholds(v->focal_depth, ctx->dof_cache);
            ctx->dof_cache_depth = v->focal_depth; ctx->dof_cache_valid = true;
        }
        memcpy(vp.dof_t, ctx->dof_cache.dof_t, sizeof vp.dof_t);
        vp.dof_on = ctx->dof_cache.dof_on; vp.dof_const_radius = ctx->dof_cache.dof_const_radius;
    }
    return SWEGL_B200_OK;
}

// the rows the kernels draw: DoF needs colour + depth of a 5-row halo around the band, rendered redundantly
static ViewParams draw_params(const ViewParams &vp, bool dof)
{
    ViewParams d = vp;
    if (dof) { d.band0 = max(vp.vy, vp.band0 - 5); d.band1 = min(vp.vy + vp.vh, vp.band1 + 5); }
    return d;
}

// the DoF pass stages its source windows by TMA: (re)encode the two tensor maps when the viewport size or the buffers changed
static bool ensure_dof_maps(swegl_b200_ctx *ctx, int vw, int vh)
{
    if (!ctx->dof_tma_policy) return false;
    if (ctx->dof_tm_w != vw || ctx->dof_tm_h != vh) {
        ctx->dof_tm_ok = make_dof_tensor_maps(ctx->d_tmp_color, ctx->d_depth, vw, vh, &ctx->dof_tm_color, &ctx->dof_tm_depth);
        ctx->dof_tm_w = vw; ctx->dof_tm_h = vh;
    }
    return ctx->dof_tm_ok;
}

// enqueue one viewport's work on the stream (no synchronisation: usable under stream capture): upload of the staging
// slot (whole block when the frame data is new, else just the ViewParams), then the kernel sequence.
// with_world: v_world is not (completely) there yet for this frame -- the vertex stage computes it.
static uint32_t issue_view(swegl_b200_ctx *ctx, const ViewParams &out, const ViewParams &vp, swegl_b200_ctx::Slot &sl, bool with_frame, bool with_world,
                           bool dof, bool count_covered, bool timing, bool sync_counters, Counters *counters_out, bool synced = false)
{
    cudaStream_t st = ctx->stream;
    uint32_t launches = 0;
    if (timing) cudaEventRecord(ctx->ev[0], st);
    if (with_frame && ctx->frame_animated) {
        // animated frame: the slot carries FrameParams (with the time stamp), ViewParams and the lights; the node matrices
        // in between are produced on the device by k_animate (scene_t::animate + hierarchy product), ahead of the vertex stage
        cudaMemcpyAsync(ctx->d_block, sl.block, ctx->off_nw, cudaMemcpyHostToDevice, st);
        if (ctx->block_bytes > ctx->off_lights)
            cudaMemcpyAsync(ctx->d_block + ctx->off_lights, sl.block + ctx->off_lights, ctx->block_bytes - ctx->off_lights, cudaMemcpyHostToDevice, st);
        launch_animate(ctx->anim, ctx->d_fp(), reinterpret_cast<float *>(ctx->d_block + ctx->off_nw), reinterpret_cast<float *>(ctx->d_block + ctx->off_nn), st);
        launches++;
    }
    else if (with_frame) cudaMemcpyAsync(ctx->d_block, sl.block, ctx->block_bytes, cudaMemcpyHostToDevice, st);
    else cudaMemcpyAsync(ctx->d_block + ctx->off_vp, sl.block + ctx->off_vp, sizeof(ViewParams), cudaMemcpyHostToDevice, st);
    // frame protocol (FrameSync): rank 0 clears the other ranks' rows of its screen and announces the frame; the others
    // wait for that before their first store into it, skip the background, and raise their flag at the end
    FrameSync *const own_sync = ctx->own_sync();
    FrameSync *const tgt_sync = ctx->color_target ? reinterpret_cast<FrameSync *>(ctx->color_target + (size_t)ctx->sw * ctx->sh) : own_sync;
    const bool skip_bg = synced && ctx->sync_rank > 0 && !dof;
    if (synced && ctx->sync_rank == 0) { launch_sync_clear(out, ctx->d_vp(), ctx->d_screen, ctx->sw, own_sync, !dof, st); launches++; }
    // a band of a sharded frame is a latency chain on an otherwise idle GPU: its kernels release their successors early
    // (common.cuh pdl_trigger); whole views, which share the GPU with other contexts' frames, do not
    const uint32_t early = (vp.band0 > vp.vy || vp.band1 < vp.vy + vp.vh) ? 1u : 0u;
    ctx->pools.early_trigger = early; ctx->cull.early_trigger = early;
    DeviceScene ds = ctx->ds;
    ds.early_trigger = early;
    if (view_culled(ctx, vp)) {                             // sort-first band: skip what cannot reach it (common.cuh)
        ds.cl_live = ctx->cull.cl_live; ds.mark_need = ctx->cull.mark_need; ds.vert_need = ctx->cull.vert_need;
        ds.live_list = ctx->cull.live_list; ds.mark_list = ctx->cull.mark_list; ds.vert_list = ctx->cull.vert_list;
        ds.cull_counts = ctx->cull.counts;
        launch_cull(ds, ctx->d_vp(), ctx->cull, st); launches += 2;
    }
    launch_vertex(ds, ctx->d_vp(), ctx->pools.counters, with_world, st); launches++;
    launch_mark(ds, st); launches++;
    if (timing) cudaEventRecord(ctx->ev[1], st);
    launch_setup(ds, ctx->d_vp(), ctx->d_fp(), ctx->pools, st); launches++;
    if (timing) cudaEventRecord(ctx->ev[2], st);
    launch_spans(ctx->d_vp(), ctx->pools, ctx->dense_spans, ctx->shared_gpu && !early, st); launches++;
    if (timing) cudaEventRecord(ctx->ev[3], st);
    uint32_t *screen_out = ctx->color_target ? ctx->color_target : ctx->d_screen;
    uint32_t *color = dof ? ctx->d_tmp_color - ((size_t)vp.vy * vp.vw + vp.vx) : screen_out;
    const int color_pitch = dof ? vp.vw : ctx->sw;
    uint32_t *post_dst = screen_out + (size_t)vp.vy * ctx->sw + vp.vx;     // the viewport's origin in the destination screen
    const int out0 = out.band0 - vp.vy, out1 = out.band1 - vp.vy;           // the rows the post pass produces (the band without its halo)
    // the first kernel that stores into rank 0's screen (with DoF-R, k_fragments already stores the constant tiles)
    if (synced && ctx->sync_rank > 0) { launch_sync_wait_ready(ctx->d_vp(), tgt_sync, own_sync, st); launches++; }
    // asynchronous frames publish their counters from inside k_fragments; synchronous ones copy them at the end
    if (vp.n_layers > 0) {
        launch_fragments_layers(ctx->ds, vp, ctx->d_vp(), ctx->d_fp(), ctx->pools, color, color_pitch, ctx->d_depth, count_covered,
                                sync_counters ? nullptr : counters_out, st);
        launches++;
        if (dof) { launch_dof_classify(vp, ctx->d_vp(), ctx->pools, out0, out1, post_dst, ctx->sw, st); launches++; }
    } else {
        launch_fragments(ctx->ds, vp, ctx->d_vp(), ctx->d_fp(), ctx->pools, color, color_pitch, ctx->d_depth, count_covered,
                         sync_counters ? nullptr : counters_out, skip_bg, ctx->fast_shading, dof, out0, out1, post_dst, ctx->sw, ctx->dense_spans, st);
        launches++;
    }
    if (timing) cudaEventRecord(ctx->ev[4], st);
    if (dof) {
        const bool tma = ensure_dof_maps(ctx, vp.vw, vp.vh);
        launch_dof(vp, ctx->d_vp(), ctx->pools, &ctx->dof_tm_color, &ctx->dof_tm_depth, tma, ctx->d_tmp_color, vp.vw, ctx->d_depth,
                   post_dst, ctx->sw, out0, out1, st);
        launches++;
    }
    if (timing) cudaEventRecord(ctx->ev[5], st);
    if (synced) {
        if (ctx->sync_rank > 0) launch_sync_signal(ctx->d_vp(), tgt_sync, ctx->sync_rank, st);
        else launch_sync_wait_done(ctx->d_vp(), own_sync, ctx->sync_world, st);
        launches++;
    }
    if (sync_counters) cudaMemcpyAsync(counters_out, ctx->pools.counters, sizeof(Counters), cudaMemcpyDeviceToHost, st);
    return launches;
}

// pick the staging slot of this render call and put the ViewParams into it
static int stage_view(swegl_b200_ctx *ctx, const ViewParams &vp, int &si, bool &with_frame)
{
    with_frame = ctx->frame_dirty;
    si = with_frame ? ctx->frame_slot : ctx->next_slot;
    auto &sl = ctx->slots[si];
    int rc = acquire_slot(ctx, sl);          // no-op right after begin_frame acquired it
    if (rc) return rc;
    memcpy(sl.block + ctx->off_vp, &vp, sizeof(ViewParams));
    // every rendered view gets a stamp no earlier view had (0 = the initial tile_stamp contents, skipped on wrap-around)
    if (++ctx->stamp == 0) ctx->stamp = 1;
    reinterpret_cast<ViewParams *>(sl.block + ctx->off_vp)->stamp = ctx->stamp;
    reinterpret_cast<ViewParams *>(sl.block + ctx->off_vp)->sync_seq = ctx->sync_seq;
    return SWEGL_B200_OK;
}

static int finish_view(swegl_b200_ctx *ctx, int si)
{
    auto &sl = ctx->slots[si];
    CK(cudaEventRecord(sl.done, ctx->stream));
    sl.pending = true;
    ctx->next_slot = si ^ 1;
    ctx->last_slot = si;
    sl.ticket = 0;
    ctx->frame_dirty = false;
    return SWEGL_B200_OK;
}

// asynchronous frame: replay (or first capture) the CUDA graph of this viewport configuration
static int render_async(swegl_b200_ctx *ctx, const ViewParams &out, bool dof)
{
    cudaStream_t st = ctx->stream;
    const ViewParams vp = draw_params(out, dof);
    // the frame protocol covers banded, opaque views of the asynchronous path (a synchronous frame may be redone when a
    // pool overflows, which would break the ranks' lock step)
    const bool banded = out.band0 > out.vy || out.band1 < out.vy + out.vh;
    const bool synced = ctx->sync_rank >= 0 && banded && out.n_layers == 0;
    if (synced) {
        if (ctx->sync_rank > 0 && !ctx->color_target) return fail(ctx, SWEGL_B200_ERR_STATE, "frame sync: rank > 0 needs a colour target (set_color_target)");
        if (ctx->sync_rank == 0 && ctx->color_target) return fail(ctx, SWEGL_B200_ERR_STATE, "frame sync: rank 0 assembles the frame in its own screen");
        ctx->sync_seq++;
    }
    int si; bool with_frame;
    int rc = stage_view(ctx, vp, si, with_frame);
    if (rc) return rc;
    auto &sl = ctx->slots[si];
    // a culled view transforms only the vertex blocks it needs (and always with the world transform); the first view
    // that covers every vertex completes v_world for the rest of the frame
    const bool culled = view_culled(ctx, vp);
    const bool with_world = !ctx->world_complete;
    const bool tma = dof && ensure_dof_maps(ctx, vp.vw, vp.vh);
    if (!ctx->graphs_enabled) {
        issue_view(ctx, out, vp, sl, with_frame, with_world, dof, false, false, false, sl.counters, synced);
    } else {
        const uint64_t tgt = (uint64_t)reinterpret_cast<uintptr_t>(ctx->color_target);
        const int32_t key[15] = { out.vx, out.vy, out.vw, out.vh, out.band0, out.band1, out.light_mode, out.tex_mode, (dof ? 1 : 0) | (out.n_layers << 1),
                                  ctx->sw, ctx->sh, (with_frame ? 1 : 0) | (synced ? 2 + 4 * ctx->sync_rank + 256 * ctx->sync_world : 0),
                                  (ctx->dense_spans ? 1 : 0) | (with_world ? 2 : 0) | (ctx->fast_shading ? 4 : 0) | (tma ? 8 : 0) | (with_frame && ctx->frame_animated ? 16 : 0) | (ctx->shared_gpu ? 32 : 0),
                                  (int32_t)(tgt & 0xFFFFFFFFu), (int32_t)(tgt >> 32) };
        swegl_b200_ctx::ViewGraph *vg = nullptr;
        for (auto &g : ctx->view_graphs) if (memcmp(g.key, key, sizeof key) == 0) { vg = &g; break; }
        if (!vg) {
            if (ctx->view_graphs.size() >= 64) { CK(cudaStreamSynchronize(st)); drop_graphs(ctx); }
            swegl_b200_ctx::ViewGraph ng{};
            memcpy(ng.key, key, sizeof key);
            ctx->view_graphs.push_back(ng);
            vg = &ctx->view_graphs.back();
        }
        if (!vg->exec[si]) {
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            issue_view(ctx, out, vp, sl, with_frame, with_world, dof, false, false, false, sl.counters, synced);
            CK(cudaStreamEndCapture(st, &g));
            cudaError_t e = cudaGraphInstantiate(&vg->exec[si], g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) { ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return SWEGL_B200_ERR_CUDA; }
        }
        CK(cudaGraphLaunch(vg->exec[si], st));
    }
    CK(cudaGetLastError());
    if (!culled) ctx->world_complete = true;
    return finish_view(ctx, si);
}

static int render_common(swegl_b200_ctx *ctx, const swegl_b200_viewport_desc *v, bool sync, swegl_b200_stats *stats)
{
    if (!ctx || !v) return SWEGL_B200_ERR_ARG;
    if (!ctx->have_scene || !ctx->have_frame || !ctx->d_screen)
        return fail(ctx, SWEGL_B200_ERR_STATE, "render before upload_scene / set_screen / begin_frame");
    CK(cudaSetDevice(ctx->device));
    ViewParams out;
    int rc = build_view(ctx, v, out);
    if (rc) return rc;
    const bool dof = v->post_mode == SWEGL_B200_POST_DOF;
    if (stats) memset(stats, 0, sizeof *stats);

    if (!sync && !stats) {                                   // fire and forget
        rc = render_async(ctx, out, dof);
        if (rc == SWEGL_B200_OK) { ctx->last_vp = out; ctx->have_vp = true; ctx->last_dof = dof; }
        return rc;
    }
    // synchronous frame: direct launches, counters checked, pools grown and the frame redone if needed
    const bool timing = ctx->timing && stats;
    const ViewParams vp = draw_params(out, dof);
    uint32_t grows = 0, launches = 0;
    for (;;) {
        int si; bool with_frame;
        rc = stage_view(ctx, vp, si, with_frame);
        if (rc) return rc;
        launches = issue_view(ctx, out, vp, ctx->slots[si], with_frame, !ctx->world_complete, dof, stats != nullptr, timing, true, ctx->h_counters);
        CK(cudaGetLastError());
        if (!view_culled(ctx, vp)) ctx->world_complete = true;
        rc = finish_view(ctx, si);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->slots[si].pending = false;
        choose_span_kernel(ctx, *ctx->h_counters);
        if (!ctx->h_counters->overflow) break;
        if (++grows > 8) return fail(ctx, SWEGL_B200_ERR_CAPACITY, "span/chunk/fragment pools keep overflowing");
        rc = grow_pools_for(ctx, *ctx->h_counters);          // the frame data is already on the device: the redo is view-only
        if (rc) return rc;
    }
    if (stats) {
        const Counters &c = *ctx->h_counters;
        stats->n_setup_triangles = c.n_slots; stats->n_spans = c.n_rows; stats->n_chunks = c.n_chunks;
        stats->n_covered = c.n_covered; stats->n_launches = launches; stats->pool_grows = grows;
        stats->n_busy_tiles = c.n_busy;
#ifdef FRAG_PROBE_MAXLIST
        stats->n_launches = c.pad; stats->pool_grows = c.dof_queue;     // probe build: bins with more than FRAG_LIST_CAP / 32 pieces
#endif
        if (timing) {
            cudaEventElapsedTime(&stats->ms_vertex, ctx->ev[0], ctx->ev[1]);
            cudaEventElapsedTime(&stats->ms_setup, ctx->ev[1], ctx->ev[2]);
            cudaEventElapsedTime(&stats->ms_raster, ctx->ev[2], ctx->ev[3]);
            cudaEventElapsedTime(&stats->ms_fragment, ctx->ev[3], ctx->ev[4]);
            cudaEventElapsedTime(&stats->ms_post, ctx->ev[4], ctx->ev[5]);
            cudaEventElapsedTime(&stats->ms_total, ctx->ev[0], ctx->ev[5]);
        }
    }
    ctx->last_vp = out; ctx->have_vp = true; ctx->last_dof = dof;
    return SWEGL_B200_OK;
}

int swegl_b200_render_viewport_device(swegl_b200_ctx *ctx, const swegl_b200_viewport_desc *vp, swegl_b200_stats *stats)
{
    return render_common(ctx, vp, false, stats);
}

int swegl_b200_render_viewport(swegl_b200_ctx *ctx, const swegl_b200_viewport_desc *v, void *pixels, int32_t pitch_bytes,
                               float *zbuffer, swegl_b200_stats *stats)
{
    if (!pixels || pitch_bytes < 4) return fail(ctx, SWEGL_B200_ERR_ARG, "render_viewport: null pixels");
    if (ctx && ctx->color_target)
        return fail(ctx, SWEGL_B200_ERR_STATE, "render_viewport: a colour target is set (the finished pixels are in the target's screen, not here); "
                                               "use render_viewport_device, or set_color_target(NULL) first");
    // the frame and its read-back are queued back to back; one synchronisation at the end.  If the frame
    // overflowed a pool (first frames of a new scene) it is redone through the synchronous path.
    int rc = render_common(ctx, v, false, stats);
    if (rc) return rc;
    const ViewParams &vp = ctx->last_vp;
    cudaStream_t st = ctx->stream;
    const int rows = vp.band1 - vp.band0;
    for (auto &hi : ctx->host_images) if (hi.pixels == pixels) hi.pixels = nullptr;     // the whole band is rewritten: forget what was tracked
    for (int attempt = 0; attempt < 2; attempt++) {
        CK(cudaMemcpy2DAsync((char *)pixels + (size_t)vp.band0 * pitch_bytes + (size_t)vp.vx * 4, (size_t)pitch_bytes,
                             ctx->d_screen + (size_t)vp.band0 * ctx->sw + vp.vx, (size_t)ctx->sw * 4,
                             (size_t)vp.vw * 4, (size_t)rows, cudaMemcpyDeviceToHost, st));
        if (zbuffer)
            CK(cudaMemcpyAsync(zbuffer + (size_t)(vp.band0 - vp.vy) * vp.vw, ctx->d_depth + (size_t)(vp.band0 - vp.vy) * vp.vw,
                               (size_t)rows * vp.vw * 4, cudaMemcpyDeviceToHost, st));
        rc = swegl_b200_synchronize(ctx);
        if (rc != SWEGL_B200_ERR_CAPACITY || attempt) break;
        ctx->err.clear();
        rc = render_common(ctx, v, true, stats);             // pools are larger now: redo synchronously
        if (rc) return rc;
    }
    return rc;
}

// what DoF-R turns untouched background into (fragment.cu dof_background_value, on the host; same comparisons)
static uint32_t post_background(const ViewParams &vp, bool dof)
{
    if (!dof) return 0u;
    uint32_t zb = MAXZ_BITS; float z; memcpy(&z, &zb, 4);
    volatile float d = vp.focal_distance - z;
    const float t = fabsf(d);
    uint32_t radius; bool counts;
    if (vp.dof_const_radius >= 0) { radius = (uint32_t)vp.dof_const_radius; counts = true; }
    else {
        radius = (t >= vp.dof_t[0]) + (t >= vp.dof_t[1]) + (t >= vp.dof_t[2]) + (t >= vp.dof_t[3]) + (t >= vp.dof_t[4]);
        counts = t > vp.dof_on;
    }
    return (radius != 0 && counts) ? 0xFF000000u : 0u;
}

// The device->host copies of an asynchronous frame.  They are issued late -- when the next frame has been queued, or in
// swegl_b200_wait -- because by then the frame's kernels are done and k_fragments has published the bounding box of what
// was drawn: outside that box (grown by the blur radius with DoF-R) the frame is one constant, and if the host image
// already holds that constant there (the library remembers what it last left in each image), only the union of the old
// and the new box has to cross PCIe.
static int issue_readback(swegl_b200_ctx *ctx, swegl_b200_ctx::OutBuf &ob)
{
    if (ob.d2h_issued) return SWEGL_B200_OK;
    const ViewParams &vp = ob.vp;
    CK(cudaEventSynchronize(ob.ready));
    const int row_a = vp.band0 - vp.vy, row_b = vp.band1 - vp.vy;            // viewport-relative rows of this view
    int cx0 = 0, cx1 = vp.vw, cy0 = row_a, cy1 = row_b;                       // the rectangle to copy; default: everything
    const auto &sl = ctx->slots[ob.slot];
    if (ob.partial_ok && ctx->partial_readback && sl.ticket == ob.ticket && !sl.counters->overflow) {
        const Counters &c = *sl.counters;
        int fx0 = 0, fx1 = 0, fy0 = 0, fy1 = 0;                               // this frame's non-constant rectangle (empty: nothing drawn)
        if (c.bb_x1 > c.bb_x0 && c.bb_y1 > c.bb_y0) {
            const int grow = ob.dof ? 8 : 0;                                  // blur radius <= 5
            fx0 = std::max(0, ((int)c.bb_x0 - grow) & ~3); fx1 = std::min(vp.vw, ((int)c.bb_x1 + grow + 3) & ~3);
            fy0 = std::max(row_a, (int)c.bb_y0 - grow); fy1 = std::min(row_b, (int)c.bb_y1 + grow);
        }
        const uint32_t bg = post_background(vp, ob.dof);
        swegl_b200_ctx::HostImage *hi = nullptr;
        for (auto &h : ctx->host_images) if (h.pixels == ob.pixels) { hi = &h; break; }
        const bool known = hi && hi->zbuffer == ob.zbuffer && hi->pitch_bytes == ob.pitch_bytes && hi->vx == vp.vx && hi->vy == vp.vy && hi->vw == vp.vw
                        && hi->vh == vp.vh && hi->band0 == vp.band0 && hi->band1 == vp.band1 && hi->bg == bg;
        if (known) {
            const bool old_empty = hi->x1 <= hi->x0 || hi->y1 <= hi->y0, new_empty = fx1 <= fx0 || fy1 <= fy0;
            if (old_empty && new_empty) { cx0 = cx1 = cy0 = cy1 = 0; }
            else if (old_empty) { cx0 = fx0; cx1 = fx1; cy0 = fy0; cy1 = fy1; }
            else if (new_empty) { cx0 = hi->x0; cx1 = hi->x1; cy0 = hi->y0; cy1 = hi->y1; }
            else { cx0 = std::min(fx0, hi->x0); cx1 = std::max(fx1, hi->x1); cy0 = std::min(fy0, hi->y0); cy1 = std::max(fy1, hi->y1); }
        }
        if (!hi) {
            if (ctx->host_images.size() >= 8) {                              // forget the least recently used image
                size_t k = 0;
                for (size_t i = 1; i < ctx->host_images.size(); i++) if (ctx->host_images[i].age < ctx->host_images[k].age) k = i;
                ctx->host_images.erase(ctx->host_images.begin() + k);
            }
            ctx->host_images.emplace_back();
            hi = &ctx->host_images.back();
        }
        hi->pixels = ob.pixels; hi->zbuffer = ob.zbuffer; hi->pitch_bytes = ob.pitch_bytes;
        hi->vx = vp.vx; hi->vy = vp.vy; hi->vw = vp.vw; hi->vh = vp.vh; hi->band0 = vp.band0; hi->band1 = vp.band1; hi->bg = bg;
        hi->x0 = fx0; hi->x1 = fx1; hi->y0 = fy0; hi->y1 = fy1;
        hi->age = ob.ticket;
    } else {
        for (auto &h : ctx->host_images) if (h.pixels == ob.pixels) h.pixels = nullptr;
    }
    CK(cudaStreamWaitEvent(ctx->copy_stream, ob.ready, 0));
    if (cx1 > cx0 && cy1 > cy0) {
        const size_t w_bytes = (size_t)(cx1 - cx0) * 4, n_rows = (size_t)(cy1 - cy0);
        CK(cudaMemcpy2DAsync((char *)ob.pixels + (size_t)(vp.vy + cy0) * ob.pitch_bytes + (size_t)(vp.vx + cx0) * 4, (size_t)ob.pitch_bytes,
                             ob.color + (size_t)(cy0 - row_a) * vp.vw + cx0, (size_t)vp.vw * 4, w_bytes, n_rows, cudaMemcpyDeviceToHost, ctx->copy_stream));
        ctx->readback_bytes += w_bytes * n_rows;
        if (ob.zbuffer) {
            CK(cudaMemcpy2DAsync(ob.zbuffer + (size_t)cy0 * vp.vw + cx0, (size_t)vp.vw * 4, ob.depth + (size_t)(cy0 - row_a) * vp.vw + cx0, (size_t)vp.vw * 4,
                                 w_bytes, n_rows, cudaMemcpyDeviceToHost, ctx->copy_stream));
            ctx->readback_bytes += w_bytes * n_rows;
        }
    }
    ctx->readback_frames++;
    CK(cudaEventRecord(ob.copied, ctx->copy_stream));
    ob.d2h_issued = true;
    return SWEGL_B200_OK;
}

int swegl_b200_render_viewport_async(swegl_b200_ctx *ctx, const swegl_b200_viewport_desc *v, void *pixels, int32_t pitch_bytes,
                                     float *zbuffer, uint64_t *ticket)
{
    if (!ctx || !pixels || pitch_bytes < 4 || !ticket) return fail(ctx, SWEGL_B200_ERR_ARG, "render_viewport_async: null argument");
    if (ctx->color_target)
        return fail(ctx, SWEGL_B200_ERR_STATE, "render_viewport_async: a colour target is set (the finished pixels are in the target's screen, not here)");
    constexpr int N_OUT = swegl_b200_ctx::N_OUT;
    auto &ob = ctx->out[ctx->ticket_seq % N_OUT];
    auto &prev = ctx->out[(ctx->ticket_seq + N_OUT - 1) % N_OUT];
    // this staging image's previous frame (N_OUT submits ago) must have its copies queued before it is overwritten
    if (ob.in_flight && !ob.d2h_issued) { int rc0 = issue_readback(ctx, ob); if (rc0) return rc0; }
    int rc = render_common(ctx, v, false, nullptr);
    if (rc) return rc;
    const ViewParams &vp = ctx->last_vp;
    const int rows = vp.band1 - vp.band0;
    const size_t need = (size_t)vp.vw * rows;
    if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ob.ready) { CK(cudaEventCreateWithFlags(&ob.ready, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ob.copied, cudaEventDisableTiming)); }
    if (ob.cap < need) {
        if (ob.in_flight) CK(cudaEventSynchronize(ob.copied));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(dalloc(ob.color, need)); CK(dalloc(ob.depth, need));
        ob.cap = need; ob.stage_valid = false;
    }
    if (!ob.rec) { CK(cudaMalloc((void **)&ob.rec, 2 * sizeof(StageRect))); ob.stage_valid = false; }
    cudaStream_t st = ctx->stream;
    if (ob.in_flight) CK(cudaStreamWaitEvent(st, ob.copied, 0));        // the frame N_OUT submits ago has left this image
    {
        // only the part of the view that differs from what the staging image already holds is copied (k_stage_rect): the image
        // must hold a complete frame of the same view geometry over the same background, and the view must publish its box
        const bool dof = ctx->last_dof;
        const uint32_t bg = post_background(vp, dof);
        const ViewParams &o = ob.vp;
        const bool same = ob.stage_valid && vp.n_layers == 0 && o.vx == vp.vx && o.vy == vp.vy && o.vw == vp.vw && o.vh == vp.vh
                       && o.band0 == vp.band0 && o.band1 == vp.band1 && ob.stage_bg == bg;
        launch_stage_rect(ob.color, ctx->d_screen + (size_t)vp.band0 * ctx->sw + vp.vx, ctx->sw, vp.vw, vp.band0 - vp.vy, vp.band1 - vp.vy,
                          dof ? 8 : 0, ctx->pools.counters, ob.rec, ob.rec_par, !same, st);
        CK(cudaGetLastError());
        ob.rec_par ^= 1; ob.stage_valid = vp.n_layers == 0; ob.stage_bg = bg;
    }
    if (zbuffer)
        CK(cudaMemcpyAsync(ob.depth, ctx->d_depth + (size_t)(vp.band0 - vp.vy) * vp.vw, need * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaEventRecord(ob.ready, st));
    ob.in_flight = true; ob.d2h_issued = false;
    ob.pixels = pixels; ob.pitch_bytes = pitch_bytes; ob.zbuffer = zbuffer; ob.vp = vp; ob.dof = ctx->last_dof;
    ob.partial_ok = vp.n_layers == 0;                                    // (the layer kernel does not publish a bounding box)
    ob.slot = ctx->last_slot;
    ob.ticket = ++ctx->ticket_seq;
    ctx->slots[ob.slot].ticket = ob.ticket;
    *ticket = ob.ticket;
    // the frame before this one: its kernels are done or about to be, and this frame is already queued behind them, so
    // waiting for it here does not idle the GPU
    if (prev.in_flight && !prev.d2h_issued) { rc = issue_readback(ctx, prev); if (rc) return rc; }
    return SWEGL_B200_OK;
}

int swegl_b200_wait(swegl_b200_ctx *ctx, uint64_t ticket)
{
    if (!ctx || ticket == 0 || ticket > ctx->ticket_seq) return fail(ctx, SWEGL_B200_ERR_ARG, "wait: unknown ticket");
    CK(cudaSetDevice(ctx->device));
    auto &ob = ctx->out[(ticket - 1) % swegl_b200_ctx::N_OUT];
    // a later frame through the same staging image implies this one is done (same stream order)
    if (ob.in_flight) {
        int rc0 = issue_readback(ctx, ob);
        if (rc0) return rc0;
        CK(cudaEventSynchronize(ob.copied));
        ob.in_flight = false;
    }
    // the frame's kernels are complete.  If its staging slot was not reused since, its pool counters are still
    // unexamined: do that now; if it was, acquire_slot() already did and noted a failure
    if (ob.ticket == ticket && ctx->slots[ob.slot].ticket == ticket) {
        int rc = check_slot_overflow(ctx, ctx->slots[ob.slot]);
        if (rc && rc != SWEGL_B200_ERR_CAPACITY) return rc;
    }
    for (size_t i = 0; i < ctx->failed_tickets.size(); i++)
        if (ctx->failed_tickets[i] == ticket) {
            ctx->failed_tickets.erase(ctx->failed_tickets.begin() + i);
            for (auto &h : ctx->host_images) h.pixels = nullptr;        // an incomplete frame went into a host image
            return fail(ctx, SWEGL_B200_ERR_CAPACITY, "the frame overflowed the span/chunk/fragment pools (now enlarged): submit it again "
                                                      "(begin_frame with its data, then render_viewport_async)");
        }
    ctx->err.clear();
    return SWEGL_B200_OK;
}

int swegl_b200_set_partial_readback(swegl_b200_ctx *ctx, int enabled)
{
    if (!ctx) return SWEGL_B200_ERR_ARG;
    ctx->partial_readback = enabled != 0;
    ctx->host_images.clear();
    return SWEGL_B200_OK;
}

int swegl_b200_invalidate_host_image(swegl_b200_ctx *ctx, const void *pixels)
{
    if (!ctx) return SWEGL_B200_ERR_ARG;
    for (auto &h : ctx->host_images) if (!pixels || h.pixels == pixels) h.pixels = nullptr;
    return SWEGL_B200_OK;
}

int swegl_b200_readback_stats(swegl_b200_ctx *ctx, uint64_t out[2], int reset)
{
    if (!ctx || !out) return SWEGL_B200_ERR_ARG;
    out[0] = ctx->readback_bytes; out[1] = ctx->readback_frames;
    if (reset) ctx->readback_bytes = ctx->readback_frames = 0;
    return SWEGL_B200_OK;
}

int swegl_b200_set_shading(swegl_b200_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 1) return fail(ctx, SWEGL_B200_ERR_ARG, "set_shading: mode must be SWEGL_B200_SHADING_EXACT or SWEGL_B200_SHADING_FAST");
    ctx->fast_shading = mode == SWEGL_B200_SHADING_FAST;        // (part of the graph key: no re-capture needed)
    return SWEGL_B200_OK;
}

int swegl_b200_set_shared_gpu(swegl_b200_ctx *ctx, int shared)
{
    if (!ctx) return SWEGL_B200_ERR_ARG;
    ctx->shared_gpu = shared != 0;                              // (part of the graph key: no re-capture needed)
    return SWEGL_B200_OK;
}

uint64_t swegl_b200_frame_hash(const uint32_t *words, size_t n_words)
{
    uint64_t h = 1469598103934665603ull;                        // FNV-1a 64 over 32-bit words
    for (size_t i = 0; i < n_words; i++) { h ^= words[i]; h *= 1099511628211ull; }
    return h;
}

int swegl_b200_set_color_target(swegl_b200_ctx *ctx, void *device_screen)
{
    if (!ctx || !ctx->d_screen) return fail(ctx, SWEGL_B200_ERR_STATE, "set_color_target before set_screen");
    ctx->color_target = static_cast<uint32_t *>(device_screen);
    return SWEGL_B200_OK;
}

int swegl_b200_export_screen(swegl_b200_ctx *ctx, void *handle64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI promises a 64-byte handle");
    if (!ctx || !handle64 || !ctx->d_screen) return fail(ctx, SWEGL_B200_ERR_STATE, "export_screen before set_screen");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->d_screen));
    memcpy(handle64, &h, sizeof h);
    return SWEGL_B200_OK;
}

int swegl_b200_import_screen(swegl_b200_ctx *ctx, const void *handle64, void **peer_screen)
{
    if (!ctx || !handle64 || !peer_screen) return fail(ctx, SWEGL_B200_ERR_ARG, "import_screen: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    for (size_t i = 0; i < ctx->imported.size(); i++)
        if (!memcmp(&ctx->imported_handles[i], &h, sizeof h)) { *peer_screen = ctx->imported[i]; return SWEGL_B200_OK; }
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->imported.push_back(p);
    ctx->imported_handles.push_back(h);
    *peer_screen = p;
    return SWEGL_B200_OK;
}

int swegl_b200_device_buffers(swegl_b200_ctx *ctx, void **screen_dev, void **depth_dev)
{
    if (!ctx) return SWEGL_B200_ERR_ARG;
    if (screen_dev) *screen_dev = ctx->d_screen;
    if (depth_dev) *depth_dev = ctx->d_depth;
    return SWEGL_B200_OK;
}

int swegl_b200_read_screen(swegl_b200_ctx *ctx, int32_t y0, int32_t y1, void *pixels, int32_t pitch_bytes)
{
    if (!ctx || !pixels || !ctx->d_screen || y0 < 0 || y1 > ctx->sh || y0 > y1 || pitch_bytes < ctx->sw * 4)
        return fail(ctx, SWEGL_B200_ERR_ARG, "read_screen: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(pixels, (size_t)pitch_bytes, ctx->d_screen + (size_t)y0 * ctx->sw, (size_t)ctx->sw * 4,
                         (size_t)ctx->sw * 4, (size_t)(y1 - y0), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SWEGL_B200_OK;
}

int swegl_b200_read_rect(swegl_b200_ctx *ctx, int32_t x, int32_t y, int32_t w, int32_t h, void *pixels, int32_t pitch_bytes)
{
    if (!ctx || !pixels || !ctx->d_screen || x < 0 || y < 0 || w <= 0 || h <= 0 || x + w > ctx->sw || y + h > ctx->sh || pitch_bytes < w * 4)
        return fail(ctx, SWEGL_B200_ERR_ARG, "read_rect: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync((char *)pixels + (size_t)y * pitch_bytes + (size_t)x * 4, (size_t)pitch_bytes, ctx->d_screen + (size_t)y * ctx->sw + x,
                         (size_t)ctx->sw * 4, (size_t)w * 4, (size_t)h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SWEGL_B200_OK;
}

int swegl_b200_read_depth_rows(swegl_b200_ctx *ctx, int32_t row0, int32_t row1, float *zbuffer)
{
    if (!ctx || !zbuffer || !ctx->have_vp) return fail(ctx, SWEGL_B200_ERR_ARG, "read_depth_rows: nothing rendered");
    const ViewParams &vp = ctx->last_vp;
    if (row0 < 0 || row1 > vp.vh || row0 > row1) return fail(ctx, SWEGL_B200_ERR_ARG, "read_depth_rows: bad rows");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(zbuffer + (size_t)row0 * vp.vw, ctx->d_depth + (size_t)row0 * vp.vw, (size_t)(row1 - row0) * vp.vw * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SWEGL_B200_OK;
}

int swegl_b200_enable_peer(swegl_b200_ctx *ctx, int peer_device)
{
    if (!ctx || peer_device < 0) return fail(ctx, SWEGL_B200_ERR_ARG, "enable_peer: bad device");
    if (peer_device == ctx->device) return SWEGL_B200_OK;
    CK(cudaSetDevice(ctx->device));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return fail(ctx, SWEGL_B200_ERR_UNSUPPORTED, "enable_peer: the devices have no peer access (NVLink / PCIe P2P)");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    CK(e);
    return SWEGL_B200_OK;
}

int swegl_b200_device_of(const swegl_b200_ctx *ctx) { return ctx ? ctx->device : -1; }
int swegl_b200_scene_opaque(const swegl_b200_ctx *ctx) { return (ctx && ctx->have_scene) ? (ctx->opaque ? 1 : 0) : -1; }

int swegl_b200_read_depth(swegl_b200_ctx *ctx, float *zbuffer)
{
    if (!ctx || !zbuffer || !ctx->have_vp) return fail(ctx, SWEGL_B200_ERR_ARG, "read_depth: nothing rendered");
    CK(cudaSetDevice(ctx->device));
    const ViewParams &vp = ctx->last_vp;
    CK(cudaMemcpyAsync(zbuffer, ctx->d_depth, (size_t)vp.vw * vp.vh * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SWEGL_B200_OK;
}

int swegl_b200_set_frame_sync(swegl_b200_ctx *ctx, int rank, int world)
{
    if (!ctx || !ctx->d_screen) return fail(ctx, SWEGL_B200_ERR_STATE, "set_frame_sync before set_screen");
    if (rank >= 0 && (world < 1 || world > SYNC_MAX_RANKS || rank >= world)) return fail(ctx, SWEGL_B200_ERR_ARG, "set_frame_sync: bad rank / world");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx);
    CK(cudaMemset(ctx->own_sync(), 0, sizeof(FrameSync)));
    preload_sync_kernels();
    ctx->sync_rank = rank < 0 ? -1 : rank; ctx->sync_world = rank < 0 ? 0 : world; ctx->sync_seq = 0;
    return SWEGL_B200_OK;
}

int swegl_b200_frame_sync_status(swegl_b200_ctx *ctx, uint32_t *timeouts, float *rank0_own_ms)
{
    if (!ctx || !ctx->d_screen) return fail(ctx, SWEGL_B200_ERR_ARG, "frame_sync_status: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    FrameSync fs;
    CK(cudaMemcpy(&fs, ctx->own_sync(), sizeof fs, cudaMemcpyDeviceToHost));
    if (timeouts) *timeouts = fs.error;
    if (rank0_own_ms) *rank0_own_ms = fs.t_own_end > fs.t_begin ? (float)((double)(fs.t_own_end - fs.t_begin) * 1e-6) : 0.0f;
    return SWEGL_B200_OK;
}

int swegl_b200_set_band_culling(swegl_b200_ctx *ctx, int policy)
{
    if (!ctx || policy < -1 || policy > 1) return fail(ctx, SWEGL_B200_ERR_ARG, "set_band_culling: policy must be -1, 0 or 1");
    if (ctx->have_scene && ctx->cull.n_clusters == 0 && policy != 0)
        return fail(ctx, SWEGL_B200_ERR_STATE, "set_band_culling: the scene was uploaded with culling off; call before upload_scene");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx);                                       // the decision is baked into the captured frames
    ctx->cull_policy = policy;
    return SWEGL_B200_OK;
}

int swegl_b200_selftest_division(swegl_b200_ctx *ctx, uint64_t n_pairs, uint32_t seed, uint64_t out[2])
{
    if (!ctx || !out) return fail(ctx, SWEGL_B200_ERR_ARG, "selftest_division: null argument");
    CK(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    CK(cudaMalloc(&d, 16));
    CK(cudaMemsetAsync(d, 0, 16, ctx->stream));
    launch_selftest_division(n_pairs, seed, d, ctx->stream);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CK(e);
    return SWEGL_B200_OK;
}

int swegl_b200_selftest_filter(swegl_b200_ctx *ctx, uint64_t n_samples, uint32_t seed, uint64_t out[1])
{
    if (!ctx || !out) return fail(ctx, SWEGL_B200_ERR_ARG, "selftest_filter: null argument");
    CK(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    CK(cudaMalloc(&d, 8));
    CK(cudaMemsetAsync(d, 0, 8, ctx->stream));
    launch_selftest_filter(n_samples, seed, d, ctx->stream);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CK(e);
    return SWEGL_B200_OK;
}

int swegl_b200_cull_counts(swegl_b200_ctx *ctx, uint32_t counts[6])
{
    if (!ctx || !counts) return SWEGL_B200_ERR_ARG;
    if (!ctx->have_scene || !ctx->have_vp) return fail(ctx, SWEGL_B200_ERR_STATE, "cull_counts before a rendered view");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const CullTables &ct = ctx->cull;
    counts[0] = ct.n_clusters; counts[1] = ct.n_vblocks; counts[2] = counts[3] = ct.n_clusters; counts[4] = ct.n_vblocks; counts[5] = 0;
    const bool dof = ctx->last_dof;
    if (!view_culled(ctx, draw_params(ctx->last_vp, dof))) return SWEGL_B200_OK;
    std::vector<uint8_t> f((size_t)2 * ct.n_clusters + ct.n_vblocks);
    CK(cudaMemcpy(f.data(), ctx->d_cull_flags, f.size(), cudaMemcpyDeviceToHost));
    counts[2] = counts[3] = counts[4] = 0; counts[5] = 1;
    for (uint32_t c = 0; c < ct.n_clusters; c++) { counts[2] += f[c]; counts[3] += f[ct.n_clusters + c]; }
    for (uint32_t j = 0; j < ct.n_vblocks; j++) counts[4] += f[(size_t)2 * ct.n_clusters + j];
    return SWEGL_B200_OK;
}

int swegl_b200_read_vertices(swegl_b200_ctx *ctx, float *v_world, float *v_viewport, float *normal_world, uint8_t *yes)
{
    if (!ctx || !ctx->have_vp) return fail(ctx, SWEGL_B200_ERR_STATE, "read_vertices: nothing rendered");
    CK(cudaSetDevice(ctx->device));
    const uint32_t nv = ctx->ds.n_vertices;
    cudaStream_t st = ctx->stream;
    std::vector<uint8_t> y(nv);
    if (v_world) CK(cudaMemcpyAsync(v_world, ctx->d_v_world, (size_t)12 * nv, cudaMemcpyDeviceToHost, st));
    if (v_viewport) CK(cudaMemcpyAsync(v_viewport, ctx->d_v_ndc, (size_t)12 * nv, cudaMemcpyDeviceToHost, st));
    if (normal_world) CK(cudaMemcpyAsync(normal_world, ctx->d_n_world, (size_t)12 * nv, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(y.data(), ctx->d_yes, nv, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (v_viewport) {
        // the device keeps NDC and applies frustum_to_viewport on the fly in k_setup; reproduce the
        // reference's in-place state (vertex_shaders.hpp:72-84) for marked vertices.  Same two fp32 ops
        // (this TU is built with -ffp-contract=off on the host side).
        const ViewParams &vp = ctx->last_vp;
        for (uint32_t i = 0; i < nv; i++)
            if (y[i]) {
                volatile float mx = vp.vp_m00 * v_viewport[3 * i];
                volatile float my = vp.vp_m11 * v_viewport[3 * i + 1];
                v_viewport[3 * i] = mx + vp.vp_m03;
                v_viewport[3 * i + 1] = my + vp.vp_m13;
            }
    }
    if (yes) memcpy(yes, y.data(), nv);
    return SWEGL_B200_OK;
}

} // extern "C"
