// fragment.cu — the per-pixel half of fill_half_triangle (renderer.cpp:486-499) and the pixel
// shaders (swegl/render/pixel_shaders.hpp, src/render/pixel_shaders.cpp), plus the DoF-R post pass.
//
// k_fragments: one warp owns one 32-pixel, 1-row bin of the viewport; lane = pixel.  The warp walks the
// bin's chunk list, each lane replays the span interpolator from the chunk checkpoint to its own
// column (<=31 fp32 add pairs, the same additions the CPU does), and keeps the nearest fragment
// in registers: key = depth bits << 32 | slot id, so equal depths resolve to the earlier draw,
// exactly like the serial `if (z >= *zb) continue;` (renderer.cpp:491).  Only the winner is shaded
// (deferred), and the bin is written once as a 128-byte colour segment and a 128-byte depth
// segment; background pixels get the clear values (viewport.cpp:88-113), so there is no clear pass
// and no atomics on the framebuffer.
#include "common.cuh"

namespace sb {

SB_DEV V3 ld3(const float *p) { return v3(p[0], p[1], p[2]); }

template <int TEX>
SB_DEV uint32_t shade_texture(const SlotShade *sh, const Prim &pr, const uint32_t *texels,
                              bool lower, bool lor, float pl, float pr_, float u)
{
    if (TEX == SWEGL_B200_TEX_PLAIN) return pr.color;                       // pixel_shaders.hpp:28
    // long side: t0 + (t2-t0)*p ; short side: upper t0 + (t1-t0)*p, lower t1 + (t2-t1)*p
    float t0x = sh->t0[0], t0y = sh->t0[1], t1x = sh->t1[0], t1y = sh->t1[1], t2x = sh->t2[0], t2y = sh->t2[1];
    float ldx = fsub(t2x, t0x), ldy = fsub(t2y, t0y);                       // side_long_t_dir
    float sbx = lower ? t1x : t0x, sby = lower ? t1y : t0y;                 // side_short_t
    float sdx = lower ? fsub(t2x, t1x) : fsub(t1x, t0x);                    // side_short_t_dir
    float sdy = lower ? fsub(t2y, t1y) : fsub(t1y, t0y);
    float tlx, tly, tdx, tdy;                                               // pixel_shaders.cpp:334-346
    if (lor) {
        tlx = fadd(sbx, fmul(sdx, pl)); tly = fadd(sby, fmul(sdy, pl));
        tdx = fsub(fadd(t0x, fmul(ldx, pr_)), tlx); tdy = fsub(fadd(t0y, fmul(ldy, pr_)), tly);
    } else {
        tlx = fadd(t0x, fmul(ldx, pl)); tly = fadd(t0y, fmul(ldy, pl));
        tdx = fsub(fadd(sbx, fmul(sdx, pr_)), tlx); tdy = fsub(fadd(sby, fmul(sdy, pr_)), tly);
    }
    float tx = fadd(tlx, fmul(tdx, u)), ty = fadd(tly, fmul(tdy, u));
    const uint32_t *bm = texels + pr.tex_off;
    if (TEX == SWEGL_B200_TEX_NEAREST) {
        // pixel_shader_texture::shade, pixel_shaders.cpp:275-281 (unsigned modulo)
        unsigned tw = (unsigned)pr.tw, th = (unsigned)pr.th;
        unsigned uu = (unsigned)f2i(tx) % tw, vv = (unsigned)f2i(ty) % th;
        return __ldg(&bm[vv * tw + uu]);
    }
    // pixel_shader_texture_bilinear::shade, pixel_shaders.cpp:348-384: t.x picks the ROW, t.y the COLUMN
    float v = tx, uq = ty;
    float u1 = fsub(uq, 0.5f), u2 = fadd(uq, 0.5f), v1 = fsub(v, 0.5f), v2 = fadd(v, 0.5f);
    uq = floorf(u2); v = floorf(v2);
    int tw = pr.tw, th = pr.th;
    int v1m = (f2i(v1) + th) % th; if (v1m < 0) { v1m %= th; if (v1m < 0) v1m += th; }   // UB guard
    int v2m = v1m + 1; if (v2m == th) v2m = 0;
    v1m *= tw; v2m *= tw;
    int u1m = (f2i(u1) + tw) % tw; if (u1m < 0) { u1m %= tw; if (u1m < 0) u1m += tw; }
    int u2m = u1m + 1; if (u2m == tw) u2m = 0;
    uint32_t p00 = __ldg(&bm[v1m + u1m]), p10 = __ldg(&bm[v2m + u1m]);
    uint32_t p01 = __ldg(&bm[v1m + u2m]), p11 = __ldg(&bm[v2m + u2m]);
    float w00 = fmul(fsub(uq, u1), fsub(v, v1)), w10 = fmul(fsub(uq, u1), fsub(v2, v));
    float w01 = fmul(fsub(u2, uq), fsub(v, v1)), w11 = fmul(fsub(u2, uq), fsub(v2, v));
    uint32_t out = 0;
    #pragma unroll
    for (int c = 0; c < 4; c++) {
        int sft = 8 * c;
        float acc = fmul((float)((p00 >> sft) & 0xFF), w00);                // pixel_colors * float, colors.cpp:27-30
        acc = fadd(acc, fmul((float)((p10 >> sft) & 0xFF), w10));           // _mm_add_ps, left to right
        acc = fadd(acc, fmul((float)((p01 >> sft) & 0xFF), w01));
        acc = fadd(acc, fmul((float)((p11 >> sft) & 0xFF), w11));
        out |= ((uint32_t)f2i(roundf(acc)) & 0xFFu) << sft;                 // (unsigned char)round(), colors.cpp:19-25
    }
    return out;
}

template <int LIGHT>
SB_DEV int shade_light(const SlotShade *sh, const ViewParams &vp, const FrameParams &fp,
                       bool lower, bool lor, float pl, float pr_, float u)
{
    if (LIGHT == SWEGL_B200_LIGHT_FLAT) return f2i(sh->flat_light);         // pixel_shaders.hpp:36-39
    // pixel_shader_lights_phong: prepare_for_{upper,lower}_triangle (pixel_shaders.cpp:106-151),
    // prepare_for_scanline (:152-158), shade (:159-205)
    V3 w0 = ld3(sh->w0), w1 = ld3(sh->w1), w2 = ld3(sh->w2);
    V3 n0 = ld3(sh->n0), n1 = ld3(sh->n1), n2 = ld3(sh->n2);
    V3 lgb = w0, lgd = sub(w2, w0);                                         // long side
    V3 shb = lower ? w1 : w0, shd = lower ? sub(w2, w1) : sub(w1, w0);      // short side
    V3 nlgb = n0, nlgd = sub(n2, n0);
    V3 nshb = lower ? n1 : n0, nshd = lower ? sub(n2, n1) : sub(n1, n0);
    V3 vl = lor ? shb : lgb, vld = lor ? shd : lgd, vr = lor ? lgb : shb, vrd = lor ? lgd : shd;
    V3 nl = lor ? nshb : nlgb, nld = lor ? nshd : nlgd, nr = lor ? nlgb : nshb, nrd = lor ? nlgd : nshd;
    V3 v = add(vl, mul(vld, pl));
    V3 vdir = sub(add(vr, mul(vrd, pr_)), v);
    V3 n = add(nl, mul(nld, pl));
    V3 ndir = sub(add(nr, mul(nrd, pr_)), n);

    V3 center = add(v, mul(vdir, u));
    V3 normal = normalize(add(n, mul(ndir, u)));
    V3 camv = normalize(sub(v3(vp.cam[0], vp.cam[1], vp.cam[2]), center));
    float sun = -dot(normal, v3(fp.sun[0], fp.sun[1], fp.sun[2]));
    if (sun < 0.0f) sun = 0.0f; else sun = fmul(sun, fp.sun_intensity);
    float dyn = point_lights_sum(fp, center, normal, camv);
    return f2i(fmul(65536.0f, fadd(fadd(fp.ambient, sun), dyn)));
}

template <int LIGHT, int TEX>
SB_DEV uint32_t shade(const SlotShade *sh, const Prim &pr, const uint32_t *texels, const ViewParams &vp,
                      const FrameParams &fp, bool lower, bool lor, float pl, float pr_, float u)
{
    uint32_t c = shade_texture<TEX>(sh, pr, texels, lower, lor, pl, pr_, u);
    if (LIGHT == SWEGL_B200_LIGHT_NONE) return c;
    // pixel_shader_light_and_texture::shade, pixel_shaders.hpp:159-178
    int li = shade_light<LIGHT>(sh, vp, fp, lower, lor, pl, pr_, u);
    float light = fmul(__int2float_rn(li), 1.0f / 65536.0f);               // (float)(li / 65536.0)
    uint32_t b = c & 0xFF, g = (c >> 8) & 0xFF, r = (c >> 16) & 0xFF;
    if (light < 1.0f) {
        b = (uint32_t)f2i(fmul((float)b, light)) & 0xFF;
        g = (uint32_t)f2i(fmul((float)g, light)) & 0xFF;
        r = (uint32_t)f2i(fmul((float)r, light)) & 0xFF;
    } else {
        light = __fsqrt_rn(__fsqrt_rn(light));
        b = (255u - ((uint32_t)f2i(fdiv((float)(255 - (int)b), light)) & 0xFF)) & 0xFF;
        g = (255u - ((uint32_t)f2i(fdiv((float)(255 - (int)g), light)) & 0xFF)) & 0xFF;
        r = (255u - ((uint32_t)f2i(fdiv((float)(255 - (int)r), light)) & 0xFF)) & 0xFF;
    }
    return (c & 0xFF000000u) | (r << 16) | (g << 8) | b;
}

static constexpr int FRAG_TPB = 256;

template <int LIGHT, int TEX>
__global__ void __launch_bounds__(FRAG_TPB) k_fragments(DeviceScene s, const __grid_constant__ ViewParams vp,
                                                        const __grid_constant__ FrameParams fp, Pools pl,
                                                        uint32_t *__restrict__ color, int color_pitch,
                                                        float *__restrict__ depth, int count_covered)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n_bins = (uint32_t)vp.nbx * (uint32_t)(vp.band1 - vp.band0);
    const uint32_t warps = (gridDim.x * FRAG_TPB) >> 5;
    const Span *spans = reinterpret_cast<const Span *>(pl.rows);
    const uint64_t KEY_INIT = (uint64_t)MAXZ_BITS << 32;

    for (uint32_t bin = (blockIdx.x * FRAG_TPB + threadIdx.x) >> 5; bin < n_bins; bin += warps) {
        const int row = (int)(bin / (uint32_t)vp.nbx) + (vp.band0 - vp.vy);      // viewport-relative row
        const int bx = (int)(bin % (uint32_t)vp.nbx);
        const int binx0 = vp.vx + (bx << 5);
        const int x = binx0 + lane, y = vp.vy + row;
        int32_t *headp = pl.bin_head + (size_t)row * vp.nbx + bx;
        int32_t c = *headp;
        if (c >= 0 && lane == 0) *headp = -1;                               // ready for the next frame

        uint64_t best = KEY_INIT;
        float best_u = 0.f;
        uint32_t best_span = 0;
        while (c >= 0) {
            const Chunk ch = pl.chunks[c];
            const Span sp = spans[ch.span];
            const int x1 = (int)(sp.x1x2 & 0xFFFFu), x2 = (int)(sp.x1x2 >> 16);
            const int xs = max(x1, binx0), xe = min(x2, binx0 + 32);
            const int k = x - xs, n = xe - xs;
            float top = ch.top, bot = ch.bottom;
            for (int j = 0; j < n - 1; j++)                                 // qpixel.Step() x (x - xs)
                if (j < k) { top = fadd(top, sp.topstep); bot = fadd(bot, sp.bottomstep); }
            const float u = fdiv(top, bot);                                 // interpolator.hpp:98
            const float z = fadd(sp.v0, fmul(sp.v1, u));                    // value(0)
            if (k >= 0 && k < n && z >= NEAR_Z) {                           // renderer.cpp:488-492
                uint64_t key = ((uint64_t)__float_as_uint(z) << 32) | (sp.slot_flags >> 2);
                if (key < best) { best = key; best_u = u; best_span = ch.span; }
            }
            c = ch.next;
        }

        const bool inside = x < vp.vx + vp.vw;
        const bool hit = best != KEY_INIT;
        uint32_t out = 0;                                                   // background, viewport.cpp:95-103
        if (hit) {
            const Span sp = spans[best_span];
            const uint32_t slot = sp.slot_flags >> 2;
            const bool lower = (sp.slot_flags >> 1) & 1u, lor = sp.slot_flags & 1u;
            const SlotShade *sh = &pl.shades[slot];
            const Prim pr = s.prims[sh->prim];
            out = shade<LIGHT, TEX>(sh, pr, s.texels, vp, fp, lower, lor, sp.pl, sp.pr, best_u);
        }
        if (inside) {
            color[(size_t)y * color_pitch + x] = out;
            depth[(size_t)row * vp.vw + (x - vp.vx)] = __uint_as_float((uint32_t)(best >> 32));
        }
        if (count_covered) {
            unsigned m = __ballot_sync(0xFFFFFFFFu, hit && inside);
            if (lane == 0 && m) atomicAdd(&pl.counters->n_covered, (uint32_t)__popc(m));
        }
    }
}

// ----------------------------------------------------------------------------------------
// DoF-R: the repaired post_shader_depth_box (post_shaders.hpp:63-111).
// 32x8 output tile per CTA, (32+10)x(8+10) source tile staged in shared memory as
// colour | (blur != 0) << 24 so a tap costs one LDS.
// ----------------------------------------------------------------------------------------
static constexpr int DOF_TX = 32, DOF_TY = 8, DOF_R = 5;
static constexpr int DOF_SW = DOF_TX + 2 * DOF_R, DOF_SH = DOF_TY + 2 * DOF_R;

SB_DEV float blur_factor(float depth, float focal_distance, float focal_depth)
{
    // remap_clipped(1.0f, focal_depth, 0.0f, 5.0f, |focal_distance - z|), lerp.hpp:24-43
    float t = fabsf(fsub(focal_distance, depth));
    float a = 1.0f, b = focal_depth, xq;
    if (a == b) xq = 0.5f; else if (t <= a) xq = 0.0f; else if (t >= b) xq = 1.0f; else xq = fdiv(fsub(t, a), fsub(b, a));
    float r;
    if (xq <= 0.0f) r = 0.0f; else if (xq >= 5.0f) r = 5.0f; else r = fadd(0.0f, fmul(5.0f, xq));
    return r;
}

__global__ void __launch_bounds__(DOF_TX * DOF_TY) k_dof(const uint32_t *__restrict__ src, int src_pitch,
                                                         const float *__restrict__ depth, uint32_t *__restrict__ dst,
                                                         int dst_pitch, int w, int h, int row0, int row1,
                                                         float focal_distance, float focal_depth)
{
    __shared__ uint32_t tile[DOF_SH][DOF_SW];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int ox = blockIdx.x * DOF_TX, oy = row0 + blockIdx.y * DOF_TY;
    for (int i = ty * DOF_TX + tx; i < DOF_SW * DOF_SH; i += DOF_TX * DOF_TY) {
        int sy = i / DOF_SW, sx = i % DOF_SW;
        int gx = ox + sx - DOF_R, gy = oy + sy - DOF_R;
        uint32_t v = 0;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
            uint32_t c = src[(size_t)gy * src_pitch + gx];
            float bf = blur_factor(depth[(size_t)gy * w + gx], focal_distance, focal_depth);
            v = (c & 0x00FFFFFFu) | (bf != 0.0f ? 0x01000000u : 0u);
        }
        tile[sy][sx] = v;
    }
    __syncthreads();
    const int x = ox + tx, y = oy + ty;
    if (x >= w || y >= row1 || y >= h) return;
    const uint32_t own = src[(size_t)y * src_pitch + x];
    const int radius = f2i(blur_factor(depth[(size_t)y * w + x], focal_distance, focal_depth));
    uint32_t out = own;
    if (radius != 0) {
        int b = 0, g = 0, r = 0, count = 0;
        const int j0 = max(0, y - radius), j1 = min(h, y + radius);
        const int i0 = max(0, x - radius), i1 = min(w, x + radius);
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
                uint32_t p = tile[j - oy + DOF_R][i - ox + DOF_R];
                if (p >> 24) { count++; b += p & 0xFF; g += (p >> 8) & 0xFF; r += (p >> 16) & 0xFF; }
            }
        if (count) out = (uint32_t)(b / count) | ((uint32_t)(g / count) << 8) | ((uint32_t)(r / count) << 16) | 0xFF000000u;
    }
    dst[(size_t)y * dst_pitch + x] = out;
}

// ----------------------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------------------
template <int LIGHT, int TEX>
static void launch_frag_t(const DeviceScene &s, const ViewParams &vp, const FrameParams &fp, const Pools &p,
                          uint32_t *color, int color_pitch, float *depth, bool count_covered, cudaStream_t st)
{
    uint32_t n_bins = (uint32_t)vp.nbx * (uint32_t)(vp.band1 - vp.band0);
    uint32_t blocks = (n_bins + (FRAG_TPB / 32) - 1) / (FRAG_TPB / 32);
    if (!blocks) return;
    k_fragments<LIGHT, TEX><<<blocks, FRAG_TPB, 0, st>>>(s, vp, fp, p, color, color_pitch, depth, count_covered ? 1 : 0);
}

void launch_fragments(const DeviceScene &s, const ViewParams &vp, const FrameParams &fp, const Pools &p,
                      uint32_t *color, int color_pitch, float *depth, bool count_covered, cudaStream_t st)
{
#define SB_CASE(L, T) if (vp.light_mode == L && vp.tex_mode == T) { launch_frag_t<L, T>(s, vp, fp, p, color, color_pitch, depth, count_covered, st); return; }
    SB_CASE(0, 0) SB_CASE(0, 1) SB_CASE(0, 2)
    SB_CASE(1, 0) SB_CASE(1, 1) SB_CASE(1, 2)
    SB_CASE(2, 0) SB_CASE(2, 1) SB_CASE(2, 2)
#undef SB_CASE
}

void launch_dof(const uint32_t *src, int src_pitch, const float *depth, uint32_t *dst, int dst_pitch,
                int w, int h, int row0, int row1, float focal_distance, float focal_depth, cudaStream_t st)
{
    dim3 block(DOF_TX, DOF_TY);
    dim3 grid((w + DOF_TX - 1) / DOF_TX, (row1 - row0 + DOF_TY - 1) / DOF_TY);
    if (grid.x && grid.y)
        k_dof<<<grid, block, 0, st>>>(src, src_pitch, depth, dst, dst_pitch, w, h, row0, row1, focal_distance, focal_depth);
}

} // namespace sb
