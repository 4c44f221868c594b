// fragment.cu — the per-pixel half of fill_half_triangle (renderer.cpp:486-499) and the pixel
// shaders (swegl/render/pixel_shaders.hpp, src/render/pixel_shaders.cpp), plus the DoF-R post pass.
//
// k_fragments: a warp owns a stretch of 32-pixel, 1-row bins of the viewport; lane = pixel.  The warp walks a
// bin's chunk list, each lane reads its pixel's interpolator progress from the fragment stream k_spans
// wrote (the same fp32 additions and division the CPU does), and keeps the nearest fragment
// in registers: key = depth bits << 32 | slot id, so equal depths resolve to the earlier draw,
// exactly like the serial `if (z >= *zb) continue;` (renderer.cpp:491).  Only the winner is shaded
// (deferred), and the bin is written once as a 128-byte colour segment and a 128-byte depth
// segment; background pixels get the clear values (viewport.cpp:88-113), so there is no clear pass
// and no atomics on the framebuffer.
#include "common.cuh"

namespace sb {

SB_DEV V3 ld3(const float *p) { return v3(p[0], p[1], p[2]); }

// Conversions between small integers and floats without the XU pipe.  I2F / F2I / FRND issue at a quarter of the FP32
// rate on sm_100 and the bilinear filter alone needs 16 + 8 of them per pixel (ncu: XU pipe 88 % busy in k_fragments
// before this); for values below 2^22 the same results come out of the FP32 adder:
//   2^23 + n is exact for an integer 0 <= n < 2^23 and its low mantissa bits ARE n.
static constexpr float MAGIC23 = 8388608.0f;            // 2^23, bits 0x4B000000
// (float)byte k of `word`: one PRMT builds the bits of 2^23 + byte, one exact subtraction removes the 2^23
template <int K> SB_DEV float byte_to_float(uint32_t word)
{
    return __fsub_rn(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440u | K)), MAGIC23);
}
SB_DEV float small_uint_to_float(uint32_t n)            // n < 2^23
{
    return __fsub_rn(__uint_as_float(0x4B000000u | n), MAGIC23);
}
// trunc(a) for 0 <= a < 2^22 as the low mantissa bits of RZ(a + 2^23); the caller masks what it needs
SB_DEV uint32_t trunc_bits_small(float a) { return __float_as_uint(__fadd_rz(a, MAGIC23)); }

// (unsigned char)round(x), colors.cpp:19-25 (round half away from zero), for the common case 0 <= x < 2^22
// as trunc + exact fraction test; anything else takes roundf
SB_DEV uint32_t round_to_byte(float a)
{
    if (!(a >= 0.0f && a < 4194304.0f)) return (uint32_t)f2i(roundf(a)) & 0xFFu;
    const float r = __fadd_rz(a, MAGIC23);                                  // 2^23 + trunc(a)
    const float fr = fsub(a, __fsub_rn(r, MAGIC23));                        // exact
    return (__float_as_uint(r) + (fr >= 0.5f ? 1u : 0u)) & 0xFFu;
}
// floorf(x) for |x| < 2^22: RD(x + 1.5 * 2^23) lies in [2^23, 2^24) where the spacing is 1
SB_DEV float floor_small(float x)
{
    if (fabsf(x) < 4194304.0f) return __fsub_rn(__fadd_rd(x, 12582912.0f), 12582912.0f);
    return floorf(x);
}
// a mod n as the reference computes texel rows/columns, with the negative (UB in the reference) case wrapped
SB_DEV int wrap_index(int a, int n, int mask)
{
    if (mask >= 0) return a & mask;                                         // power-of-two size
    int r = a % n; if (r < 0) r += n;
    return r;
}

// The texture part of a pixel in two steps, so that the four texel loads of the bilinear filter can be in flight while
// the lighting arithmetic (which does not need them) runs: tex_fetch computes the addresses and issues the loads,
// tex_filter weighs the texels.  (Texels come from DRAM more often than not: a 16 MB texture, no mip maps.)
struct TexFetch { uint32_t p00, p10, p01, p11; float fu, fv; };     // the four texels + the texture coordinate (column, row) the weights derive from

template <int TEX>
SB_DEV TexFetch tex_fetch(const SpanShade *ss, const Prim &pr, const uint32_t *texels, float u)
{
    TexFetch f;
    f.p00 = pr.color; f.p10 = f.p01 = f.p11 = 0u; f.fu = f.fv = 0.f;
    if (TEX == SWEGL_B200_TEX_PLAIN) return f;                              // pixel_shaders.hpp:28
    // t = t_left + t_dir * progress   (pixel_shaders.cpp:277, 352)
    float tx = fadd(ss->t_left[0], fmul(ss->t_dir[0], u)), ty = fadd(ss->t_left[1], fmul(ss->t_dir[1], u));
    const uint32_t *bm = texels + pr.tex_off;
    if (TEX == SWEGL_B200_TEX_NEAREST) {
        // pixel_shader_texture::shade, pixel_shaders.cpp:275-281 (unsigned modulo)
        unsigned tw = (unsigned)pr.tw, th = (unsigned)pr.th;
        unsigned uu = pr.tw_mask >= 0 ? ((unsigned)f2i(tx) & (unsigned)pr.tw_mask) : (unsigned)f2i(tx) % tw;
        unsigned vv = pr.th_mask >= 0 ? ((unsigned)f2i(ty) & (unsigned)pr.th_mask) : (unsigned)f2i(ty) % th;
        f.p00 = __ldg(&bm[vv * tw + uu]);
        return f;
    }
    // pixel_shader_texture_bilinear::shade, pixel_shaders.cpp:348-384: t.x picks the ROW, t.y the COLUMN
    const float v1 = fsub(tx, 0.5f), u1 = fsub(ty, 0.5f);
    int tw = pr.tw, th = pr.th;
    int v1m = wrap_index(f2i(v1) + th, th, pr.th_mask);                     // ((int)v1 + theight) % theight; UB guard
    int v2m = v1m + 1; if (v2m == th) v2m = 0;
    v1m *= tw; v2m *= tw;
    int u1m = wrap_index(f2i(u1) + tw, tw, pr.tw_mask);
    int u2m = u1m + 1; if (u2m == tw) u2m = 0;
    f.p00 = __ldg(&bm[v1m + u1m]); f.p10 = __ldg(&bm[v2m + u1m]);
    f.p01 = __ldg(&bm[v1m + u2m]); f.p11 = __ldg(&bm[v2m + u2m]);
    f.fu = ty; f.fv = tx;                                                   // the filter recomputes its weights from t
    return f;
}

template <int TEX>
SB_DEV uint32_t tex_filter(const TexFetch &f)
{
    if (TEX != SWEGL_B200_TEX_BILINEAR) return f.p00;
    float v = f.fv, uq = f.fu;
    float u1 = fsub(uq, 0.5f), u2 = fadd(uq, 0.5f), v1 = fsub(v, 0.5f), v2 = fadd(v, 0.5f);
    uq = floor_small(u2); v = floor_small(v2);
    const uint32_t p00 = f.p00, p10 = f.p10, p01 = f.p01, p11 = f.p11;
    float w00 = fmul(fsub(uq, u1), fsub(v, v1)), w10 = fmul(fsub(uq, u1), fsub(v2, v));
    float w01 = fmul(fsub(u2, uq), fsub(v, v1)), w11 = fmul(fsub(u2, uq), fsub(v2, v));
    uint32_t out = 0;
    #define SB_BILINEAR_CHANNEL(C) { \
        float acc = fmul(byte_to_float<C>(p00), w00);                       /* pixel_colors * float, colors.cpp:27-30 */ \
        acc = fadd(acc, fmul(byte_to_float<C>(p10), w10));                  /* _mm_add_ps, left to right */ \
        acc = fadd(acc, fmul(byte_to_float<C>(p01), w01)); \
        acc = fadd(acc, fmul(byte_to_float<C>(p11), w11)); \
        out |= round_to_byte(acc) << (8 * C); }                             /* (unsigned char)round(), colors.cpp:19-25 */
    SB_BILINEAR_CHANNEL(0) SB_BILINEAR_CHANNEL(1) SB_BILINEAR_CHANNEL(2) SB_BILINEAR_CHANNEL(3)
    #undef SB_BILINEAR_CHANNEL
    return out;
}

template <int TEX>
SB_DEV uint32_t shade_texture(const SpanShade *ss, const Prim &pr, const uint32_t *texels, float u)
{
    return tex_filter<TEX>(tex_fetch<TEX>(ss, pr, texels, u));
}

template <int LIGHT>
SB_DEV int shade_light(const SpanShade *ss, float flat_light, const ViewParams &vp, const FrameParams &fp, float u)
{
    if (LIGHT == SWEGL_B200_LIGHT_FLAT) return f2i(flat_light);             // pixel_shaders.hpp:36-39
    // pixel_shader_lights_phong::shade (pixel_shaders.cpp:159-205) on the span constants of prepare_for_scanline
    const V3 v = ld3(ss->v), vdir = ld3(ss->vdir), n = ld3(ss->n), ndir = ld3(ss->ndir);

    V3 center = add(v, mul(vdir, u));
    V3 normal = normalize(add(n, mul(ndir, u)));
    V3 camv = normalize(sub(v3(vp.cam[0], vp.cam[1], vp.cam[2]), center));
    float sun = -dot(normal, v3(fp.sun[0], fp.sun[1], fp.sun[2]));
    if (sun < 0.0f) sun = 0.0f; else sun = fmul(sun, fp.sun_intensity);
    float dyn = point_lights_sum(fp, center, normal, camv);
    return f2i(fmul(65536.0f, fadd(fadd(fp.ambient, sun), dyn)));
}

template <int LIGHT, int TEX>
SB_DEV uint32_t shade(const SpanShade *ss, float flat_light, const Prim &pr, const uint32_t *texels, const ViewParams &vp,
                      const FrameParams &fp, float u)
{
    const TexFetch tf = tex_fetch<TEX>(ss, pr, texels, u);                  // texel loads issued ...
    if (LIGHT == SWEGL_B200_LIGHT_NONE) return tex_filter<TEX>(tf);
    // pixel_shader_light_and_texture::shade, pixel_shaders.hpp:159-178
    int li = shade_light<LIGHT>(ss, flat_light, vp, fp, u);                 // ... in flight under the lighting arithmetic ...
    uint32_t c = tex_filter<TEX>(tf);                                       // ... consumed here
    float light = fmul(__int2float_rn(li), 1.0f / 65536.0f);               // (float)(li / 65536.0)
    uint32_t b = c & 0xFF, g = (c >> 8) & 0xFF, r = (c >> 16) & 0xFF;
    if (light < 1.0f) {
        if (light >= 0.0f) {
            // 0 <= c * light < 255: truncation through the adder (see trunc_bits_small), no I2F / F2I
            b = trunc_bits_small(fmul(small_uint_to_float(b), light)) & 0xFF;
            g = trunc_bits_small(fmul(small_uint_to_float(g), light)) & 0xFF;
            r = trunc_bits_small(fmul(small_uint_to_float(r), light)) & 0xFF;
        } else {
            // |light| <= 32768 and c <= 255, so the product is always inside int range: plain truncation
            b = (uint32_t)__float2int_rz(fmul((float)b, light)) & 0xFF;
            g = (uint32_t)__float2int_rz(fmul((float)g, light)) & 0xFF;
            r = (uint32_t)__float2int_rz(fmul((float)r, light)) & 0xFF;
        }
    } else {
        light = __fsqrt_rn(__fsqrt_rn(light));                              // light >= 1: 0 <= (255 - c) / light <= 255
        const SharedDivisor dl = shared_divisor(light);
        b = (255u - (trunc_bits_small(div_by(small_uint_to_float(255u - b), dl)) & 0xFF)) & 0xFF;
        g = (255u - (trunc_bits_small(div_by(small_uint_to_float(255u - g), dl)) & 0xFF)) & 0xFF;
        r = (255u - (trunc_bits_small(div_by(small_uint_to_float(255u - r), dl)) & 0xFF)) & 0xFF;
    }
    return (c & 0xFF000000u) | (r << 16) | (g << 8) | b;
}


// chunks of a bin whose records are fetched up front, all bins of the stretch at once; longer lists continue serially
#ifndef FRAG_LIST_CAP_V
#define FRAG_LIST_CAP_V 12
#endif
static constexpr int FRAG_LIST_CAP = FRAG_LIST_CAP_V;

// per-warp staging: the stretch's chunk records (pass 1) and the queue of winning fragments (pass 1 -> 2 -> 3)
struct FragWarp {
    float u[FRAG_STRETCH * 32];         // interpolator progress of the winner; overwritten by its colour in pass 2
    uint32_t span[FRAG_STRETCH * 32];   // winning span, 0xFFFFFFFF = background
    uint32_t slot[FRAG_STRETCH * 32];   // winning slot (-> SlotShade)
    uint4 rec[FRAG_STRETCH * FRAG_LIST_CAP * 2];    // Chunk records of bin b at [b * FRAG_LIST_CAP ..], 2 x 16 B each
    int32_t cnt[FRAG_STRETCH], cursor[FRAG_STRETCH];   // staged list lengths, continuation of longer lists
    uint8_t idx[FRAG_STRETCH * 32];     // compacted list of covered pixels
};

// One warp owns a stretch of FRAG_STRETCH bins (256 pixels) of one scanline:
//   1. one load fetches the bin heads (and resets them for the next frame),
//   2. empty bins are cleared in bulk with 128-bit stores (colour 0, depth 0x7F7F7F7F: viewport.cpp:88-113),
//   3. each non-empty bin is resolved with lane = pixel as described at the top of this file.
// the part of ViewParams that is fixed for a captured frame graph (rectangle, band), passed by value so that the
// first loads of the kernel do not wait for the parameter block
struct FragGeom { int32_t vx, vy, vw, band0, band1, nbx, ntx, n_tiles, skip_bg; };   // skip_bg: background colour is not stored (FrameSync)

// clear values (viewport.cpp:88-113) for one row of a tile: colour 0, depth 0x7F7F7F7F
SB_DEV void clear_tile_row(uint32_t *crow, float *drow, int px_left, int lane, bool with_color)
{
    const float maxz = __uint_as_float(MAXZ_BITS);
    const bool aligned = ((reinterpret_cast<uintptr_t>(crow) | reinterpret_cast<uintptr_t>(drow)) & 15) == 0;
    #pragma unroll
    for (int px = lane << 2; px < FRAG_STRETCH * 32; px += 128) {
        if (aligned && px + 4 <= px_left) {
            if (with_color) *reinterpret_cast<uint4 *>(crow + px) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<float4 *>(drow + px) = make_float4(maxz, maxz, maxz, maxz);
        } else {
            #pragma unroll 1
            for (int k = 0; k < 4; k++)
                if (px + k < px_left) { if (with_color) crow[px + k] = 0; drow[px + k] = maxz; }
        }
    }
}

// Grid: n_tiles "busy" CTAs followed by ceil(n_tiles / FRAG_ROWS) "clear" CTAs.
//  * busy CTA i works on busy_list[i] (the tiles k_spans put chunks into, so all the expensive tiles start at once
//    at the head of the grid instead of wherever the scene happens to sit on the screen); i >= n_busy exits;
//  * clear CTA j: warp w takes tile j * FRAG_ROWS + w and, unless it was busy, fills it with the clear values
//    (16 independent 128-bit stores per lane; no bin heads are read for tiles nothing was drawn into).
#ifndef FRAG_MINB
#define FRAG_MINB 8         // 32 registers: all 64 warps of an SM resident; measured faster than 48 registers / 40 warps
#endif
template <int LIGHT, int TEX>
__global__ void __launch_bounds__(FRAG_TPB, FRAG_MINB) k_fragments(DeviceScene s, const ViewParams *__restrict__ vpp,
                                                        const FrameParams *__restrict__ fpp, Pools pl, FragGeom g,
                                                        uint32_t *__restrict__ color, int color_pitch,
                                                        float *__restrict__ depth, int count_covered,
                                                        Counters *__restrict__ h_counters_out)
{
    static_assert(FRAG_ROWS * FRAG_STRETCH <= 32, "one warp-wide load fetches the bin heads of the whole CTA tile");
    __shared__ ViewParams vp;                   // per-frame constants, staged only by tiles that shade something
    __shared__ FrameParams fp;
    __shared__ FragWarp fwarp[FRAG_ROWS];
    __shared__ int32_t s_head[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int band_rows = g.band1 - g.band0;
    pdl_trigger();
    pdl_wait();                                                             // k_spans' bins, chunks, fragment stream, tile list
    if ((int)blockIdx.x >= g.n_tiles) {
        // ---- clear CTA ----
        const int t = ((int)blockIdx.x - g.n_tiles) * FRAG_ROWS + warp;
        if (t >= g.n_tiles) return;
        if (pl.tile_stamp[t] == vpp->stamp) return;                         // a busy CTA owns this tile
        const int ty = t / g.ntx, tx = t - ty * g.ntx;
        const int row0 = (g.band0 - g.vy) + ty * FRAG_ROWS, bx0 = tx * FRAG_STRETCH;
        const int rows_here = min(FRAG_ROWS, band_rows - ty * FRAG_ROWS);
        const int px_left = g.vw - (bx0 << 5);
        for (int r = 0; r < rows_here; r++)
            clear_tile_row(color + (size_t)(g.vy + row0 + r) * color_pitch + g.vx + (bx0 << 5),
                           depth + (size_t)(row0 + r) * g.vw + (bx0 << 5), px_left, lane, !g.skip_bg);
        {   // occupancy map for the DoF pass
            const int r = lane / FRAG_STRETCH, b = lane % FRAG_STRETCH;
            if (r < rows_here && bx0 + b < g.nbx) pl.bin_used[(size_t)(row0 + r) * g.nbx + bx0 + b] = 0;
        }
        return;
    }
    // ---- busy CTA ----
    // k_setup / k_spans are done: publish their counters (pool demand, overflow flags) to the pinned slot the host
    // polls, instead of a D2H copy node at the end of the graph.  (n_covered is only final after this kernel; the
    // synchronous stats path copies the counters itself.)
    if (h_counters_out && blockIdx.x == 0 && threadIdx.x < sizeof(Counters) / 4)
        reinterpret_cast<uint32_t *>(h_counters_out)[threadIdx.x] = reinterpret_cast<const uint32_t *>(pl.counters)[threadIdx.x];
    if (blockIdx.x >= pl.counters->n_busy) return;
    const int t = (int)pl.busy_list[blockIdx.x];
    const int ty = t / g.ntx, tx = t - ty * g.ntx;
    const int row0 = (g.band0 - g.vy) + ty * FRAG_ROWS;                     // viewport-relative first row of the tile
    const int rows_here = min(FRAG_ROWS, band_rows - ty * FRAG_ROWS);
    const int bx0 = tx * FRAG_STRETCH;
    const int nb = min(FRAG_STRETCH, g.nbx - bx0);
    const int px_left = g.vw - (bx0 << 5);                                  // pixels from the stretch start to the row end
    const uint64_t KEY_INIT = (uint64_t)MAXZ_BITS << 32;
    if (warp == 0) {        // lane -> (row, bin) of the tile: fetch and reset the heads, leave the occupancy map for the DoF pass
        const int r = lane / FRAG_STRETCH, b = lane % FRAG_STRETCH;
        int32_t h = -1;
        if (r < rows_here && b < nb) {
            const size_t bi = (size_t)(row0 + r) * g.nbx + bx0 + b;
            h = pl.bin_head[bi];
            if (h >= 0) pl.bin_head[bi] = -1;
            pl.bin_used[bi] = h >= 0;
        }
        s_head[lane] = h;
    }
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += FRAG_TPB) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    for (int w = threadIdx.x; w < (int)(sizeof(FrameParams) / 4); w += FRAG_TPB) reinterpret_cast<uint32_t *>(&fp)[w] = reinterpret_cast<const uint32_t *>(fpp)[w];
    __syncthreads();
    if (warp >= rows_here) return;
    const int row = row0 + warp;
    const int y = g.vy + row;
    const int32_t head = lane < FRAG_STRETCH ? s_head[warp * FRAG_STRETCH + lane] : -1;
    unsigned mask = __ballot_sync(0xFFFFFFFFu, head >= 0);

    uint32_t *crow = color + (size_t)y * color_pitch + g.vx + (bx0 << 5);
    float *drow = depth + (size_t)row * g.vw + (bx0 << 5);
    const bool aligned = ((reinterpret_cast<uintptr_t>(crow) | reinterpret_cast<uintptr_t>(drow)) & 15) == 0;
    // ---- empty bins ----
    if (mask != 0xFFFFFFFFu) {
        const float maxz = __uint_as_float(MAXZ_BITS);
        for (int i = lane; i < nb * 8; i += 32) {
            const int b = i >> 3, px = (b << 5) + ((i & 7) << 2);
            if ((mask >> b) & 1u) continue;
            if (aligned && px + 4 <= px_left) {
                if (!g.skip_bg) *reinterpret_cast<uint4 *>(crow + px) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<float4 *>(drow + px) = make_float4(maxz, maxz, maxz, maxz);
            } else {
                for (int k = 0; k < 4; k++)
                    if (px + k < px_left) { if (!g.skip_bg) crow[px + k] = 0; drow[px + k] = maxz; }
            }
        }
    }
    // ---- non-empty bins, pass 1: depth resolve.
    //  a. the first lanes walk one bin list each (FRAG_STRETCH pointer chases side by side instead of one after the
    //     other) and leave the chunk records in shared memory;
    //  b. lane = pixel: per bin, the nearest fragment of every pixel is kept in registers.  Depth is written at once,
    //     and the winners of the whole stretch are queued in shared memory so that shading (pass 2) runs on dense
    //     batches of 32 covered pixels. ----
    FragWarp &fw = fwarp[threadIdx.x >> 5];
    if (lane < FRAG_STRETCH) {
        int n = 0;
        int32_t c = head;
        while (c >= 0 && n < FRAG_LIST_CAP) {
            const uint4 *src = reinterpret_cast<const uint4 *>(&pl.chunks[c]);
            const uint4 r0 = src[0], r1 = src[1];
            fw.rec[2 * (lane * FRAG_LIST_CAP + n)] = r0; fw.rec[2 * (lane * FRAG_LIST_CAP + n) + 1] = r1;
            n++;
            c = (int32_t)r1.z;                                                  // Chunk::next
        }
        fw.cnt[lane] = n; fw.cursor[lane] = c;
    }
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    uint32_t n_hit = 0;
    const unsigned used = mask;
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        uint64_t best = KEY_INIT;
        float best_u = 0.f;
        uint32_t best_span = 0xFFFFFFFFu;
        const int n = fw.cnt[b];
        const uint4 *rec = &fw.rec[2 * b * FRAG_LIST_CAP];
        #pragma unroll 2
        for (int k = 0; k < n; k++) {
            const uint4 r0 = rec[2 * k];                                        // frag0, xs_xe, v0, v1
            const unsigned xs = r0.y & 0xFFu, wd = (r0.y >> 8) - xs;
            if ((unsigned)lane - xs < wd) {
                const float u = pl.frag_u[r0.x + (uint32_t)lane];               // qpixel.ualpha, replayed by k_spans
                const float z = fadd(__uint_as_float(r0.z), fmul(__uint_as_float(r0.w), u));   // value(0), renderer.cpp:488
                if (z >= NEAR_Z) {                                              // renderer.cpp:489-492
                    const uint2 r1 = *reinterpret_cast<const uint2 *>(&rec[2 * k + 1]);       // slot, span
                    const uint64_t key = ((uint64_t)__float_as_uint(z) << 32) | r1.x;
                    if (key < best) { best = key; best_u = u; best_span = r1.y; }
                }
            }
        }
        for (int32_t c = fw.cursor[b]; c >= 0;) {                               // a list longer than FRAG_LIST_CAP: the rest, serially
            const Chunk ch = pl.chunks[c];
            c = ch.next;
            const int xs = (int)(ch.xs_xe & 0xFFu), xe = (int)(ch.xs_xe >> 8);
            if (lane >= xs && lane < xe) {
                const float u = pl.frag_u[ch.frag0 + (uint32_t)lane];
                const float z = fadd(ch.v0, fmul(ch.v1, u));
                if (z >= NEAR_Z) {
                    const uint64_t key = ((uint64_t)__float_as_uint(z) << 32) | ch.slot;
                    if (key < best) { best = key; best_u = u; best_span = ch.span; }
                }
            }
        }
        const int px = (b << 5) + lane;
        const bool inside = px < px_left;
        const bool hit = best_span != 0xFFFFFFFFu && inside;
        if (inside) drow[px] = __uint_as_float((uint32_t)(best >> 32));
        fw.u[px] = best_u; fw.span[px] = hit ? best_span : 0xFFFFFFFFu; fw.slot[px] = (uint32_t)best;
        const unsigned hm = __ballot_sync(0xFFFFFFFFu, hit);
        if (hit) fw.idx[n_hit + __popc(hm & lt)] = (uint8_t)px;
        n_hit += __popc(hm);
    }
    __syncwarp();
    // ---- pass 2: shade the queued winners, 32 at a time (deferred: only the visible fragment of a pixel is shaded) ----
    for (uint32_t k = lane; k < n_hit; k += 32) {
        const int px = fw.idx[k];
        const uint32_t span = fw.span[px];
        const SlotShade *sh = &pl.shades[fw.slot[px]];
        const uint4 bind = *reinterpret_cast<const uint4 *>(&sh->color);       // colour, tex_off, tw, th
        Prim pr;
        pr.color = bind.x; pr.tex_off = bind.y; pr.tw = (int32_t)bind.z; pr.th = (int32_t)bind.w;
        pr.tw_mask = (pr.tw & (pr.tw - 1)) == 0 ? pr.tw - 1 : -1;
        pr.th_mask = (pr.th & (pr.th - 1)) == 0 ? pr.th - 1 : -1;
        const uint32_t out = shade<LIGHT, TEX>(&pl.span_shades[span], sh->flat_light, pr, s.texels, vp, fp, fw.u[px]);
        fw.u[px] = __uint_as_float(out);
    }
    __syncwarp();
    // ---- pass 3: colour write-back, one 128-byte segment per bin (background pixels of a used bin get 0) ----
    mask = used;
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const int px = (b << 5) + lane;
        if (px < px_left) crow[px] = fw.span[px] != 0xFFFFFFFFu ? __float_as_uint(fw.u[px]) : 0u;
    }
    if (count_covered && lane == 0 && n_hit) atomicAdd(&pl.counters->n_covered, n_hit);
}

// ----------------------------------------------------------------------------------------
// Transparency layers (viewport_t::m_got_transparency; renderer.cpp:500-550, viewport.cpp:43-86).
//
// The reference keeps, per pixel, the opaque colour/z and up to L transparent fragments sorted
// far -> near, updated fragment by fragment in draw order:
//   * a fragment must pass `z < zbuffer` (the OPAQUE depth so far) to be looked at all,
//   * an opaque one (shaded alpha == 255) replaces the base and drops every layer with layer_z >= z,
//   * a transparent one is inserted by depth (equal depth: the later draw is nearer); when all L layers
//     are used the farthest is discarded.
// The end state does not depend on the draw order except through ties, so it has a closed form that a
// parallel resolve can evaluate: base = min (z, slot) over opaque fragments; layers = the L nearest
// transparent fragments under the order (z ascending, slot descending) among those with z < base z.
// (A fragment discarded for capacity is farther than every kept one, so it would also fall to any later
// opaque fragment that removes a kept one; and a transparent fragment that failed the z test at its
// time has z >= the final base z.)  flatten() then blends layer[1..] onto layer[0] and the result onto
// the screen with colors.cpp:39-57's blend(), which is exact integer arithmetic.
//
// lane = pixel, as in k_fragments, without the shading queue (every kept layer is shaded as well).
// ----------------------------------------------------------------------------------------
static constexpr int MAX_LAYERS = 8;

SB_DEV uint32_t blend_px(uint32_t back, uint32_t front)                    // colors.cpp:39-57
{
    const int alpha = (int)(front >> 24), back_a = (int)(back >> 24);
    const uint32_t new_alpha = (uint32_t)(255 - ((255 - alpha) * (255 - back_a) / 255));
    if (alpha == 255) return (front & 0x00FFFFFFu) | (new_alpha << 24);
    if (alpha == 0) return (back & 0x00FFFFFFu) | (new_alpha << 24);
    uint32_t out = new_alpha << 24;
    #pragma unroll
    for (int c = 0; c < 3; c++) {
        // (unsigned char)(back * ((256 - alpha) / 256.0) + front * (alpha / 256.0)): every term is a multiple of
        // 1/256 below 2^16, so the double expression is exact and the truncation is a shift
        const uint32_t bk = (back >> (8 * c)) & 0xFFu, fr = (front >> (8 * c)) & 0xFFu;
        out |= (((bk * (uint32_t)(256 - alpha) + fr * (uint32_t)alpha) >> 8) & 0xFFu) << (8 * c);
    }
    return out;
}

template <int LIGHT, int TEX>
__global__ void __launch_bounds__(FRAG_TPB) k_fragments_layers(DeviceScene s, const ViewParams *__restrict__ vpp,
                                                               const FrameParams *__restrict__ fpp, Pools pl,
                                                               uint32_t *__restrict__ color, int color_pitch,
                                                               float *__restrict__ depth, int count_covered,
                                                               Counters *__restrict__ h_counters_out)
{
    pdl_trigger();
    pdl_wait();
    if (h_counters_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < sizeof(Counters) / 4)
        reinterpret_cast<uint32_t *>(h_counters_out)[threadIdx.x] = reinterpret_cast<const uint32_t *>(pl.counters)[threadIdx.x];
    __shared__ ViewParams vp;
    __shared__ FrameParams fp;
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += FRAG_TPB) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    for (int w = threadIdx.x; w < (int)(sizeof(FrameParams) / 4); w += FRAG_TPB) reinterpret_cast<uint32_t *>(&fp)[w] = reinterpret_cast<const uint32_t *>(fpp)[w];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int row = (vp.band0 - vp.vy) + blockIdx.y * FRAG_ROWS + (threadIdx.x >> 5);
    if (row >= vp.band1 - vp.vy) return;
    const int y = vp.vy + row;
    const int bx0 = blockIdx.x * FRAG_STRETCH;
    const int nb = min(FRAG_STRETCH, vp.nbx - bx0);
    const int L = vp.n_layers;
    const uint64_t KEY_INIT = (uint64_t)MAXZ_BITS << 32;

    int32_t *headp = pl.bin_head + (size_t)row * vp.nbx + bx0 + lane;
    int32_t head = -1;
    if (lane < nb) {
        head = *headp;
        if (head >= 0) *headp = -1;
        pl.bin_used[(size_t)row * vp.nbx + bx0 + lane] = head >= 0;
    }
    unsigned mask = __ballot_sync(0xFFFFFFFFu, head >= 0);
    uint32_t *crow = color + (size_t)y * color_pitch + vp.vx + (bx0 << 5);
    float *drow = depth + (size_t)row * vp.vw + (bx0 << 5);
    const int px_left = vp.vw - (bx0 << 5);
    const float maxz = __uint_as_float(MAXZ_BITS);
    for (int b = 0; b < nb; b++) {
        const int px = (b << 5) + lane;
        if (!((mask >> b) & 1u)) {                                          // empty bin: clear values (viewport.cpp:88-113)
            if (px < px_left) { crow[px] = 0; drow[px] = maxz; }
            continue;
        }
        int32_t c = __shfl_sync(0xFFFFFFFFu, head, b);
        uint64_t best = KEY_INIT;                                           // opaque base: z bits << 32 | slot
        float best_u = 0.f;
        uint32_t best_span = 0xFFFFFFFFu;
        // kept transparent fragments, nearest first: key = z bits << 32 | ~slot (a later draw at equal depth is nearer)
        uint64_t tkey[MAX_LAYERS]; float tu[MAX_LAYERS]; uint32_t tspan[MAX_LAYERS];
        int tn = 0;
        while (c >= 0) {
            const Chunk ch = pl.chunks[c];
            c = ch.next;
            const int xs = (int)(ch.xs_xe & 0xFFu), xe = (int)(ch.xs_xe >> 8);
            if (lane < xs || lane >= xe) continue;
            const float u = pl.frag_u[ch.frag0 + (uint32_t)lane];
            const float z = fadd(ch.v0, fmul(ch.v1, u));
            if (!(z >= NEAR_Z)) continue;                                   // renderer.cpp:489
            bool opaque = ch.alpha_class == ALPHA_OPAQUE;
            if (ch.alpha_class == ALPHA_PER_FRAGMENT) {                     // the texel decides (renderer.cpp:505)
                const Prim pr = s.prims[pl.shades[ch.slot].prim];
                opaque = (shade_texture<TEX>(&pl.span_shades[ch.span], pr, s.texels, u) >> 24) == 255u;
            }
            if (opaque) {
                const uint64_t key = ((uint64_t)__float_as_uint(z) << 32) | ch.slot;
                if (key < best) { best = key; best_u = u; best_span = ch.span; }
                continue;
            }
            const uint64_t key = ((uint64_t)__float_as_uint(z) << 32) | (0xFFFFFFFFu - ch.slot);
            if (tn == L && key > tkey[L - 1]) continue;                     // farther than all L kept ones
            int i = tn < L ? tn : L - 1;                                    // insertion sort; the farthest falls off
            for (; i > 0 && tkey[i - 1] > key; i--) { tkey[i] = tkey[i - 1]; tu[i] = tu[i - 1]; tspan[i] = tspan[i - 1]; }
            tkey[i] = key; tu[i] = u; tspan[i] = ch.span;
            if (tn < L) tn++;
        }
        const bool inside = px < px_left;
        const bool hit = best_span != 0xFFFFFFFFu && inside;
        const uint32_t zb = (uint32_t)(best >> 32);
        uint32_t out = 0;                                                   // cleared screen
        if (hit) {
            const SlotShade *sh = &pl.shades[(uint32_t)best];
            out = shade<LIGHT, TEX>(&pl.span_shades[best_span], sh->flat_light, s.prims[sh->prim], s.texels, vp, fp, best_u);
        }
        if (inside) {
            int m = 0;                                                      // layers strictly in front of the base
            while (m < tn && (uint32_t)(tkey[m] >> 32) < zb) m++;
            uint32_t layer0 = 0;
            for (int i = m - 1; i >= 0; i--) {                              // far -> near: viewport.cpp:45-58
                const uint32_t slot = 0xFFFFFFFFu - (uint32_t)tkey[i];
                const SlotShade *sh = &pl.shades[slot];
                const uint32_t cl = shade<LIGHT, TEX>(&pl.span_shades[tspan[i]], sh->flat_light, s.prims[sh->prim], s.texels, vp, fp, tu[i]);
                if (i == m - 1) layer0 = cl;
                else if ((cl >> 24) != 0) layer0 = blend_px(layer0, cl);
            }
            if ((layer0 >> 24) != 0) out = blend_px(out, layer0);          // viewport.cpp:60-84
            crow[px] = out;
            drow[px] = __uint_as_float(zb);
        }
        if (count_covered) {
            const unsigned hm = __ballot_sync(0xFFFFFFFFu, hit);
            if (lane == 0 && hm) atomicAdd(&pl.counters->n_covered, (uint32_t)__popc(hm));
        }
    }
}

// ----------------------------------------------------------------------------------------
// DoF-R: the repaired post_shader_depth_box (post_shaders.hpp:63-111).
//
// out(x,y) = src(x,y)                                   if r == 0 or no tap counts
//          = (sum b / n, sum g / n, sum r / n, 255)      over taps (i,j) in [x-r,x+r) x [y-r,y+r), clipped
//                                                        to the viewport, whose blur factor is != 0
// with r = (int)blur(depth(x,y)) in 0..5.  All integer arithmetic -> pixel-identical.
//
// One CTA produces a 64x32 tile.  It stages the (64+9)x(32+9) source window in shared memory as
// packed per-pixel contributions (b | g<<21 | r<<42 in a u64, the tap count in a u32), turns
// them into a summed-area table with two short serial scans (rows, then columns), and every
// output pixel is then 4 corner lookups instead of up to 100 taps.  HBM traffic is the
// algorithmic 12 B/pixel; the halo over-fetch (80x41 staged for 64x32 outputs) is served by L2.
// The blur radius needs no per-pixel division: r(t) is a monotone step function of
// t = |focal_distance - z|, so the host finds its 5 step positions by exact bisection over the
// float bit patterns (abi.cu dof_thresholds) and the kernel just compares.
// ----------------------------------------------------------------------------------------
static constexpr int DOF_OW = 64, DOF_OH = 32, DOF_LO = 5, DOF_HI = 4;
static constexpr int DOF_SH = DOF_OH + DOF_LO + DOF_HI;      // 41 source rows
static constexpr int DOF_X0 = 8;                             // the staged window starts 8 columns left of the tile (16-byte aligned)
static constexpr int DOF_WW = DOF_OW + 16;                   // 80 staged columns = 20 x 128-bit loads per row
static constexpr int DOF_PW = DOF_WW + 1;                    // SAT row stride in entries (col 0 = zero border; odd -> no bank conflicts)
static constexpr int DOF_THREADS = 256;

struct DofClass { uint32_t radius; bool counts; };
// blur radius and "this pixel is a tap" for depth z, from the host-derived thresholds (common.cuh ViewParams)
SB_DEV DofClass dof_classify(const ViewParams &vp, float z)
{
    const float t = fabsf(fsub(vp.focal_distance, z));
    DofClass c;
    if (vp.dof_const_radius >= 0) { c.radius = (uint32_t)vp.dof_const_radius; c.counts = true; return c; }   // focal_depth == 1
    c.radius = (t >= vp.dof_t[0]) + (t >= vp.dof_t[1]) + (t >= vp.dof_t[2]) + (t >= vp.dof_t[3]) + (t >= vp.dof_t[4]);
    c.counts = t > vp.dof_on;
    return c;
}

// constant fill of a tile whose whole window is untouched background (colour 0, depth 0x7F7F7F7F)
SB_DEV void dof_fill_background(const ViewParams &vp, uint32_t *__restrict__ dst, int dst_pitch, int ox, int oy, int w, int h, int row1, int tid)
{
    const DofClass bgc = dof_classify(vp, __uint_as_float(MAXZ_BITS));
    // radius 0: copy (0); else every tap counts iff its blur != 0 -> average of zeros with alpha 255, or no taps -> copy
    const uint32_t v = (bgc.radius != 0 && bgc.counts) ? 0xFF000000u : 0u;
    const bool vst = ((dst_pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ox + DOF_OW <= w;
    if (vst) {
        for (int q = tid; q < DOF_OW * DOF_OH / 4; q += DOF_THREADS) {
            const int ty = q / (DOF_OW / 4), tx = (q % (DOF_OW / 4)) * 4;
            const int y = oy + ty;
            if (y < row1 && y < h) *reinterpret_cast<uint4 *>(dst + (size_t)y * dst_pitch + ox + tx) = make_uint4(v, v, v, v);
        }
    } else {
        for (int p = tid; p < DOF_OW * DOF_OH; p += DOF_THREADS) {
            const int ty = p / DOF_OW, tx = p % DOF_OW;
            const int x = ox + tx, y = oy + ty;
            if (y < row1 && y < h && x < w) dst[(size_t)y * dst_pitch + x] = v;
        }
    }
}

__global__ void __launch_bounds__(DOF_THREADS) k_dof(const ViewParams *__restrict__ vpp, const uint8_t *__restrict__ bin_used, int nbx,
                                                     const uint32_t *__restrict__ src, int src_pitch,
                                                     const float *__restrict__ depth, uint32_t *__restrict__ dst,
                                                     int dst_pitch, int w, int h, int row0, int row1)
{
    __shared__ unsigned long long s64[(DOF_SH + 1) * DOF_PW];
    __shared__ uint32_t s32[(DOF_SH + 1) * DOF_PW];
    __shared__ uint8_t srad[DOF_SH * DOF_WW];               // blur radius (0..5) of every staged pixel
    __shared__ uint32_t magic[128];                          // ceil(2^28 / n): exact v / n for v < 2^15, n <= 100
    __shared__ ViewParams vp;
    const int tid = threadIdx.x;
    const int ox = blockIdx.x * DOF_OW, oy = row0 + blockIdx.y * DOF_OH;
    for (int k = tid; k < (int)(sizeof(ViewParams) / 4); k += DOF_THREADS) reinterpret_cast<uint32_t *>(&vp)[k] = reinterpret_cast<const uint32_t *>(vpp)[k];
    pdl_wait();                                                             // k_fragments' colour, depth and occupancy map

    // ---- k_fragments left one byte per 32-column bin saying whether it drew anything there this frame.  If no bin
    //      under the window [ox-8, ox+72) x [oy-5, oy+36) did, the window is pure background: no loads at all. ----
    {
        const int b0 = max(0, (ox - DOF_X0) >> 5), b1 = min(nbx - 1, (ox - DOF_X0 + DOF_WW - 1) >> 5);
        const int nbw = b1 - b0 + 1;                         // <= 4
        bool used = false;
        for (int i = tid; i < nbw * DOF_SH; i += DOF_THREADS) {
            const int sy = i / nbw, gy = oy + sy - DOF_LO;
            if (gy >= 0 && gy < h) used = used || bin_used[(size_t)gy * nbx + b0 + (i - sy * nbw)];
        }
        if (!__syncthreads_or(used)) {                       // (the barrier also makes `vp` visible)
            dof_fill_background(vp, dst, dst_pitch, ox, oy, w, h, row1, tid);
            return;
        }
    }

    // ---- stage the source window [ox-8, ox+72) x [oy-5, oy+36): 20 x 41 quads of 4 pixels, 128-bit loads, all of a
    //      thread's loads issued before any is consumed.  Needs 16-byte aligned rows (w, pitches multiples of 4). ----
    constexpr int QPR = DOF_WW / 4, NQ = QPR * DOF_SH;       // 20 quads per row, 820 quads
    constexpr int QPT = (NQ + DOF_THREADS - 1) / DOF_THREADS;   // 4 per thread
    const bool vec_ok = ((w | src_pitch | dst_pitch) & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(depth)) & 15) == 0;
    float4 dz[QPT]; uint4 cc[QPT];
    uint32_t inmask = 0;                                     // 4 bits per quad: which of its pixels exist
    #pragma unroll
    for (int k = 0; k < QPT; k++) {
        const int qi = tid + k * DOF_THREADS;
        const int sy = qi / QPR, sxq = qi - sy * QPR;
        const int gy = oy + sy - DOF_LO, gx = ox - DOF_X0 + 4 * sxq;
        dz[k] = make_float4(0.f, 0.f, 0.f, 0.f); cc[k] = make_uint4(0u, 0u, 0u, 0u);
        if (qi < NQ && gy >= 0 && gy < h) {
            if (vec_ok) {
                if (gx >= 0 && gx < w) {
                    dz[k] = __ldg(reinterpret_cast<const float4 *>(depth + (size_t)gy * w + gx));
                    cc[k] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)gy * src_pitch + gx));
                    inmask |= 0xFu << (4 * k);
                }
            } else {
                float d4[4] = { 0.f, 0.f, 0.f, 0.f }; uint32_t c4[4] = { 0u, 0u, 0u, 0u };
                #pragma unroll
                for (int e = 0; e < 4; e++)
                    if (gx + e >= 0 && gx + e < w) {
                        d4[e] = __ldg(&depth[(size_t)gy * w + gx + e]); c4[e] = __ldg(&src[(size_t)gy * src_pitch + gx + e]);
                        inmask |= 1u << (4 * k + e);
                    }
                dz[k] = make_float4(d4[0], d4[1], d4[2], d4[3]); cc[k] = make_uint4(c4[0], c4[1], c4[2], c4[3]);
            }
        }
    }
    // a window that only sees untouched background (colour 0, depth 0x7F7F7F7F) blurs to a constant
    bool all_bg = true;
    #pragma unroll
    for (int k = 0; k < QPT; k++) {
        const uint32_t m = (inmask >> (4 * k)) & 0xFu;
        const bool bg = (cc[k].x | cc[k].y | cc[k].z | cc[k].w) == 0u
                     && __float_as_uint(dz[k].x) == MAXZ_BITS && __float_as_uint(dz[k].y) == MAXZ_BITS
                     && __float_as_uint(dz[k].z) == MAXZ_BITS && __float_as_uint(dz[k].w) == MAXZ_BITS;
        all_bg = all_bg && (m == 0u || (m == 0xFu && bg));
    }
    if (__syncthreads_and(all_bg)) {                         // bins were touched, but only by background-coloured pixels
        dof_fill_background(vp, dst, dst_pitch, ox, oy, w, h, row1, tid);
        return;
    }

    if (tid < 128) magic[tid] = tid ? (uint32_t)(((1u << 28) + tid - 1) / tid) : 0u;
    for (int i = tid; i < DOF_PW; i += DOF_THREADS) { s64[i] = 0; s32[i] = 0; }                 // zero row 0
    for (int i = tid; i <= DOF_SH; i += DOF_THREADS) { s64[i * DOF_PW] = 0; s32[i * DOF_PW] = 0; } // zero column 0
    #pragma unroll
    for (int k = 0; k < QPT; k++) {
        const int qi = tid + k * DOF_THREADS;
        if (qi >= NQ) continue;
        const int sy = qi / QPR, sx0 = 4 * (qi - sy * QPR);
        const float d4[4] = { dz[k].x, dz[k].y, dz[k].z, dz[k].w };
        const uint32_t c4[4] = { cc[k].x, cc[k].y, cc[k].z, cc[k].w };
        #pragma unroll
        for (int e = 0; e < 4; e++) {
            unsigned long long v = 0; uint32_t n = 0, rad = 0;
            if ((inmask >> (4 * k + e)) & 1u) {
                const DofClass dc = dof_classify(vp, d4[e]);
                rad = dc.radius;
                if (dc.counts) {
                    const uint32_t c = c4[e];
                    v = (unsigned long long)(c & 0xFF) | ((unsigned long long)((c >> 8) & 0xFF) << 21)
                      | ((unsigned long long)((c >> 16) & 0xFF) << 42);
                    n = 1;
                }
            }
            s64[(sy + 1) * DOF_PW + sx0 + e + 1] = v;
            s32[(sy + 1) * DOF_PW + sx0 + e + 1] = n;
            srad[sy * DOF_WW + sx0 + e] = (uint8_t)rad;
        }
    }
    __syncthreads();
    if (tid < DOF_SH) {                                      // inclusive scan along each row
        unsigned long long a = 0; uint32_t n = 0;
        const int base = (tid + 1) * DOF_PW;
        #pragma unroll 8
        for (int sx = 1; sx <= DOF_WW; sx++) {
            a += s64[base + sx]; n += s32[base + sx];
            s64[base + sx] = a; s32[base + sx] = n;
        }
    }
    __syncthreads();
    if (tid < DOF_WW) {                                      // then down each column
        unsigned long long a = 0; uint32_t n = 0;
        const int col = tid + 1;
        #pragma unroll 8
        for (int sy = 1; sy <= DOF_SH; sy++) {
            a += s64[sy * DOF_PW + col]; n += s32[sy * DOF_PW + col];
            s64[sy * DOF_PW + col] = a; s32[sy * DOF_PW + col] = n;
        }
    }
    __syncthreads();

    // ---- outputs: consecutive lanes take consecutive pixels (conflict-free SAT reads, coalesced stores).
    //      Staged column of output pixel tx is tx + DOF_X0; its SAT column index is that + 1. ----
    #pragma unroll 2
    for (int p = tid; p < DOF_OW * DOF_OH; p += DOF_THREADS) {
        const int ty = p / DOF_OW, tx = p % DOF_OW;
        const int x = ox + tx, y = oy + ty;
        if (y >= row1 || y >= h || x >= w) continue;
        const int radius = srad[(ty + DOF_LO) * DOF_WW + tx + DOF_X0];
        uint32_t out;
        bool have = false;
        if (radius != 0) {
            // window rows [y-r, y+r) -> SAT rows (ty+5-r, ty+5+r]; columns likewise with the +8 staging offset; the zero
            // border and the zeros stored for out-of-viewport pixels implement the max(0,..)/min(w|h,..) clipping
            const int J0 = ty + DOF_LO - radius, J1 = ty + DOF_LO + radius;
            const int I0 = tx + DOF_X0 - radius, I1 = tx + DOF_X0 + radius;
            const uint32_t count = (s32[J1 * DOF_PW + I1] + s32[J0 * DOF_PW + I0])
                                 - (s32[J0 * DOF_PW + I1] + s32[J1 * DOF_PW + I0]);
            if (count) {
                const unsigned long long sum = (s64[J1 * DOF_PW + I1] + s64[J0 * DOF_PW + I0])
                                             - (s64[J0 * DOF_PW + I1] + s64[J1 * DOF_PW + I0]);
                // exact floor(v / count) for v < 2^15, count <= 100: (v * ceil(2^28 / count)) >> 28, as one IMAD.HI
                const uint32_t m = magic[count];
                const uint32_t lo = (uint32_t)sum, hi = (uint32_t)(sum >> 32);
                const uint32_t b = __umulhi((lo & 0x1FFFFFu) << 4, m);
                const uint32_t g = __umulhi((((lo >> 21) | (hi << 11)) & 0x1FFFFFu) << 4, m);
                const uint32_t r = __umulhi(((hi >> 10) & 0x1FFFFFu) << 4, m);
                out = b | (g << 8) | (r << 16) | 0xFF000000u;
                have = true;
            }
        }
        if (!have) out = __ldg(&src[(size_t)y * src_pitch + x]);
        dst[(size_t)y * dst_pitch + x] = out;
    }
}

// ----------------------------------------------------------------------------------------
// Frame protocol of the band-sharded single frame (FrameSync, common.cuh): the assembling GPU (rank 0) clears the other
// ranks' rows of its screen itself -- local HBM writes -- and announces the frame; the other ranks then store only the
// tiles they drew into (their last kernel's ordinary stores, travelling over NVLink) and raise a flag in rank 0's
// memory; rank 0's stream ends with a kernel that waits for the flags.  No collective, no host round trip, and the
// NVLink ingress of rank 0 carries the covered pixels instead of the whole frame.
// ----------------------------------------------------------------------------------------
static constexpr long long SYNC_TIMEOUT_CYCLES = 4000000000ll;      // ~2 s (first frames include graph instantiation on the peer): a lost peer must not hang the GPU

SB_DEV uint32_t ld_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
SB_DEV unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
SB_DEV void st_sys(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// rank 0, first kernel of the frame: clear the viewport's rows outside the own band, then ready = seq (last CTA done)
__global__ void __launch_bounds__(256) k_sync_clear(const ViewParams *__restrict__ vpp, uint32_t *__restrict__ screen, int pitch,
                                                    FrameSync *own, int vx, int vy, int vw, int vh, int band0, int band1, int do_clear)
{
    pdl_trigger();
    if (blockIdx.x == 0 && threadIdx.x == 0) own->t_begin = global_ns();
    if (do_clear) {
        const int rows_above = band0 - vy, rows_out = vh - (band1 - band0);
        const bool vec = ((vx | vw | pitch) & 3) == 0 && (reinterpret_cast<uintptr_t>(screen) & 15) == 0;
        if (vec) {
            const int w4 = vw >> 2;
            const long long n = (long long)rows_out * w4;
            for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
                const int r = (int)(i / w4), c = (int)(i - (long long)r * w4);
                const int y = r < rows_above ? vy + r : band1 + (r - rows_above);
                *reinterpret_cast<uint4 *>(screen + (size_t)y * pitch + vx + 4 * c) = make_uint4(0, 0, 0, 0);
            }
        } else {
            const long long n = (long long)rows_out * vw;
            for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
                const int r = (int)(i / vw), c = (int)(i - (long long)r * vw);
                const int y = r < rows_above ? vy + r : band1 + (r - rows_above);
                screen[(size_t)y * pitch + vx + c] = 0;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&own->clear_ctas, 1u) == gridDim.x - 1) {
            own->clear_ctas = 0;
            __threadfence_system();
            st_sys(&own->ready, vpp->sync_seq);
        }
    }
}

// other ranks, in front of the kernel that stores into rank 0's screen
// (the dependent kernel is released only after the wait: its CTAs would otherwise sit on the SMs, and when the "ranks"
// are contexts of ONE device -- the in-process test -- rank 0's clear kernel could not be scheduled under them)
__global__ void k_sync_wait_ready(const ViewParams *__restrict__ vpp, const FrameSync *peer, FrameSync *own)
{
    if (threadIdx.x == 0) {
        const uint32_t seq = vpp->sync_seq;
        const long long t0 = clock64();
        while ((int32_t)(ld_sys(&peer->ready) - seq) < 0)
            if (clock64() - t0 > SYNC_TIMEOUT_CYCLES) { atomicAdd(&own->error, 1u); break; }
    }
    __syncwarp();
    pdl_trigger();
    pdl_wait();                                                             // keeps completion transitive along the chain
}

// other ranks, after their last kernel: everything that kernel stored is ordered before the flag
__global__ void k_sync_signal(const ViewParams *__restrict__ vpp, FrameSync *peer, int rank)
{
    pdl_wait();
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_sys(&peer->done[rank], vpp->sync_seq);
    }
}

// rank 0, last kernel of the frame
__global__ void k_sync_wait_done(const ViewParams *__restrict__ vpp, FrameSync *own, int world)
{
    pdl_wait();
    const int r = threadIdx.x;
    if (r == 0) own->t_own_end = global_ns();                               // rank 0's own band is finished here
    if (r >= 1 && r < world) {
        const uint32_t seq = vpp->sync_seq;
        const long long t0 = clock64();
        while ((int32_t)(ld_sys(&own->done[r]) - seq) < 0)
            if (clock64() - t0 > SYNC_TIMEOUT_CYCLES) { atomicAdd(&own->error, 1u); break; }
    }
    __threadfence_system();
}

// ----------------------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------------------
// With lazy module loading the first use of a kernel may have to synchronise with the device -- which never happens while
// another context's wait kernel is spinning for this very launch.  Load the protocol's kernels up front.
void preload_sync_kernels()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_sync_clear); cudaFuncGetAttributes(&a, k_sync_wait_ready);
    cudaFuncGetAttributes(&a, k_sync_signal); cudaFuncGetAttributes(&a, k_sync_wait_done);
}
void launch_sync_clear(const ViewParams &vp, const ViewParams *d_vp, uint32_t *screen, int pitch, FrameSync *own, bool do_clear, cudaStream_t st)
{
    const int grid = do_clear ? 148 * 4 : 1;
    launch_chain(k_sync_clear, grid, 256, st, false, d_vp, screen, pitch, own, (int)vp.vx, (int)vp.vy, (int)vp.vw, (int)vp.vh, (int)vp.band0, (int)vp.band1,
                 do_clear ? 1 : 0);
}
void launch_sync_wait_ready(const ViewParams *d_vp, const FrameSync *peer, FrameSync *own, cudaStream_t st)
{
    launch_chain(k_sync_wait_ready, 1, 32, st, true, d_vp, peer, own);
}
void launch_sync_signal(const ViewParams *d_vp, FrameSync *peer, int rank, cudaStream_t st)
{
    launch_chain(k_sync_signal, 1, 32, st, true, d_vp, peer, rank);
}
void launch_sync_wait_done(const ViewParams *d_vp, FrameSync *own, int world, cudaStream_t st)
{
    launch_chain(k_sync_wait_done, 1, 32, st, true, d_vp, own, world);
}

template <int LIGHT, int TEX>
static void launch_frag_t(const DeviceScene &s, const ViewParams &vp, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p,
                          uint32_t *color, int color_pitch, float *depth, bool count_covered, Counters *h_counters_out, bool skip_bg, cudaStream_t st)
{
    const int nty = (vp.band1 - vp.band0 + FRAG_ROWS - 1) / FRAG_ROWS, n_tiles = vp.ntx * nty;
    if (n_tiles <= 0) return;
    const FragGeom g = { vp.vx, vp.vy, vp.vw, vp.band0, vp.band1, vp.nbx, vp.ntx, n_tiles, skip_bg ? 1 : 0 };
    const unsigned grid = (unsigned)n_tiles + (unsigned)((n_tiles + FRAG_ROWS - 1) / FRAG_ROWS);
    launch_chain(k_fragments<LIGHT, TEX>, grid, FRAG_TPB, st, true, s, d_vp, d_fp, p, g, color, color_pitch, depth, count_covered ? 1 : 0, h_counters_out);
}

void launch_fragments(const DeviceScene &s, const ViewParams &vp, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p,
                      uint32_t *color, int color_pitch, float *depth, bool count_covered, Counters *h_counters_out, bool skip_bg_color,
                      cudaStream_t st)
{
#define SB_CASE(L, T) if (vp.light_mode == L && vp.tex_mode == T) { launch_frag_t<L, T>(s, vp, d_vp, d_fp, p, color, color_pitch, depth, count_covered, h_counters_out, skip_bg_color, st); return; }
    SB_CASE(0, 0) SB_CASE(0, 1) SB_CASE(0, 2)
    SB_CASE(1, 0) SB_CASE(1, 1) SB_CASE(1, 2)
    SB_CASE(2, 0) SB_CASE(2, 1) SB_CASE(2, 2)
#undef SB_CASE
}

template <int LIGHT, int TEX>
static void launch_frag_layers_t(const DeviceScene &s, const ViewParams &vp, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p,
                                 uint32_t *color, int color_pitch, float *depth, bool count_covered, Counters *h_counters_out, cudaStream_t st)
{
    dim3 grid((vp.nbx + FRAG_STRETCH - 1) / FRAG_STRETCH, (vp.band1 - vp.band0 + FRAG_ROWS - 1) / FRAG_ROWS);
    launch_chain(k_fragments_layers<LIGHT, TEX>, grid, FRAG_TPB, st, true, s, d_vp, d_fp, p, color, color_pitch, depth, count_covered ? 1 : 0, h_counters_out);
}
void launch_fragments_layers(const DeviceScene &s, const ViewParams &vp, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p,
                             uint32_t *color, int color_pitch, float *depth, bool count_covered, Counters *h_counters_out, cudaStream_t st)
{
#define SB_LCASE(L, T) if (vp.light_mode == L && vp.tex_mode == T) { launch_frag_layers_t<L, T>(s, vp, d_vp, d_fp, p, color, color_pitch, depth, count_covered, h_counters_out, st); return; }
    SB_LCASE(0, 0) SB_LCASE(0, 1) SB_LCASE(0, 2) SB_LCASE(1, 0) SB_LCASE(1, 1) SB_LCASE(1, 2) SB_LCASE(2, 0) SB_LCASE(2, 1) SB_LCASE(2, 2)
#undef SB_LCASE
}

// ----------------------------------------------------------------------------------------
// self-test of div_by() against __fdiv_rn (swegl_b200_selftest_division): every thread draws operand pairs from a
// counter-based generator -- raw bit patterns, so every exponent, sign, zero, denormal, infinity and NaN occurs, and a
// share with the exponents pulled into the range shading uses and with extreme significands -- and counts quotients
// whose bits differ.
// ----------------------------------------------------------------------------------------
SB_DEV uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__global__ void k_selftest_division(uint64_t n_pairs, uint32_t seed, unsigned long long *mismatches, unsigned long long *fast_path)
{
    unsigned long long bad = 0, fast = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t h0 = mix32((uint32_t)i ^ seed), h1 = mix32(h0 + (uint32_t)(i >> 32) + 0x9e3779b9u), h2 = mix32(h1 ^ 0x85ebca6bu);
        uint32_t ab = h0, bb = h1;
        const uint32_t mode = h2 & 7u;
        if (mode >= 2u) {                                   // exponents 127 +- 40: the fast path's territory
            ab = (ab & 0x807FFFFFu) | ((87u + (h2 >> 8) % 81u) << 23);
            bb = (bb & 0x807FFFFFu) | ((87u + (h2 >> 16) % 81u) << 23);
        }
        if (mode == 3u) bb |= 0x007FFFFFu;                  // significand of all ones (the hard case of reciprocal-based division)
        if (mode == 4u) bb &= 0xFF800000u;                  // power of two
        if (mode == 5u) ab = (ab & 0xFFFFFF00u) | (h2 >> 24 & 1u);   // short significands
        const float a = __uint_as_float(ab), b = __uint_as_float(bb);
        const SharedDivisor d = shared_divisor(b);
        const float q = div_by(a, d), want = __fdiv_rn(a, b);
        const bool same = __float_as_uint(q) == __float_as_uint(want) || (q != q && want != want);
        bad += same ? 0u : 1u;
        const float aa = fabsf(a);
        fast += (d.ok && aa >= 0x1p-70f && aa <= 0x1p70f) ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) { bad += __shfl_down_sync(0xFFFFFFFFu, bad, o); fast += __shfl_down_sync(0xFFFFFFFFu, fast, o); }
    if ((threadIdx.x & 31) == 0) { if (bad) atomicAdd(mismatches, bad); atomicAdd(fast_path, fast); }
}
void launch_selftest_division(uint64_t n_pairs, uint32_t seed, unsigned long long *d_out2, cudaStream_t st)
{
    k_selftest_division<<<148 * 8, 256, 0, st>>>(n_pairs, seed, d_out2, d_out2 + 1);
}

void launch_dof(const ViewParams *d_vp, const uint8_t *bin_used, int nbx, const uint32_t *src, int src_pitch, const float *depth,
                uint32_t *dst, int dst_pitch, int w, int h, int row0, int row1, cudaStream_t st)
{
    dim3 grid((w + DOF_OW - 1) / DOF_OW, (row1 - row0 + DOF_OH - 1) / DOF_OH);
    if (grid.x && grid.y)
        launch_chain(k_dof, grid, DOF_THREADS, st, true, d_vp, bin_used, nbx, src, src_pitch, depth, dst, dst_pitch, w, h, row0, row1);
}

} // namespace sb
