// fragment.cu — the per-pixel half of fill_half_triangle (renderer.cpp:486-499) and the pixel
// shaders (swegl/render/pixel_shaders.hpp, src/render/pixel_shaders.cpp), plus the DoF-R post pass.
//
// k_fragments: one warp owns one 32-pixel, 1-row bin of the viewport; lane = pixel.  The warp walks the
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

// (unsigned char)round(x), colors.cpp:19-25 (round half away from zero), for the common case 0 <= x < 2^23
// as trunc + exact fraction test; anything else takes roundf
SB_DEV uint32_t round_to_byte(float a)
{
    if (!(a >= 0.0f && a < 8388608.0f)) return (uint32_t)f2i(roundf(a)) & 0xFFu;
    const int i = __float2int_rz(a);
    const float fr = fsub(a, (float)i);                                     // exact
    return (uint32_t)(i + (fr >= 0.5f ? 1 : 0)) & 0xFFu;
}
// a mod n as the reference computes texel rows/columns, with the negative (UB in the reference) case wrapped
SB_DEV int wrap_index(int a, int n, int mask)
{
    if (mask >= 0) return a & mask;                                         // power-of-two size
    int r = a % n; if (r < 0) r += n;
    return r;
}

template <int TEX>
SB_DEV uint32_t shade_texture(const SpanShade *ss, const Prim &pr, const uint32_t *texels, float u)
{
    if (TEX == SWEGL_B200_TEX_PLAIN) return pr.color;                       // pixel_shaders.hpp:28
    // t = t_left + t_dir * progress   (pixel_shaders.cpp:277, 352)
    float tx = fadd(ss->t_left[0], fmul(ss->t_dir[0], u)), ty = fadd(ss->t_left[1], fmul(ss->t_dir[1], u));
    const uint32_t *bm = texels + pr.tex_off;
    if (TEX == SWEGL_B200_TEX_NEAREST) {
        // pixel_shader_texture::shade, pixel_shaders.cpp:275-281 (unsigned modulo)
        unsigned tw = (unsigned)pr.tw, th = (unsigned)pr.th;
        unsigned uu = pr.tw_mask >= 0 ? ((unsigned)f2i(tx) & (unsigned)pr.tw_mask) : (unsigned)f2i(tx) % tw;
        unsigned vv = pr.th_mask >= 0 ? ((unsigned)f2i(ty) & (unsigned)pr.th_mask) : (unsigned)f2i(ty) % th;
        return __ldg(&bm[vv * tw + uu]);
    }
    // pixel_shader_texture_bilinear::shade, pixel_shaders.cpp:348-384: t.x picks the ROW, t.y the COLUMN
    float v = tx, uq = ty;
    float u1 = fsub(uq, 0.5f), u2 = fadd(uq, 0.5f), v1 = fsub(v, 0.5f), v2 = fadd(v, 0.5f);
    uq = floorf(u2); v = floorf(v2);
    int tw = pr.tw, th = pr.th;
    int v1m = wrap_index(f2i(v1) + th, th, pr.th_mask);                     // ((int)v1 + theight) % theight; UB guard
    int v2m = v1m + 1; if (v2m == th) v2m = 0;
    v1m *= tw; v2m *= tw;
    int u1m = wrap_index(f2i(u1) + tw, tw, pr.tw_mask);
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
        out |= round_to_byte(acc) << sft;                                   // (unsigned char)round(), colors.cpp:19-25
    }
    return out;
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
    uint32_t c = shade_texture<TEX>(ss, pr, texels, u);
    if (LIGHT == SWEGL_B200_LIGHT_NONE) return c;
    // pixel_shader_light_and_texture::shade, pixel_shaders.hpp:159-178
    int li = shade_light<LIGHT>(ss, flat_light, vp, fp, u);
    float light = fmul(__int2float_rn(li), 1.0f / 65536.0f);               // (float)(li / 65536.0)
    uint32_t b = c & 0xFF, g = (c >> 8) & 0xFF, r = (c >> 16) & 0xFF;
    if (light < 1.0f) {
        // |light| <= 32768 and c <= 255, so the product is always inside int range: plain truncation
        b = (uint32_t)__float2int_rz(fmul((float)b, light)) & 0xFF;
        g = (uint32_t)__float2int_rz(fmul((float)g, light)) & 0xFF;
        r = (uint32_t)__float2int_rz(fmul((float)r, light)) & 0xFF;
    } else {
        light = __fsqrt_rn(__fsqrt_rn(light));
        b = (255u - ((uint32_t)__float2int_rz(fdiv((float)(255 - (int)b), light)) & 0xFF)) & 0xFF;   // light >= 1
        g = (255u - ((uint32_t)__float2int_rz(fdiv((float)(255 - (int)g), light)) & 0xFF)) & 0xFF;
        r = (255u - ((uint32_t)__float2int_rz(fdiv((float)(255 - (int)r), light)) & 0xFF)) & 0xFF;
    }
    return (c & 0xFF000000u) | (r << 16) | (g << 8) | b;
}

static constexpr int FRAG_TPB = 256;
static constexpr int FRAG_ROWS = FRAG_TPB / 32;     // one warp per row of the CTA's 8-row group
static constexpr int FRAG_STRETCH = 8;              // bins (of 32 pixels) one warp owns along its row

// One warp owns a stretch of FRAG_STRETCH bins (256 pixels) of one scanline:
//   1. one load fetches the bin heads (and resets them for the next frame),
//   2. empty bins are cleared in bulk with 128-bit stores (colour 0, depth 0x7F7F7F7F: viewport.cpp:88-113),
//   3. each non-empty bin is resolved with lane = pixel as described at the top of this file.
template <int LIGHT, int TEX>
__global__ void __launch_bounds__(FRAG_TPB) k_fragments(DeviceScene s, const ViewParams *__restrict__ vpp,
                                                        const FrameParams *__restrict__ fpp, Pools pl,
                                                        uint32_t *__restrict__ color, int color_pitch,
                                                        float *__restrict__ depth, int count_covered,
                                                        Counters *__restrict__ h_counters_out)
{
    // k_setup / k_spans are done: publish their counters (pool demand, overflow flags) to the pinned slot the host
    // polls, instead of a D2H copy node at the end of the graph.  (n_covered is only final after this kernel; the
    // synchronous stats path copies the counters itself.)
    if (h_counters_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < sizeof(Counters) / 4)
        reinterpret_cast<uint32_t *>(h_counters_out)[threadIdx.x] = reinterpret_cast<const uint32_t *>(pl.counters)[threadIdx.x];
    __shared__ ViewParams vp;                   // per-frame constants staged once per CTA
    __shared__ FrameParams fp;
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += FRAG_TPB) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    for (int w = threadIdx.x; w < (int)(sizeof(FrameParams) / 4); w += FRAG_TPB) reinterpret_cast<uint32_t *>(&fp)[w] = reinterpret_cast<const uint32_t *>(fpp)[w];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int row = (vp.band0 - vp.vy) + blockIdx.y * FRAG_ROWS + (threadIdx.x >> 5);   // viewport-relative
    if (row >= vp.band1 - vp.vy) return;
    const int y = vp.vy + row;
    const int bx0 = blockIdx.x * FRAG_STRETCH;
    const int nb = min(FRAG_STRETCH, vp.nbx - bx0);
    const Span *spans = pl.spans;
    const uint64_t KEY_INIT = (uint64_t)MAXZ_BITS << 32;

    int32_t *headp = pl.bin_head + (size_t)row * vp.nbx + bx0 + lane;
    int32_t head = -1;
    if (lane < nb) { head = *headp; if (head >= 0) *headp = -1; }
    unsigned mask = __ballot_sync(0xFFFFFFFFu, head >= 0);

    uint32_t *crow = color + (size_t)y * color_pitch + vp.vx + (bx0 << 5);
    float *drow = depth + (size_t)row * vp.vw + (bx0 << 5);
    const int px_left = vp.vw - (bx0 << 5);                                 // pixels from the stretch start to the row end
    const bool aligned = ((reinterpret_cast<uintptr_t>(crow) | reinterpret_cast<uintptr_t>(drow)) & 15) == 0;
    // ---- empty bins ----
    if (mask != 0xFFFFFFFFu) {
        const float maxz = __uint_as_float(MAXZ_BITS);
        for (int i = lane; i < nb * 8; i += 32) {
            const int b = i >> 3, px = (b << 5) + ((i & 7) << 2);
            if ((mask >> b) & 1u) continue;
            if (aligned && px + 4 <= px_left) {
                *reinterpret_cast<uint4 *>(crow + px) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<float4 *>(drow + px) = make_float4(maxz, maxz, maxz, maxz);
            } else {
                for (int k = 0; k < 4; k++)
                    if (px + k < px_left) { crow[px + k] = 0; drow[px + k] = maxz; }
            }
        }
    }
    // ---- non-empty bins ----
    uint32_t covered = 0;
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        int32_t c = __shfl_sync(0xFFFFFFFFu, head, b);
        const int binx0 = vp.vx + ((bx0 + b) << 5);
        const int x = binx0 + lane;

        uint64_t best = KEY_INIT;
        float best_u = 0.f;
        uint32_t best_span = 0;
        while (c >= 0) {
            const Chunk ch = pl.chunks[c];
            const Span sp = spans[ch.span];
            const int x1 = (int)(sp.x1x2 & 0xFFFFu), x2 = (int)(sp.x1x2 >> 16);
            if (x >= x1 && x < x2) {
                const float2 tb = pl.frag_tb[sp.frag_base + (uint32_t)(x - x1)];   // qpixel state, replayed by k_spans
                const float u = fdiv(tb.x, tb.y);                               // progress(), interpolator.hpp:98
                const float z = fadd(sp.v0, fmul(sp.v1, u));                    // value(0), renderer.cpp:488
                if (z >= NEAR_Z) {                                              // renderer.cpp:489-492
                    uint64_t key = ((uint64_t)__float_as_uint(z) << 32) | (sp.slot_flags >> 2);
                    if (key < best) { best = key; best_u = u; best_span = ch.span; }
                }
            }
            c = ch.next;
        }

        const bool inside = x < vp.vx + vp.vw;
        const bool hit = best != KEY_INIT;
        uint32_t out = 0;                                                   // background, viewport.cpp:95-103
        if (hit) {
            const uint32_t slot = spans[best_span].slot_flags >> 2;
            const SlotShade *sh = &pl.shades[slot];
            const Prim pr = s.prims[sh->prim];
            out = shade<LIGHT, TEX>(&pl.span_shades[best_span], sh->flat_light, pr, s.texels, vp, fp, best_u);
        }
        if (inside) {
            crow[(b << 5) + lane] = out;
            drow[(b << 5) + lane] = __uint_as_float((uint32_t)(best >> 32));
        }
        if (count_covered) covered += __popc(__ballot_sync(0xFFFFFFFFu, hit && inside));
    }
    if (count_covered && lane == 0 && covered) atomicAdd(&pl.counters->n_covered, covered);
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
// algorithmic 12 B/pixel; the 46 % halo over-fetch is served by L2.
// ----------------------------------------------------------------------------------------
static constexpr int DOF_OW = 64, DOF_OH = 32, DOF_LO = 5, DOF_HI = 4;
static constexpr int DOF_SW = DOF_OW + DOF_LO + DOF_HI;      // 73 source columns
static constexpr int DOF_SH = DOF_OH + DOF_LO + DOF_HI;      // 41 source rows
static constexpr int DOF_PW = 75;                            // SAT row stride in entries (col 0 = zero border; odd -> no bank conflicts)
static constexpr int DOF_THREADS = 256;

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

__global__ void __launch_bounds__(DOF_THREADS) k_dof(const ViewParams *__restrict__ vpp, const uint32_t *__restrict__ src, int src_pitch,
                                                     const float *__restrict__ depth, uint32_t *__restrict__ dst,
                                                     int dst_pitch, int w, int h, int row0, int row1)
{
    const float focal_distance = vpp->focal_distance, focal_depth = vpp->focal_depth;
    __shared__ unsigned long long s64[(DOF_SH + 1) * DOF_PW];
    __shared__ uint32_t s32[(DOF_SH + 1) * DOF_PW];
    __shared__ uint8_t srad[DOF_SH * DOF_SW];               // blur radius (0..5) of every staged pixel
    __shared__ uint32_t magic[128];                          // ceil(2^28 / n): exact v / n for v < 2^15, n <= 100
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ox = blockIdx.x * DOF_OW, oy = row0 + blockIdx.y * DOF_OH;

    // ---- stage the source window: warp `warp` loads rows warp, warp+8, ...; lanes sweep the 73 columns.
    //      All global loads of a thread are issued before any is consumed (memory-level parallelism).
    constexpr int ROWS_PER_WARP = (DOF_SH + 7) / 8;         // 6
    constexpr int COLS_PER_LANE = (DOF_SW + 31) / 32;       // 3
    float dz[ROWS_PER_WARP][COLS_PER_LANE]; uint32_t cc[ROWS_PER_WARP][COLS_PER_LANE];
    const float maxz = __uint_as_float(MAXZ_BITS);
    #pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int sy = warp + rr * 8, gy = oy + sy - DOF_LO;
        #pragma unroll
        for (int cq = 0; cq < COLS_PER_LANE; cq++) {
            const int sx = lane + cq * 32, gx = ox + sx - DOF_LO;
            const bool in = sy < DOF_SH && sx < DOF_SW && gx >= 0 && gx < w && gy >= 0 && gy < h;
            dz[rr][cq] = in ? __ldg(&depth[(size_t)gy * w + gx]) : __int_as_float(0x7FC00000);   // NaN = not a pixel
            cc[rr][cq] = in ? __ldg(&src[(size_t)gy * src_pitch + gx]) : 0u;
        }
    }
    // a window that only sees untouched background (colour 0, depth 0x7F7F7F7F) blurs to a constant
    bool all_bg = true;
    #pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++)
        #pragma unroll
        for (int cq = 0; cq < COLS_PER_LANE; cq++) {
            const bool pixel = dz[rr][cq] == dz[rr][cq];
            all_bg = all_bg && (!pixel || (cc[rr][cq] == 0u && __float_as_uint(dz[rr][cq]) == MAXZ_BITS));
        }
    if (__syncthreads_and(all_bg)) {
        const float bf = blur_factor(maxz, focal_distance, focal_depth);
        const int radius = f2i(bf);
        // r == 0: copy (0); else every tap counts iff bf != 0 -> average of zeros with alpha 255, or no taps -> copy
        const uint32_t v = (radius != 0 && bf != 0.0f) ? 0xFF000000u : 0u;
        for (int p = tid; p < DOF_OW * DOF_OH; p += DOF_THREADS) {
            const int ty = p / DOF_OW, tx = p % DOF_OW;
            const int x = ox + tx, y = oy + ty;
            if (y < row1 && y < h && x < w) dst[(size_t)y * dst_pitch + x] = v;
        }
        return;
    }

    if (tid < 128) magic[tid] = tid ? (uint32_t)(((1u << 28) + tid - 1) / tid) : 0u;
    for (int i = tid; i < DOF_PW; i += DOF_THREADS) { s64[i] = 0; s32[i] = 0; }                 // zero row 0
    for (int i = tid; i <= DOF_SH; i += DOF_THREADS) { s64[i * DOF_PW] = 0; s32[i * DOF_PW] = 0; } // zero column 0
    #pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int sy = warp + rr * 8;
        #pragma unroll
        for (int cq = 0; cq < COLS_PER_LANE; cq++) {
            const int sx = lane + cq * 32;
            if (sy >= DOF_SH || sx >= DOF_SW) continue;
            unsigned long long v = 0; uint32_t n = 0; uint32_t rad = 0;
            if (dz[rr][cq] == dz[rr][cq]) {
                const float bf = blur_factor(dz[rr][cq], focal_distance, focal_depth);
                rad = (uint32_t)f2i(bf);
                if (bf != 0.0f) {
                    const uint32_t c = cc[rr][cq];
                    v = (unsigned long long)(c & 0xFF) | ((unsigned long long)((c >> 8) & 0xFF) << 21)
                      | ((unsigned long long)((c >> 16) & 0xFF) << 42);
                    n = 1;
                }
            }
            s64[(sy + 1) * DOF_PW + sx + 1] = v;
            s32[(sy + 1) * DOF_PW + sx + 1] = n;
            srad[sy * DOF_SW + sx] = (uint8_t)rad;
        }
    }
    __syncthreads();
    if (tid < DOF_SH) {                                      // inclusive scan along each row
        unsigned long long a = 0; uint32_t n = 0;
        const int base = (tid + 1) * DOF_PW;
        #pragma unroll 8
        for (int sx = 1; sx <= DOF_SW; sx++) {
            a += s64[base + sx]; n += s32[base + sx];
            s64[base + sx] = a; s32[base + sx] = n;
        }
    }
    __syncthreads();
    if (tid < DOF_SW) {                                      // then down each column
        unsigned long long a = 0; uint32_t n = 0;
        const int col = tid + 1;
        #pragma unroll 8
        for (int sy = 1; sy <= DOF_SH; sy++) {
            a += s64[sy * DOF_PW + col]; n += s32[sy * DOF_PW + col];
            s64[sy * DOF_PW + col] = a; s32[sy * DOF_PW + col] = n;
        }
    }
    __syncthreads();

    // ---- outputs: consecutive lanes take consecutive pixels (conflict-free SAT reads, coalesced stores)
    #pragma unroll 2
    for (int p = tid; p < DOF_OW * DOF_OH; p += DOF_THREADS) {
        const int ty = p / DOF_OW, tx = p % DOF_OW;
        const int x = ox + tx, y = oy + ty;
        if (y >= row1 || y >= h || x >= w) continue;
        const int radius = srad[(ty + DOF_LO) * DOF_SW + tx + DOF_LO];
        uint32_t out;
        bool have = false;
        if (radius != 0) {
            // SAT index of source pixel (gx, gy) is (gx - ox + 6, gy - oy + 6); the zero border and the
            // zeros stored for out-of-viewport pixels implement the max(0,..)/min(w|h,..) clipping
            const int J0 = ty + DOF_LO - radius, J1 = ty + DOF_LO + radius;
            const int I0 = tx + DOF_LO - radius, I1 = tx + DOF_LO + radius;
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
// launchers
// ----------------------------------------------------------------------------------------
template <int LIGHT, int TEX>
static void launch_frag_t(const DeviceScene &s, const ViewParams &vp, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p,
                          uint32_t *color, int color_pitch, float *depth, bool count_covered, Counters *h_counters_out, cudaStream_t st)
{
    dim3 grid((vp.nbx + FRAG_STRETCH - 1) / FRAG_STRETCH, (vp.band1 - vp.band0 + FRAG_ROWS - 1) / FRAG_ROWS);
    if (!grid.x || !grid.y) return;
    k_fragments<LIGHT, TEX><<<grid, FRAG_TPB, 0, st>>>(s, d_vp, d_fp, p, color, color_pitch, depth, count_covered ? 1 : 0, h_counters_out);
}

void launch_fragments(const DeviceScene &s, const ViewParams &vp, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p,
                      uint32_t *color, int color_pitch, float *depth, bool count_covered, Counters *h_counters_out, cudaStream_t st)
{
#define SB_CASE(L, T) if (vp.light_mode == L && vp.tex_mode == T) { launch_frag_t<L, T>(s, vp, d_vp, d_fp, p, color, color_pitch, depth, count_covered, h_counters_out, st); return; }
    SB_CASE(0, 0) SB_CASE(0, 1) SB_CASE(0, 2)
    SB_CASE(1, 0) SB_CASE(1, 1) SB_CASE(1, 2)
    SB_CASE(2, 0) SB_CASE(2, 1) SB_CASE(2, 2)
#undef SB_CASE
}

void launch_dof(const ViewParams *d_vp, const uint32_t *src, int src_pitch, const float *depth, uint32_t *dst, int dst_pitch,
                int w, int h, int row0, int row1, cudaStream_t st)
{
    dim3 grid((w + DOF_OW - 1) / DOF_OW, (row1 - row0 + DOF_OH - 1) / DOF_OH);
    if (grid.x && grid.y)
        k_dof<<<grid, DOF_THREADS, 0, st>>>(d_vp, src, src_pitch, depth, dst, dst_pitch, w, h, row0, row1);
}

} // namespace sb
