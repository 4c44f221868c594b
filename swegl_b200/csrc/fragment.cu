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
#include <cstddef>
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

// pixel_shader_light_and_texture::shade, pixel_shaders.hpp:159-178: the texture colour scaled by the light
SB_DEV uint32_t combine_light(uint32_t c, int li)
{
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

template <int LIGHT, int TEX>
SB_DEV uint32_t shade(const SpanShade *ss, float flat_light, const Prim &pr, const uint32_t *texels, const ViewParams &vp,
                      const FrameParams &fp, float u)
{
    const TexFetch tf = tex_fetch<TEX>(ss, pr, texels, u);                  // texel loads issued ...
    if (LIGHT == SWEGL_B200_LIGHT_NONE) return tex_filter<TEX>(tf);
    int li = shade_light<LIGHT>(ss, flat_light, vp, fp, u);                 // ... in flight under the lighting arithmetic ...
    return combine_light(tex_filter<TEX>(tf), li);                          // ... consumed here
}


// ----------------------------------------------------------------------------------------
// Phong lighting inside the colour tolerance (+-1 LSB per 8-bit channel, alpha exact; coverage and depth stay
// bit-exact -- they never pass through here).  The exact path above spends ~650 of its ~1000 instructions per pixel on
// IEEE-correct divisions, square roots and an fp64 pow in the reference's unfused evaluation order.  This version uses
// FMA, rsqrt.approx / rcp.approx and five fp32 squarings; its relative error on the light sum is ~1e-6, which moves a
// truncated colour channel by at most one step -- EXCEPT next to the shader's own discontinuities, where a tiny
// difference flips a branch with a visible jump (pixel_shaders.cpp:173-203, pixel_shaders.hpp:159-178):
//   diffuse < 0.05 (the light is dropped),  alignment < 0 (dropped, its specular term with it),  degenerate vectors,
//   light == 1 (dark branch truncates c*light -> c-1, bright branch gives c+1),  light < 0 or 65536*light beyond int.
// A pixel within a guard band of any of these reports `false` and is shaded by the exact path instead.  The texture
// filter is always the exact one (texel selection is discontinuous, and an exact `c` keeps the bound at 1 step).
// ----------------------------------------------------------------------------------------
SB_DEV float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SB_DEV float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

SB_DEV bool phong_light_fast(const SpanShade *ss, const ViewParams &vp, const FrameParams &fp, float u, int &li)
{
    const float4 *sq = reinterpret_cast<const float4 *>(ss);
    const float4 q0 = sq[0];          // v.xyz, vdir.x
    const float4 q1 = sq[1];        // vdir.yz, n.xy
    const float4 q2 = sq[2];        // n.z, ndir.xyz
    const float cx = __fmaf_rn(q0.w, u, q0.x), cy = __fmaf_rn(q1.x, u, q0.y), cz = __fmaf_rn(q1.y, u, q0.z);
    float nx = __fmaf_rn(q2.y, u, q1.z), ny = __fmaf_rn(q2.z, u, q1.w), nz = __fmaf_rn(q2.w, u, q2.x);
    const float nn = __fmaf_rn(nx, nx, __fmaf_rn(ny, ny, nz * nz));
    const float ni = rsqrt_fast(nn);
    nx *= ni; ny *= ni; nz *= ni;
    float ex = vp.cam[0] - cx, ey = vp.cam[1] - cy, ez = vp.cam[2] - cz;
    const float ee = __fmaf_rn(ex, ex, __fmaf_rn(ey, ey, ez * ez));
    const float ei = rsqrt_fast(ee);
    ex *= ei; ey *= ei; ez *= ei;
    bool ok = nn > 1e-30f && ee > 1e-30f;                                   // (NaN compares false)
    float sun = -__fmaf_rn(nx, fp.sun[0], __fmaf_rn(ny, fp.sun[1], nz * fp.sun[2]));
    sun = sun < 0.0f ? 0.0f : sun * fp.sun_intensity;
    float dyn = 0.0f;
    for (uint32_t i = 0; i < fp.n_lights; i++) {
        const float4 L = __ldg(&fp.lights[i]);
        float lx = cx - L.x, ly = cy - L.y, lz = cz - L.z;
        const float d2 = __fmaf_rn(lx, lx, __fmaf_rn(ly, ly, lz * lz));
        const float inv = rcp_fast(d2);
        float diffuse = L.w * inv;
        ok = ok && d2 > 1e-30f && fabsf(diffuse - 0.05f) > 2e-6f;
        if (!(diffuse >= 0.05f)) continue;
        const float il = rsqrt_fast(d2);
        lx *= il; ly *= il; lz *= il;
        const float al = -__fmaf_rn(nx, lx, __fmaf_rn(ny, ly, nz * lz));
        ok = ok && fabsf(al) > 2e-5f;
        if (!(al > 0.0f)) continue;
        diffuse *= al;
        const float a2 = al + al;
        const float rx = __fmaf_rn(nx, a2, lx), ry = __fmaf_rn(ny, a2, ly), rz = __fmaf_rn(nz, a2, lz);
        float sp = __fmaf_rn(rx, ex, __fmaf_rn(ry, ey, rz * ez));
        if (sp > 0.0f) {
            sp *= sp; sp *= sp; sp *= sp; sp *= sp; sp *= sp;               // ^32
            dyn += __fmaf_rn(sp * 16.0f, inv, diffuse);
        } else {
            dyn += diffuse;
        }
    }
    const float total = fp.ambient + sun + dyn;
    ok = ok && total >= 0.0f && total < 16384.0f && fabsf(total - 1.0f) > 2e-4f;
    li = __float2int_rz(65536.0f * total);
    return ok;
}

// ----------------------------------------------------------------------------------------
// The bilinear filter of the fast kernels: the SAME arithmetic as tex_fetch / tex_filter above (bit-identical texels, weights,
// products, sums and rounding), minus the per-operation range checks.  The caller guarantees |t| < 2^21 for both texture
// coordinates (a pixel outside takes the exact shader), and then
//   * (int)(t - 0.5) is a plain truncation (no cvttss2si overflow case), floor(t + 0.5) is RD(x + 1.5 * 2^23) - 1.5 * 2^23;
//   * the four weights are >= 0 (rounding is monotone: floor(RN(t + 0.5)) >= RN(t - 0.5)) and sum to 1 +- 2^-2, so every
//     channel sum lies in [0, 2^10): (unsigned char)round(acc) = trunc(RZ(acc + 0.5)) -- RZ, not RN, so that
//     0.5 - 2^-25 + 0.5 does not round up to 1 -- read off the low mantissa bits of RZ(. + 2^23);
//   * below |t| < 2^12 the weights sum to 1 +- 2^-11, so four texels of alpha 255 give 255 +- 0.13 -> 255 without computing it.
// ----------------------------------------------------------------------------------------
template <int K> SB_DEV float byte_to_float_r(uint32_t word, uint32_t k4b)
{
    return __fsub_rn(__uint_as_float(__byte_perm(word, k4b, 0x7440u | K)), MAGIC23);
}
SB_DEV uint32_t round_bits_small(float acc)     // 0x4B0000bb with bb = (unsigned char)round(acc), for 0 <= acc < 2^22
{
    return __float_as_uint(__fadd_rz(__fadd_rz(acc, 0.5f), MAGIC23));
}
SB_DEV TexFetch tex_fetch_bilinear_fast(float tx, float ty, const Prim &pr, const uint32_t *texels)
{
    TexFetch f;
    const uint32_t *bm = texels + pr.tex_off;
    const float v1 = fsub(tx, 0.5f), u1 = fsub(ty, 0.5f);
    const int tw = pr.tw, th = pr.th;
    int v1m = wrap_index(__float2int_rz(v1) + th, th, pr.th_mask);
    int v2m = v1m + 1; if (v2m == th) v2m = 0;
    v1m *= tw; v2m *= tw;
    const int u1m = wrap_index(__float2int_rz(u1) + tw, tw, pr.tw_mask);
    int u2m = u1m + 1; if (u2m == tw) u2m = 0;
    f.p00 = __ldg(&bm[v1m + u1m]); f.p10 = __ldg(&bm[v2m + u1m]);
    f.p01 = __ldg(&bm[v1m + u2m]); f.p11 = __ldg(&bm[v2m + u2m]);
    f.fu = ty; f.fv = tx;
    return f;
}
// k4b = 0x4B000000 handed in as a RUN-TIME value (FragGeom::k4b, a kernel parameter): as a literal, ptxas makes it PRMT's
// immediate and spends an instruction per conversion on materialising the selector in a register instead
SB_DEV uint32_t tex_filter_bilinear_fast(const TexFetch &f, const uint32_t k4b)
{
    const float u1 = fsub(f.fu, 0.5f), u2 = fadd(f.fu, 0.5f), v1 = fsub(f.fv, 0.5f), v2 = fadd(f.fv, 0.5f);
    const float uq = __fsub_rn(__fadd_rd(u2, 12582912.0f), 12582912.0f), v = __fsub_rn(__fadd_rd(v2, 12582912.0f), 12582912.0f);
    const uint32_t p00 = f.p00, p10 = f.p10, p01 = f.p01, p11 = f.p11;
    const float w00 = fmul(fsub(uq, u1), fsub(v, v1)), w10 = fmul(fsub(uq, u1), fsub(v2, v));
    const float w01 = fmul(fsub(u2, uq), fsub(v, v1)), w11 = fmul(fsub(u2, uq), fsub(v2, v));
    #define SB_FAST_CHANNEL(C) round_bits_small(fadd(fadd(fadd(fmul(byte_to_float_r<C>(p00, k4b), w00), fmul(byte_to_float_r<C>(p10, k4b), w10)), \
                                                          fmul(byte_to_float_r<C>(p01, k4b), w01)), fmul(byte_to_float_r<C>(p11, k4b), w11)))
    const uint32_t b = SB_FAST_CHANNEL(0), g = SB_FAST_CHANNEL(1), r = SB_FAST_CHANNEL(2);
    uint32_t a = 0xFFu;
    if ((p00 & p10 & p01 & p11) < 0xFF000000u || !(fabsf(f.fu) < 4096.0f && fabsf(f.fv) < 4096.0f)) a = SB_FAST_CHANNEL(3);
    #undef SB_FAST_CHANNEL
    // low bytes of b, g, r, a -> one word
    return __byte_perm(__byte_perm(b, g, 0x0040u), __byte_perm(r, a, 0x0040u), 0x5410u);
}

// one winner -> its colour
template <int LIGHT, int TEX, int FAST>
SB_DEV uint32_t shade_winner(const Pools &pl, const uint32_t *texels, const ViewParams &vp, const FrameParams &fp, uint32_t span, uint32_t slot, float u, uint32_t k4b);

// the exact shader, out of line: the fast kernels call it for the few pixels next to a discontinuity
template <int LIGHT, int TEX>
static __device__ __noinline__ uint32_t shade_exact_call(const SpanShade *ss, float flat_light, uint4 bind, const uint32_t *texels,
                                                         const ViewParams *vp, const FrameParams *fp, float u)
{
    Prim pr;
    pr.color = bind.x; pr.tex_off = bind.y; pr.tw = (int32_t)bind.z; pr.th = (int32_t)bind.w;
    pr.tw_mask = (pr.tw & (pr.tw - 1)) == 0 ? pr.tw - 1 : -1;
    pr.th_mask = (pr.th & (pr.th - 1)) == 0 ? pr.th - 1 : -1;
    return shade<LIGHT, TEX>(ss, flat_light, pr, texels, *vp, *fp, u);
}

template <int LIGHT, int TEX, int FAST>
SB_DEV uint32_t shade_winner(const Pools &pl, const uint32_t *texels, const ViewParams &vp, const FrameParams &fp, uint32_t span, uint32_t slot, float u, uint32_t k4b)
{
    const SpanShade *ss = &pl.span_shades[span];
    const SlotShade *sh = &pl.shades[slot];
    const uint4 bind = *reinterpret_cast<const uint4 *>(&sh->color);       // colour, tex_off, tw, th
    Prim pr;
    pr.color = bind.x; pr.tex_off = bind.y; pr.tw = (int32_t)bind.z; pr.th = (int32_t)bind.w;
    pr.tw_mask = (pr.tw & (pr.tw - 1)) == 0 ? pr.tw - 1 : -1;
    pr.th_mask = (pr.th & (pr.th - 1)) == 0 ? pr.th - 1 : -1;
#ifdef FRAG_PROBE_NOSHADE
    return 0xFF00FF00u ^ bind.x ^ __float_as_uint(u) ^ __float_as_uint(ss->v[0]);      // timing probe only: what the kernel costs without the shader
#else
    if (FAST && TEX == SWEGL_B200_TEX_BILINEAR) {
        // t = t_left + t_dir * progress (pixel_shaders.cpp:352); a coordinate beyond 2^21 (or NaN) takes the exact shader
        const float tx = fadd(ss->t_left[0], fmul(ss->t_dir[0], u)), ty = fadd(ss->t_left[1], fmul(ss->t_dir[1], u));
        int li;
        if (fabsf(tx) < 2097152.0f && fabsf(ty) < 2097152.0f) {
            const TexFetch tf = tex_fetch_bilinear_fast(tx, ty, pr, texels);   // texel loads in flight under the lighting
            if (phong_light_fast(ss, vp, fp, u, li)) return combine_light(tex_filter_bilinear_fast(tf, k4b), li);
        }
        return shade_exact_call<LIGHT, TEX>(ss, 0.0f, bind, texels, &vp, &fp, u);
    }
    if (FAST) {
        const TexFetch tf = tex_fetch<TEX>(ss, pr, texels, u);             // texel loads in flight under the lighting
        int li;
        if (phong_light_fast(ss, vp, fp, u, li)) return combine_light(tex_filter<TEX>(tf), li);
        return shade_exact_call<LIGHT, TEX>(ss, 0.0f, bind, texels, &vp, &fp, u);
    }
    return shade<LIGHT, TEX>(ss, sh->flat_light, pr, texels, vp, fp, u);
#endif
}

// chunks of a bin whose records are fetched up front, all bins of the stretch at once; longer lists continue serially
static constexpr int FRAG_STAGE = 16;           // pieces of a bin staged in shared memory per round

// per-warp staging of the slot records of the current round
struct FragWarp {
    uint4 rec[FRAG_STRETCH][FRAG_STAGE * 2];    // Chunk records (2 x 16 B each) of the row's bins: the first FRAG_STAGE of every bin are staged together
};
// a load hint: bring the line into L1 / L2 without a destination register (the resolve and shading loads that follow
// find it there instead of paying a round trip to L2 / HBM each)
SB_DEV void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

// the part of ViewParams that is fixed for a captured frame graph (rectangle, band), passed by value so that the
// first loads of the kernel do not wait for the parameter block
struct FragGeom {
    int32_t vx, vy, vw, band0, band1, nbx, ntx, nty, n_tiles;
    int32_t skip_bg;                    // background colour is not stored (FrameSync: rank 0 cleared it already)
    // when DoF-R follows (k_dof), k_fragments also decides which DoF output tiles are constant, and fills those
    int32_t dof, ndx, n_dof;            // DoF tiles per row / in total (grid anchored at band0, like the fragment tiles)
    int32_t out0, out1;                 // viewport-relative rows [out0, out1) the post pass outputs
    int32_t dof_pitch;                  // pitch of the post pass's destination
    uint32_t k4b;                       // 0x4B000000 (the bits of 2^23), see tex_filter_bilinear_fast
};

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

// ---- DoF-R tile classification (one warp per DoF output tile d) ----
// The post pass reads, for output tile (i, j), the source window [64i-8, 64i+72) x [32j-5, 32j+36) (rows relative to the
// first drawn row).  If no fragment tile under that window received a chunk this frame (tile_stamp, written by k_spans),
// the whole window is background and the output is one constant: it is stored here, at full store rate, and k_dof never
// sees the tile.  Otherwise the tile goes on k_dof's work list.
struct DofClass { uint32_t radius; bool counts; };
SB_DEV DofClass dof_classify(const ViewParams &vp, float z);
SB_DEV uint32_t dof_background_value(const ViewParams &vp);

SB_DEV void dof_tile_duty(const Pools &pl, const FragGeom &g, uint32_t stamp, uint32_t fill, uint32_t *__restrict__ dof_dst, uint32_t d, int lane)
{
    const int j = (int)d / g.ndx, i = (int)d - j * g.ndx;
    const int c0 = max(0, (i * DOF_OW - 8) >> 7), c1 = min(g.ntx - 1, (i * DOF_OW + DOF_OW + 7) >> 7);       // fragment tile columns under the window
    constexpr int RPT = DOF_OH / FRAG_ROWS;                                                                  // fragment tile rows per DoF tile row
    const int r0 = max(0, j * RPT - 1), r1 = min(g.nty - 1, j * RPT + RPT);
    bool busy = false;
    {
        const int rr = r0 + (lane >> 1), cc = c0 + (lane & 1);
        if (rr <= r1 && cc <= c1) busy = pl.tile_stamp[rr * g.ntx + cc] == stamp;
    }
    static_assert(2 * (DOF_OH / FRAG_ROWS + 2) <= 32, "one lane per fragment tile under a DoF window");
    if (__any_sync(0xFFFFFFFFu, busy)) {
        if (lane == 0) pl.dof_list[atomicAdd(&pl.counters->n_dof_busy, 1u)] = d;
        return;
    }
    const int anchor = g.band0 - g.vy;
    const int x0 = i * DOF_OW, wd = min(DOF_OW, g.vw - x0);
    const int ya = max(anchor + j * DOF_OH, g.out0), yb = min(anchor + (j + 1) * DOF_OH, g.out1);
    if (yb <= ya) return;
    const bool vec = (reinterpret_cast<uintptr_t>(dof_dst) & 15) == 0 && (g.dof_pitch & 3) == 0 && wd == DOF_OW;
    if (vec) {
        // 16 lanes cover a 64-pixel row with 128-bit stores: the warp does two rows per step
        static_assert(DOF_OW / 4 == 16, "two rows of a DoF tile per warp step");
        const uint4 v4 = make_uint4(fill, fill, fill, fill);
        uint4 *q = reinterpret_cast<uint4 *>(dof_dst + (size_t)(ya + (lane >> 4)) * g.dof_pitch + x0) + (lane & 15);
        const size_t step = (size_t)(g.dof_pitch >> 1);                     // two rows, in 16-byte units
        for (int y = ya + (lane >> 4); y < yb; y += 2, q += step) *q = v4;
    } else {
        for (int k = lane; k < (yb - ya) * wd; k += 32)
            dof_dst[(size_t)(ya + k / wd) * g.dof_pitch + x0 + k % wd] = fill;
    }
}

// ----------------------------------------------------------------------------------------
// k_fragments: ONE wave of persistent CTAs; every CTA takes item after item from a queue (Counters::frag_queue) until it
// is empty.  Item p bundles up to three independent duties, so that the streaming stores of the first two drain to
// L2 / HBM underneath the arithmetic of the third:
//   * clear duty   tiles 8p .. 8p+7    : warp w looks at fragment tile 8p+w (8 rows x 128 px); unless k_spans put
//                                       something into it (tile_stamp) it gets the clear values (viewport.cpp:88-113) --
//                                       depth always, colour unless the frame protocol cleared it already, or, with
//                                       DoF-R behind it, only when a DoF window that will really be computed can see it;
//   * DoF duty     tiles 8p .. 8p+7    : warp w classifies DoF output tile 8p+w (dof_tile_duty) -- constant tiles are
//                                       stored from here, the others go on k_dof's list;
//   * busy duty    p < n_busy          : warp w takes row w of busy tile busy_list[p]: depth resolve, deferred shading
//                                       of the winners, one 128-byte colour / depth store per bin (see the file header).
// A frame has no more empty CTAs in front of the work (8 100 launches for ~1 500 busy tiles at 4K before), the expensive
// tiles are spread dynamically, and the background never waits behind them.
// ----------------------------------------------------------------------------------------
#ifndef FRAG_MINB
#define FRAG_MINB 5         // 48 registers, 40 warps per SM: measured against 6 (40 registers) and 8 (32 registers, spills in the shader)
#endif
#ifndef FRAG_COVER_MIN
#define FRAG_COVER_MIN 3    // k_fragments<.., CROWD = 1>: staged rounds of at least this many pieces are resolved per covered lane
#endif
#ifndef FRAG_TURN
#define FRAG_TURN 2         // pieces per lane and turn of the CROWD resolve (their stream loads go out together); 1 / 2 / 3 / 4 measured: BrainStem 88.5 / 81.0 / 79.9 / 81.9 us, 8K sphere 336 / 341 / 348 / 356 us
#endif
#ifndef FRAG_RESOLVE_UNROLL
#define FRAG_RESOLVE_UNROLL 4
#endif
static constexpr int RESOLVE_UNROLL = FRAG_RESOLVE_UNROLL;     // chunks of a bin whose fragment-stream loads are in flight together
#ifndef FRAG_CTAS_PER_SM
#ifndef FRAG_CTAS_PER_SM
#define FRAG_CTAS_PER_SM FRAG_MINB
#endif
#endif
SB_DEV unsigned long long global_ns();
#ifdef FRAG_PROBE_TIMELINE
// probe build only: per CTA, the %globaltimer at kernel entry, after the dependency wait, and at the end of every item
__device__ unsigned long long g_frag_timeline[2048 * 16];
#define FRAG_TL(slot, v) do { if (threadIdx.x == 0 && blockIdx.x < 2048 && (slot) < 16) g_frag_timeline[blockIdx.x * 16 + (slot)] = (v); } while (0)
#define FRAG_PH(slot) do { if (tl_first) FRAG_TL(slot, global_ns()); } while (0)
#else
#define FRAG_PH(slot) do { } while (0)
#define FRAG_TL(slot, v) do { } while (0)
#endif
template <int LIGHT, int TEX, int FAST, int CROWD>
__global__ void __launch_bounds__(FRAG_TPB, FRAG_MINB) k_fragments(DeviceScene s, const ViewParams *__restrict__ vpp,
                                                        const FrameParams *__restrict__ fpp, Pools pl, FragGeom g,
                                                        uint32_t *__restrict__ color, int color_pitch,
                                                        float *__restrict__ depth, uint32_t *__restrict__ dof_dst, int count_covered,
                                                        Counters *__restrict__ h_counters_out)
{
    static_assert(FAST == 0 || LIGHT == SWEGL_B200_LIGHT_PHONG, "only Phong lighting has a fast variant");
    __shared__ ViewParams vp;
    __shared__ FrameParams fp;
    __shared__ FragWarp fwarp[FRAG_ROWS];
    __shared__ uint32_t s_next[2];
    __shared__ uint32_t s_nb[4], s_cost[2];                                 // entries of the busy list's cost classes; start clock and tile of the current item
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int band_rows = g.band1 - g.band0, anchor = g.band0 - g.vy;
    pdl_trigger(pl.early_trigger);
    FRAG_TL(0, global_ns());
    // the parameter block is written by the copy at the head of the chain, not by a kernel: readable before the wait
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += FRAG_TPB) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    for (int w = threadIdx.x; w < (int)(sizeof(FrameParams) / 4); w += FRAG_TPB) reinterpret_cast<uint32_t *>(&fp)[w] = reinterpret_cast<const uint32_t *>(fpp)[w];
    __syncthreads();
    pdl_wait();                                                             // k_spans' bins, chunks, fragment stream, tile list
    Counters *const cn = pl.counters;
    const uint32_t n_busy = cn->n_busy;
    const uint32_t stamp = vp.stamp;
    if (threadIdx.x < 4) s_nb[threadIdx.x] = cn->n_busy_b[threadIdx.x];
    __syncthreads();
    FRAG_TL(1, global_ns());
#ifdef FRAG_PROBE_TIMELINE
    int tl_slot = 2;
#endif
    // CTA-level items: [0, n_busy) = the busy tiles, then n_groups streaming groups (8 fragment tiles to clear and 8 DoF tiles
    // to classify, one of each per warp).  Busy tiles come first: the CTAs whose tile was light finish early and take the
    // streaming groups, whose stores then drain underneath the arithmetic of the heavy tiles.
    uint32_t n_groups = (uint32_t)(g.n_tiles + FRAG_ROWS - 1) / FRAG_ROWS;
    if (g.dof) n_groups = max(n_groups, (uint32_t)(g.n_dof + FRAG_ROWS - 1) / FRAG_ROWS);
    const uint32_t n_pops = n_busy + n_groups;
    const uint32_t dof_fill = g.dof ? dof_background_value(vp) : 0u;
    const uint64_t KEY_INIT = (uint64_t)MAXZ_BITS << 32;
    FragWarp &fw = fwarp[warp];
    uint32_t my_covered = 0;
    uint32_t tx0 = 0xFFFFFFFFu, ty0 = 0xFFFFFFFFu, tx1 = 0, ty1 = 0;      // bounding box of the busy tiles this CTA worked on (tile units)

    // The first item of every CTA is its own index (no atomic: a 4K frame has about as many busy tiles as there are CTAs,
    // and a thousand CTAs hitting one counter in the same microsecond would queue up behind each other); later ones come
    // from the queue, one atomic per CTA and item, fetched while the current item is worked on.
    uint32_t p = blockIdx.x;
    for (int it = 0; p < n_pops; it ^= 1) {
        uint32_t next_raw = 0;
        if (threadIdx.x == 0) next_raw = atomicAdd(&cn->frag_queue, 1u);    // consumed at the end of the iteration: the round trip is free

        if (p >= n_busy) {
            // ---- a streaming group ----
            const uint32_t q = (p - n_busy) * FRAG_ROWS + (uint32_t)warp;   // this warp's fragment tile / DoF tile
#ifndef FRAG_PROBE_NOCLEAR
            if (q < (uint32_t)g.n_tiles && pl.tile_stamp[q] != stamp) {     // clear duty: nothing was drawn into fragment tile q
                const int ty = (int)q / g.ntx, tx = (int)q - ty * g.ntx;
                bool with_color = !g.skip_bg;
                if (g.dof) {
                    // tmp colour is only ever read through the window of a DoF tile that is computed, i.e. one with a busy
                    // fragment tile under its window: two tiles under one window are at most 1 column and 5 rows apart
                    constexpr int NDY = DOF_OH / FRAG_ROWS + 1;              // 5
                    bool near = false;
                    for (int k = lane; k < 3 * (2 * NDY + 1); k += 32) {
                        const int yy = ty + k / 3 - NDY, xx = tx + k % 3 - 1;
                        if (yy >= 0 && yy < g.nty && xx >= 0 && xx < g.ntx) near = near || pl.tile_stamp[yy * g.ntx + xx] == stamp;
                    }
                    with_color = __any_sync(0xFFFFFFFFu, near);
                }
                const int row0 = anchor + ty * FRAG_ROWS, bx0 = tx * FRAG_STRETCH;
                const int rows_here = min(FRAG_ROWS, band_rows - ty * FRAG_ROWS);
                const int px_left = g.vw - (bx0 << 5);
                uint32_t *c0 = color + (size_t)(g.vy + row0) * color_pitch + g.vx + (bx0 << 5);
                float *d0 = depth + (size_t)row0 * g.vw + (bx0 << 5);
                if (px_left >= FRAG_STRETCH * 32 && ((color_pitch | g.vw) & 3) == 0
                    && ((reinterpret_cast<uintptr_t>(c0) | reinterpret_cast<uintptr_t>(d0)) & 15) == 0) {
                    // the common case, a whole tile of 16-byte aligned rows: lane = 4 pixels, one 128-bit store per row and plane
                    const float maxz = __uint_as_float(MAXZ_BITS);
                    uint4 *cq = reinterpret_cast<uint4 *>(c0) + lane;
                    float4 *dq = reinterpret_cast<float4 *>(d0) + lane;
                    const int cstep = color_pitch >> 2, dstep = g.vw >> 2;
                    #pragma unroll
                    for (int r = 0; r < FRAG_ROWS; r++) {
                        if (r < rows_here) {
                            if (with_color) cq[(size_t)r * cstep] = make_uint4(0, 0, 0, 0);
                            dq[(size_t)r * dstep] = make_float4(maxz, maxz, maxz, maxz);
                        }
                    }
                } else {
                    for (int r = 0; r < rows_here; r++)
                        clear_tile_row(c0 + (size_t)r * color_pitch, d0 + (size_t)r * g.vw, px_left, lane, with_color);
                }
            }
            if (g.dof && q < (uint32_t)g.n_dof) dof_tile_duty(pl, g, stamp, dof_fill, dof_dst, q, lane);   // DoF duty
#endif
        } else do {
            // ---- a busy tile: the CTA's warps take its rows (their spans, pieces and texels are neighbours in memory) ----
#ifdef FRAG_PROBE_TIMELINE
            const bool tl_first = tl_slot == 2;
#endif
            FRAG_PH(8);
            uint32_t q = p, cls = 0;                                        // item p of the heaviest-first order: entry q of cost class cls
            #pragma unroll
            for (uint32_t k = 0; k < 3; k++) { const uint32_t nk = s_nb[k]; if (cls == k && q >= nk) { q -= nk; cls = k + 1; } }
            const uint32_t t = pl.busy_list[cls * pl.busy_stride + q];      // tile row << 16 | tile column (k_spans)
            const int ty = (int)(t >> 16), tx = (int)(t & 0xFFFFu);
            if (t == 0xFFFFFFFEu) break;                                    // (never: makes the stamp below wait for the load)
            if (threadIdx.x == 0) { s_cost[0] = (uint32_t)clock(); s_cost[1] = (uint32_t)(ty * g.ntx + tx); }
            FRAG_PH(9);
            tx0 = min(tx0, (uint32_t)tx); tx1 = max(tx1, (uint32_t)tx + 1); ty0 = min(ty0, (uint32_t)ty); ty1 = max(ty1, (uint32_t)ty + 1);
            if (warp >= min(FRAG_ROWS, band_rows - ty * FRAG_ROWS)) break;
            const int row = anchor + ty * FRAG_ROWS + warp;                 // viewport-relative
            const int y = g.vy + row;
            const int bx0 = tx * FRAG_STRETCH;
            const int nb = min(FRAG_STRETCH, g.nbx - bx0);
            const int px_left = g.vw - (bx0 << 5);                          // pixels from the stretch start to the row end
            // the bins' piece counts (and overflow lists), fetched and reset for the next frame
            const size_t bin0 = (size_t)row * g.nbx + bx0;
            int32_t cnt = 0, over = -1;
            if (lane < nb) {
                cnt = pl.bin_cnt[bin0 + lane];
                if (cnt > 0) pl.bin_cnt[bin0 + lane] = 0;
                if (cnt > BIN_SLOTS) { over = pl.bin_head[bin0 + lane]; pl.bin_head[bin0 + lane] = -1; }
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, cnt > 0);
            FRAG_PH(10);
            uint32_t *crow = color + (size_t)y * color_pitch + g.vx + (bx0 << 5);
            float *drow = depth + (size_t)row * g.vw + (bx0 << 5);
            const bool aligned = ((reinterpret_cast<uintptr_t>(crow) | reinterpret_cast<uintptr_t>(drow)) & 15) == 0;
            // ---- empty bins ----
            if (mask != (1u << nb) - 1u) {
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
            if (!mask) break;
            // ---- non-empty bins.  The first FRAG_STAGE slot records of EVERY bin of the row are fetched at once (lane = half a
            //      record, up to four independent 128-bit loads per lane), and every record's fragment-stream segment, span
            //      constants and colour binding are prefetched into L1 right away: the row pays these round trips once, side by
            //      side, instead of one after the other in front of every bin ----
            __syncwarp();                                                   // fw of the previous item is consumed
            #pragma unroll
            for (int b = 0; b < FRAG_STRETCH; b++) {
                const int nk = min(__shfl_sync(0xFFFFFFFFu, cnt, b), FRAG_STAGE);
                if (lane < 2 * nk) fw.rec[b][lane] = reinterpret_cast<const uint4 *>(pl.bin_slots + (bin0 + b) * BIN_SLOTS)[lane];
            }
            __syncwarp();
            FRAG_PH(11);
#ifndef FRAG_NO_PREFETCH
            #pragma unroll
            for (int h = 0; h < FRAG_STAGE / 8; h++) {
                const int b = lane >> 3, k = (lane & 7) + 8 * h;           // FRAG_STRETCH * 8 == 32 lanes
                if (k < min(__shfl_sync(0xFFFFFFFFu, cnt, b), FRAG_STAGE)) {
                    const uint4 r0 = fw.rec[b][2 * k];
                    const uint2 r1 = *reinterpret_cast<const uint2 *>(&fw.rec[b][2 * k + 1]);
                    prefetch_l1(&pl.frag_u[r0.x + (r0.y & 0xFFu)]);
                    prefetch_l1(&pl.frag_u[r0.x + (r0.y >> 8) - 1u]);
                    prefetch_l1(&pl.span_shades[r1.y]);
                    prefetch_l1(&pl.shades[r1.x].color);
                }
            }
#endif
            // ---- depth resolve, bin by bin: lane = pixel, the nearest fragment of every pixel is kept in registers ----
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                uint64_t best = KEY_INIT;
                float best_u = 0.f;
                uint32_t best_span = 0xFFFFFFFFu;
                const int n = min(__shfl_sync(0xFFFFFFFFu, cnt, b), BIN_SLOTS);
                const Chunk *slots = pl.bin_slots + (bin0 + b) * BIN_SLOTS;
                const uint4 *rec = fw.rec[b];
                for (int base = 0; base < n; base += FRAG_STAGE) {
                    const int m = min(FRAG_STAGE, n - base);
                    if (base) {                                                     // a bin with more than FRAG_STAGE pieces: the next round
                        __syncwarp();
                        if (lane < 2 * m) fw.rec[b][lane] = reinterpret_cast<const uint4 *>(slots + base)[lane];
                        __syncwarp();
                    }
                    if (CROWD && m >= FRAG_COVER_MIN) {
                        // A crowded round (a dense mesh: many pieces of a few pixels each; CROWD is chosen per frame, like
                        // the span kernel, from the previous frame's scanline count -- the mere presence of this path costs
                        // the lean kernel a microsecond on a frame of few, wide pieces).  Taking the pieces one after the
                        // other costs the whole warp a pass per piece although each touches a few lanes; instead every lane
                        // first notes WHICH of the staged pieces cover it (one bit per piece, from the broadcast xs | xe words),
                        // then walks only its own pieces -- the warp makes as many turns as its most overdrawn pixel has
                        // pieces, FRAG_TURN pieces per turn so that their stream loads go out together.  Same keys, same minimum.
                        uint32_t cover = 0;
                        #pragma unroll 4
                        for (int k = 0; k < m; k++) {
                            const uint32_t xx = reinterpret_cast<const uint32_t *>(&rec[2 * k])[1];
                            const unsigned xs = xx & 0xFFu, wd = (xx >> 8) - xs;
                            cover |= ((unsigned)lane - xs < wd ? 1u : 0u) << k;
                        }
                        while (__any_sync(0xFFFFFFFFu, cover != 0u)) {
                            bool in[FRAG_TURN]; int pc[FRAG_TURN]; uint4 r0[FRAG_TURN]; float u[FRAG_TURN];
                            #pragma unroll
                            for (int j = 0; j < FRAG_TURN; j++) {
                                in[j] = cover != 0u;
                                pc[j] = in[j] ? __ffs(cover) - 1 : 0;
                                cover &= cover - 1u;
                                r0[j] = rec[2 * pc[j]];
                            }
                            #pragma unroll
                            for (int j = 0; j < FRAG_TURN; j++) u[j] = pl.frag_u[in[j] ? r0[j].x + (uint32_t)lane : 0u];
                            #pragma unroll
                            for (int j = 0; j < FRAG_TURN; j++) {
                                const float z = fadd(__uint_as_float(r0[j].z), fmul(__uint_as_float(r0[j].w), u[j]));   // value(0), renderer.cpp:488
                                if (in[j] && z >= NEAR_Z) {                         // renderer.cpp:489-492
                                    const uint2 r1 = *reinterpret_cast<const uint2 *>(&rec[2 * pc[j] + 1]);
                                    const uint64_t key = ((uint64_t)__float_as_uint(z) << 32) | r1.x;
                                    if (key < best) { best = key; best_u = u[j]; best_span = r1.y; }
                                }
                            }
                        }
                        continue;
                    }
                    // groups of RESOLVE_UNROLL pieces whose fragment-stream loads are issued back to back, WITHOUT a remainder
                    // loop: a bin holds ~3 pieces on average, and pieces taken one by one pay one L2 round trip each.  Every lane
                    // loads for every piece of the group (a lane outside the piece, or a piece beyond the bin's last, reads
                    // entry 0 instead), so no branch separates the loads.
                    for (int k0 = 0; k0 < m; k0 += RESOLVE_UNROLL) {
                        uint4 r0[RESOLVE_UNROLL]; float u[RESOLVE_UNROLL]; bool in[RESOLVE_UNROLL];
                        #pragma unroll
                        for (int j = 0; j < RESOLVE_UNROLL; j++) {
                            r0[j] = rec[2 * min(k0 + j, m - 1)];                    // frag0, xs_xe, v0, v1
                            const unsigned xs = r0[j].y & 0xFFu, wd = (r0[j].y >> 8) - xs;
                            in[j] = (unsigned)lane - xs < wd && k0 + j < m;
                        }
                        #pragma unroll
                        for (int j = 0; j < RESOLVE_UNROLL; j++)
                            u[j] = pl.frag_u[in[j] ? r0[j].x + (uint32_t)lane : 0u];   // qpixel.ualpha, replayed by k_spans
                        #pragma unroll
                        for (int j = 0; j < RESOLVE_UNROLL; j++) {
                            const float z = fadd(__uint_as_float(r0[j].z), fmul(__uint_as_float(r0[j].w), u[j]));   // value(0), renderer.cpp:488
                            if (in[j] && z >= NEAR_Z) {                             // renderer.cpp:489-492
                                const uint2 r1 = *reinterpret_cast<const uint2 *>(&rec[2 * (k0 + j) + 1]);         // slot, span
                                const uint64_t key = ((uint64_t)__float_as_uint(z) << 32) | r1.x;
                                if (key < best) { best = key; best_u = u[j]; best_span = r1.y; }
                            }
                        }
                    }
                }
                for (int32_t c = __shfl_sync(0xFFFFFFFFu, over, b); c >= 0;) {      // more than BIN_SLOTS pieces: the rest, serially
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
#ifdef FRAG_PROBE_TIMELINE
                if (tl_first && __any_sync(0xFFFFFFFFu, best != KEY_INIT || true)) FRAG_TL(12 + (b & 1) * 2, global_ns());
#endif
                if (inside) drow[px] = __uint_as_float((uint32_t)(best >> 32));
                // depth and colour of the bin at once: lane = pixel, the winner is shaded by its own lane (only the visible
                // fragment of a pixel is ever shaded).  Resolve and shading alternate bin by bin, so the warps of an SM drift out
                // of phase -- some wait for loads while others compute.  (Measured alternatives: winners of the whole 128-pixel
                // stretch queued and shaded in dense batches of 32 -- equal; all bins of the row resolved first, their winners
                // parked in shared memory and the texels prefetched before any shading -- 10 % slower on the truck, 2x on
                // BrainStem: 12 KB more shared memory per CTA takes it from the L1 the texels live in.)
                uint32_t out = 0u;                                                  // background pixels of a used bin get 0
                if (hit) out = shade_winner<LIGHT, TEX, FAST>(pl, s.texels, vp, fp, best_span, (uint32_t)best, best_u, g.k4b);
                if (inside) crow[px] = out;
#ifdef FRAG_PROBE_TIMELINE
                if (tl_first && __any_sync(0xFFFFFFFFu, out != 0x12345u)) FRAG_TL(13 + (b & 1) * 2, global_ns());
#endif
                my_covered += __popc(__ballot_sync(0xFFFFFFFFu, hit));
            }
            __syncwarp();                                                   // fw is reused by the next item
        } while (0);

        if (threadIdx.x == 0) s_next[it] = gridDim.x + next_raw;
        // every warp is done with item p; the next index is visible.  (The barrier is worth keeping: with a ring of item
        // indices in shared memory and every warp running ahead on its own -- no barrier, measured -- the truck frame went
        // from 45 to 60 us and the 8K sphere from 368 to 549 us: the rows of one tile share spans, slot records and texels in
        // L1, and warps that drift apart onto different tiles evict each other's.)
        __syncthreads();
#ifdef FRAG_PROBE_TIMELINE
        FRAG_TL(tl_slot, (global_ns() << 1) | (p >= n_busy ? 1ull : 0ull)); tl_slot++;
#endif
        if (threadIdx.x == 0 && p < n_busy) {                               // what this tile cost: the next frame's hint for the order
            const uint32_t cost = ((uint32_t)clock() - s_cost[0]) >> 6;
            pl.tile_cost[s_cost[1]] = cost;
            atomicAdd(&pl.cost_acc[0], cost);
        }
        p = s_next[it];
    }
    if (count_covered && lane == 0 && my_covered) atomicAdd(&cn->n_covered, my_covered);
    // The frame's bounding box (for the host's partial read-back) is the union over the CTAs; the last CTA to finish
    // publishes the counters of the frame (pool demand and overflow flags of k_setup / k_spans, covered pixels, the box)
    // to the pinned slot the host polls -- instead of a D2H copy node at the end of the graph.
    __syncthreads();
    if (threadIdx.x == 0) {
        if (tx1) {
            atomicMin(&cn->bb_x0, tx0 * (FRAG_STRETCH * 32)); atomicMin(&cn->bb_y0, (uint32_t)anchor + ty0 * FRAG_ROWS);
            atomicMax(&cn->bb_x1, min((uint32_t)g.vw, tx1 * (FRAG_STRETCH * 32)));
            atomicMax(&cn->bb_y1, min((uint32_t)(anchor + band_rows), (uint32_t)anchor + ty1 * FRAG_ROWS));
        }
        __threadfence();
        if (atomicAdd(&cn->frag_done, 1u) == gridDim.x - 1) {
            __threadfence();
            if (n_busy) pl.cost_acc[1] = atomicExch(&pl.cost_acc[0], 0u) / n_busy;     // mean tile cost of this frame
            if (h_counters_out)
                for (int w = 0; w < (int)(sizeof(Counters) / 4); w++)
                    reinterpret_cast<volatile uint32_t *>(h_counters_out)[w] = reinterpret_cast<volatile const uint32_t *>(cn)[w];
        }
    }
}

// DoF tile classification on its own (the transparency-layer kernel has no persistent warps to fold it into)
__global__ void __launch_bounds__(256) k_dof_classify(const ViewParams *__restrict__ vpp, Pools pl, FragGeom g, uint32_t *__restrict__ dof_dst)
{
    pdl_trigger(pl.early_trigger);
    __shared__ ViewParams vp;
    for (int w = threadIdx.x; w < (int)(sizeof(ViewParams) / 4); w += 256) reinterpret_cast<uint32_t *>(&vp)[w] = reinterpret_cast<const uint32_t *>(vpp)[w];
    __syncthreads();
    pdl_wait();
    const uint32_t d = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (d < (uint32_t)g.n_dof) dof_tile_duty(pl, g, vp.stamp, dof_background_value(vp), dof_dst, d, threadIdx.x & 31);
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
    pdl_trigger(pl.early_trigger);
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

    // the bins' piece counts and overflow lists (common.cuh BIN_SLOTS), fetched and reset for the next frame
    const size_t bin0 = (size_t)row * vp.nbx + bx0;
    int32_t cnt = 0, head = -1;
    if (lane < nb) {
        cnt = pl.bin_cnt[bin0 + lane];
        if (cnt > 0) pl.bin_cnt[bin0 + lane] = 0;
        if (cnt > BIN_SLOTS) { head = pl.bin_head[bin0 + lane]; pl.bin_head[bin0 + lane] = -1; }
    }
    unsigned mask = __ballot_sync(0xFFFFFFFFu, cnt > 0);
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
        const int n_in = min(__shfl_sync(0xFFFFFFFFu, cnt, b), BIN_SLOTS);
        uint64_t best = KEY_INIT;                                           // opaque base: z bits << 32 | slot
        float best_u = 0.f;
        uint32_t best_span = 0xFFFFFFFFu;
        // kept transparent fragments, nearest first: key = z bits << 32 | ~slot (a later draw at equal depth is nearer)
        uint64_t tkey[MAX_LAYERS]; float tu[MAX_LAYERS]; uint32_t tspan[MAX_LAYERS];
        int tn = 0;
        for (int k = 0; k < n_in || c >= 0; k++) {                          // the bin's own slots, then its overflow list
            Chunk ch;
            if (k < n_in) ch = pl.bin_slots[(bin0 + b) * BIN_SLOTS + k];
            else { ch = pl.chunks[c]; c = ch.next; }
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
// k_dof is a persistent grid over the list of DoF output tiles (64 x 32) that k_fragments found something under
// (dof_tile_duty; every other tile is a constant k_fragments stored already).  Per tile:
//   1. the (64+16) x (32+9) source window of colour and depth arrives in shared memory by TMA (two 2-D tensor loads,
//      cp.async.bulk.tensor; coordinates outside the viewport are zero-filled by the hardware, which replaces all edge
//      predication: a depth word of 0 means "no such pixel").  The load of the NEXT tile's window is issued as soon as
//      this tile's window has been consumed, so it lands underneath steps 3-5;
//   2. lane = consecutive pixel: every staged pixel becomes ONE packed 64-bit contribution
//          b | g << 15 | r << 30 | 1 << 45          (zero when the pixel's blur factor is 0 or it does not exist)
//      15 bits hold a box sum of <= 100 bytes, 7 bits its count -- a box of (2r)^2 <= 100 taps; the prefix sums
//      themselves overflow their fields, but the four-corner difference is exact modulo 2^64;
//   3. summed-area table: one thread per row, then one thread per column (the conflict-free order for a row stride
//      of 81 entries);
//   4. per output pixel 4 corner lookups, exact v / n by a multiply-high; radius-0 / empty boxes copy the source;
//   5. a window that turns out to hold only background (the tile list is conservative) is a constant fill.
// The blur radius needs no per-pixel division: r(t) is a monotone step function of t = |focal_distance - z|, so the
// host finds its 5 step positions by exact bisection over the float bit patterns (abi.cu dof_thresholds).
// Without TMA (viewport width not a multiple of 4 pixels: the tensor's row pitch must be a multiple of 16 bytes) the
// window is staged by predicated loads instead; everything downstream is the same.
// ----------------------------------------------------------------------------------------
static constexpr int DOF_LO = 5, DOF_HI = 4;
static constexpr int DOF_SH = DOF_OH + DOF_LO + DOF_HI;      // 41 source rows
static constexpr int DOF_X0 = 8;                             // the staged window starts 8 columns left of the tile (16-byte aligned)
static constexpr int DOF_WW = DOF_OW + 16;                   // 80 staged columns
static constexpr int DOF_PW = DOF_WW + 1;                    // SAT row stride in entries (col 0 = zero border; odd -> rows on different banks)
static constexpr int DOF_THREADS = 256;
#ifndef DOF_CTAS_PER_SM
#define DOF_CTAS_PER_SM 4         // persistent CTAs per SM (64 registers, 45 KB of shared memory each)
#endif
static constexpr int DOF_NPX = DOF_SH * DOF_WW;              // 3280 staged pixels

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
// what a window of untouched background (colour 0, depth 0x7F7F7F7F) blurs to:
// radius 0: copy (0); else every tap counts iff its blur != 0 -> average of zeros with alpha 255, or no taps -> copy
SB_DEV uint32_t dof_background_value(const ViewParams &vp)
{
    const DofClass bgc = dof_classify(vp, __uint_as_float(MAXZ_BITS));
    return (bgc.radius != 0 && bgc.counts) ? 0xFF000000u : 0u;
}

struct DofGeom {
    int32_t vw, vh;             // the viewport (source arrays are [vh][src_pitch] colour, [vh][vw] depth)
    int32_t anchor;             // viewport-relative row of the DoF tile grid's origin (= first drawn row)
    int32_t out0, out1;         // viewport-relative rows [out0, out1) to produce
    int32_t ndx;                // DoF tiles per row
    int32_t src_pitch, dst_pitch;
    int32_t use_tma;
};

struct __align__(128) DofSmem {
    uint32_t raw_c[DOF_NPX];                                    // staged colour window (TMA destination: 128-byte aligned)
    alignas(128) uint32_t raw_z[DOF_NPX];                       // staged depth window (bit patterns)
    alignas(16) unsigned long long sat[(DOF_SH + 1) * DOF_PW];  // packed contributions -> summed-area table (row 0 / column 0: zero border)
    uint8_t rad[DOF_OW * DOF_OH];                               // blur radius of the output pixels
    uint32_t magic[128];                                        // ceil(2^28 / n): exact v / n for v < 2^15, n <= 100
    ViewParams vp;
    alignas(8) unsigned long long bar;                          // mbarrier of the window loads
    uint32_t item[2];                                           // DoF tile of this / the next iteration, 0xFFFFFFFF = none
};

SB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// both window loads of DoF tile d, completing on `bar`
SB_DEV void dof_issue_window(const CUtensorMap *tm_color, const CUtensorMap *tm_depth, DofSmem &sm, const DofGeom &g, uint32_t d)
{
    const int j = (int)d / g.ndx, i = (int)d - j * g.ndx;
    const int x = i * DOF_OW - DOF_X0, y = g.anchor + j * DOF_OH - DOF_LO;
    const uint32_t bar = smem_u32(&sm.bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // earlier generic-proxy accesses to the window are ordered before the refill
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2 * DOF_NPX * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(sm.raw_c)), "l"(tm_color), "r"(bar), "r"(x), "r"(y) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(sm.raw_z)), "l"(tm_depth), "r"(bar), "r"(x), "r"(y) : "memory");
}
SB_DEV void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000ll) __trap();                 // ~2 s: a tensor load that never lands must not hang the GPU
    }
}

__global__ void __launch_bounds__(DOF_THREADS, DOF_CTAS_PER_SM) k_dof(const __grid_constant__ CUtensorMap tm_color, const __grid_constant__ CUtensorMap tm_depth,
                                                        const ViewParams *__restrict__ vpp, Pools pl, DofGeom g,
                                                        const uint32_t *__restrict__ src, const float *__restrict__ depth, uint32_t *__restrict__ dst)
{
    extern __shared__ __align__(128) unsigned char dof_smem_raw[];
    DofSmem &sm = *reinterpret_cast<DofSmem *>(dof_smem_raw);
    const int tid = threadIdx.x;
    pdl_trigger(pl.early_trigger);
    // ---- prologue, independent of the predecessor ----
    for (int k = tid; k < (int)(sizeof(ViewParams) / 4); k += DOF_THREADS) reinterpret_cast<uint32_t *>(&sm.vp)[k] = reinterpret_cast<const uint32_t *>(vpp)[k];
    if (tid < 128) sm.magic[tid] = tid ? (uint32_t)(((1u << 28) + tid - 1) / tid) : 0u;
    for (int i = tid; i < DOF_PW; i += DOF_THREADS) sm.sat[i] = 0;                              // zero row 0 ...
    for (int i = tid; i <= DOF_SH; i += DOF_THREADS) sm.sat[i * DOF_PW] = 0;                    // ... and column 0, for good
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&sm.bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();                                                             // k_fragments' colour, depth and tile list
    const ViewParams &vp = sm.vp;
    Counters *const cn = pl.counters;
    const uint32_t n = cn->n_dof_busy;
    if (tid == 0) {
        const uint32_t i0 = atomicAdd(&cn->dof_queue, 1u);
        const uint32_t d0 = i0 < n ? pl.dof_list[i0] : 0xFFFFFFFFu;
        sm.item[0] = d0;
        if (d0 != 0xFFFFFFFFu && g.use_tma) dof_issue_window(&tm_color, &tm_depth, sm, g, d0);
    }
    __syncthreads();
    const uint32_t fill = dof_background_value(vp);
    // the classification constants live in registers for the whole kernel (13 pixels per thread and tile use them)
    const float fd = vp.focal_distance, th0 = vp.dof_t[0], th1 = vp.dof_t[1], th2 = vp.dof_t[2], th3 = vp.dof_t[3], th4 = vp.dof_t[4], t_on = vp.dof_on;
    const int const_radius = vp.dof_const_radius;
    uint32_t phase = 0;
    int cur = 0;
    for (;;) {
        const uint32_t d = sm.item[cur];
        if (d == 0xFFFFFFFFu) break;
        uint32_t nxt_raw = 0;
        if (tid == 0) nxt_raw = atomicAdd(&cn->dof_queue, 1u);              // the next index travels under the wait for the window
        const int j = (int)d / g.ndx, i = (int)d - j * g.ndx;
        const int ox = i * DOF_OW, oy = g.anchor + j * DOF_OH;              // viewport-relative origin of the output tile
        if (g.use_tma) {
            mbar_wait(&sm.bar, phase);
            phase ^= 1u;
        } else {
            for (int p = tid; p < DOF_NPX; p += DOF_THREADS) {
                const int sy = p / DOF_WW, sx = p - sy * DOF_WW;
                const int gx = ox - DOF_X0 + sx, gy = oy - DOF_LO + sy;
                uint32_t c = 0u, z = 0u;
                if (gx >= 0 && gx < g.vw && gy >= 0 && gy < g.vh) {
                    c = __ldg(&src[(size_t)gy * g.src_pitch + gx]);
                    z = __float_as_uint(__ldg(&depth[(size_t)gy * g.vw + gx]));
                }
                sm.raw_c[p] = c; sm.raw_z[p] = z;
            }
            __syncthreads();
        }
        // ---- staged pixels -> packed contributions (+ the radius of the output pixels); is there anything but background? ----
        bool bg = true;
        #pragma unroll 2
        for (int p = tid; p < DOF_NPX; p += DOF_THREADS) {
            const int sy = p / DOF_WW, sx = p - sy * DOF_WW;
            const uint32_t c = sm.raw_c[p], zb = sm.raw_z[p];
            unsigned long long v = 0;
            uint32_t radius = 0;
            if (zb != 0u) {                                                 // the pixel exists (a real depth is >= 0.001 or 0x7F7F7F7F)
                bg = bg && c == 0u && zb == MAXZ_BITS;
                const float t = fabsf(fsub(fd, __uint_as_float(zb)));      // dof_classify with the constants in registers
                radius = (t >= th0) + (t >= th1) + (t >= th2) + (t >= th3) + (t >= th4);
                bool counts = t > t_on;
                if (const_radius >= 0) { radius = (uint32_t)const_radius; counts = true; }
                if (counts) {
                    const uint32_t b = c & 0xFFu, gg = (c >> 8) & 0xFFu, r = (c >> 16) & 0xFFu;
                    const uint32_t lo = b | (gg << 15) | (r << 30), hi = (r >> 2) | (1u << 13);
                    v = ((unsigned long long)hi << 32) | lo;
                }
            }
            sm.sat[(sy + 1) * DOF_PW + sx + 1] = v;
            const int ty = sy - DOF_LO, tx = sx - DOF_X0;
            if ((unsigned)ty < (unsigned)DOF_OH && (unsigned)tx < (unsigned)DOF_OW) sm.rad[ty * DOF_OW + tx] = (uint8_t)radius;
        }
        const bool all_bg = __syncthreads_and(bg);                          // (also: the raw window is consumed, the contributions are visible)
        if (tid == 0) {
            const uint32_t dn = nxt_raw < n ? pl.dof_list[nxt_raw] : 0xFFFFFFFFu;
            sm.item[cur ^ 1] = dn;
            if (dn != 0xFFFFFFFFu && g.use_tma) dof_issue_window(&tm_color, &tm_depth, sm, g, dn);   // lands under the scans and the outputs
        }
        const int ya = max(oy, g.out0), yb = min(min(oy + DOF_OH, g.out1), g.vh);
        if (all_bg) {
            // bins were touched nearby, but this window only sees background-coloured pixels: a constant
            const int wd = min(DOF_OW, g.vw - ox);
            const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (g.dst_pitch & 3) == 0 && wd == DOF_OW;
            if (vec) {
                const uint4 v4 = make_uint4(fill, fill, fill, fill);
                for (int k = tid; k < (yb - ya) * (DOF_OW / 4); k += DOF_THREADS)
                    *reinterpret_cast<uint4 *>(dst + (size_t)(ya + k / (DOF_OW / 4)) * g.dst_pitch + ox + (k % (DOF_OW / 4)) * 4) = v4;
            } else {
                for (int k = tid; k < (yb - ya) * wd; k += DOF_THREADS)
                    dst[(size_t)(ya + k / wd) * g.dst_pitch + ox + k % wd] = fill;
            }
        } else {
            if (tid < DOF_SH) {                                             // inclusive scan along each row
                unsigned long long a = 0;
                unsigned long long *row = &sm.sat[(tid + 1) * DOF_PW];
                #pragma unroll 8
                for (int sx = 1; sx <= DOF_WW; sx++) { a += row[sx]; row[sx] = a; }
            }
            __syncthreads();
            if (tid < DOF_WW) {                                             // then down each column
                unsigned long long a = 0;
                unsigned long long *col = &sm.sat[tid + 1];
                #pragma unroll 8
                for (int sy = 1; sy <= DOF_SH; sy++) { a += col[sy * DOF_PW]; col[sy * DOF_PW] = a; }
            }
            __syncthreads();
            // ---- outputs: consecutive lanes take consecutive pixels (conflict-free SAT reads, coalesced stores).
            //      Staged column of output pixel tx is tx + DOF_X0; SAT entry (J, I) = sum over staged rows < J, columns < I. ----
            #pragma unroll 2
            for (int p = tid; p < DOF_OW * DOF_OH; p += DOF_THREADS) {
                const int ty = p / DOF_OW, tx = p % DOF_OW;
                const int x = ox + tx, y = oy + ty;
                if (y < ya || y >= yb || x >= g.vw) continue;
                const int radius = sm.rad[p];
                uint32_t out;
                bool have = false;
                if (radius != 0) {
                    // window rows [y-r, y+r) -> SAT rows (ty+5-r, ty+5+r]; columns likewise with the +8 staging offset; the zero
                    // border and the zeros stored for out-of-viewport pixels implement the max(0,..)/min(w|h,..) clipping
                    const int J0 = ty + DOF_LO - radius, J1 = ty + DOF_LO + radius;
                    const int I0 = tx + DOF_X0 - radius, I1 = tx + DOF_X0 + radius;
                    const unsigned long long S = (sm.sat[J1 * DOF_PW + I1] + sm.sat[J0 * DOF_PW + I0])
                                               - (sm.sat[J0 * DOF_PW + I1] + sm.sat[J1 * DOF_PW + I0]);
                    const uint32_t lo = (uint32_t)S, hi = (uint32_t)(S >> 32);
                    const uint32_t count = (hi >> 13) & 0x7Fu;
                    if (count) {
                        // exact floor(v / count) for v < 2^15, count <= 100: (v * ceil(2^28 / count)) >> 28, as one IMAD.HI
                        const uint32_t m = sm.magic[count];
                        const uint32_t b = __umulhi((lo & 0x7FFFu) << 4, m);
                        const uint32_t gg = __umulhi(((lo >> 15) & 0x7FFFu) << 4, m);
                        const uint32_t r = __umulhi((((lo >> 30) | (hi << 2)) & 0x7FFFu) << 4, m);
                        out = b | (gg << 8) | (r << 16) | 0xFF000000u;
                        have = true;
                    }
                }
                if (!have) out = __ldg(&src[(size_t)y * g.src_pitch + x]);
                dst[(size_t)y * g.dst_pitch + x] = out;
            }
        }
        __syncthreads();                                                    // SAT, radii and item[] are reused by the next tile
        cur ^= 1;
    }
}

// ----------------------------------------------------------------------------------------
// Staging of a finished view for the pipelined read-back (swegl_b200_render_viewport_async): the view's rows of the screen
// go into a staging image the next frames do not touch.  Outside the bounding box of what was drawn (grown by the blur
// radius under DoF-R) a frame is one constant, so an image that already holds a complete earlier frame of the same view
// and background only needs the union of that frame's box and this one's -- 7 MB instead of 33 MB on the 4K truck
// frame.  The boxes never leave the device: this frame's comes from the counters k_fragments left, the image's previous
// one from a two-entry record (read entry `par`, write entry `par ^ 1`: no CTA can see the new box too early).
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stage_rect(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, int src_pitch, int vw,
                                                     int row_a, int row_b, int grow, const Counters *__restrict__ cn, StageRect *rec, int par, int full)
{
    int fx0 = 0, fx1 = 0, fy0 = 0, fy1 = 0;                                 // this frame's box, as the host derives it (abi.cu issue_readback)
    const uint32_t bx0 = cn->bb_x0, bx1 = cn->bb_x1, by0 = cn->bb_y0, by1 = cn->bb_y1;
    if (bx1 > bx0 && by1 > by0) {
        fx0 = max(0, ((int)bx0 - grow) & ~3); fx1 = min(vw, ((int)bx1 + grow + 3) & ~3);
        fy0 = max(row_a, (int)by0 - grow); fy1 = min(row_b, (int)by1 + grow);
    }
    if (cn->overflow) { full = 1; fx0 = 0; fx1 = vw; fy0 = row_a; fy1 = row_b; }   // an incomplete frame: nothing is known about it
    int cx0 = 0, cx1 = vw, cy0 = row_a, cy1 = row_b;
    if (!full) {
        const StageRect o = rec[par];
        const bool old_empty = o.x1 <= o.x0 || o.y1 <= o.y0, new_empty = fx1 <= fx0 || fy1 <= fy0;
        if (old_empty && new_empty) { cx0 = cx1 = cy0 = cy1 = 0; }
        else if (old_empty) { cx0 = fx0; cx1 = fx1; cy0 = fy0; cy1 = fy1; }
        else if (new_empty) { cx0 = o.x0; cx1 = o.x1; cy0 = o.y0; cy1 = o.y1; }
        else { cx0 = min(fx0, o.x0); cx1 = max(fx1, o.x1); cy0 = min(fy0, o.y0); cy1 = max(fy1, o.y1); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { StageRect n; n.x0 = fx0; n.y0 = fy0; n.x1 = fx1; n.y1 = fy1; rec[par ^ 1] = n; }
    const int w = cx1 - cx0, h = cy1 - cy0;
    if (w <= 0 || h <= 0) return;
    const uint32_t *s0 = src + (size_t)(cy0 - row_a) * src_pitch + cx0;
    uint32_t *d0 = dst + (size_t)(cy0 - row_a) * vw + cx0;
    if (((src_pitch | vw | cx0 | w) & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        // item = 128 uint4 (2 KB) of one row, one warp per item, four independent 128-bit copies per lane
        const int w4 = w >> 2, segs = (w4 + 127) >> 7, n_items = h * segs;
        const int lane = threadIdx.x & 31, n_warps = (int)(gridDim.x * blockDim.x) >> 5;
        for (int it = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); it < n_items; it += n_warps) {
            const int r = it / segs, c = ((it - r * segs) << 7) + lane;
            const uint4 *sp = reinterpret_cast<const uint4 *>(s0 + (size_t)r * src_pitch);
            uint4 *dp = reinterpret_cast<uint4 *>(d0 + (size_t)r * vw);
            uint4 v[4];
            #pragma unroll
            for (int k = 0; k < 4; k++) if (c + 32 * k < w4) v[k] = sp[c + 32 * k];
            #pragma unroll
            for (int k = 0; k < 4; k++) if (c + 32 * k < w4) dp[c + 32 * k] = v[k];
        }
    } else {
        const size_t n = (size_t)w * h;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const int r = (int)(i / w), c = (int)(i - (size_t)r * w);
            d0[(size_t)r * vw + c] = s0[(size_t)r * src_pitch + c];
        }
    }
}

// ----------------------------------------------------------------------------------------
// Frame protocol of the band-sharded single frame (FrameSync, common.cuh): the assembling GPU (rank 0) clears the other
// ranks' rows of its screen itself -- local HBM writes -- and announces the frame; the other ranks then store only the
// tiles they drew into (their last kernel's ordinary stores, travelling over NVLink) and raise a flag in rank 0's
// memory; rank 0's stream ends with a kernel that waits for the flags.  No collective, no host round trip, and the
// NVLink ingress of rank 0 carries the covered pixels instead of the whole frame.
// ----------------------------------------------------------------------------------------
static int g_num_sms = 0;                 // of the device the contexts run on (one device per process, or equal devices)
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
    pdl_trigger(1u);
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
    pdl_trigger(1u);
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
// per device, once (swegl_b200_create): k_dof needs more than the default 48 KB of dynamic shared memory
void configure_kernels()
{
    cudaFuncSetAttribute(k_dof, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DofSmem));
    cudaFuncSetAttribute(k_dof, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    g_num_sms = 0;
}
void preload_sync_kernels()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_sync_clear); cudaFuncGetAttributes(&a, k_sync_wait_ready);
    cudaFuncGetAttributes(&a, k_sync_signal); cudaFuncGetAttributes(&a, k_sync_wait_done);
}
static int num_sms();
void launch_stage_rect(uint32_t *dst, const uint32_t *src, int src_pitch, int vw, int row_a, int row_b, int grow, const Counters *counters,
                       StageRect *rec, int par, bool full, cudaStream_t st)
{
    launch_chain(k_stage_rect, num_sms() * 4, 256, st, false, dst, src, src_pitch, vw, row_a, row_b, grow, counters, rec, par, full ? 1 : 0);
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

// geometry shared by k_fragments, k_dof_classify and k_dof: `vp` = the rows that are drawn, [out0, out1) = the rows the post
// pass produces (the band without its 5-row halo)
static FragGeom make_geom(const ViewParams &vp, bool skip_bg, bool dof, int out_row0, int out_row1, int dof_pitch)
{
    FragGeom g{};
    g.vx = vp.vx; g.vy = vp.vy; g.vw = vp.vw; g.band0 = vp.band0; g.band1 = vp.band1; g.nbx = vp.nbx; g.ntx = vp.ntx;
    g.nty = (vp.band1 - vp.band0 + FRAG_ROWS - 1) / FRAG_ROWS;
    g.n_tiles = g.ntx * g.nty;
    g.skip_bg = skip_bg ? 1 : 0;
    g.dof = dof ? 1 : 0;
    g.ndx = (vp.vw + DOF_OW - 1) / DOF_OW;
    g.n_dof = dof ? g.ndx * ((vp.band1 - vp.band0 + DOF_OH - 1) / DOF_OH) : 0;
    g.out0 = out_row0; g.out1 = out_row1; g.dof_pitch = dof_pitch;
    g.k4b = 0x4B000000u;
    return g;
}

static int num_sms()
{
    if (!g_num_sms) {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        g_num_sms = n > 0 ? n : 148;
    }
    return g_num_sms;
}

template <int LIGHT, int TEX, int FAST>
static void launch_frag_t(const DeviceScene &s, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p, const FragGeom &g,
                          uint32_t *color, int color_pitch, float *depth, uint32_t *dof_dst, bool count_covered, Counters *h_counters_out, bool crowded, cudaStream_t st)
{
    if (g.n_tiles <= 0) return;
    // one wave of persistent CTAs (fewer when the viewport has fewer row items than that)
    const unsigned wave = (unsigned)(num_sms() * FRAG_CTAS_PER_SM);
    const unsigned need = ((unsigned)g.n_tiles * FRAG_ROWS + FRAG_ROWS - 1) / FRAG_ROWS;
    const unsigned grid = need < wave ? (need ? need : 1u) : wave;
    if (crowded) launch_chain(k_fragments<LIGHT, TEX, FAST, 1>, grid, FRAG_TPB, st, true, s, d_vp, d_fp, p, g, color, color_pitch, depth, dof_dst, count_covered ? 1 : 0, h_counters_out);
    else launch_chain(k_fragments<LIGHT, TEX, FAST, 0>, grid, FRAG_TPB, st, true, s, d_vp, d_fp, p, g, color, color_pitch, depth, dof_dst, count_covered ? 1 : 0, h_counters_out);
}

void launch_fragments(const DeviceScene &s, const ViewParams &vp, const ViewParams *d_vp, const FrameParams *d_fp, const Pools &p,
                      uint32_t *color, int color_pitch, float *depth, bool count_covered, Counters *h_counters_out, bool skip_bg_color,
                      bool fast, bool dof, int out_row0, int out_row1, uint32_t *dof_dst, int dof_pitch, bool crowded, cudaStream_t st)
{
    const FragGeom g = make_geom(vp, skip_bg_color, dof, out_row0, out_row1, dof_pitch);
#define SB_CASE(L, T) if (vp.light_mode == L && vp.tex_mode == T) { launch_frag_t<L, T, 0>(s, d_vp, d_fp, p, g, color, color_pitch, depth, dof_dst, count_covered, h_counters_out, crowded, st); return; }
#define SB_FAST(T) if (fast && vp.light_mode == SWEGL_B200_LIGHT_PHONG && vp.tex_mode == T) { launch_frag_t<SWEGL_B200_LIGHT_PHONG, T, 1>(s, d_vp, d_fp, p, g, color, color_pitch, depth, dof_dst, count_covered, h_counters_out, crowded, st); return; }
    SB_FAST(0) SB_FAST(1) SB_FAST(2)
    SB_CASE(0, 0) SB_CASE(0, 1) SB_CASE(0, 2)
    SB_CASE(1, 0) SB_CASE(1, 1) SB_CASE(1, 2)
    SB_CASE(2, 0) SB_CASE(2, 1) SB_CASE(2, 2)
#undef SB_CASE
#undef SB_FAST
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

// the DoF tile classification as a kernel of its own (behind k_fragments_layers; k_fragments does it itself)
void launch_dof_classify(const ViewParams &vp, const ViewParams *d_vp, const Pools &p, int out_row0, int out_row1, uint32_t *dof_dst, int dof_pitch, cudaStream_t st)
{
    const FragGeom g = make_geom(vp, false, true, out_row0, out_row1, dof_pitch);
    if (g.n_dof > 0) launch_chain(k_dof_classify, (unsigned)(g.n_dof + 7) / 8, 256, st, true, d_vp, p, g, dof_dst);
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
// self-test of the fast kernels' bilinear filter against the exact one (swegl_b200_selftest_filter): random texels, texture
// coordinates drawn over every binade below 2^21 with extra weight on integers, halves and their float neighbours
__global__ void k_selftest_filter(uint64_t n, uint32_t seed, unsigned long long *mismatches, uint32_t k4b)
{
    unsigned long long bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t h0 = mix32((uint32_t)i ^ seed), h1 = mix32(h0 + (uint32_t)(i >> 32) + 0x9e3779b9u), h2 = mix32(h1 ^ 0x85ebca6bu), h3 = mix32(h2 + 0xc2b2ae35u);
        TexFetch f;
        f.p00 = mix32(h3 ^ 1u); f.p10 = mix32(h3 ^ 2u); f.p01 = mix32(h3 ^ 3u); f.p11 = mix32(h3 ^ 4u);
        if (h3 & 1u) { f.p00 |= 0xFF000000u; f.p10 |= 0xFF000000u; f.p01 |= 0xFF000000u; f.p11 |= 0xFF000000u; }    // opaque texels: the alpha shortcut
        if ((h3 & 6u) == 2u) f.p00 = f.p10 = f.p01 = f.p11 = 0xFFFFFFFFu;                                           // saturated channels
        float c[2];
        for (int k = 0; k < 2; k++) {
            const uint32_t h = k ? h1 : h0, mode = (h2 >> (4 * k)) & 15u;
            const int e = (int)(h >> 8) % 23 - 2;                            // binade 2^-2 .. 2^20
            float t = ldexpf(1.0f + (float)(h & 0x7FFFFFu) * (1.0f / 8388608.0f), e);
            if (mode == 1u) t = floorf(t);
            if (mode == 2u) t = floorf(t) + 0.5f;
            if (mode == 3u) t = __uint_as_float(__float_as_uint(floorf(t) + 0.5f) - 1u);
            if (mode == 4u) t = __uint_as_float(__float_as_uint(floorf(t) + 0.5f) + 1u);
            if (mode == 5u) t = __uint_as_float(__float_as_uint(floorf(t)) - 1u);
            if (mode == 6u) t = (float)(h & 0xFFFu) * (1.0f / 4096.0f);      // [0, 1)
            if (mode == 7u) t = -t;
            if (mode == 8u) t = 0.0f;
            c[k] = t;
        }
        f.fu = c[0]; f.fv = c[1];
        if (!(fabsf(f.fu) < 2097152.0f && fabsf(f.fv) < 2097152.0f)) continue;
        bad += tex_filter<SWEGL_B200_TEX_BILINEAR>(f) != tex_filter_bilinear_fast(f, k4b) ? 1u : 0u;
        // the index arithmetic of the fetch: plain truncation == the guarded f2i inside the range
        bad += (__float2int_rz(fsub(f.fu, 0.5f)) != f2i(fsub(f.fu, 0.5f)) || __float2int_rz(fsub(f.fv, 0.5f)) != f2i(fsub(f.fv, 0.5f))) ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) bad += __shfl_down_sync(0xFFFFFFFFu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}
#ifdef FRAG_PROBE_TIMELINE
extern "C" int swegl_b200_probe_timeline(unsigned long long *out, int clear)
{
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, g_frag_timeline, sizeof(g_frag_timeline));
    if (clear) { static unsigned long long z[2048 * 16]; cudaMemcpyToSymbol(g_frag_timeline, z, sizeof(z)); }
    return 0;
}
#endif
void launch_selftest_filter(uint64_t n, uint32_t seed, unsigned long long *d_out, cudaStream_t st)
{
    k_selftest_filter<<<148 * 8, 256, 0, st>>>(n, seed, d_out, 0x4B000000u);
}

void launch_selftest_division(uint64_t n_pairs, uint32_t seed, unsigned long long *d_out2, cudaStream_t st)
{
    k_selftest_division<<<148 * 8, 256, 0, st>>>(n_pairs, seed, d_out2, d_out2 + 1);
}

// Tensor maps of the post pass's two sources, both dense [vh][vw] arrays of 32-bit words; box = the 80 x 41 window of one tile.
// cuTensorMapEncodeTiled comes from the driver through the runtime (no link against libcuda).
bool make_dof_tensor_maps(const uint32_t *src, const float *depth, int vw, int vh, CUtensorMap *tm_color, CUtensorMap *tm_depth)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<encode_fn>(fn);
    }
    if (!encode || (vw & 3) || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(depth)) & 15)) return false;
    const cuuint64_t dims[2] = { (cuuint64_t)vw, (cuuint64_t)vh };
    const cuuint64_t strides[1] = { (cuuint64_t)vw * 4 };                   // bytes between rows: a multiple of 16
    const cuuint32_t box[2] = { (cuuint32_t)DOF_WW, (cuuint32_t)DOF_SH };
    const cuuint32_t estr[2] = { 1, 1 };
    if (encode(tm_color, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t *>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    // (the depth words are moved as integers: a float tensor could not promise to keep every bit pattern)
    if (encode(tm_depth, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<float *>(depth), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    return true;
}

void launch_dof(const ViewParams &vp, const ViewParams *d_vp, const Pools &p, const CUtensorMap *tm_color, const CUtensorMap *tm_depth, bool use_tma,
                const uint32_t *src, int src_pitch, const float *depth, uint32_t *dst, int dst_pitch, int out_row0, int out_row1, cudaStream_t st)
{
    DofGeom g{};
    g.vw = vp.vw; g.vh = vp.vh; g.anchor = vp.band0 - vp.vy; g.out0 = out_row0; g.out1 = out_row1;
    g.ndx = (vp.vw + DOF_OW - 1) / DOF_OW;
    g.src_pitch = src_pitch; g.dst_pitch = dst_pitch; g.use_tma = use_tma ? 1 : 0;
    const int n_dof = g.ndx * ((vp.band1 - vp.band0 + DOF_OH - 1) / DOF_OH);
    if (n_dof <= 0) return;
    const unsigned wave = (unsigned)(num_sms() * DOF_CTAS_PER_SM);
    const unsigned grid = (unsigned)n_dof < wave ? (unsigned)n_dof : wave;
    CUtensorMap zero{};
    launch_chain_smem(k_dof, grid, DOF_THREADS, sizeof(DofSmem), st, true, use_tma ? *tm_color : zero, use_tma ? *tm_depth : zero, d_vp, p, g, src, depth, dst);
}

} // namespace sb
