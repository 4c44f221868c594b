// image_decode.cpp — PNG and JPEG decoders for the textures embedded in .glb files, without libpng / libjpeg.
//
// SURVEY §8f N4.  The reference decodes embedded images with libpng and libjpeg (src/misc/image.cpp:93-258,
// called from src/data/gltf.cpp:77-111) into row-major uint32 texels with bytes b,g,r,a, alpha 255 where the file
// has none.  Those libraries are not in this image; this file produces the SAME texels from the same bytes:
//
//   PNG   zlib inflate + the five scanline filters + the transformations read_png_file asks libpng for
//         (16 -> 8 bit strip, palette -> RGB, gray 1/2/4 -> 8, tRNS -> alpha, filler 0xFF, gray -> RGB), then the
//         r<->b swap of image.cpp:164-171.  Lossless, so "the same" is unambiguous.
//   JPEG  baseline / extended sequential / progressive Huffman, 8 bit, 1 or 3 components.  Lossy formats are only
//         "the same" up to the decoder's arithmetic, so the arithmetic is libjpeg's (which libjpeg-turbo, PIL's and
//         every distribution's decoder, reproduces bit for bit): the 13-bit fixed-point "islow" inverse DCT of
//         jidctint.c, "fancy" triangle-filter chroma upsampling of jdsample.c (h2v1, h2v2), and the 16-bit fixed-point
//         YCbCr -> RGB tables of jdcolor.c.  read_jpeg_file (image.cpp:181-252) uses libjpeg's defaults, which are
//         exactly these.
//
// tests/test_image_decode.py checks the decoded texels of the bundled models' images against the digests stored in
// assets/*.scenepack (made with PIL = libpng / libjpeg-turbo) and a set of small fixtures covering the other modes.
// Host code only; part of libswegl_b200.so so that one library serves the loader side too.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <zlib.h>

#include "swegl_b200.h"

namespace {

struct Err { const char *msg; };
[[noreturn]] void fail(const char *m) { throw Err{m}; }

uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
uint32_t be16(const uint8_t *p) { return (uint32_t)p[0] << 8 | p[1]; }
uint32_t pack_bgra(int r, int g, int b, int a) { return (uint32_t)b | (uint32_t)g << 8 | (uint32_t)r << 16 | (uint32_t)a << 24; }

// ======================================================================================== PNG
int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

void decode_png(const uint8_t *d, size_t n, std::vector<uint32_t> &out, int &w, int &h)
{
    size_t pos = 8;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool have_trns = false, have_ihdr = false;
    while (pos + 12 <= n) {
        const uint32_t len = be32(d + pos);
        const uint8_t *ty = d + pos + 4, *body = d + pos + 8;
        if (pos + 12 + (size_t)len > n) fail("png: truncated chunk");
        if (!memcmp(ty, "IHDR", 4)) {
            if (len < 13) fail("png: bad IHDR");
            w = (int)be32(body); h = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            have_ihdr = true;
        } else if (!memcmp(ty, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(ty, "tRNS", 4)) { trns.assign(body, body + len); have_trns = true; }
        else if (!memcmp(ty, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(ty, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w <= 0 || h <= 0 || w > 32768 || h > 32768) fail("png: no usable IHDR");
    if (interlace > 1) fail("png: unknown interlace method");
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) fail("png: bad colour type / bit depth");
    if ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) fail("png: bad bit depth for colour type");
    if (ctype == 3 && (depth == 16 || plte.empty())) fail("png: bad palette image");
    const int bpp_bits = channels * depth, bpp = (bpp_bits + 7) / 8;       // filter unit in bytes
    // the image as one pass, or as the seven reduced images of Adam7 (PNG spec 8.2; libpng's png_read_image de-interlaces
    // them transparently, src/misc/image.cpp:93-170): pass = every dx-th column from x0 of every dy-th row from y0
    struct Pass { int x0, y0, dx, dy, pw, ph; size_t stride, offset; };
    static const int A7[7][4] = { {0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2} };
    std::vector<Pass> passes;
    size_t raw_size = 0;
    for (int k = 0; k < (interlace ? 7 : 1); k++) {
        Pass ps;
        ps.x0 = interlace ? A7[k][0] : 0; ps.y0 = interlace ? A7[k][1] : 0; ps.dx = interlace ? A7[k][2] : 1; ps.dy = interlace ? A7[k][3] : 1;
        ps.pw = (w - ps.x0 + ps.dx - 1) / ps.dx; ps.ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        if (ps.pw <= 0 || ps.ph <= 0) continue;                             // an empty pass has no bytes at all, not even filter bytes
        ps.stride = ((size_t)ps.pw * bpp_bits + 7) / 8;
        ps.offset = raw_size;
        raw_size += (ps.stride + 1) * (size_t)ps.ph;
        passes.push_back(ps);
    }
    // deflate cannot expand by more than ~1032:1: an IHDR that promises more than the IDAT bytes can hold is refused before
    // anything is allocated (a few-byte file must not make the decoder allocate gigabytes)
    if (raw_size > (size_t)1040 * idat.size() + 65536) fail("png: image data too short for the declared size");
    std::vector<uint8_t> raw(raw_size);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) fail("png: inflate failed");
    out.resize((size_t)w * h);
    // tRNS for gray / RGB: one colour key (16-bit samples in the chunk), compared BEFORE 16 -> 8 stripping as libpng does
    int key[3] = { -1, -1, -1 };
    if (have_trns && ctype == 0 && trns.size() >= 2) key[0] = (int)be16(trns.data());
    if (have_trns && ctype == 2 && trns.size() >= 6) for (int c = 0; c < 3; c++) key[c] = (int)be16(trns.data() + 2 * c);
    for (const Pass &ps : passes) {
        const size_t stride = ps.stride;
        // unfilter in place (PNG spec 9.2), prior row = zeros for the first row of the pass
        std::vector<uint8_t> zero(stride, 0);
        for (int y = 0; y < ps.ph; y++) {
            uint8_t *row = raw.data() + ps.offset + (stride + 1) * (size_t)y + 1;
            const uint8_t *up = y ? row - (stride + 1) : zero.data();
            const int f = row[-1];
            for (size_t i = 0; i < stride; i++) {
                const int a = i >= (size_t)bpp ? row[i - bpp] : 0, b = up[i], c = i >= (size_t)bpp ? up[i - bpp] : 0;
                int add;
                switch (f) {
                    case 0: add = 0; break;
                    case 1: add = a; break;
                    case 2: add = b; break;
                    case 3: add = (a + b) >> 1; break;
                    case 4: add = paeth(a, b, c); break;
                    default: fail("png: bad filter type");
                }
                row[i] = (uint8_t)(row[i] + add);
            }
        }
        for (int y = 0; y < ps.ph; y++) {
            const uint8_t *row = raw.data() + ps.offset + (stride + 1) * (size_t)y + 1;
            for (int x = 0; x < ps.pw; x++) {
                int s[4] = { 0, 0, 0, 0 }, s8[4];
                for (int c = 0; c < channels; c++) {
                    if (depth == 8) s[c] = row[(size_t)x * channels + c];
                    else if (depth == 16) s[c] = (int)be16(row + 2 * ((size_t)x * channels + c));
                    else { const size_t bit = (size_t)x * depth; s[c] = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1); }
                }
                for (int c = 0; c < channels; c++)
                    s8[c] = depth == 16 ? s[c] >> 8 : depth == 8 ? s[c] : s[c] * (255 / ((1 << depth) - 1));   // strip_16 / expand 1,2,4
                int r, g, b, a = 255;
                if (ctype == 3) {
                    const size_t i = (size_t)s[0];
                    if (3 * i + 2 >= plte.size()) fail("png: palette index out of range");
                    r = plte[3 * i]; g = plte[3 * i + 1]; b = plte[3 * i + 2];
                    if (have_trns && i < trns.size()) a = trns[i];
                } else if (ctype == 0 || ctype == 4) {
                    r = g = b = s8[0];
                    if (ctype == 4) a = s8[1];
                    else if (have_trns && s[0] == key[0]) a = 0;
                } else {
                    r = s8[0]; g = s8[1]; b = s8[2];
                    if (ctype == 6) a = s8[3];
                    else if (have_trns && s[0] == key[0] && s[1] == key[1] && s[2] == key[2]) a = 0;
                }
                out[(size_t)(ps.y0 + y * ps.dy) * w + (ps.x0 + x * ps.dx)] = pack_bgra(r, g, b, a);
            }
        }
    }
}

// ======================================================================================== JPEG
const uint8_t ZIGZAG[64 + 16] = {                       // jpeg_natural_order (+16 guard entries as in jutils.c)
     0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
    63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63 };

struct Huff {
    bool present = false;
    uint8_t bits[17] = {0}, vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    void build()
    {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7FFFFFFF;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int wb = 0, hb = 0;                 // blocks that carry image samples: ceil(ceil(W*h/hmax)/8) x ...
    int wb_pad = 0, hb_pad = 0;         // padded to whole MCUs (interleaved scans address these)
    int dw = 0, dh = 0;                 // downsampled_width / height in samples
    std::vector<int16_t> coef;          // wb_pad * hb_pad * 64
    std::vector<uint8_t> plane;         // (wb_pad*8) x (hb_pad*8) samples after the inverse DCT
    int dc_tbl = 0, ac_tbl = 0, last_dc = 0;
};

struct BitReader {
    const uint8_t *p, *end;
    uint32_t acc = 0; int cnt = 0;
    int marker = 0;                     // a marker met inside the entropy-coded segment (decoding then feeds zeros)
    void fill()
    {
        while (cnt <= 24) {
            int byte = 0;
            if (!marker && p < end) {
                byte = *p++;
                if (byte == 0xFF) {
                    int nx = p < end ? *p : 0xD9;
                    while (nx == 0xFF && p + 1 < end) { p++; nx = *p; }                 // fill bytes
                    if (nx == 0) p++;                                                   // stuffed zero
                    else { marker = nx; p++; byte = 0; }
                }
            }
            acc |= (uint32_t)byte << (24 - cnt);
            cnt += 8;
        }
    }
    int get(int n)                      // n in 0..16
    {
        if (!n) return 0;
        if (cnt < n) fill();
        const int v = (int)(acc >> (32 - n));
        acc <<= n; cnt -= n;
        return v;
    }
    int bit() { return get(1); }
    void reset() { acc = 0; cnt = 0; }
};

int huff_decode(BitReader &br, const Huff &t)
{
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | br.bit();
        if (t.maxcode[l] >= 0 && code <= t.maxcode[l] && code >= t.mincode[l]) return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    fail("jpeg: bad Huffman code");
}
inline int extend(int r, int s) { return r < (1 << (s - 1)) ? r - (1 << s) + 1 : r; }       // HUFF_EXTEND

struct Jpeg {
    int W = 0, H = 0, ncomp = 0, hmax = 1, vmax = 1;
    bool progressive = false;
    uint16_t qt[4][64]; bool qt_ok[4] = { false, false, false, false };
    Huff dc[4], ac[4];
    Component comp[3];
    int restart_interval = 0;
    int adobe_transform = -1;
    bool jfif = false;
};

// ---- inverse DCT: jidctint.c jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2), output clamped to 0..255 ----
typedef int64_t jlong;               // libjpeg's JLONG is `long`: no overflow on corrupt coefficients either
constexpr jlong F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
                  F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
inline jlong descale(jlong x, int n) { return (x + ((jlong)1 << (n - 1))) >> n; }
inline uint8_t clamp_sample(jlong v) { v += 128; return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

void idct_islow(const int16_t *coef, const uint16_t *q, uint8_t *out, int out_stride)
{
    jlong ws[64];
    for (int c = 0; c < 8; c++) {
        const int16_t *in = coef + c;
        const uint16_t *qq = q + c;
        jlong *w = ws + c;
        if (!in[8] && !in[16] && !in[24] && !in[32] && !in[40] && !in[48] && !in[56]) {
            const jlong dcv = ((jlong)in[0] * qq[0]) * 4;              // << PASS1_BITS
            for (int r = 0; r < 8; r++) w[8 * r] = dcv;
            continue;
        }
        jlong z2 = (jlong)in[16] * qq[16], z3 = (jlong)in[48] * qq[48];
        jlong z1 = (z2 + z3) * F_0_541;
        jlong tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
        z2 = (jlong)in[0] * qq[0]; z3 = (jlong)in[32] * qq[32];
        jlong tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;          // << CONST_BITS
        const jlong tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = (jlong)in[56] * qq[56]; tmp1 = (jlong)in[40] * qq[40]; tmp2 = (jlong)in[24] * qq[24]; tmp3 = (jlong)in[8] * qq[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; jlong z4 = tmp1 + tmp3;
        const jlong z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
        z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        w[0] = descale(tmp10 + tmp3, 11); w[56] = descale(tmp10 - tmp3, 11);
        w[8] = descale(tmp11 + tmp2, 11); w[48] = descale(tmp11 - tmp2, 11);
        w[16] = descale(tmp12 + tmp1, 11); w[40] = descale(tmp12 - tmp1, 11);
        w[24] = descale(tmp13 + tmp0, 11); w[32] = descale(tmp13 - tmp0, 11);
    }
    for (int r = 0; r < 8; r++) {
        const jlong *w = ws + 8 * r;
        uint8_t *o = out + (size_t)r * out_stride;
        // (jidctint.c's all-zero-AC row shortcut gives the same value as the full computation)
        jlong z2 = w[2], z3 = w[6];
        jlong z1 = (z2 + z3) * F_0_541;
        jlong tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
        jlong tmp0 = (w[0] + w[4]) * 8192, tmp1 = (w[0] - w[4]) * 8192;
        const jlong tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; jlong z4 = tmp1 + tmp3;
        const jlong z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
        z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        o[0] = clamp_sample(descale(tmp10 + tmp3, 18)); o[7] = clamp_sample(descale(tmp10 - tmp3, 18));
        o[1] = clamp_sample(descale(tmp11 + tmp2, 18)); o[6] = clamp_sample(descale(tmp11 - tmp2, 18));
        o[2] = clamp_sample(descale(tmp12 + tmp1, 18)); o[5] = clamp_sample(descale(tmp12 - tmp1, 18));
        o[3] = clamp_sample(descale(tmp13 + tmp0, 18)); o[4] = clamp_sample(descale(tmp13 - tmp0, 18));
    }
}

// ---- one scan ----
struct Scan { int n = 0, ci[3] = {0, 0, 0}, Ss = 0, Se = 63, Ah = 0, Al = 0; };

void decode_block_sequential(BitReader &br, Jpeg &j, Component &c, int16_t *blk)
{
    int s = huff_decode(br, j.dc[c.dc_tbl]);
    if (s > 16) fail("jpeg: bad DC magnitude category");
    if (s) { const int r = br.get(s); s = extend(r, s); }
    c.last_dc += s;
    blk[0] = (int16_t)c.last_dc;
    const Huff &at = j.ac[c.ac_tbl];
    for (int k = 1; k < 64; k++) {
        const int rs = huff_decode(br, at), r = rs >> 4, sz = rs & 15;
        if (sz) {
            k += r;
            const int v = extend(br.get(sz), sz);
            blk[ZIGZAG[k]] = (int16_t)v;
        } else {
            if (r != 15) break;
            k += 15;
        }
    }
}

void decode_scan(Jpeg &j, const Scan &sc, BitReader &br)
{
    for (int i = 0; i < sc.n; i++) j.comp[sc.ci[i]].last_dc = 0;
    uint32_t eobrun = 0;
    const bool interleaved = sc.n > 1;
    int mcus_x, mcus_y;
    if (interleaved) { mcus_x = (j.W + 8 * j.hmax - 1) / (8 * j.hmax); mcus_y = (j.H + 8 * j.vmax - 1) / (8 * j.vmax); }
    else { mcus_x = j.comp[sc.ci[0]].wb; mcus_y = j.comp[sc.ci[0]].hb; }
    int todo = j.restart_interval;
    const int p1 = 1 << sc.Al, m1 = -(1 << sc.Al);
    for (int my = 0; my < mcus_y; my++)
        for (int mx = 0; mx < mcus_x; mx++) {
            if (j.restart_interval && todo == 0) {
                // byte-align, expect RSTn
                br.reset();
                if (!br.marker) {                       // the marker has not been met yet: scan forward to it
                    while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] >= 0xD0 && br.p[1] <= 0xD7)) br.p++;
                    if (br.p + 1 < br.end) br.p += 2;
                } else if (br.marker < 0xD0 || br.marker > 0xD7) fail("jpeg: restart marker expected");
                br.marker = 0;
                for (int i = 0; i < sc.n; i++) j.comp[sc.ci[i]].last_dc = 0;
                eobrun = 0;
                todo = j.restart_interval;
            }
            for (int i = 0; i < sc.n; i++) {
                Component &c = j.comp[sc.ci[i]];
                const int bw = interleaved ? c.h : 1, bh = interleaved ? c.v : 1;
                for (int by = 0; by < bh; by++)
                    for (int bx = 0; bx < bw; bx++) {
                        const int X = interleaved ? mx * c.h + bx : mx, Y = interleaved ? my * c.v + by : my;
                        int16_t *blk = c.coef.data() + ((size_t)Y * c.wb_pad + X) * 64;
                        if (!j.progressive) { decode_block_sequential(br, j, c, blk); continue; }
                        if (sc.Ss == 0) {
                            if (sc.Ah == 0) {                                           // DC first (jdphuff.c decode_mcu_DC_first)
                                int s = huff_decode(br, j.dc[c.dc_tbl]);
                                if (s > 16) fail("jpeg: bad DC magnitude category");
    if (s) { const int r = br.get(s); s = extend(r, s); }
                                c.last_dc += s;
                                blk[0] = (int16_t)(c.last_dc * (1 << sc.Al));
                            } else if (br.bit()) blk[0] |= (int16_t)p1;                 // DC refine
                            continue;
                        }
                        const Huff &at = j.ac[c.ac_tbl];
                        if (sc.Ah == 0) {                                               // AC first
                            if (eobrun > 0) { eobrun--; continue; }
                            for (int k = sc.Ss; k <= sc.Se; k++) {
                                const int rs = huff_decode(br, at), r = rs >> 4, sz = rs & 15;
                                if (sz) {
                                    k += r;
                                    const int v = extend(br.get(sz), sz);
                                    blk[ZIGZAG[k]] = (int16_t)(v * (1 << sc.Al));
                                } else if (r == 15) k += 15;
                                else {
                                    eobrun = 1u << r;
                                    if (r) eobrun += (uint32_t)br.get(r);
                                    eobrun--;
                                    break;
                                }
                            }
                            continue;
                        }
                        // AC refine (jdphuff.c decode_mcu_AC_refine)
                        int k = sc.Ss;
                        if (eobrun == 0) {
                            for (; k <= sc.Se; k++) {
                                const int rs = huff_decode(br, at);
                                int r = rs >> 4, s = rs & 15;
                                if (s) { s = br.bit() ? p1 : m1; }
                                else if (r != 15) {
                                    eobrun = 1u << r;
                                    if (r) eobrun += (uint32_t)br.get(r);
                                    break;
                                }
                                do {
                                    int16_t *cf = blk + ZIGZAG[k];
                                    if (*cf != 0) {
                                        if (br.bit() && (*cf & p1) == 0) *cf = (int16_t)(*cf >= 0 ? *cf + p1 : *cf + m1);
                                    } else if (--r < 0) break;
                                    k++;
                                } while (k <= sc.Se);
                                if (s) blk[ZIGZAG[k]] = (int16_t)s;
                            }
                        }
                        if (eobrun > 0) {
                            for (; k <= sc.Se; k++) {
                                int16_t *cf = blk + ZIGZAG[k];
                                if (*cf != 0 && br.bit() && (*cf & p1) == 0) *cf = (int16_t)(*cf >= 0 ? *cf + p1 : *cf + m1);
                            }
                            eobrun--;
                        }
                    }
            }
            if (j.restart_interval) todo--;
        }
}

// ---- upsampling (jdsample.c, do_fancy_upsampling) of one component to full resolution ----
void upsample(const Jpeg &j, const Component &c, std::vector<uint8_t> &full, int FW, int FH)
{
    const int ps = c.wb_pad * 8;
    full.assign((size_t)FW * FH, 0);
    const int hx = j.hmax / c.h, vx = j.vmax / c.v;
    auto in = [&](int y, int x) -> int { return c.plane[(size_t)y * ps + x]; };
    if (hx == 1 && vx == 1) {
        for (int y = 0; y < j.H; y++) memcpy(&full[(size_t)y * FW], &c.plane[(size_t)y * ps], (size_t)j.W);
        return;
    }
    const int dw = c.dw, dh = c.dh;
    if (hx == 2 && vx == 1 && j.hmax % c.h == 0) {                         // h2v1_fancy_upsample
        for (int y = 0; y < dh && y < FH; y++) {
            uint8_t *o = &full[(size_t)y * FW];
            if (dw == 1) { o[0] = (uint8_t)in(y, 0); if (FW > 1) o[1] = (uint8_t)in(y, 0); continue; }
            for (int x = 0; x < dw; x++) {
                const int v = in(y, x);
                const int a = x == 0 ? v : (3 * v + in(y, x - 1) + 1) >> 2;
                const int b = x == dw - 1 ? v : (3 * v + in(y, x + 1) + 2) >> 2;
                if (2 * x < FW) o[2 * x] = (uint8_t)a;
                if (2 * x + 1 < FW) o[2 * x + 1] = (uint8_t)b;
            }
        }
        return;
    }
    if (hx == 2 && vx == 2 && j.hmax % c.h == 0 && j.vmax % c.v == 0) {     // h2v2_fancy_upsample
        for (int y = 0; y < dh; y++)
            for (int v = 0; v < 2; v++) {
                const int oy = 2 * y + v;
                if (oy >= FH) continue;
                const int y1 = v == 0 ? (y > 0 ? y - 1 : 0) : (y < dh - 1 ? y + 1 : dh - 1);     // edge rows are duplicated (jdmainct.c)
                uint8_t *o = &full[(size_t)oy * FW];
                auto colsum = [&](int x) { return 3 * in(y, x) + in(y1, x); };
                if (dw == 1) { const int t = colsum(0); o[0] = (uint8_t)((t * 4 + 8) >> 4); if (FW > 1) o[1] = (uint8_t)((t * 4 + 7) >> 4); continue; }
                for (int x = 0; x < dw; x++) {
                    const int t = colsum(x);
                    const int a = x == 0 ? (t * 4 + 8) >> 4 : (t * 3 + colsum(x - 1) + 8) >> 4;
                    const int b = x == dw - 1 ? (t * 4 + 7) >> 4 : (t * 3 + colsum(x + 1) + 7) >> 4;
                    if (2 * x < FW) o[2 * x] = (uint8_t)a;
                    if (2 * x + 1 < FW) o[2 * x + 1] = (uint8_t)b;
                }
            }
        return;
    }
    fail("jpeg: unsupported chroma subsampling (supported: 1x1, 2x1, 2x2)");
}

void decode_jpeg(const uint8_t *d, size_t n, std::vector<uint32_t> &out, int &w, int &h)
{
    Jpeg j;
    size_t pos = 2;
    bool have_sof = false, done = false;
    while (!done && pos + 4 <= n) {
        if (d[pos] != 0xFF) { pos++; continue; }
        const int m = d[pos + 1];
        if (m == 0xFF) { pos++; continue; }
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) { pos += 2; continue; }
        if (m == 0xD9) break;
        const size_t len = be16(d + pos + 2);
        if (len < 2 || pos + 2 + len > n) fail("jpeg: truncated segment");
        const uint8_t *b = d + pos + 4;
        const size_t bl = len - 2;
        if (m == 0xDB) {                                                    // DQT
            size_t i = 0;
            while (i < bl) {
                const int pq = b[i] >> 4, tq = b[i] & 15; i++;
                if (tq > 3 || i + (pq ? 128 : 64) > bl) fail("jpeg: bad DQT");
                for (int k = 0; k < 64; k++) { j.qt[tq][ZIGZAG[k]] = (uint16_t)(pq ? be16(b + i + 2 * k) : b[i + k]); }
                i += pq ? 128 : 64;
                j.qt_ok[tq] = true;
            }
        } else if (m == 0xC4) {                                             // DHT
            size_t i = 0;
            while (i + 17 <= bl) {
                const int tc = b[i] >> 4, th = b[i] & 15; i++;
                if (th > 3 || tc > 1) fail("jpeg: bad DHT");
                Huff &t = tc ? j.ac[th] : j.dc[th];
                int total = 0;
                t.bits[0] = 0;
                for (int l = 1; l <= 16; l++) { t.bits[l] = b[i + l - 1]; total += t.bits[l]; }
                i += 16;
                if (total > 256 || i + (size_t)total > bl) fail("jpeg: bad DHT");
                memcpy(t.vals, b + i, (size_t)total);
                i += (size_t)total;
                t.present = true;
                t.build();
            }
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {                   // SOF0/1/2
            if (bl < 6 || b[0] != 8) fail("jpeg: only 8-bit precision");
            j.progressive = m == 0xC2;
            j.H = (int)be16(b + 1); j.W = (int)be16(b + 3); j.ncomp = b[5];
            if (j.W <= 0 || j.H <= 0 || (j.ncomp != 1 && j.ncomp != 3) || bl < 6 + 3 * (size_t)j.ncomp) fail("jpeg: unsupported frame header");
            for (int i = 0; i < j.ncomp; i++) {
                Component &c = j.comp[i];
                c.id = b[6 + 3 * i]; c.h = b[7 + 3 * i] >> 4; c.v = b[7 + 3 * i] & 15; c.tq = b[8 + 3 * i];
                if (c.h < 1 || c.h > 2 || c.v < 1 || c.v > 2 || c.tq > 3) fail("jpeg: unsupported sampling factors");
                if (c.h > j.hmax) j.hmax = c.h;
                if (c.v > j.vmax) j.vmax = c.v;
            }
            if (j.ncomp == 1) { j.comp[0].h = j.comp[0].v = 1; j.hmax = j.vmax = 1; }      // a single component is never subsampled
            const int mx = (j.W + 8 * j.hmax - 1) / (8 * j.hmax), my = (j.H + 8 * j.vmax - 1) / (8 * j.vmax);
            for (int i = 0; i < j.ncomp; i++) {
                Component &c = j.comp[i];
                c.dw = (j.W * c.h + j.hmax - 1) / j.hmax; c.dh = (j.H * c.v + j.vmax - 1) / j.vmax;
                c.wb = (c.dw + 7) / 8; c.hb = (c.dh + 7) / 8;
                c.wb_pad = mx * c.h; c.hb_pad = my * c.v;
                c.coef.assign((size_t)c.wb_pad * c.hb_pad * 64, 0);
            }
            have_sof = true;
        } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            fail("jpeg: lossless / hierarchical / arithmetic-coded files are not supported");
        } else if (m == 0xDD) {
            if (bl >= 2) j.restart_interval = (int)be16(b);
        } else if (m == 0xE0) {
            if (bl >= 5 && !memcmp(b, "JFIF", 5)) j.jfif = true;
        } else if (m == 0xEE) {
            if (bl >= 12 && !memcmp(b, "Adobe", 5)) j.adobe_transform = b[11];
        } else if (m == 0xDA) {                                             // SOS + entropy-coded data
            if (!have_sof) fail("jpeg: scan before frame header");
            Scan sc;
            sc.n = b[0];
            if (sc.n < 1 || sc.n > j.ncomp || bl < 1 + 2 * (size_t)sc.n + 3) fail("jpeg: bad scan header");
            for (int i = 0; i < sc.n; i++) {
                int ci = -1;
                for (int k = 0; k < j.ncomp; k++) if (j.comp[k].id == b[1 + 2 * i]) ci = k;
                if (ci < 0) fail("jpeg: scan names an unknown component");
                sc.ci[i] = ci;
                j.comp[ci].dc_tbl = b[2 + 2 * i] >> 4; j.comp[ci].ac_tbl = b[2 + 2 * i] & 15;
                if (j.comp[ci].dc_tbl > 3 || j.comp[ci].ac_tbl > 3) fail("jpeg: bad table selector");
            }
            sc.Ss = b[1 + 2 * sc.n]; sc.Se = b[2 + 2 * sc.n]; sc.Ah = b[3 + 2 * sc.n] >> 4; sc.Al = b[3 + 2 * sc.n] & 15;
            if (!j.progressive) { sc.Ss = 0; sc.Se = 63; sc.Ah = sc.Al = 0; }
            else if (sc.Ss > sc.Se || sc.Se > 63 || (sc.Ss == 0 && sc.Se != 0) || (sc.Ss > 0 && sc.n != 1) || sc.Al > 13) fail("jpeg: bad progressive scan parameters");
            for (int i = 0; i < sc.n; i++) {
                const Component &c = j.comp[sc.ci[i]];
                const bool need_dc = !j.progressive || (sc.Ss == 0 && sc.Ah == 0), need_ac = !j.progressive || sc.Ss > 0;
                if ((need_dc && !j.dc[c.dc_tbl].present) || (need_ac && !j.ac[c.ac_tbl].present)) fail("jpeg: missing Huffman table");
            }
            BitReader br{ d + pos + 2 + len, d + n };
            decode_scan(j, sc, br);
            // continue after the entropy-coded segment: at the marker the reader stopped at, else search for one
            size_t q = (size_t)(br.p - d);
            if (br.marker) q -= 2;
            else while (q + 1 < n && !(d[q] == 0xFF && d[q + 1] != 0 && !(d[q + 1] >= 0xD0 && d[q + 1] <= 0xD7) && d[q + 1] != 0xFF)) q++;
            pos = q;
            continue;
        }
        pos += 2 + len;
    }
    if (!have_sof) fail("jpeg: no frame header");
    w = j.W; h = j.H;
    for (int i = 0; i < j.ncomp; i++) {
        Component &c = j.comp[i];
        if (!j.qt_ok[c.tq]) fail("jpeg: missing quantisation table");
        const int ps = c.wb_pad * 8;
        c.plane.assign((size_t)ps * c.hb_pad * 8, 0);
        for (int by = 0; by < c.hb_pad; by++)
            for (int bx = 0; bx < c.wb_pad; bx++)
                idct_islow(c.coef.data() + ((size_t)by * c.wb_pad + bx) * 64, j.qt[c.tq], &c.plane[(size_t)by * 8 * ps + bx * 8], ps);
    }
    out.resize((size_t)w * h);
    if (j.ncomp == 1) {
        const int ps = j.comp[0].wb_pad * 8;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) { const int g = j.comp[0].plane[(size_t)y * ps + x]; out[(size_t)y * w + x] = pack_bgra(g, g, g, 255); }
        return;
    }
    std::vector<uint8_t> full[3];
    for (int i = 0; i < 3; i++) upsample(j, j.comp[i], full[i], w, h);
    // colour space as jdapimin.c default_decompress_parms guesses it: JFIF -> YCbCr; Adobe transform 0 -> RGB; else by ids
    bool ycc = true;
    if (!j.jfif && j.adobe_transform == 0) ycc = false;
    else if (!j.jfif && j.adobe_transform < 0 && j.comp[0].id == 'R' && j.comp[1].id == 'G' && j.comp[2].id == 'B') ycc = false;
    // jdcolor.c build_ycc_rgb_table / ycc_rgb_convert, SCALEBITS 16
    int cr_r[256], cb_b[256]; int32_t cr_g[256], cb_g[256];
    for (int i = 0; i < 256; i++) {
        const int32_t x = i - 128;
        cr_r[i] = (int)((91881 * x + 32768) >> 16);
        cb_b[i] = (int)((116130 * x + 32768) >> 16);
        cr_g[i] = -46802 * x;
        cb_g[i] = -22554 * x + 32768;
    }
    auto lim = [](int v) { return v < 0 ? 0 : v > 255 ? 255 : v; };
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const int y = full[0][i], cb = full[1][i], cr = full[2][i];
        if (!ycc) { out[i] = pack_bgra(y, cb, cr, 255); continue; }
        const int r = lim(y + cr_r[cr]), g = lim(y + (int)((cb_g[cb] + cr_g[cr]) >> 16)), b = lim(y + cb_b[cb]);
        out[i] = pack_bgra(r, g, b, 255);
    }
}

thread_local char g_image_error[160] = "";

} // namespace

extern "C" int swegl_b200_decode_image(const void *data, size_t size, uint32_t **texels_bgra, int32_t *width, int32_t *height)
{
    g_image_error[0] = 0;
    if (!data || !texels_bgra || !width || !height) { strncpy(g_image_error, "decode_image: null argument", sizeof g_image_error - 1); return SWEGL_B200_ERR_ARG; }
    const uint8_t *d = static_cast<const uint8_t *>(data);
    std::vector<uint32_t> out;
    int w = 0, h = 0;
    try {
        static const uint8_t png_sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n' };
        if (size >= 8 && !memcmp(d, png_sig, 8)) decode_png(d, size, out, w, h);
        else if (size >= 4 && d[0] == 0xFF && d[1] == 0xD8) decode_jpeg(d, size, out, w, h);
        else fail("decode_image: neither PNG nor JPEG");
    } catch (const Err &e) {
        strncpy(g_image_error, e.msg, sizeof g_image_error - 1);
        return SWEGL_B200_ERR_UNSUPPORTED;
    } catch (const std::bad_alloc &) {
        strncpy(g_image_error, "decode_image: out of memory", sizeof g_image_error - 1);
        return SWEGL_B200_ERR_UNSUPPORTED;
    }
    uint32_t *p = static_cast<uint32_t *>(malloc(out.size() * 4));
    if (!p) return SWEGL_B200_ERR_UNSUPPORTED;
    memcpy(p, out.data(), out.size() * 4);
    *texels_bgra = p; *width = w; *height = h;
    return SWEGL_B200_OK;
}

extern "C" void swegl_b200_image_free(uint32_t *texels_bgra) { free(texels_bgra); }
extern "C" const char *swegl_b200_image_error(void) { return g_image_error; }
