/*
 * swegl_oracle.c — TEST INFRASTRUCTURE ONLY (see swegl_oracle.h).
 *
 * Serial CPU restatement of swegl's frame: vertex stage, cull/mark, near clip,
 * triangle setup, scanline rasteriser with z test, pixel shaders, transparency
 * layers, flatten and post pass.  Every function cites the reference file:line it
 * follows (paths relative to the swegl checkout).  Arithmetic is plain IEEE fp32
 * in the reference's evaluation order; build with -ffp-contract=off (oracle/Makefile)
 * exactly as the reference is built without FMA (Makefile:16-19).
 */
#include "swegl_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;

/* ---- float -> int as the reference's x86-64 build does it (cvttss2si: out of range
 *      or NaN gives INT_MIN) ---- */
static int f2i(float f)
{
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT_MIN;
    return (int)f;
}
static int d2i(double d)
{
    if (!(d > -2147483649.0 && d < 2147483648.0)) return INT_MIN;
    return (int)d;
}
static int ceil_i(float f) { return d2i(ceil((double)f)); }   /* (int) ceil(v->y()) renderer.cpp:390 */

/* ---- swegl/projection/points.hpp, src/projection/points.cpp ---- */
static v3 v3_add(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static v3 v3_sub(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static v3 v3_mul(v3 a, float s) { v3 r = { a.x * s, a.y * s, a.z * s }; return r; }
static v3 v3_neg(v3 a) { v3 r = { -a.x, -a.y, -a.z }; return r; }
static float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       /* points.hpp:91-94 */
static float v3_len2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }            /* points.hpp:75-78 */
static v3 v3_normalize(v3 a)                                                        /* points.hpp:71-90 */
{
    float l = (float)sqrt((double)(a.x * a.x + a.y * a.y + a.z * a.z));
    if (l != 0) { a.x /= l; a.y /= l; a.z /= l; }
    return a;
}
static v3 xform(const float *m, v3 v)                                               /* points.cpp:8-13 */
{
    v3 r;
    r.x = m[0] * v.x + m[1] * v.y + m[2]  * v.z + m[3];
    r.y = m[4] * v.x + m[5] * v.y + m[6]  * v.z + m[7];
    r.z = m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11];
    return r;
}
static v3 rotate3(const float *m9, v3 v)              /* rotate(), points.cpp:44-49: normal_t ctor normalises */
{
    v3 r;
    r.x = m9[0] * v.x + m9[1] * v.y + m9[2] * v.z;
    r.y = m9[3] * v.x + m9[4] * v.y + m9[5] * v.z;
    r.z = m9[6] * v.x + m9[7] * v.y + m9[8] * v.z;
    return v3_normalize(r);
}
static v3 cross_n(v3 l, v3 r)                                   /* cross(), points.cpp:32-37 (normalised) */
{
    v3 c = { l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x };
    return v3_normalize(c);
}

/* ---- one mesh_vertex_t (swegl/data/model.hpp:20-29) ---- */
typedef struct {
    v3 v, v_world, v_viewport;
    v2 tex;
    v3 normal, normal_world;
    int yes;
} overt;

/* ---- interpolator_g<1> (swegl/render/interpolator.hpp:52-104) ---- */
typedef struct { float top, topstep, ualpha, bottom, bottomstep, v0, v1; } interp_t;

static void interp_init_self(interp_t *q, float dist, float z1, float z2)           /* :77-90 */
{
    q->v0 = z1;
    q->v1 = z2 - z1;
    float alphastep = 1.0f / dist;
    q->bottom = 1.0f / z1;
    float invz2 = 1.0f / z2;
    q->ualpha = 0.0f;
    q->top = q->ualpha;
    q->top *= q->bottom;
    q->topstep = (1.0f * invz2 - q->top) * alphastep;
    q->bottomstep = (invz2 - q->bottom) * alphastep;
}
static void interp_displace(interp_t *q, float move)                               /* :91-95 */
{
    q->top += q->topstep * move;
    q->bottom += q->bottomstep * move;
    q->ualpha = q->top / q->bottom;
}
static void interp_step(interp_t *q)                                               /* :96-100 */
{
    q->top += q->topstep;
    q->bottom += q->bottomstep;
    q->ualpha = q->top / q->bottom;
}
static float interp_value(const interp_t *q) { return q->v0 + q->v1 * q->ualpha; }  /* :103 */

typedef struct { interp_t ip; float ratio, x; } line_side;                          /* renderer.cpp:21-26 */

/* ---- colours (swegl/render/colors.hpp, src/render/colors.cpp) ---- */
typedef union { uint32_t i; struct { uint8_t b, g, r, a; } o; } pc_t;

static uint32_t blend(uint32_t back_i, uint32_t front_i)                            /* colors.cpp:39-57 */
{
    pc_t back, front, out;
    back.i = back_i; front.i = front_i;
    int alpha = front.o.a;
    int new_alpha = 255 - ((255 - alpha) * (255 - back.o.a) / 255);
    if (alpha == 255) { out.o.b = front.o.b; out.o.g = front.o.g; out.o.r = front.o.r; out.o.a = (uint8_t)new_alpha; return out.i; }
    if (alpha == 0)   { out.o.b = back.o.b;  out.o.g = back.o.g;  out.o.r = back.o.r;  out.o.a = (uint8_t)new_alpha; return out.i; }
    out.o.b = (uint8_t)d2i(back.o.b * ((256 - alpha) / 256.0) + front.o.b * (alpha / 256.0));
    out.o.g = (uint8_t)d2i(back.o.g * ((256 - alpha) / 256.0) + front.o.g * (alpha / 256.0));
    out.o.r = (uint8_t)d2i(back.o.r * ((256 - alpha) / 256.0) + front.o.r * (alpha / 256.0));
    out.o.a = (uint8_t)new_alpha;
    return out.i;
}

/* ---- the render state ---- */
typedef struct {
    const swegl_b200_scene_desc *scene;
    const swegl_b200_frame_desc *frame;
    const swegl_b200_viewport_desc *vp;
    uint32_t *pixels; int pitch_words;
    float *zbuffer;
    int got_transparency, n_layers;
    uint32_t **layer_colors; float **layer_z;
    int band_y0, band_y1;           /* absolute rows that may be written */
    orc_dump *dump;

    /* pixel shader state: pixel_shaders.hpp:15-118 */
    const swegl_b200_primitive *prim;
    uint32_t color;                 /* pixel_shader_t::color */
    const overt *tv[3];             /* the 3 y-sorted vertices of the current triangle */
    /* lights_flat */
    float flat_light;
    /* lights_phong */
    v3 v0, v1, v2, vleft, vright, vleftdir, vrightdir, v, vdir;
    v3 n0, n1, n2, nleft, nright, nleftdir, nrightdir, n, ndir;
    /* texture / texture_bilinear */
    v2 t0, t1, t2, side_long_t_dir, side_short_t, side_short_t_dir, t_left, t_dir;
    int long_line_on_right;
    uint32_t default_bitmap;
    const uint32_t *tbitmap; int twidth, theight;
} rs_t;

static float max_z(void) { union { uint32_t i; float f; } u; u.i = 0x7F7F7F7Fu; return u.f; } /* renderer.cpp:15-19 */

/* ---- pixel shaders ---- */

static void ps_prepare_for_primitive(rs_t *s, const swegl_b200_primitive *p)
{
    /* pixel_shader_t::prepare_for_primitive, pixel_shaders.cpp:14-31 */
    const swegl_b200_material *m = (p->material_id != -1) ? &s->scene->materials[p->material_id]
                                                          : &s->scene->default_material;
    pc_t c; c.o.b = m->b; c.o.g = m->g; c.o.r = m->r; c.o.a = m->a;
    s->prim = p;
    s->color = c.i;
    /* pixel_shader_texture(_bilinear)::prepare_for_primitive, pixel_shaders.cpp:209-229, 284-302 */
    if (p->material_id == -1 || m->texture_idx == -1) {
        s->default_bitmap = c.i;       /* deviation: materials[-1] is UB in the reference */
        s->tbitmap = &s->default_bitmap;
        s->twidth = 1; s->theight = 1;
    } else {
        const swegl_b200_texture *t = &s->scene->textures[m->texture_idx];
        s->tbitmap = t->texels; s->twidth = t->width; s->theight = t->height;
    }
}

/* shared by lights_flat::prepare_for_triangle (pixel_shaders.cpp:33-84) and
 * lights_phong::shade (:159-205): the point-light sum */
static float point_lights_sum(const rs_t *s, v3 center, v3 normal, v3 camera_vector)
{
    float dyn = 0.0f;
    for (uint32_t i = 0; i < s->frame->n_point_lights; i++) {
        const float *L = &s->frame->point_lights[4 * i];
        v3 lpos = { L[0], L[1], L[2] };
        v3 ld = v3_sub(center, lpos);
        float d2 = v3_len2(ld);
        float diffuse = L[3] / d2;
        if ((double)diffuse < 0.05) continue;
        ld = v3_normalize(ld);
        float alignment = -v3_dot(normal, ld);
        if (alignment < 0.0f) continue;
        diffuse *= alignment;
        v3 refl = v3_add(ld, v3_mul(normal, alignment * 2));
        float specular = v3_dot(refl, camera_vector);
        if (specular > 0) {
            specular = (float)pow((double)specular, 32.0);
            specular = specular * 32 / 2;
            dyn += diffuse + specular / d2;
        } else {
            dyn += diffuse;
        }
    }
    return dyn;
}

static void ps_prepare_for_triangle(rs_t *s, const overt *a, const overt *b, const overt *c, int inverted)
{
    const int lm = s->vp->light_mode, tm = s->vp->tex_mode;
    s->tv[0] = a; s->tv[1] = b; s->tv[2] = c;
    if (lm == SWEGL_B200_LIGHT_FLAT) {
        /* pixel_shader_lights_flat::prepare_for_triangle, pixel_shaders.cpp:33-84 */
        v3 nw = cross_n(v3_sub(b->v_world, a->v_world), v3_sub(c->v_world, a->v_world));
        if (inverted) nw = v3_normalize(v3_neg(nw));     /* operator-(normal_t) re-normalises, points.hpp:171-174 */
        float sun = -v3_dot(nw, *(const v3 *)s->frame->sun_dir);
        if (sun < 0.0f) sun = 0.0f; else sun *= s->frame->sun_intensity;
        v3 center = v3_add(v3_add(a->v_world, b->v_world), c->v_world);
        center.x = center.x / 3; center.y = center.y / 3; center.z = center.z / 3;
        v3 cam = { s->vp->cam_pos[0], s->vp->cam_pos[1], s->vp->cam_pos[2] };
        v3 camv = v3_normalize(v3_sub(cam, center));
        float dyn = point_lights_sum(s, center, nw, camv);
        float light = s->frame->ambient + sun + dyn;
        light *= 65536;
        s->flat_light = light;
    } else if (lm == SWEGL_B200_LIGHT_PHONG) {
        /* pixel_shader_lights_phong::prepare_for_triangle, pixel_shaders.cpp:88-105 */
        s->v0 = a->v_world; s->v1 = b->v_world; s->v2 = c->v_world;
        if (!inverted) { s->n0 = a->normal_world; s->n1 = b->normal_world; s->n2 = c->normal_world; }
        else { s->n0 = v3_neg(a->normal_world); s->n1 = v3_neg(b->normal_world); s->n2 = v3_neg(c->normal_world); }
    }
    if (tm != SWEGL_B200_TEX_PLAIN) {
        /* pixel_shader_texture(_bilinear)::prepare_for_triangle, pixel_shaders.cpp:231-245, 304-318 */
        s->t0 = a->tex; s->t1 = b->tex; s->t2 = c->tex;
        s->t0.x *= s->twidth; s->t0.y *= s->theight;
        s->t1.x *= s->twidth; s->t1.y *= s->theight;
        s->t2.x *= s->twidth; s->t2.y *= s->theight;
        s->side_long_t_dir.x = s->t2.x - s->t0.x; s->side_long_t_dir.y = s->t2.y - s->t0.y;
    }
}

static void ps_prepare_for_half(rs_t *s, int lower, int lor)
{
    const int lm = s->vp->light_mode, tm = s->vp->tex_mode;
    if (lm == SWEGL_B200_LIGHT_PHONG) {
        if (!lower) {
            /* prepare_for_upper_triangle, pixel_shaders.cpp:106-126 */
            s->vleft = s->v0; s->vright = s->v0;
            if (lor) { s->vleftdir = v3_sub(s->v1, s->v0); s->vrightdir = v3_sub(s->v2, s->v0); }
            else     { s->vleftdir = v3_sub(s->v2, s->v0); s->vrightdir = v3_sub(s->v1, s->v0); }
            s->nleft = s->n0; s->nright = s->n0;
            if (lor) { s->nleftdir = v3_sub(s->n1, s->n0); s->nrightdir = v3_sub(s->n2, s->n0); }
            else     { s->nleftdir = v3_sub(s->n2, s->n0); s->nrightdir = v3_sub(s->n1, s->n0); }
        } else {
            /* prepare_for_lower_triangle, pixel_shaders.cpp:127-151 */
            if (lor) {
                s->vright = s->v0; s->vrightdir = v3_sub(s->v2, s->v0);
                s->vleft = s->v1;  s->vleftdir = v3_sub(s->v2, s->v1);
                s->nright = s->n0; s->nrightdir = v3_sub(s->n2, s->n0);
                s->nleft = s->n1;  s->nleftdir = v3_sub(s->n2, s->n1);
            } else {
                s->vleft = s->v0;  s->vleftdir = v3_sub(s->v2, s->v0);
                s->vright = s->v1; s->vrightdir = v3_sub(s->v2, s->v1);
                s->nleft = s->n0;  s->nleftdir = v3_sub(s->n2, s->n0);
                s->nright = s->n1; s->nrightdir = v3_sub(s->n2, s->n1);
            }
        }
    }
    if (tm != SWEGL_B200_TEX_PLAIN) {
        /* pixel_shaders.cpp:247-259, 320-332 */
        s->long_line_on_right = lor;
        if (!lower) { s->side_short_t = s->t0; s->side_short_t_dir.x = s->t1.x - s->t0.x; s->side_short_t_dir.y = s->t1.y - s->t0.y; }
        else        { s->side_short_t = s->t1; s->side_short_t_dir.x = s->t2.x - s->t1.x; s->side_short_t_dir.y = s->t2.y - s->t1.y; }
    }
}

static void ps_prepare_for_scanline(rs_t *s, float pl, float pr)
{
    const int lm = s->vp->light_mode, tm = s->vp->tex_mode;
    if (lm == SWEGL_B200_LIGHT_PHONG) {
        /* pixel_shaders.cpp:152-158 */
        s->v = v3_add(s->vleft, v3_mul(s->vleftdir, pl));
        s->vdir = v3_sub(v3_add(s->vright, v3_mul(s->vrightdir, pr)), s->v);
        s->n = v3_add(s->nleft, v3_mul(s->nleftdir, pl));
        s->ndir = v3_sub(v3_add(s->nright, v3_mul(s->nrightdir, pr)), s->n);
    }
    if (tm != SWEGL_B200_TEX_PLAIN) {
        /* pixel_shaders.cpp:261-273, 334-346 */
        if (s->long_line_on_right) {
            s->t_left.x = s->side_short_t.x + s->side_short_t_dir.x * pl;
            s->t_left.y = s->side_short_t.y + s->side_short_t_dir.y * pl;
            s->t_dir.x = s->t0.x + s->side_long_t_dir.x * pr - s->t_left.x;
            s->t_dir.y = s->t0.y + s->side_long_t_dir.y * pr - s->t_left.y;
        } else {
            s->t_left.x = s->t0.x + s->side_long_t_dir.x * pl;
            s->t_left.y = s->t0.y + s->side_long_t_dir.y * pl;
            s->t_dir.x = s->side_short_t.x + s->side_short_t_dir.x * pr - s->t_left.x;
            s->t_dir.y = s->side_short_t.y + s->side_short_t_dir.y * pr - s->t_left.y;
        }
    }
}

/* deviation: a negative texel row / column -- the reference then reads outside the bitmap -- is wrapped.
 * g_texel_guard counts the fragments where that happened, so that a comparison with the reference can tell "the
 * reference left its defined behaviour on this frame" from a real difference (orc_dump::n_texel_guard). */
static uint64_t g_texel_guard;
static int wrap_i(int v, int n) { if (v < 0) { g_texel_guard++; v %= n; if (v < 0) v += n; } return v; }

static uint32_t shade_texture(const rs_t *s, float progress)
{
    const int tm = s->vp->tex_mode;
    if (tm == SWEGL_B200_TEX_PLAIN)
        return s->color;                                    /* pixel_shader_t::shade, pixel_shaders.hpp:28 */
    v2 t = { s->t_left.x + s->t_dir.x * progress, s->t_left.y + s->t_dir.y * progress };
    if (tm == SWEGL_B200_TEX_NEAREST) {
        /* pixel_shader_texture::shade, pixel_shaders.cpp:275-281 (twidth/theight are unsigned there) */
        unsigned tw = (unsigned)s->twidth, th = (unsigned)s->theight;
        int u = (int)((unsigned)f2i(t.x) % tw);
        int v = (int)((unsigned)f2i(t.y) % th);
        return s->tbitmap[(unsigned)v * tw + (unsigned)u];
    }
    /* pixel_shader_texture_bilinear::shade, pixel_shaders.cpp:348-384 */
    float v = t.x;
    float u = t.y;
    float u1 = (float)(u - 0.5);
    float u2 = (float)(u + 0.5);
    float v1 = (float)(v - 0.5);
    float v2_ = (float)(v + 0.5);
    u = (float)floor(u2);
    v = (float)floor(v2_);
    int tw = s->twidth, th = s->theight;
    int v1m = wrap_i((f2i(v1) + th) % th, th);
    int v2m = v1m + 1;
    if (v2m == th) v2m = 0;
    v1m *= tw; v2m *= tw;
    int u1m = wrap_i((f2i(u1) + tw) % tw, tw);
    int u2m = u1m + 1;
    if (u2m == tw) u2m = 0;
    pc_t p00, p10, p01, p11, out;
    p00.i = s->tbitmap[v1m + u1m];
    p10.i = s->tbitmap[v2m + u1m];
    p01.i = s->tbitmap[v1m + u2m];
    p11.i = s->tbitmap[v2m + u2m];
    float w00 = (u - u1) * (v - v1), w10 = (u - u1) * (v2_ - v), w01 = (u2 - u) * (v - v1), w11 = (u2 - u) * (v2_ - v);
    /* pixel_colors * float (colors.cpp:27-30), _mm_add_ps left to right (:11-16), round (:19-25) */
    float b = ((p00.o.b * w00 + p10.o.b * w10) + p01.o.b * w01) + p11.o.b * w11;
    float g = ((p00.o.g * w00 + p10.o.g * w10) + p01.o.g * w01) + p11.o.g * w11;
    float r = ((p00.o.r * w00 + p10.o.r * w10) + p01.o.r * w01) + p11.o.r * w11;
    float a = ((p00.o.a * w00 + p10.o.a * w10) + p01.o.a * w01) + p11.o.a * w11;
    out.o.b = (uint8_t)d2i(round((double)b));
    out.o.g = (uint8_t)d2i(round((double)g));
    out.o.r = (uint8_t)d2i(round((double)r));
    out.o.a = (uint8_t)d2i(round((double)a));
    return out.i;
}

static int shade_light(const rs_t *s, float progress)
{
    if (s->vp->light_mode == SWEGL_B200_LIGHT_FLAT)
        return f2i(s->flat_light);                          /* pixel_shaders.hpp:36-39 */
    /* pixel_shader_lights_phong::shade, pixel_shaders.cpp:159-205 */
    v3 center = v3_add(s->v, v3_mul(s->vdir, progress));
    v3 normal = v3_normalize(v3_add(s->n, v3_mul(s->ndir, progress)));
    v3 cam = { s->vp->cam_pos[0], s->vp->cam_pos[1], s->vp->cam_pos[2] };
    v3 camv = v3_normalize(v3_sub(cam, center));
    float sun = -v3_dot(normal, *(const v3 *)s->frame->sun_dir);
    if (sun < 0.0f) sun = 0.0f; else sun *= s->frame->sun_intensity;
    float dyn = point_lights_sum(s, center, normal, camv);
    return f2i(65536 * (s->frame->ambient + sun + dyn));
}

static uint32_t ps_shade(const rs_t *s, float progress)
{
    if (s->vp->light_mode == SWEGL_B200_LIGHT_NONE)
        return shade_texture(s, progress);                  /* a bare texture / colour shader */
    /* pixel_shader_light_and_texture::shade, pixel_shaders.hpp:159-178 */
    pc_t c; c.i = shade_texture(s, progress);
    float light = (float)(shade_light(s, progress) / 65536.0);
    if (light < 1) {
        c.o.b = (uint8_t)f2i(c.o.b * light);
        c.o.g = (uint8_t)f2i(c.o.g * light);
        c.o.r = (uint8_t)f2i(c.o.r * light);
    } else {
        light = (float)sqrt((double)light);
        light = (float)sqrt((double)light);
        c.o.b = (uint8_t)(255 - (uint8_t)f2i((255 - c.o.b) / light));
        c.o.g = (uint8_t)(255 - (uint8_t)f2i((255 - c.o.g) / light));
        c.o.r = (uint8_t)(255 - (uint8_t)f2i((255 - c.o.r) / light));
    }
    return c.i;
}

/* ---- rasteriser ---- */

static void fill_half_triangle(rs_t *s, int y, int y_end, line_side *left, line_side *right)
{
    /* renderer.cpp:462-558 */
    const swegl_b200_viewport_desc *vp = s->vp;
    for (; y < y_end; y++) {
        int x1 = ceil_i(left->x);  if (x1 < vp->x) x1 = vp->x;
        int x2 = ceil_i(right->x); if (x2 > vp->x + vp->w) x2 = vp->x + vp->w;
        if (x1 < x2 && y >= s->band_y0 && y < s->band_y1) {
            ps_prepare_for_scanline(s, left->ip.ualpha, right->ip.ualpha);
            interp_t q;
            interp_init_self(&q, right->x - left->x, interp_value(&left->ip), interp_value(&right->ip));
            interp_displace(&q, x1 - left->x);
            uint32_t *video = &s->pixels[(size_t)y * s->pitch_words + x1];
            int off = (y - vp->y) * vp->w + (x1 - vp->x);
            float *zb = &s->zbuffer[off];
            if (s->dump) s->dump->n_spans++;
            for (; x1 < x2; x1++, video++, zb++, off++, interp_step(&q)) {
                float z = interp_value(&q);
                if ((double)z <= 0.001) continue;
                if (z >= *zb) continue;
                uint32_t new_color = ps_shade(s, q.ualpha);
                if (s->dump) s->dump->n_fragments++;
                if (!s->got_transparency) { *video = new_color; *zb = z; continue; }
                /* transparency layers, renderer.cpp:500-550 */
                int L = s->n_layers, li;
                for (li = 0; li < L; li++)
                    if (s->layer_z[li][off] == max_z() || s->layer_z[li][off] < z) break;
                if ((new_color >> 24) == 255) {
                    *video = new_color; *zb = z;
                    int i, k;
                    for (i = 0, k = li; k < L; i++, k++) {
                        s->layer_z[i][off] = s->layer_z[k][off];
                        s->layer_colors[i][off] = s->layer_colors[k][off];
                    }
                    for (; i < L; i++) { s->layer_z[i][off] = max_z(); s->layer_colors[i][off] = 0; }
                } else {
                    int all_used = s->layer_z[L - 1][off] != max_z();
                    float zz = z; uint32_t cc = new_color;
                    if (all_used) {
                        while (li-- > 0) {
                            float tz = s->layer_z[li][off]; s->layer_z[li][off] = zz; zz = tz;
                            uint32_t tc = s->layer_colors[li][off]; s->layer_colors[li][off] = cc; cc = tc;
                        }
                    } else {
                        for (; li < L; li++) {
                            float tz = s->layer_z[li][off]; s->layer_z[li][off] = zz; zz = tz;
                            uint32_t tc = s->layer_colors[li][off]; s->layer_colors[li][off] = cc; cc = tc;
                            if (zz == max_z()) break;
                        }
                    }
                }
            }
        }
        left->x += left->ratio;
        right->x += right->ratio;
        interp_step(&left->ip);
        interp_step(&right->ip);
    }
}

static void fill_triangle_2(rs_t *s, const overt *a, const overt *b, const overt *c, int front_face_visible)
{
    /* renderer.cpp:361-460 */
    const swegl_b200_viewport_desc *vp = s->vp;
    const overt *t;
    int inverted = !front_face_visible;
    if (b->v_viewport.y < a->v_viewport.y) { t = a; a = b; b = t; }
    if (c->v_viewport.y < b->v_viewport.y) { t = b; b = c; c = t; }
    if (b->v_viewport.y < a->v_viewport.y) { t = a; a = b; b = t; }
    const v3 *v0 = &a->v_viewport, *v1 = &b->v_viewport, *v2 = &c->v_viewport;
    int y0 = ceil_i(v0->y), y1 = ceil_i(v1->y), y2 = ceil_i(v2->y);
    if (y0 == y2) return;
    if (s->dump) s->dump->n_setup_triangles++;

    line_side side_long, side_short;
    side_long.ratio = (v2->x - v0->x) / (v2->y - v0->y);
    interp_init_self(&side_long.ip, v2->y - v0->y, v0->z, v2->z);
    if (y0 < vp->y) {
        interp_displace(&side_long.ip, vp->y - v0->y);
        side_long.x = v0->x + side_long.ratio * (vp->y - v0->y);
    } else {
        interp_displace(&side_long.ip, y0 - v0->y);
        side_long.x = v0->x + side_long.ratio * (y0 - v0->y);
    }
    ps_prepare_for_triangle(s, a, b, c, inverted);

    int y, y_end;
    if (y1 >= vp->y) {
        side_short.ratio = (v1->x - v0->x) / (v1->y - v0->y);
        interp_init_self(&side_short.ip, v1->y - v0->y, v0->z, v1->z);
        y = y0 > vp->y ? y0 : vp->y;
        y_end = y1 < vp->y + vp->h ? y1 : vp->y + vp->h;
        interp_displace(&side_short.ip, y - v0->y);
        side_short.x = v0->x + side_short.ratio * (y - v0->y);
        int lor = side_long.ratio > side_short.ratio;
        ps_prepare_for_half(s, 0, lor);
        if (lor) fill_half_triangle(s, y, y_end, &side_short, &side_long);
        else     fill_half_triangle(s, y, y_end, &side_long, &side_short);
    }
    if (y1 < vp->y + vp->h) {
        side_short.ratio = (v2->x - v1->x) / (v2->y - v1->y);
        interp_init_self(&side_short.ip, v2->y - v1->y, v1->z, v2->z);
        y = y1 > vp->y ? y1 : vp->y;
        y_end = y2 < vp->y + vp->h ? y2 : vp->y + vp->h;
        interp_displace(&side_short.ip, y - v1->y);
        side_short.x = v1->x + side_short.ratio * (y - v1->y);
        int lor = side_long.ratio < side_short.ratio;
        ps_prepare_for_half(s, 1, lor);
        if (lor) fill_half_triangle(s, y, y_end, &side_short, &side_long);
        else     fill_half_triangle(s, y, y_end, &side_long, &side_short);
    }
}

/* camera_to_frustum + frustum_to_viewport on one vertex, vertex_shaders.hpp:54-71, viewport.cpp:123-129 */
static void camera_to_frustum(overt *mv, const float *node_normal9, const swegl_b200_viewport_desc *vp)
{
    /* rotate() normalises in the normal_t ctor, .normalize() again, and the assignment goes through
     * normal_t::operator=(const vector_t&) (points.hpp:143-150) which normalises a third time */
    mv->normal_world = v3_normalize(v3_normalize(rotate3(node_normal9, mv->normal)));
    mv->v_viewport = xform(vp->proj, mv->v_viewport);
    if (mv->v_viewport.z != 0) {
        mv->v_viewport.x = (float)(mv->v_viewport.x / fabs((double)mv->v_viewport.z));
        mv->v_viewport.y = (float)(mv->v_viewport.y / fabs((double)mv->v_viewport.z));
    }
}
static void to_viewport(overt *mv, const swegl_b200_viewport_desc *vp)
{
    mv->v_viewport.x = vp->vp_m00 * mv->v_viewport.x + vp->vp_m03;
    mv->v_viewport.y = vp->vp_m11 * mv->v_viewport.y + vp->vp_m13;
}
static void world_to_viewport(overt *mv, const float *node_normal9, const swegl_b200_viewport_desc *vp)
{
    mv->v_viewport = xform(vp->view, mv->v_world);
    camera_to_frustum(mv, node_normal9, vp);
    to_viewport(mv, vp);
}

static int v3_eq(v3 a, v3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/* the clipped vertex of renderer.cpp:289-297 (and :299-307, :321-339) */
static void make_clip_vertex(overt *nv, const overt *from, const overt *to, float cut,
                             const float *node_normal9, const swegl_b200_viewport_desc *vp)
{
    memset(nv, 0, sizeof *nv);
    nv->v_world = v3_add(from->v_world, v3_mul(v3_sub(to->v_world, from->v_world), cut));
    nv->tex.x = from->tex.x + (to->tex.x - from->tex.x) * cut;
    nv->tex.y = from->tex.y + (to->tex.y - from->tex.y) * cut;
    if (v3_eq(to->normal, from->normal))
        nv->normal = from->normal;
    else
        nv->normal = v3_normalize(v3_add(from->normal, v3_mul(v3_sub(to->normal, from->normal), cut)));
    world_to_viewport(nv, node_normal9, vp);
}

static void fill_triangle(rs_t *s, overt *verts, uint32_t i0, uint32_t i1, uint32_t i2, const float *node_normal9)
{
    /* renderer.cpp:240-359 */
    const overt *a = &verts[i0], *b = &verts[i1], *c = &verts[i2], *t;
    if (!a->yes || !b->yes || !c->yes) return;
    if (s->dump) s->dump->n_fill_triangle++;
    int ffv = cross_n(v3_sub(b->v_viewport, a->v_viewport), v3_sub(c->v_viewport, a->v_viewport)).z < 0;
    int inverted_order = 0;
    if (b->v_viewport.z > a->v_viewport.z) { t = a; a = b; b = t; inverted_order = !inverted_order; }
    if (c->v_viewport.z > b->v_viewport.z) { t = b; b = c; c = t; inverted_order = !inverted_order; }
    if (b->v_viewport.z > a->v_viewport.z) { t = a; a = b; b = t; inverted_order = !inverted_order; }
    const v3 *v0 = &a->v_viewport, *v1 = &b->v_viewport, *v2 = &c->v_viewport;

    if ((double)v2->z >= 0.001) {
        fill_triangle_2(s, a, b, c, ffv);
    } else if ((double)v1->z < 0.001) {
        overt n1, n2;
        float cut_1 = (v0->z - 0.001f) / (v0->z - v1->z);
        make_clip_vertex(&n1, a, b, cut_1, node_normal9, s->vp);
        float cut_2 = (v0->z - 0.001f) / (v0->z - v2->z);
        make_clip_vertex(&n2, a, c, cut_2, node_normal9, s->vp);
        ffv = cross_n(v3_sub(n1.v_viewport, *v0), v3_sub(n2.v_viewport, *v0)).z < 0;
        if (inverted_order) ffv = !ffv;
        fill_triangle_2(s, a, &n1, &n2, ffv);
    } else if ((double)v2->z < 0.001) {
        overt n1, n2;
        float cut_0 = (v0->z - 0.001f) / (v0->z - v2->z);
        make_clip_vertex(&n1, a, c, cut_0, node_normal9, s->vp);
        float cut_1 = (v1->z - 0.001f) / (v1->z - v2->z);
        make_clip_vertex(&n2, b, c, cut_1, node_normal9, s->vp);
        ffv = cross_n(v3_sub(*v1, *v0), v3_sub(n2.v_viewport, *v0)).z < 0;
        if (inverted_order) ffv = !ffv;
        fill_triangle_2(s, a, b, &n2, ffv);
        ffv = cross_n(v3_sub(n2.v_viewport, *v0), v3_sub(n1.v_viewport, *v0)).z < 0;
        if (inverted_order) ffv = !ffv;
        fill_triangle_2(s, a, &n2, &n1, ffv);
    }
}

/* inside_camera_frustum / front_face_visible, renderer.cpp:58-75 */
static int inside_camera_frustum(const overt *a, const overt *b, const overt *c)
{
    const v3 *v0 = &a->v_viewport, *v1 = &b->v_viewport, *v2 = &c->v_viewport;
    return ((v0->x >= -1) || (v1->x >= -1) || (v2->x >= -1))
        && ((v0->y >= -1) || (v1->y >= -1) || (v2->y >= -1))
        && ((v0->x < 1) || (v1->x < 1) || (v2->x < 1))
        && ((v0->y < 1) || (v1->y < 1) || (v2->y < 1))
        && (((double)v0->z >= 0.001) || ((double)v1->z >= 0.001) || ((double)v2->z >= 0.001))
        && (v0->x != v1->x || v0->x != v2->x)
        && (v0->y != v1->y || v0->y != v2->y);
}
static int front_face_visible_ndc(const overt *a, const overt *b, const overt *c)
{
    return cross_n(v3_sub(b->v_viewport, a->v_viewport), v3_sub(c->v_viewport, a->v_viewport)).z > 0;
}

void orc_dof_r(const uint32_t *src, const float *depth, uint32_t *dst, int w, int h,
               float focal_distance, float focal_depth)
{
    /* DoF-R: repaired post_shader_depth_box (post_shaders.hpp:63-111), remap_clipped (lerp.hpp:24-43) */
    float *blurf = (float *)malloc(sizeof(float) * (size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        float t = fabsf(focal_distance - depth[i]);
        float a = 1.0f, b = focal_depth, x;
        if (a == b) x = 0.5f; else if (t <= a) x = 0; else if (t >= b) x = 1; else x = (t - a) / (b - a);
        float u = 0.0f, v = 5.0f, r;
        if (x <= 0) r = u; else if (x >= v) r = v; else r = u + v * x;
        blurf[i] = r;
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t i = (size_t)y * w + x;
            int radius = f2i(blurf[i]);
            uint32_t out = src[i];
            if (radius != 0) {
                int b = 0, g = 0, r = 0, count = 0;
                int j0 = y - radius > 0 ? y - radius : 0, j1 = y + radius < h ? y + radius : h;
                int i0 = x - radius > 0 ? x - radius : 0, i1 = x + radius < w ? x + radius : w;
                for (int j = j0; j < j1; j++)
                    for (int k = i0; k < i1; k++)
                        if (blurf[(size_t)j * w + k] != 0) {
                            uint32_t p = src[(size_t)j * w + k];
                            count++; b += p & 0xFF; g += (p >> 8) & 0xFF; r += (p >> 16) & 0xFF;
                        }
                if (count)
                    out = (uint32_t)(b / count) | ((uint32_t)(g / count) << 8) | ((uint32_t)(r / count) << 16) | 0xFF000000u;
            }
            dst[i] = out;
        }
    free(blurf);
}

uint64_t orc_fnv1a64_words(const uint32_t *words, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= words[i]; h *= 1099511628211ull; }
    return h;
}

int orc_render(const swegl_b200_scene_desc *scene, const swegl_b200_frame_desc *frame,
               const swegl_b200_viewport_desc *vp,
               uint32_t *pixels, int32_t pitch_bytes, int32_t screen_w, int32_t screen_h,
               float *zbuffer, orc_dump *dump)
{
    if (!scene || !frame || !vp || !pixels || !zbuffer) return SWEGL_B200_ERR_ARG;
    if (vp->w <= 0 || vp->h <= 0 || vp->x < 0 || vp->y < 0 || vp->x + vp->w > screen_w || vp->y + vp->h > screen_h)
        return SWEGL_B200_ERR_ARG;
    rs_t s; memset(&s, 0, sizeof s);
    s.scene = scene; s.frame = frame; s.vp = vp; s.pixels = pixels; s.pitch_words = pitch_bytes / 4;
    s.zbuffer = zbuffer; s.dump = dump;
    s.got_transparency = vp->transparency_layers > 0;              /* viewport.cpp:27 */
    s.n_layers = vp->transparency_layers > 1 ? vp->transparency_layers : 1;   /* viewport.cpp:37-39 */
    s.band_y0 = vp->y; s.band_y1 = vp->y + vp->h;
    if (vp->band_y0 != 0 || vp->band_y1 != 0) { s.band_y0 = vp->y + vp->band_y0; s.band_y1 = vp->y + vp->band_y1; }
    if (dump) { dump->n_fill_triangle = dump->n_setup_triangles = dump->n_spans = dump->n_fragments = dump->n_covered = dump->n_texel_guard = 0; }
    g_texel_guard = 0;
    const size_t npx = (size_t)vp->w * vp->h;

    /* ---- vertex stage: original_to_world + world_to_camera_or_frustum, vertex_shaders.hpp:16-52 ---- */
    overt *verts = (overt *)calloc(scene->n_vertices ? scene->n_vertices : 1, sizeof(overt));
    for (uint32_t p = 0; p < scene->n_primitives; p++) {
        const swegl_b200_primitive *pr = &scene->primitives[p];
        const float *M = &frame->node_world[16 * pr->node];
        const float *R = &frame->node_normal[9 * pr->node];
        for (uint32_t k = 0; k < pr->n_vertices; k++) {
            uint32_t i = pr->first_vertex + k;
            overt *mv = &verts[i];
            mv->v.x = scene->positions[3 * i]; mv->v.y = scene->positions[3 * i + 1]; mv->v.z = scene->positions[3 * i + 2];
            mv->normal.x = scene->normals[3 * i]; mv->normal.y = scene->normals[3 * i + 1]; mv->normal.z = scene->normals[3 * i + 2];
            mv->tex.x = scene->texcoords[2 * i]; mv->tex.y = scene->texcoords[2 * i + 1];
            mv->v_world = xform(M, mv->v);
            mv->yes = 0;
            mv->v_viewport = xform(vp->view, mv->v_world);
            camera_to_frustum(mv, R, vp);
        }
    }

    /* ---- viewport_t::clear, viewport.cpp:88-113 ---- */
    for (int j = s.band_y0; j < s.band_y1; j++)
        memset(&pixels[(size_t)j * s.pitch_words + vp->x], 0, 4 * (size_t)vp->w);
    memset(zbuffer, 0x7F, 4 * npx);
    s.layer_colors = (uint32_t **)calloc(s.n_layers, sizeof(uint32_t *));
    s.layer_z = (float **)calloc(s.n_layers, sizeof(float *));
    for (int l = 0; l < s.n_layers; l++) {
        s.layer_colors[l] = (uint32_t *)calloc(npx, 4);
        s.layer_z[l] = (float *)malloc(4 * npx);
        memset(s.layer_z[l], 0x7F, 4 * npx);
    }

    /* ---- per node: mark, frustum_to_viewport, paint (renderer.cpp:83-231). Primitives are
     *      flattened in node order, and marking only touches a primitive's own vertices, so
     *      walking primitives in order is the same schedule. ---- */
    uint32_t p = 0;
    while (p < scene->n_primitives) {
        uint32_t pend = p;
        while (pend < scene->n_primitives && scene->primitives[pend].node == scene->primitives[p].node) pend++;
        /* mark pass, renderer.cpp:86-185 */
        for (uint32_t q = p; q < pend; q++) {
            const swegl_b200_primitive *pr = &scene->primitives[q];
            overt *V = &verts[pr->first_vertex];
            const uint32_t *I = &scene->indices[pr->first_index];
            int double_sided = pr->material_id != -1 && scene->materials[pr->material_id].double_sided;
            uint32_t n = pr->n_indices;
            if (pr->mode == SWEGL_B200_MODE_TRIANGLE_STRIP) {
                for (uint32_t i = 2; i < n; i++) {
                    if (inside_camera_frustum(&V[I[i - 2]], &V[I[i - 1]], &V[I[i]])
                        && (double_sided || front_face_visible_ndc(&V[I[i - 2]], &V[I[i - 1 + (i & 1)]], &V[I[i - (i & 1)]])))
                        V[I[i - 2]].yes = V[I[i - 1]].yes = V[I[i]].yes = 1;
                }
            } else if (pr->mode == SWEGL_B200_MODE_TRIANGLE_FAN) {
                for (uint32_t i = 2; i < n; i++) {
                    if (inside_camera_frustum(&V[I[0]], &V[I[i - 1]], &V[I[i]])
                        && (double_sided || front_face_visible_ndc(&V[I[0]], &V[I[i - 1]], &V[I[i]])))
                        V[I[0]].yes = V[I[i - 1]].yes = V[I[i]].yes = 1;
                }
            } else if (pr->mode == SWEGL_B200_MODE_TRIANGLES) {
                for (uint32_t i = 2; i < n; i += 3) {
                    if (inside_camera_frustum(&V[I[i - 2]], &V[I[i - 1]], &V[I[i]])
                        && (double_sided || front_face_visible_ndc(&V[I[i - 2]], &V[I[i - 1]], &V[I[i]])))
                        V[I[i - 2]].yes = V[I[i - 1]].yes = V[I[i]].yes = 1;
                }
            }
        }
        /* frustum_to_viewport, vertex_shaders.hpp:72-84 */
        for (uint32_t q = p; q < pend; q++) {
            const swegl_b200_primitive *pr = &scene->primitives[q];
            for (uint32_t k = 0; k < pr->n_vertices; k++)
                if (verts[pr->first_vertex + k].yes) to_viewport(&verts[pr->first_vertex + k], vp);
        }
        /* paint, renderer.cpp:191-230 */
        for (uint32_t q = p; q < pend; q++) {
            const swegl_b200_primitive *pr = &scene->primitives[q];
            overt *V = &verts[pr->first_vertex];
            const uint32_t *I = &scene->indices[pr->first_index];
            const float *R = &frame->node_normal[9 * pr->node];
            uint32_t n = pr->n_indices;
            ps_prepare_for_primitive(&s, pr);
            if (pr->mode == SWEGL_B200_MODE_TRIANGLE_STRIP)
                for (uint32_t i = 2; i < n; i++) fill_triangle(&s, V, I[i - 2], I[i - 1 + (i & 1)], I[i - (i & 1)], R);
            else if (pr->mode == SWEGL_B200_MODE_TRIANGLE_FAN)
                for (uint32_t i = 2; i < n; i++) fill_triangle(&s, V, I[0], I[i - 1], I[i], R);
            else if (pr->mode == SWEGL_B200_MODE_TRIANGLES)
                for (uint32_t i = 2; i < n; i += 3) fill_triangle(&s, V, I[i - 2], I[i - 1], I[i], R);
        }
        p = pend;
    }

    /* ---- viewport_t::flatten, viewport.cpp:43-86 (reads screen rows j from column 0, as written) ---- */
    for (int l = 1; l < s.n_layers; l++)
        for (size_t i = 0; i < npx; i++)
            if ((s.layer_colors[l][i] >> 24) != 0)
                s.layer_colors[0][i] = blend(s.layer_colors[0][i], s.layer_colors[l][i]);
    {
        uint32_t *front = s.layer_colors[0];
        for (int j = 0; j < vp->h; j++) {
            const uint32_t *back = &pixels[(size_t)j * s.pitch_words];
            for (int i = 0; i < vp->w; i++, front++, back++) {
                if (s.got_transparency && (*front >> 24) != 0) *front = blend(*back, *front);
                else *front = *back;
            }
        }
    }
    /* ---- post shader ---- */
    if (vp->post_mode == SWEGL_B200_POST_DOF) {
        /* DoF-R on the viewport rectangle: src = rendered colour, depth = m_zbuffer */
        uint32_t *src = (uint32_t *)malloc(4 * npx), *dst = (uint32_t *)malloc(4 * npx);
        for (int j = 0; j < vp->h; j++)
            memcpy(&src[(size_t)j * vp->w], &pixels[(size_t)(j + vp->y) * s.pitch_words + vp->x], 4 * (size_t)vp->w);
        if (s.got_transparency) memcpy(src, s.layer_colors[0], 4 * npx);
        orc_dof_r(src, zbuffer, dst, vp->w, vp->h, vp->focal_distance, vp->focal_depth);
        for (int j = 0; j < vp->h; j++)
            memcpy(&pixels[(size_t)(j + vp->y) * s.pitch_words + vp->x], &dst[(size_t)j * vp->w], 4 * (size_t)vp->w);
        free(src); free(dst);
    } else if (s.got_transparency) {
        /* post_shader_t::shade -> copy_first_transparency_layer_to_screen, post_shaders.hpp:22-47 */
        const uint32_t *px = s.layer_colors[0];
        for (int j = 0; j < vp->h; j++) {
            uint32_t *screen = &pixels[(size_t)(j + vp->y) * s.pitch_words];
            for (int i = 0; i < vp->w; i++) *screen++ = *px++;
        }
    }

    if (dump) {
        for (uint32_t i = 0; i < scene->n_vertices; i++) {
            if (dump->v_world) memcpy(&dump->v_world[3 * i], &verts[i].v_world, 12);
            if (dump->v_viewport) memcpy(&dump->v_viewport[3 * i], &verts[i].v_viewport, 12);
            if (dump->normal_world) memcpy(&dump->normal_world[3 * i], &verts[i].normal_world, 12);
            if (dump->yes) dump->yes[i] = (uint8_t)verts[i].yes;
        }
        uint32_t mz = 0x7F7F7F7Fu;
        for (size_t i = 0; i < npx; i++) { uint32_t zb; memcpy(&zb, &zbuffer[i], 4); if (zb != mz) dump->n_covered++; }
        dump->n_texel_guard = g_texel_guard;
    }
    for (int l = 0; l < s.n_layers; l++) { free(s.layer_colors[l]); free(s.layer_z[l]); }
    free(s.layer_colors); free(s.layer_z); free(verts);
    return SWEGL_B200_OK;
}
