// radd.h — exact k-fold repeated fp32 addition: s <- RN(s + c), k times, without doing k additions.
//
// swegl's rasteriser is built on serial fp32 recurrences (side.x += ratio, renderer.cpp:553-554;
// topalpha += topstep / bottomalpha += bottomstep, interpolator.hpp:96-100).  x_k != x_0 + k*ratio in
// floating point, so pixel-identical coverage needs the SAME additions -- but not one at a time:
// radd() jumps k steps in O(#binades crossed), which turns every scanline and every span segment into an
// independent work item.  Compiles as C (tests/radd_bruteforce.c checks it against the k-step loop on
// millions of adversarial inputs) and as CUDA device code.
#pragma once
#include <stdint.h>
#include <string.h>
#ifdef __CUDACC__
#define RADD_FN __device__ __forceinline__
#define RADD_ADD(a, b) __fadd_rn(a, b)
#define RADD_BITS(f) __float_as_uint(f)
#define RADD_FLOAT(u) __uint_as_float(u)
static __device__ __forceinline__ float radd_rcp_approx_(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#define RADD_RCP(x) radd_rcp_approx_(x)     // MUFU.RCP, <= 1 ulp: covered by the 2^-19 safety factor below
#else
#define RADD_FN static inline
static inline float radd_add_(float a, float b) { volatile float r = a + b; return r; }
static inline uint32_t radd_bits_(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float radd_float_(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define RADD_ADD(a, b) radd_add_(a, b)
#define RADD_BITS(f) radd_bits_(f)
#define RADD_FLOAT(u) radd_float_(u)
#define RADD_RCP(x) (1.0f / (x))
#endif

// Within one binade (fixed sign and exponent) the grid spacing u is constant, so RN(s + c) moves the
// bit pattern by a constant D per step once the state is "settled" (result of an in-binade step: in the
// round-half-even tie case that makes the mantissa even, after which the tie always resolves the same
// way).  We therefore take real additions until three consecutive values r1,r2,r3 share a binade, read
// D = |r3 - r2| off the bit patterns, jump as many steps as keep the mantissa strictly inside the binade
// (so every skipped step's exact sum stayed in the binade too), and fall back to real additions to cross
// the boundary.  A step that does not change the value is a fixed point and ends the recurrence.
// a lower bound of floor(num / den) for num, den < 2^24 that is never more than ~num/den * 2^-19 + 1 short:
// one float multiply by a reciprocal instead of a ~100-cycle integer division.  Jumping fewer steps than
// allowed is always safe (the loop just takes real steps / another jump for the rest).
RADD_FN uint32_t radd_div_floor_lb(uint32_t num, uint32_t den)
{
    const float q = (float)num * RADD_RCP((float)den) * 0.99999809265136719f;   // * (1 - 2^-19)
    return (uint32_t)q;
}

RADD_FN float radd(float s, float c, uint32_t k)
{
    while (k) {
        float r1 = RADD_ADD(s, c); k--;
        if (RADD_BITS(r1) == RADD_BITS(s) || k == 0) return r1;
        float r2 = RADD_ADD(r1, c); k--;
        if (RADD_BITS(r2) == RADD_BITS(r1) || k == 0) return r2;
        float r3 = RADD_ADD(r2, c); k--;
        if (RADD_BITS(r3) == RADD_BITS(r2) || k == 0) return r3;
        s = r3;
        uint32_t b1 = RADD_BITS(r1), b2 = RADD_BITS(r2), b3 = RADD_BITS(r3);
        // same sign and exponent for all three, finite
        if ((((b1 ^ b2) | (b2 ^ b3)) & 0xFF800000u) == 0 && (b3 & 0x7F800000u) != 0x7F800000u) {
            uint32_t m3 = b3 & 0x7FFFFFu;
            if (b3 > b2) {                                  // magnitude grows
                uint32_t D = b3 - b2;
                if (b2 > b1) {
                    uint32_t n = radd_div_floor_lb(0x7FFFFFu - m3, D);
                    if (n > k) n = k;
                    b3 += n * D; k -= n;
                }
            } else {                                        // magnitude shrinks
                uint32_t D = b2 - b3;
                if (b1 > b2 && m3 >= 1) {
                    uint32_t n = radd_div_floor_lb(m3 - 1, D);
                    if (n > k) n = k;
                    b3 -= n * D; k -= n;
                }
            }
            s = RADD_FLOAT(b3);
        }
    }
    return s;
}
