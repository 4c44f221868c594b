#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "radd.h"
static uint64_t rs = 88172645463325252ull;
static uint64_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }
static float brute(float s, float c, uint32_t k) { for (uint32_t i = 0; i < k; i++) s = radd_add_(s, c); return s; }
static long fails = 0, tests = 0;
static void check(float s, float c, uint32_t k) {
    float a = brute(s, c, k), b = radd(s, c, k);
    tests++;
    if (radd_bits_(a) != radd_bits_(b) && !(a != a && b != b)) {
        if (fails < 20) printf("FAIL s=%a (%08x) c=%a (%08x) k=%u brute=%a (%08x) radd=%a (%08x)\n", s, radd_bits_(s), c, radd_bits_(c), k, a, radd_bits_(a), b, radd_bits_(b));
        fails++;
    }
}
int main(int argc, char **argv) {
    long N = argc > 1 ? atol(argv[1]) : 2000000;
    // 1. fully random bit patterns (includes NaN/inf/denormals), small k
    for (long i = 0; i < N; i++) { float s = radd_float_((uint32_t)rnd()), c = radd_float_((uint32_t)rnd()); check(s, c, rnd() % 300); }
    // 2. s and c with nearby exponents (the interesting regime), both sign combos, k up to 8192
    for (long i = 0; i < N; i++) {
        int es = 100 + rnd() % 60, de = (int)(rnd() % 40) - 32;
        int ec = es + de; if (ec < 0) ec = 0; if (ec > 254) ec = 254;
        uint32_t sb = ((uint32_t)(rnd() & 1) << 31) | ((uint32_t)es << 23) | (uint32_t)(rnd() & 0x7FFFFF);
        uint32_t cb = ((uint32_t)(rnd() & 1) << 31) | ((uint32_t)ec << 23) | (uint32_t)(rnd() & 0x7FFFFF);
        check(radd_float_(sb), radd_float_(cb), rnd() % 8192);
    }
    // 3. tie-prone increments: c with few mantissa bits (exact halves of the ulp of s)
    for (long i = 0; i < N; i++) {
        int es = 110 + rnd() % 40, de = -(int)(rnd() % 30);
        int ec = es + de;
        uint32_t mant_c = (uint32_t)(rnd() & 0x7FFFFF) & ~((1u << (rnd() % 23)) - 1);   // trailing zeros
        uint32_t sb = ((uint32_t)(rnd() & 1) << 31) | ((uint32_t)es << 23) | (uint32_t)(rnd() & 0x7FFFFF);
        uint32_t cb = ((uint32_t)(rnd() & 1) << 31) | ((uint32_t)ec << 23) | mant_c;
        check(radd_float_(sb), radd_float_(cb), rnd() % 5000);
    }
    // 4. denormal / tiny range and near-overflow
    for (long i = 0; i < N / 4; i++) {
        uint32_t sb = ((uint32_t)(rnd() & 1) << 31) | ((uint32_t)(rnd() % 4) << 23) | (uint32_t)(rnd() & 0x7FFFFF);
        uint32_t cb = ((uint32_t)(rnd() & 1) << 31) | ((uint32_t)(rnd() % 4) << 23) | (uint32_t)(rnd() & 0x7FFFFF);
        check(radd_float_(sb), radd_float_(cb), rnd() % 3000);
        sb = ((uint32_t)(rnd() & 1) << 31) | ((uint32_t)(250 + rnd() % 5) << 23) | (uint32_t)(rnd() & 0x7FFFFF);
        cb = ((uint32_t)(rnd() & 1) << 31) | ((uint32_t)(240 + rnd() % 15) << 23) | (uint32_t)(rnd() & 0x7FFFFF);
        check(radd_float_(sb), radd_float_(cb), rnd() % 3000);
    }
    // 5. rasteriser-like values: x in [0,8000), ratio in [-50,50]; 1/z style values
    for (long i = 0; i < N; i++) {
        float s = (float)((double)(rnd() % 8000000) / 1000.0) - 100.f, c = (float)(((double)(rnd() % 2000001) - 1000000.0) / 20000.0);
        if (rnd() % 8 == 0) c = (float)(int)c;                     // integral slopes
        if (rnd() % 8 == 0) c = c / 1024.f;
        check(s, c, rnd() % 8000);
        float t = (float)((double)(rnd() % 1000000) / 1e6), ts = (float)(((double)(rnd() % 2000001) - 1e6) / 1e9);
        check(t, ts, rnd() % 8000);
    }
    printf("%ld tests, %ld failures\n", tests, fails);
    return fails != 0;
}
