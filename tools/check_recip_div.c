/* Randomised check of fm_div_recip (csrc/fm_wc.cuh): a / b from y = RN(1/b) with Markstein's residual
 * correction applied twice equals IEEE division bit for bit.  Divisors: the integer, half-integer and
 * 2*(n/2)^2 forms the W&C kernel meets, plus arbitrary doubles in [0.5, 1) (the a-denominator 1 - c^2).
 * build: gcc -O2 -ffp-contract=off -o /tmp/check_recip_div tools/check_recip_div.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double divr(double a, double b, double y) {
    double q = a * y;
    double r = fma(-b, q, a);
    q = fma(r, y, q);
    r = fma(-b, q, a);
    return fma(r, y, q);
}
int main(void) {
    long bad = 0, n = 0;
    for (long it = 0; it < 400000000L; ++it) {
        uint64_t u = rnd();
        double b, a;
        switch (it & 3) {
            case 0: b = (double)(1 + (u % 400000)); break;
            case 1: b = (double)(1 + (u % 400000)) * 0.5; break;
            case 2: { double x = (double)(1 + (u % 400000)); b = x * x * 0.5; break; }
            default: { uint64_t m = (rnd() & 0xFFFFFFFFFFFFFull) | 0x3FE0000000000000ull; memcpy(&b, &m, 8); }
        }
        const double y = 1.0 / b;
        uint64_t v = rnd();
        switch ((it >> 2) & 3) {
            case 0: a = (double)(v % 400000); break;
            case 1: { uint64_t m = (v & 0xFFFFFFFFFFFFFull) | 0x3FF0000000000000ull; memcpy(&a, &m, 8); a -= 1.0; break; }
            case 2: { uint64_t m = (v & 0xFFFFFFFFFFFFFull) | ((uint64_t)(0x3C0 + (v >> 58)) << 52); memcpy(&a, &m, 8); break; }
            default: { uint64_t m = (v & 0xFFFFFFFFFFFFFull) | 0x3FF0000000000000ull; memcpy(&a, &m, 8); a = -(a - 1.5) * 3.0; }
        }
        const double t = a / b, g = divr(a, b, y);
        if (memcmp(&t, &g, 8)) { if (bad < 5) printf("differs: a=%a b=%a ieee=%a recip=%a\n", a, b, t, g); ++bad; }
        ++n;
    }
    printf("cases=%ld differences=%ld\n", n, bad);
    return bad != 0;
}
