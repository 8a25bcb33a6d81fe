// tools/markstein_check_patterns.c — the same check on adversarial pairs: significands near all-ones / all-zeros, runs of
// ones and zeros, exponents over 2^-50 .. 2^20.  gcc -O2 -mfma -ffp-contract=off ... -lm && ./a.out 500000000 <seed>
// (4 seeds x 5e8 pairs: 0 mismatches)
// general pairs: random significands (including patterns near all-ones / all-zeros), |a| <= n, moderate exponents
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s[2];
static inline uint64_t rnd(void) { uint64_t a = s[0], b = s[1]; s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26); return s[1] + b; }
static inline double mk(uint64_t sig, int e) { uint64_t b = ((uint64_t)(1023 + e) << 52) | (sig & 0xFFFFFFFFFFFFFull); double d; memcpy(&d, &b, 8); return d; }
static uint64_t sig_pattern(void)
{
    uint64_t r = rnd();
    switch (rnd() % 6) {
    case 0: return r;
    case 1: return ~0ull << (rnd() % 52);            // ones then zeros
    case 2: return (1ull << (rnd() % 52)) - 1;        // zeros then ones
    case 3: return 0xFFFFFFFFFFFFFull - (r & 0xFF);   // near all ones
    case 4: return r & 0xFF;                          // near power of two
    default: return r & (~0ull << (rnd() % 40));
    }
}
int main(int argc, char **argv)
{
    long long N = argc > 1 ? atoll(argv[1]) : 100000000LL;
    s[0] = 0x9E3779B97F4A7C15ull ^ (argc > 2 ? strtoull(argv[2], 0, 0) : 0); s[1] = 0xD1B54A32D192ED03ull;
    long long bad = 0, shown = 0;
    for (long long i = 0; i < N; i++) {
        double n = mk(sig_pattern(), (int)(rnd() % 41) - 20);
        double a = mk(sig_pattern(), (int)(rnd() % 41) - 20 - (int)(rnd() % 30));
        if (fabs(a) > n) a = a / 4096.0;
        if (rnd() & 1) a = -a;
        double r = 1.0 / n, q_ref = a / n, q0 = a * r, rem = fma(-q0, n, a), q = fma(rem, r, q0);
        if (memcmp(&q, &q_ref, 8) != 0) { bad++; if (shown++ < 5) printf("mismatch a=%a n=%a ref=%a got=%a\n", a, n, q_ref, q); }
    }
    printf("%lld pairs, %lld mismatches\n", N, bad);
    return bad != 0;
}
