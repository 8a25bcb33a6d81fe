// tools/markstein_check.c — host check of dev_math.cuh: normalised3_shared_rcp (three divisions by one divisor from one
// correctly rounded reciprocal + Markstein's correction) against plain IEEE division, on vectors like the direction pass
// sees them.  gcc -O2 -mfma -ffp-contract=off tools/markstein_check.c -lm && ./a.out 300000000   (3e8 vectors: 0 mismatches)
// Does q = fma(fma(-q0, n, a), y, q0), q0 = a*y, y = RN(1/n) equal RN(a/n) ?  (Markstein's final correction with a shared reciprocal)
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t rnd(void) { uint64_t a = s[0], b = s[1]; s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26); return s[1] + b; }
static inline double u01(void) { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
int main(int argc, char **argv)
{
    long long N = argc > 1 ? atoll(argv[1]) : 100000000LL;
    long long bad = 0, worst_cases = 0;
    for (long long i = 0; i < N; i++) {
        double x, y, z;
        int mode = (int)(rnd() % 4);
        if (mode == 0) { x = u01() * 2 - 1; y = u01() * 2 - 1; z = u01() * 2 - 1; }
        else if (mode == 1) { double sc = ldexp(1.0, (int)(rnd() % 80) - 40); x = (u01() * 2 - 1) * sc; y = (u01() * 2 - 1) * sc; z = (u01() * 2 - 1) * sc; }
        else if (mode == 2) { x = 1.0; y = (u01() * 2 - 1) * 0.2; z = (u01() * 2 - 1) * 0.2; }   // the beam-frame direction: (bs, bs + s*i, ...)
        else { // an already (nearly) unit vector: the second normalisation
            x = u01() * 2 - 1; y = u01() * 2 - 1; z = u01() * 2 - 1; double m = sqrt(x * x + y * y + z * z); x /= m; y /= m; z /= m; }
        volatile double n2 = x * x + y * y + z * z;
        double n = sqrt(n2);
        if (!(n > 0)) continue;
        double r = 1.0 / n;              // correctly rounded reciprocal
        double c[3] = {x, y, z};
        for (int k = 0; k < 3; k++) {
            double a = c[k];
            double q_ref = a / n;
            double q0 = a * r;
            double rem = fma(-q0, n, a);
            double q = fma(rem, r, q0);
            if (memcmp(&q, &q_ref, 8) != 0) { bad++; if (worst_cases++ < 5) printf("mismatch a=%a n=%a ref=%a got=%a\n", a, n, q_ref, q); }
        }
    }
    printf("%lld vectors, %lld mismatching quotients\n", N, bad);
    return bad != 0;
}
