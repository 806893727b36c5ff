/* host check of the 4-operation exact divide of csrc/common.cuh (tb2_rcp_of / tb2_div_by):
 * the same operation sequence with the host's fused multiply-add against the IEEE divide, on the
 * generator of csrc/debug.cu (k_div_check).  Test infrastructure only. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t xs(uint64_t *s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; *s = x; return x; }
static inline double mk(uint64_t mant, int e)
{
    uint64_t u = ((uint64_t)(e + 1023) << 52) | (mant & 0xFFFFFFFFFFFFFULL);
    double d; memcpy(&d, &u, 8); return d;
}

/* returns the number of mismatches over n pairs; *q0_bad counts quotients the first (uncorrected)
 * estimate already missed -- informational */
long div_check(uint64_t seed, long n, long *q0_bad, double *example4)
{
    uint64_t s = 0x1234567ULL ^ (seed * 0x9E3779B97F4A7C15ULL);
    long bad = 0; *q0_bad = 0;
    for (long it = 0; it < n; ++it) {
        const uint64_t r0 = xs(&s), r1 = xs(&s), r2 = xs(&s);
        uint64_t mb = r0;
        switch (r2 & 7) {
        case 0: mb = 0xFFFFFFFFFFFFFULL; break;
        case 1: mb = 0; break;
        case 2: mb = 0xFFFFFFFFFFFFFULL - (r0 & 15); break;
        case 3: mb = r0 & 15; break;
        default: break;
        }
        const double b = mk(mb, (int)((r2 >> 8) % 21) - 10);
        double a;
        if ((r2 >> 16) & 1) {
            const double q = mk(r1, (int)((r2 >> 20) % 31) - 20);
            a = b * q;
            int64_t u; memcpy(&u, &a, 8); u += (int64_t)((r2 >> 32) % 5) - 2; memcpy(&a, &u, 8);
        } else {
            a = mk(r1, (int)((r2 >> 20) % 43) - 30);
        }
        if ((r2 >> 40) % 97 == 0) a = 0.0;
        const double want = a / b;
        const double y = 1.0 / b, ylo = fma(-b, y, 1.0) * y;       /* tb2_rcp_of */
        const double q0 = fma(a, y, a * ylo);                        /* tb2_div_by */
        const double r = fma(-b, q0, a);
        const double got = fma(r, y, q0);
        if (memcmp(&want, &got, 8)) {
            if (!bad) { example4[0] = a; example4[1] = b; example4[2] = want; example4[3] = got; }
            ++bad;
        }
        if (memcmp(&want, &q0, 8)) ++*q0_bad;
    }
    return bad;
}
