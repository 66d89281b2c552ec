/* exact_div_check.c -- exhaustive check of the FMA quotient used by the flow_warp coordinate pipeline
 * (unsamflow_b200/csrc/warp_coord.cuh: div_by_size).
 *
 * Claim: for an integer divisor n (= image size - 1), y = RN(1/n) and ANY fp32 t,
 *     q0 = RN(t*y);  r = t - n*q0 (one FMA, exact);  q1 = RN(q0 + r*y)
 * equals the correctly rounded IEEE quotient RN(t/n).  The quotient is invariant under scaling t by powers of two
 * (no over/underflow in the kernel's range |t| < 2^100, tiny quotients vanish in the following "- 1"), so it is
 * enough to try every significand: all 2^24 floats of [1, 4).
 *
 *   gcc -O2 -mfma -fopenmp -ffp-contract=off exact_div_check.c -o exact_div_check -lm
 *   ./exact_div_check LO HI      -> prints "bad=<count>" over divisors LO..HI (exit 1 if any mismatch)
 * Round 2 ran it over 1..16400: bad=0 (the kernel uses the IEEE sequence beyond that).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
    const int lo = argc > 1 ? atoi(argv[1]) : 1, hi = argc > 2 ? atoi(argv[2]) : 300;
    long bad = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : bad)
    for (int n = lo; n <= hi; ++n) {
        const float b = (float)n, y = 1.0f / b;
        for (uint32_t m = 0; m < (1u << 24); ++m) {
            const uint32_t bits = (127u << 23) + m; /* [1, 4) */
            float t;
            memcpy(&t, &bits, 4);
            const float q0 = t * y;
            const float r0 = fmaf(-b, q0, t);
            const float q1 = fmaf(r0, y, q0);
            if (q1 != t / b) bad++;
        }
    }
    printf("bad=%ld\n", bad);
    return bad != 0;
}
