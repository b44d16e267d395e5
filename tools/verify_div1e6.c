/*
 * tools/verify_div1e6.c -- exhaustive proof obligation for cb::div_1e6 (cb_device.cuh).
 *
 * utils.currentMillis() (reference lib/utils.js:198-204) divides the nanosecond part of
 * process.hrtime() -- an integer in [0, 1e9) -- by 1e6 in IEEE double precision.  The
 * CUDA path replaces the division by   q = x*R;  r = fma(-q, 1e6, x);  q' = fma(r, R, q)
 * with R = RN(1e-6).  This program checks, for EVERY one of the 10^9 possible dividends
 * (or a strided subset: argv[1] = stride), that q' is bit-identical to x / 1e6.
 *
 * build: gcc -O2 -mfma -ffp-contract=off -o verify_div1e6 verify_div1e6.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int
main(int argc, char **argv)
{
	const double R = 1e-6;
	uint64_t stride = argc > 1 ? strtoull(argv[1], NULL, 10) : 1, bad = 0, n = 0;
	for (uint64_t i = 0; i < 1000000000ull; i += stride) {
		double x = (double)i;
		double ref = x / 1000000.0;
		double q = x * R;
		double r = fma(-q, 1000000.0, x);
		double got = fma(r, R, q);
		if (memcmp(&ref, &got, 8) != 0) {
			if (bad < 10)
				printf("MISMATCH x=%llu ref=%a got=%a\n", (unsigned long long)i, ref, got);
			bad++;
		}
		n++;
	}
	printf("checked %llu dividends, %llu mismatches\n", (unsigned long long)n, (unsigned long long)bad);
	return bad != 0;
}
