/*
 * tools/libm_exhaustive.c -- exhaustive check of csrc/pwn_libm.h against the host's glibc:
 * every float with |x| < 2^20, both signs (2 466 250 752 inputs) for sinf, cosf, sincosf, and the
 * non-positive ones above -87 for expf.
 *
 *   gcc -O2 -fopenmp -ffp-contract=off -I pwnfps_b200/csrc -o /tmp/libm_ex tools/libm_exhaustive.c -lm
 *   /tmp/libm_ex 1                                             # the "_fma" build against this host's libm
 *   GLIBC_TUNABLES=glibc.cpu.hwcaps=-AVX2,-FMA /tmp/libm_ex 0   # the "_sse2" build, glibc forced onto it
 *
 * Result on the build box (glibc 2.39-0ubuntu8.5, Xeon Sapphire Rapids), recorded in
 * profiles/r01_libm_exhaustive.txt: 0 mismatches for both builds; the two builds differ from each
 * other on 12 (sin) / 22 (cos) / 1 (exp) inputs.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <omp.h>
#include "pwn_libm.h"
static const uint64_t TAB[32] = PWN_EXP2F_TAB;
int main(int argc, char **argv)
{
	int fma_ = argc > 1 ? atoi(argv[1]) : 1;
	/* all floats with |x| < 2^20 (both signs): exponent field < 127+20 */
	unsigned long long bad_s = 0, bad_c = 0, bad_sc = 0, bad_e = 0, tot = 0;
	uint32_t first_s = 0, first_c = 0, first_e = 0;
	#pragma omp parallel for reduction(+:bad_s,bad_c,bad_sc,bad_e,tot) schedule(dynamic, 1<<16)
	for(long long b = 0; b < (long long)((127u+20u) << 23); b++)
	{
		for(int sg = 0; sg < 2; sg++)
		{
			uint32_t u = (uint32_t)b | ((uint32_t)sg << 31);
			float x; memcpy(&x, &u, 4);
			float s = sinf(x), c = cosf(x), s2, c2, ps, pc;
			sincosf(x, &s2, &c2);
			pwn_sincosf(x, &ps, &pc, fma_);
			tot++;
			if(pwn_asuint(ps) != pwn_asuint(s)) { bad_s++; if(!first_s) first_s = u; }
			if(pwn_asuint(pc) != pwn_asuint(c)) { bad_c++; if(!first_c) first_c = u; }
			if(pwn_asuint(s2) != pwn_asuint(s) || pwn_asuint(c2) != pwn_asuint(c)) bad_sc++;
			if(sg == 1 || b == 0)
			{
				float e = expf(x), pe = pwn_expf(x, TAB, fma_);
				/* the trace only needs x <= 0; below -88 any tiny value blends identically */
				if(x >= -87.0f && pwn_asuint(pe) != pwn_asuint(e)) { bad_e++; if(!first_e) first_e = u; }
			}
		}
	}
	printf("fma=%d tested %llu: sin mismatches %llu (first %08x), cos %llu (first %08x), sincosf!=sinf/cosf %llu, expf %llu (first %08x)\n",
		fma_, tot, bad_s, first_s, bad_c, first_c, bad_sc, bad_e, first_e);
	return 0;
}
