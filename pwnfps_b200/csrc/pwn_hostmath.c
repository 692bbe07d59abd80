/*
 * pwn_hostmath.c -- dump of the host CPU's rcpps / rsqrtps as lookup tables.
 *
 * The reference computes iavel = _mm_rcp_ps(avel) (trace.h:231) and normalises
 * with _mm_rsqrt_ps (util.h:43): 12-bit hardware approximations whose exact bits
 * are vendor specific.  A B200 cannot execute them, so the device reproduces
 * them from tables dumped here, on the very CPU the reference host would run on.
 *
 * Model: rcpps(2^k * (1+f)) == 2^-k * rcpps(1+f), and rcpps(1+f) depends only
 * on the top `bits` bits of f; likewise rsqrtps with the exponent parity as one
 * more index bit.  The smallest such `bits` is searched (8..23) and the model is
 * then verified: exhaustively over all 2^23 mantissas at the base exponents, and
 * over every other exponent at the bucket edges (first/last mantissa of each
 * bucket, capped at 65536 buckets per exponent).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>

#include <math.h>

#include "pwntrace.h"
#include "pwn_libm.h"

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t hw_rcp(uint32_t xbits) { return f2u(_mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(u2f(xbits))))); }
static inline uint32_t hw_rsqrt(uint32_t xbits) { return f2u(_mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(u2f(xbits))))); }

#define NMANT (1u << 23)

/* smallest bits such that full[m] == full[m with low (23-bits) bits cleared] for all m */
static int min_bits(const uint32_t *full)
{
	int bits;
	for(bits = 8; bits <= 23; bits++)
	{
		uint32_t mask = ~((1u << (23 - bits)) - 1u) & (NMANT - 1u);
		uint32_t m;
		int ok = 1;
		for(m = 0; m < NMANT; m++)
			if(full[m] != full[m & mask]) { ok = 0; break; }
		if(ok) return bits;
	}
	return -1;
}

int pwn_host_math_probe(int *rcp_bits_out, int *rsqrt_bits_out, uint32_t *rcp_tab, size_t rcp_cap, uint32_t *rsqrt_tab, size_t rsqrt_cap)
{
	uint32_t *full = malloc(sizeof(uint32_t) * NMANT);
	uint32_t *full2 = malloc(sizeof(uint32_t) * NMANT);
	uint32_t m;
	int e, rc = PWN_OK;
	if(full == NULL || full2 == NULL) { free(full); free(full2); return PWN_E_ARG; }

	/* ---- rcpps over [1,2) ---- */
	for(m = 0; m < NMANT; m++) full[m] = hw_rcp(0x3F800000u | m);
	int rb = min_bits(full);
	if(rb < 0) { rc = PWN_E_HOSTMATH; goto done; }
	{
		int sh = 23 - rb;
		uint32_t nb = 1u << rb, step = nb > 65536u ? nb / 65536u : 1u, b;
		/* exponent independence: biased exponents whose result stays normal */
		for(e = 2; e <= 252 && rc == PWN_OK; e++)
		for(b = 0; b < nb; b += step)
		{
			uint32_t m0 = b << sh, m1 = m0 | ((1u << sh) - 1u);
			uint32_t adj = (uint32_t)(e - 127) << 23;
			if(hw_rcp(((uint32_t)e << 23) | m0) != full[m0] - adj || hw_rcp(((uint32_t)e << 23) | m1) != full[m0] - adj)
			{ rc = PWN_E_HOSTMATH; break; }
		}
		if(rc != PWN_OK) goto done;
		if(rcp_bits_out) *rcp_bits_out = rb;
		if(rcp_tab != NULL)
		{
			if(rcp_cap < nb) { rc = PWN_E_ARG; goto done; }
			for(b = 0; b < nb; b++) rcp_tab[b] = full[b << sh];
		}
	}

	/* ---- rsqrtps over [1,2) and [2,4) ---- */
	for(m = 0; m < NMANT; m++) { full[m] = hw_rsqrt(0x3F800000u | m); full2[m] = hw_rsqrt(0x40000000u | m); }
	int sb0 = min_bits(full), sb1 = min_bits(full2);
	if(sb0 < 0 || sb1 < 0) { rc = PWN_E_HOSTMATH; goto done; }
	{
		int sb = sb0 > sb1 ? sb0 : sb1;
		int sh = 23 - sb;
		uint32_t nb = 1u << sb, step = nb > 65536u ? nb / 65536u : 1u, b;
		for(e = 1; e <= 254 && rc == PWN_OK; e++)
		{
			int p = (e - 127) & 1;              /* exponent parity */
			int k = (e - 127 - p) / 2;          /* rsqrt(2^(2k+p) y) = 2^-k rsqrt(2^p y) */
			const uint32_t *base = p ? full2 : full;
			uint32_t adj = (uint32_t)k << 23;
			for(b = 0; b < nb; b += step)
			{
				uint32_t m0 = b << sh, m1 = m0 | ((1u << sh) - 1u);
				if(hw_rsqrt(((uint32_t)e << 23) | m0) != base[m0] - adj || hw_rsqrt(((uint32_t)e << 23) | m1) != base[m0] - adj)
				{ rc = PWN_E_HOSTMATH; break; }
			}
		}
		if(rc != PWN_OK) goto done;
		if(rsqrt_bits_out) *rsqrt_bits_out = sb;
		if(rsqrt_tab != NULL)
		{
			if(rsqrt_cap < 2u*nb) { rc = PWN_E_ARG; goto done; }
			for(b = 0; b < nb; b++) { rsqrt_tab[b] = full[b << sh]; rsqrt_tab[nb + b] = full2[b << sh]; }
		}
	}

done:
	free(full);
	free(full2);
	return rc;
}

/*
 * screen.h:31-57: the per-frame ray basis.  Source-order float arithmetic (this file
 * is compiled without -ffast-math and with -ffp-contract=off); tests/ checks it
 * bit for bit against the reference's own shipped-flags build over many sizes.
 * cam: rows x,y,z,w (defs.h:46-52).
 */
void pwn_ray_setup(const float cam[16], int dimx, int dimy, float rayb[4], float rdx[4], float rdy[4])
{
	int i;
	float xrat = -1.0f;
	float yrat = -((float)dimy)/(float)dimx;
	float xsrat = 2.0f * xrat / (float)dimx;
	float ysrat = 2.0f * yrat / (float)dimy;
	for(i = 0; i < 4; i++)
	{
		rayb[i] = cam[8+i] + ((-xrat)*cam[0+i] + (-yrat)*cam[4+i]);
		rdx[i] = xsrat*cam[0+i];
		rdy[i] = ysrat*cam[4+i];
	}
}

/*
 * Which glibc build of sinf/cosf/expf is this host running (the "_fma" or the "_sse2" one, see
 * pwn_libm.h)?  Asked of the host libm itself: first on the inputs where the two builds round
 * differently (found by an exhaustive search over |x| < 2^20), then on a spread of 2^18 inputs
 * over the ranges the trace produces.  *exact = 1 when every sample agrees with the chosen
 * restatement, 0 when the host's libm is something else (parity of the bounce shading is then
 * only as good as 1-2 ulp).
 */
int pwn_host_libm_probe(int *fma_build, int *exact)
{
	static const uint64_t tab[32] = PWN_EXP2F_TAB;
	/* volatile: the compiler must not fold sinf(constant) with its own (correctly rounded) arithmetic */
	static volatile uint32_t disc_sin[] = {0x4255b0a9u}, disc_cos[] = {0x418a3adbu}, disc_exp[] = {0xc27c65d9u};
	int votes[2] = {0, 0}, v, ok = 1;
	uint32_t i;
	for(v = 0; v < 2; v++)
	{
		float x, s, c;
		x = u2f(disc_sin[0]); pwn_sincosf(x, &s, &c, v); votes[v] += (f2u(s) == f2u(sinf(x)));
		x = u2f(disc_cos[0]); pwn_sincosf(x, &s, &c, v); votes[v] += (f2u(c) == f2u(cosf(x)));
		x = u2f(disc_exp[0]); votes[v] += (f2u(pwn_expf(x, tab, v)) == f2u(expf(x)));
	}
	v = votes[1] >= votes[0] ? 1 : 0;
	if(votes[v] != 3) ok = 0;
	/* a spread: multiplicative-hash walk over [0, 2^18) mapped to (-400, 400) and (-87, 0] */
	for(i = 0; i < (1u << 18) && ok; i++)
	{
		const uint32_t h = i*2654435761u;
		const float x = ((float)(h >> 8)/16777216.0f - 0.5f)*800.0f;
		const float e = -((float)(h >> 8)/16777216.0f)*87.0f;
		float s, c;
		pwn_sincosf(x, &s, &c, v);
		if(f2u(s) != f2u(sinf(x)) || f2u(c) != f2u(cosf(x)) || f2u(pwn_expf(e, tab, v)) != f2u(expf(e))) ok = 0;
	}
	if(fma_build) *fma_build = v;
	if(exact) *exact = ok;
	return PWN_OK;
}

/* test hook (host only): the restatement of pwn_libm.h over an array, build `fma_build` */
void pwn_host_libm_eval(const float *in, float *s, float *c, float *e, size_t n, int fma_build)
{
	static const uint64_t tab[32] = PWN_EXP2F_TAB;
	size_t i;
	for(i = 0; i < n; i++)
	{
		pwn_sincosf(in[i], &s[i], &c[i], fma_build);
		e[i] = pwn_expf(in[i], tab, fma_build);
	}
}
