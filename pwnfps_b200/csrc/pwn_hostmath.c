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

#include "pwntrace.h"

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
