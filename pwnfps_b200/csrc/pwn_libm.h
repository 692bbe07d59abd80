/*
 * pwn_libm.h -- sinf / cosf / sincosf / expf exactly as the reference host computes them.
 *
 * The reference calls libm in two places on the hot path: the wavy floor normal
 * (sinf, cosf, sincosf -- trace.h:42-46) and the fog blend (expf -- trace.h:97).  libm is a
 * third-party dependency that is not in /root/reference: GNU libc 2.39 (Ubuntu
 * 2.39-0ubuntu8.x) on the build box and on the GPU box.  This header restates the published
 * algorithms of that version (sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,s_sincosf.c,
 * s_sincosf.h,e_expf.c}, which glibc took from ARM's optimized-routines): double-precision
 * polynomials after a table-free (< 120) or 192-bit (>= 120) range reduction, rounded once to
 * float at the end.
 *
 * x86-64 glibc picks one of two builds of each function at load time: the "_fma" build
 * (compiled with -mfma -mavx2, so a*b+c contracts to one rounding) when the CPU has FMA and
 * AVX2, else the "_sse2" build (every product rounded).  The two differ in the last double
 * bit and, on a handful of inputs, in the final float.  Both are implemented here; the order
 * of operations of the _fma build is the one found by disassembling libm.so.6 (noted per
 * statement).  pwn_libm_probe() (pwn_hostmath.c) asks the host libm which one it is.
 *
 * One text, three users: the CUDA kernel (device), pwn_hostmath.c (host probe) and the CPU
 * tests, which check it against the host's glibc over every float the trace can produce.
 */
#ifndef PWN_LIBM_H
#define PWN_LIBM_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PWN_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define PWN_HD static inline
#endif

/* one rounding (the _fma build) or two (the _sse2 build) */
#if defined(__CUDA_ARCH__)
#define PWN_MULADD(FMA, a, b, c) ((FMA) ? __fma_rn((a), (b), (c)) : __dadd_rn(__dmul_rn((a), (b)), (c)))
#define PWN_DMUL(a, b) __dmul_rn((a), (b))
#define PWN_DSUB(a, b) __dsub_rn((a), (b))
#define PWN_DADD(a, b) __dadd_rn((a), (b))
#else
#define PWN_MULADD(FMA, a, b, c) ((FMA) ? fma((a), (b), (c)) : ((a)*(b) + (c)))
#define PWN_DMUL(a, b) ((a)*(b))
#define PWN_DSUB(a, b) ((a) - (b))
#define PWN_DADD(a, b) ((a) + (b))
#endif

PWN_HD uint32_t pwn_asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
PWN_HD uint64_t pwn_asuint64(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
PWN_HD double pwn_asdouble(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

/* __sincosf_table[0] (s_sincosf_data.c); table[1] is the same with c0..c4 negated */
#define PWN_SC_HPI_INV 0x1.45F306DC9C883p+23   /* 2/pi * 2^24 */
#define PWN_SC_HPI 0x1.921FB54442D18p0
#define PWN_SC_C0 0x1p0
#define PWN_SC_C1 -0x1.ffffffd0c621cp-2
#define PWN_SC_C2 0x1.55553e1068f19p-5
#define PWN_SC_C3 -0x1.6c087e89a359dp-10
#define PWN_SC_C4 0x1.99343027bf8c3p-16
#define PWN_SC_S1 -0x1.555545995a603p-3
#define PWN_SC_S2 0x1.1107605230bc4p-7
#define PWN_SC_S3 -0x1.994eb3774cf24p-13
#define PWN_SC_PI63 0x1.921FB54442D18p-62

/* sinf_poly, sine branch (s_sincosf.h): s + x7*s1 */
PWN_HD double pwn_sin_poly(double x, double x2, int fma_)
{
	double x3 = PWN_DMUL(x, x2);
	double s1 = PWN_MULADD(fma_, x2, PWN_SC_S3, PWN_SC_S2);   /* p->s2 + x2*p->s3 */
	double x7 = PWN_DMUL(x3, x2);
	double s = PWN_MULADD(fma_, x3, PWN_SC_S1, x);            /* x + x3*p->s1 */
	return PWN_MULADD(fma_, s1, x7, s);                       /* s + x7*s1 */
}

/* sinf_poly, cosine branch with table[0]; table[1] negates every coefficient, hence the result */
PWN_HD double pwn_cos_poly(double x2, int fma_)
{
	double x4 = PWN_DMUL(x2, x2);
	double c2 = PWN_MULADD(fma_, x2, PWN_SC_C4, PWN_SC_C3);   /* p->c3 + x2*p->c4 */
	double c1 = PWN_MULADD(fma_, x2, PWN_SC_C1, PWN_SC_C0);   /* p->c0 + x2*p->c1 */
	double x6 = PWN_DMUL(x4, x2);
	double c = PWN_MULADD(fma_, x4, PWN_SC_C2, c1);           /* c1 + x4*p->c2 */
	return PWN_MULADD(fma_, c2, x6, c);                       /* c + x6*c2 */
}

/* __inv_pio4 (s_sincosf_data.c): 4/pi in 192 bits, 8 bits per step */
PWN_HD uint32_t pwn_inv_pio4(int i)
{
#if defined(__CUDA_ARCH__)
	/* selected arithmetically from three 64-bit words: no table in local or constant memory */
	const uint64_t w0 = 0xa2f9836e4e441529ull, w1 = 0xfc2757d1f534ddc0ull, w2 = 0xdb6295993c439041ull;
	/* entry i = bits [8i-24 .. 8i+8) of the 192-bit string w0|w1|w2 (entries 0..2 are partial) */
	int hi = 8*i + 8;            /* number of leading bits consumed */
	/* take the 32 bits ending at bit position hi (counted from the top) */
	uint64_t a, b;
	int sh;
	if(hi <= 64) { a = 0; b = w0; sh = 64 - hi; }
	else if(hi <= 128) { a = w0; b = w1; sh = 128 - hi; }
	else { a = w1; b = w2; sh = 192 - hi; }
	uint64_t v = sh == 0 ? b : ((b >> sh) | (a << (64 - sh)));
	return (uint32_t)v;
#else
	static const uint32_t t[24] = {
		0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
		0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
		0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041 };
	return t[i];
#endif
}

/* reduce_fast (s_sincosf.h, !TOINT_INTRINSICS): |x| < 120 */
PWN_HD double pwn_reduce_fast(double x, int *np, int fma_)
{
	double r = PWN_DMUL(x, PWN_SC_HPI_INV);
	int n = ((int32_t)r + 0x800000) >> 24;
	*np = n;
	/* x - n*hpi: vfnmadd in the _fma build */
	return PWN_MULADD(fma_, -(double)n, PWN_SC_HPI, x);
}

/* reduce_large (s_sincosf.h): 120 <= |x| < inf */
PWN_HD double pwn_reduce_large(uint32_t xi, int *np)
{
	const int i0 = (int)((xi >> 26) & 15);
	int shift = (int)((xi >> 23) & 7);
	uint64_t n, res0, res1, res2;
	xi = (xi & 0xffffff) | 0x800000;
	xi <<= shift;
	res0 = (uint32_t)(xi * pwn_inv_pio4(i0));
	res1 = (uint64_t)xi * pwn_inv_pio4(i0 + 4);
	res2 = (uint64_t)xi * pwn_inv_pio4(i0 + 8);
	res0 = (res2 >> 32) | (res0 << 32);
	res0 += res1;
	n = (res0 + (1ULL << 61)) >> 62;
	res0 -= n << 62;
	double x = (double)(int64_t)res0;
	*np = (int)n;
	return PWN_DMUL(x, PWN_SC_PI63);
}

#define PWN_ABSTOP12(u) (((u) >> 20) & 0x7ff)

/*
 * sin and cos of y together.  glibc's sinf(), cosf() and sincosf() are three entry points to
 * this one computation (s_sinf.c, s_cosf.c, s_sincosf.c): the quadrant n picks which polynomial
 * is the sine and which the cosine (n & 1), p->sign[q & 3] = {1,-1,-1,1} flips the sine
 * polynomial's argument, q & 2 selects the negated cosine table; q = n, or n + signbit(y) after
 * the large reduction.  (Checked against the _fma builds' disassembly for all three.)
 */
PWN_HD void pwn_sincosf(float y, float *sinp, float *cosp, int fma_)
{
	const uint32_t iy = pwn_asuint(y);
	const uint32_t top = PWN_ABSTOP12(iy);
	double x = (double)y;
	if(top < 0x3f4)   /* |y| < pi/4 */
	{
		if(top < 0x398)   /* |y| < 2^-12 */
		{
			*sinp = y;
			*cosp = 1.0f;
			return;
		}
		const double x2 = PWN_DMUL(x, x);
		*sinp = (float)pwn_sin_poly(x, x2, fma_);
		*cosp = (float)pwn_cos_poly(x2, fma_);
		return;
	}
	if(top >= 0x7f8)   /* inf / nan: __math_invalidf */
	{
		*sinp = *cosp = y - y;
		return;
	}
	int n, q;
	if(top < 0x42f)    /* |y| < 120 */
	{
		x = pwn_reduce_fast(x, &n, fma_);
		q = n;
	}
	else
	{
		x = pwn_reduce_large(iy, &n);
		q = n + (int)(iy >> 31);
	}
	const double sgn = ((q & 3) == 1 || (q & 3) == 2) ? -1.0 : 1.0;
	const double x2 = PWN_DMUL(x, x);
	const double s = pwn_sin_poly(PWN_DMUL(x, sgn), x2, fma_);
	double c = pwn_cos_poly(x2, fma_);
	if(q & 2) c = -c;   /* table[1]: every cosine coefficient negated, hence the result */
	if(n & 1) { *sinp = (float)c; *cosp = (float)s; }
	else { *sinp = (float)s; *cosp = (float)c; }
}

PWN_HD float pwn_sinf(float y, int fma_)
{
	float s, c;
	pwn_sincosf(y, &s, &c, fma_);
	return s;
}

PWN_HD float pwn_cosf(float y, int fma_)
{
	float s, c;
	pwn_sincosf(y, &s, &c, fma_);
	return c;
}

/* __exp2f_data.tab[i] = asuint64(2^(i/32)) - (i << 47)   (e_exp2f_data.c) */
#define PWN_EXP2F_TAB { \
	0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
	0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
	0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
	0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
	0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
	0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
	0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
	0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull }

#define PWN_EXP_SHIFT 0x1.8p+52
#define PWN_EXP_INVLN2N 0x1.71547652b82fep+5      /* 32/ln2 */
#define PWN_EXP_C0 0x1.c6af84b912394p-20           /* poly_scaled */
#define PWN_EXP_C1 0x1.ebfce50fac4f3p-13
#define PWN_EXP_C2 0x1.62e42ff0c52d6p-6

/*
 * expf (e_expf.c, !TOINT_INTRINSICS).  `tab` is PWN_EXP2F_TAB wherever the caller keeps it.
 * Valid for x <= 0 (the trace passes -0.6*fog); below -0x1.9fe368p6 the result underflows to 0.
 */
PWN_HD float pwn_expf(float x, const uint64_t *tab, int fma_)
{
	const uint32_t top = PWN_ABSTOP12(pwn_asuint(x));
	const double xd = (double)x;
	if(!(fabsf(x) < 88.0f))   /* top >= 0x42b: |x| >= 88 or NaN, as one float compare */
	{
		if(pwn_asuint(x) == 0xff800000u) return 0.0f;
		if(top >= 0x7f8) return x + x;
		if(x > 0x1.62e42ep6f) return x*0x1p127f;     /* overflow (not reachable from the trace) */
		if(x < -0x1.9fe368p6f) return 0.0f;          /* __math_uflowf */
		/* -103.97 .. -88: glibc falls through to the main path (denormal results) */
	}
	double kd, r;
	if(fma_)
	{
		kd = fma(PWN_EXP_INVLN2N, xd, PWN_EXP_SHIFT);        /* z + SHIFT, z never rounded */
		const double k0 = PWN_DSUB(kd, PWN_EXP_SHIFT);
		r = fma(PWN_EXP_INVLN2N, xd, -k0);                    /* vfmsub: z - kd */
	}
	else
	{
		const double z = PWN_DMUL(PWN_EXP_INVLN2N, xd);
		kd = PWN_DADD(z, PWN_EXP_SHIFT);
		r = PWN_DSUB(z, PWN_DSUB(kd, PWN_EXP_SHIFT));
	}
	const uint64_t ki = pwn_asuint64(kd);
	uint64_t t = tab[ki & 31];
	t += ki << 47;
	const double s = pwn_asdouble(t);
	const double z2 = PWN_MULADD(fma_, PWN_EXP_C0, r, PWN_EXP_C1);
	const double r2 = PWN_DMUL(r, r);
	double yv = PWN_MULADD(fma_, PWN_EXP_C2, r, 1.0);
	yv = PWN_MULADD(fma_, z2, r2, yv);
	yv = PWN_DMUL(yv, s);
	return (float)yv;
}

#endif /* PWN_LIBM_H */
