/*
 * oracle/pwn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain scalar C, IEEE source-order arithmetic, no FMA) of the
 * pwnfps per-pixel portal ray cast: ray setup, cell walk, portal crossing, sphere
 * test, shading, two mirror bounces, colour pack and the focal-blur pass.  Each
 * function cites the reference file:line it follows.  Unlike the reference it is
 * not limited to a 64x64 grid / 26 portals, and it reports per-segment hit ids.
 *
 * Pinning (DESIGN.md "Oracle"): the reference ships no tests or golden vectors
 * (SURVEY 4), so this restatement is pinned against the reference's own code
 * compiled headless here (oracle/_ref, built by oracle/make_ref.py): bit-exact
 * frames / zbuf / ids on level.txt and the synthetic levels (tests/test_oracle_*)
 * and against the committed fixtures in tests/golden/ generated from that build.
 *
 * The only non-portable arithmetic is what the reference itself delegates to the
 * host CPU: rcpps (trace.h:231), rsqrtps (util.h:43), cvtps2dq (util.h:51) and
 * glibc sinf/cosf/expf (trace.h:43-46,97).  They are used here as-is.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link,
 * load or call this file.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off,
 * deliberately NOT -ffast-math; the foldings gcc applies to the reference under
 * the shipped -ffast-math are mirrored explicitly under ORC_MIRROR_FASTMATH).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <xmmintrin.h>
#include <emmintrin.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef ORC_MIRROR_FASTMATH
#define ORC_MIRROR_FASTMATH 1
#endif

/* defs.h:1-19 */
#define ORC_EPSILON 0.0000000000001f
#define ORC_REFLECT_BLUR 0.03f
#define ORC_REFLECT 2

/* defs.h:25-33 */
enum { FXP = 0, FZP, FXN, FZN, FYP, FYN };

typedef struct { float a[4]; } v4;

typedef struct
{
	int32_t x1, z1, x2, z2, rot12, c1, c2;  /* defs.h:87-94 (double1/2 unused) */
} orc_portal;

typedef struct
{
	float r, refl;      /* defs.h:73-79 */
	float pos[4];
	float col[4];       /* b g r a */
} orc_sphere;

typedef struct
{
	int32_t cell;     /* cz*W + cx at termination; -1 never terminated */
	int16_t ldir;     /* FXP..FYN, -1 sphere */
	int16_t kind;     /* 0 trace_hit_wall, 1 sphere (trace_hit_bounce), 2 exhausted */
	int32_t nstep;
	int32_t nportal;
} orc_ids;

typedef struct
{
	int w, h;               /* grid dims (reference: 64 x 64, defs.h:105) */
	int32_t *cells;         /* [h][w] cell codes: reference chars, or 0x100+p wide portal */
	int nport;
	orc_portal *pm;
	int nsph;
	orc_sphere *sph;
	/* per-cell sphere lists, level.h:1-19,64-81 */
	int *parts_off;         /* [h*w+1] */
	int *parts_idx;
	float sec_current;      /* defs.h:23 */
	int reflect;            /* REFLECT, defs.h:8 */
	int reflect_blur;       /* REFLECT_BLUR defined? defs.h:6 */
} orc_level;

static orc_level g_lv = {0, 0, NULL, 0, NULL, 0, NULL, NULL, NULL, 0.0f, ORC_REFLECT, 1};

#define FOGACC(f, c, b) ((f) + ((c) - (b)))

#define IS_PORTAL(c) (((c) >= 'A' && (c) <= 'Z') || (c) >= 0x100)
#define PIDX(c) ((c) >= 0x100 ? (c) - 0x100 : (c) - 'A')

/* ---- host-CPU primitives the reference delegates to ------------------------- */
static inline float cpu_rcp(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }
static inline float cpu_rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }
static inline int32_t cpu_cvt_rn(float x) { return _mm_cvtss_si32(_mm_set_ss(x)); }   /* cvtss2si: RN-even, 0x80000000 on NaN/overflow */
static inline int32_t cpu_cvt_tr(float x) { return _mm_cvttss_si32(_mm_set_ss(x)); }  /* (int)float as gcc emits it */

/* ---- util.h:1-16 --------------------------------------------------------------- */
static inline uint32_t orc_randi(uint32_t *seed)
{
	*seed = (*seed * 25739u) + 4u;
	*seed &= 0x7FFFFFFFu;
	return *seed;
}

static inline float orc_randfs(uint32_t *seed)
{
	uint32_t r = orc_randi(seed) % 3759u;
#if ORC_MIRROR_FASTMATH
	/* gcc -ffast-math folds (r/3759.0f)*2.0f-1.0f into r*fl(2/3759)-1 */
	return (float)r * (2.0f/3759.0f) - 1.0f;
#else
	return ((float)r / 3759.0f) * 2.0f - 1.0f;
#endif
}

/* util.h:18-30: (x*x' + z*z') + (y*y' + w*w') */
static inline float orc_dot(const float *a, const float *b)
{
	float m0 = a[0]*b[0], m1 = a[1]*b[1], m2 = a[2]*b[2], m3 = a[3]*b[3];
	return (m0 + m2) + (m1 + m3);
}

/* util.h:32-46: v * rsqrtps(v.v) */
static inline void orc_normalise(float *v)
{
	float s = cpu_rsqrt(orc_dot(v, v));
	v[0] *= s; v[1] *= s; v[2] *= s; v[3] *= s;
}

/* util.h:48-59 */
static inline uint32_t orc_col_ftoint(const float *v)
{
	uint32_t ret = 0;
	int i;
	for(i = 0; i < 4; i++)
	{
		int32_t q = cpu_cvt_rn(v[i] * 255.0f);
		if(q < -32768) q = -32768;        /* packssdw */
		if(q > 32767) q = 32767;
		if(q < 0) q = 0;                  /* packuswb */
		if(q > 255) q = 255;
		ret |= ((uint32_t)q) << (8*i);
	}
	return ret;
}

/* util.h:151-158, generalised from 64 to (w,h) */
static inline int32_t orc_get_cell(const orc_level *lv, int cx, int cz)
{
	if(cx < 0 || cx >= lv->w) cx = 0;
	if(cz < 0 || cz >= lv->h) cz = 0;
	return lv->cells[cz*lv->w + cx];
}

/* ---- trace.h ----------------------------------------------------------------------- */
typedef struct
{
	orc_ids *ids;   /* 3 records for this pixel, or NULL */
} orc_pix;

static void orc_trace_ray(const orc_level *lv, int hitctr, uint32_t *seed, float *dist,
	const float *ifrom, const float *iray, const float *icol, float *out, orc_pix *px);

#define ORC_REC(k, ld) do { if(px->ids != NULL && hitctr >= 0 && hitctr < 3) { \
	orc_ids *o_ = &px->ids[hitctr]; o_->cell = cz*lv->w + cx; o_->ldir = (int16_t)(ld); \
	o_->kind = (k); o_->nstep = nstep; o_->nportal = nport; } } while(0)

/* trace.h:3-106 */
static void orc_hit_bounce(const orc_level *lv, int hitctr, uint32_t *seed, const float *iray, const float *ipos,
	const float *inorm, int ldir, float refl, float *dist, float fog, const float *icolv, float *out, orc_pix *px)
{
	float ray[4], pos[4], col[4], bcol[4];
	int i;

	memcpy(col, icolv, sizeof(col));
	if(hitctr < 0 || hitctr >= lv->reflect || refl == 0.0f)   /* trace.h:6-7 */
	{
		memcpy(out, col, sizeof(col));
		return;
	}

	memcpy(ray, iray, sizeof(ray));
	memcpy(pos, ipos, sizeof(pos));
	switch(ldir)   /* trace.h:12-75 */
	{
		case FXP: ray[0] = -ray[0]; pos[0] -= 0.001f; break;
		case FXN: ray[0] = -ray[0]; pos[0] += 0.001f; break;
		case FZP: ray[2] = -ray[2]; pos[2] -= 0.001f; break;
		case FZN: ray[2] = -ray[2]; pos[2] += 0.001f; break;
		case FYP: ray[1] = -ray[1]; pos[1] -= 0.001f; break;

		case FYN: {   /* wavy floor, trace.h:39-56 */
			float norm[4];
			pos[1] -= 0.001f;
			float ang = (((float)M_PI)*2.0f)*(
				sinf((((float)M_PI)*0.5f)*pos[0])
				+ cosf((((float)M_PI)*0.5f)*pos[2])
				+ lv->sec_current*1.0f);
#ifdef ORC_EXP_SINCOS
			sincosf(ang, &norm[0], &norm[2]); norm[1] = 38.0f; norm[3] = 0.0f;
#else
			norm[0] = sinf(ang); norm[1] = 38.0f; norm[2] = cosf(ang); norm[3] = 0.0f;
#endif
			orc_normalise(norm);
			/* shipped-flags build keeps (a+b)+c: addss/addss in trace_hit_bounce of oracle/_ref */
			float rmul = -2.0f * (0.0f + ray[0]*norm[0] + ray[1]*norm[1] + ray[2]*norm[2]);
			for(i = 0; i < 4; i++) ray[i] = rmul*norm[i] + ray[i];
			orc_normalise(ray);
		} break;

		default: {    /* sphere, trace.h:58-73 */
			float norm[4];
			for(i = 0; i < 4; i++) pos[i] = pos[i] - 0.001f*ray[i];
			memcpy(norm, inorm, sizeof(norm));
			float rmul = -2.0f * (0.0f + ray[0]*norm[0] + ray[1]*norm[1] + ray[2]*norm[2]);
			for(i = 0; i < 4; i++) ray[i] = rmul*norm[i] + ray[i];
			orc_normalise(ray);
		} break;
	}

	if(lv->reflect_blur)   /* trace.h:77-84: 5 draws, x y - z - */
	{
		const float rblur = ORC_REFLECT_BLUR;
		ray[0] += orc_randfs(seed) * rblur;
		ray[1] += orc_randfs(seed) * rblur;
		orc_randfs(seed);
		ray[2] += orc_randfs(seed) * rblur;
		orc_randfs(seed);
	}

	float odist = *dist;   /* trace.h:86-87 */
	*dist = 0;

	memcpy(bcol, col, sizeof(col));
	orc_trace_ray(lv, hitctr+1, seed, dist, pos, ray, bcol, col, px);   /* trace.h:90 */
	for(i = 0; i < 4; i++)
		col[i] = refl*col[i] + (1.0f-refl)*bcol[i];

	if(fog != 0.0f)   /* trace.h:95-101 */
	{
		fog = expf(-0.6f*fog);
		for(i = 0; i < 4; i++)
			col[i] = fog*col[i] + (1.0f-fog)*1.0f;
	}

	*dist = odist;
	memcpy(out, col, sizeof(col));
}

/* trace.h:108-154 */
static void orc_hit_wall(const orc_level *lv, int hitctr, uint32_t *seed, const float *ipos, const float *iray,
	int ldir, float *dist, const float *icol, float fog, const float *scol, float *out, orc_pix *px)
{
	float diffuse, col[4];
	int i;

	for(i = 0; i < 4; i++) col[i] = icol[i]*scol[i];

	switch(ldir)
	{
		default:
		case FXP: diffuse = iray[0]; break;
		case FYP: diffuse = iray[1]; break;
		case FZP: diffuse = iray[2]; break;
		case FXN: diffuse = -iray[0]; break;
		case FYN: diffuse = -iray[1]; break;
		case FZN: diffuse = -iray[2]; break;
	}
	if(diffuse < 0.0f) diffuse = 0.0f;

	const float amb = 0.1f;
	diffuse = (1.0f - amb)*diffuse + amb;
	for(i = 0; i < 4; i++) col[i] *= diffuse;

	orc_hit_bounce(lv, hitctr, seed, iray, ipos, NULL, ldir,
		(ldir == FYN ? 0.7f : 0.25f), dist, fog, col, out, px);
}

/* trace.h:156-184 */
static inline void orc_through(int *ldir, float *cdist, const float *wdist, float *pos, const float *ray, int gx, int gy, int gz)
{
	float t;
	if(wdist[1] < wdist[0] && wdist[1] < wdist[2]) { t = wdist[1]; *ldir = (gy < 0 ? FYN : FYP); }
	else if(wdist[0] < wdist[2]) { t = wdist[0]; *ldir = (gx < 0 ? FXN : FXP); }
	else { t = wdist[2]; *ldir = (gz < 0 ? FZN : FZP); }
	*cdist += t;
	pos[0] = t*ray[0] + pos[0];
	pos[1] = t*ray[1] + pos[1];
	pos[2] = t*ray[2] + pos[2];
	pos[3] = t*ray[3] + pos[3];
}

static const float COL_CEIL[4]  = {30.0f, 30.0f, 0.0f, 0.0f};   /* defs.h:17-19, lanes b g r a */
static const float COL_FLOOR[4] = {1.0f, 1.0f, 1.0f, 0.0f};
static const float COL_WALL[4]  = {0.8f, 0.8f, 1.0f, 0.0f};
static const float COL_BADPORTAL[4] = {5.0f, 0.0f, 5.0f, 0.0f}; /* trace.h:558 */

/* trace.h:186-679 */
static void orc_trace_ray(const orc_level *lv, int hitctr, uint32_t *seed, float *dist,
	const float *ifrom, const float *iray, const float *icol, float *out, orc_pix *px)
{
	int i;
	float ray[4], pos[4], iavel[4], wdist[4];
	int gx, gy, gz;
	float cdist = 0.0f, fog = 0.0f, fogbeg = 0.0f;
	int nstep = 0, nport = 0;

	float aux_dist = -1.0f, aux_refl = 0.25f;
	const int aux_dir = -1;
	float aux_pos[4] = {0, 0, 0, 0}, aux_norm[4] = {0, 0, 0, 0}, aux_col[4] = {1.0f, 1.0f, 1.0f, 1.0f};

	memcpy(ray, iray, sizeof(ray));
	memcpy(pos, ifrom, sizeof(pos));
	orc_normalise(ray);   /* trace.h:212 */

	int cx = cpu_cvt_tr(ifrom[0]);   /* trace.h:215-216 */
	int cz = cpu_cvt_tr(ifrom[2]);

	for(i = 0; i < 3; i++)   /* trace.h:220-222 */
		if(ray[i] > -ORC_EPSILON && ray[i] < ORC_EPSILON) ray[i] = (ray[i] < 0.0f ? -ORC_EPSILON : ORC_EPSILON);

	gx = (iray[0] < 0 ? -1 : 1);   /* trace.h:225-227: signs of the un-normalised input */
	gy = (iray[1] < 0 ? -1 : 1);
	gz = (iray[2] < 0 ? -1 : 1);

	for(i = 0; i < 4; i++) iavel[i] = cpu_rcp(fabsf(ray[i]));   /* trace.h:230-231 */

	wdist[0] = pos[0] - (float)cx;   /* trace.h:235-241 */
	wdist[1] = pos[1] - 0.0f;
	wdist[2] = pos[2] - (float)cz;
	wdist[3] = pos[3] - 0.0f;
	if(ray[0] >= 0.0f) wdist[0] = 1.0f - wdist[0];
	if(ray[1] >= 0.0f) wdist[1] = 1.0f - wdist[1];
	if(ray[2] >= 0.0f) wdist[2] = 1.0f - wdist[2];
	for(i = 0; i < 4; i++) wdist[i] *= iavel[i];

	int32_t cell = orc_get_cell(lv, cx, cz);
	int maxsteps = 1000;
	int ldir = FYN;

#define SPHERE_CHECK(FOGCELL) \
	if(aux_dist != -1.0f && cdist > aux_dist) { \
		*dist = aux_dist; \
		if((FOGCELL) && aux_dist > fogbeg) fog += aux_dist - fogbeg; \
		ORC_REC(1, -1); \
		orc_hit_bounce(lv, hitctr, seed, ray, aux_pos, aux_norm, aux_dir, aux_refl, dist, fog, aux_col, out, px); \
		return; }
#define WALL_HIT(COLV) do { ORC_REC(0, ldir); \
	orc_hit_wall(lv, hitctr, seed, pos, ray, ldir, dist, icol, fog, (COLV), out, px); return; } while(0)
#define STEP_XZ() \
	if(ldir == FXN || ldir == FXP) { \
		float t_ = wdist[0]; wdist[0] -= t_; wdist[1] -= t_; wdist[2] -= t_; wdist[3] -= t_; \
		wdist[0] = iavel[0]; cx += gx; \
	} else if(ldir == FZN || ldir == FZP) { \
		float t_ = wdist[2]; wdist[0] -= t_; wdist[1] -= t_; wdist[2] -= t_; wdist[3] -= t_; \
		wdist[2] = iavel[2]; cz += gz; }

	while(maxsteps-- > 0)
	{
		nstep++;

		/* spheres of this cell, trace.h:252-296 */
		if(cx >= 0 && cx < lv->w && cz >= 0 && cz < lv->h && lv->parts_off != NULL)
		for(i = lv->parts_off[cz*lv->w + cx]; i < lv->parts_off[cz*lv->w + cx + 1]; i++)
		{
			const orc_sphere *pt = &lv->sph[lv->parts_idx[i]];
			float sph_rad2 = pt->r * pt->r;
			float sph_dist;
			float sp[4];
			int k;
			for(k = 0; k < 4; k++) sp[k] = pt->pos[k] - pos[k];
			float sph_dist2 = orc_dot(sp, sp);
			float sph_dot = orc_dot(sp, ray);
			if(sph_dot > 0.0f)
			{
				float sph_calcrad2 = (sph_dist2 - sph_dot*sph_dot);
				if(sph_calcrad2 < sph_rad2)
				{
					float sph_sdist2 = 1.0f - sph_calcrad2/sph_rad2;
					sph_dist = sqrtf(sph_dist2) - sqrtf(sph_sdist2);
					if(aux_dist == -1.0f || sph_dist + cdist < aux_dist)
					{
						aux_dist = sph_dist + cdist;
						for(k = 0; k < 4; k++) aux_pos[k] = pos[k] + sph_dist*ray[k];
						for(k = 0; k < 4; k++) aux_norm[k] = aux_pos[k] - pt->pos[k];
						orc_normalise(aux_norm);
						float diff = -orc_dot(ray, aux_norm);
						if(diff < 0.0f) diff = 0.0f;
						float amb = 0.2f;
						aux_refl = pt->refl;
						diff = amb + (1.0f - amb)*diff;
						for(k = 0; k < 4; k++) aux_col[k] = diff*pt->col[k];
					}
				}
			}
		}

		int32_t this_cell = cell;
		switch(this_cell)
		{
			case ';': case '$': case '"':   /* trace.h:302-352 */
				if(this_cell == '$') fogbeg = cdist;
				orc_through(&ldir, &cdist, wdist, pos, ray, gx, gy, gz);
				SPHERE_CHECK(this_cell == '$')
				if(this_cell == '$') fog = FOGACC(fog, cdist, fogbeg);
				if(ldir == FYN || ldir == FYP)
				{
					*dist = cdist;
					WALL_HIT(gy > 0 ? COL_CEIL : COL_FLOOR);
				}
				STEP_XZ()
				cell = orc_get_cell(lv, cx, cz);
				if(this_cell == '"' && (cell == '#' || cell == '&'))
				{
					pos[1] += 1.0f;
					if(gy < 0) wdist[1] += iavel[1];
					else wdist[1] -= iavel[1];
				}
				break;

			case '#': case '&':   /* trace.h:354-441 */
				if(gy > 0) wdist[1] += iavel[1];
				if(this_cell == '&') fogbeg = cdist;
				orc_through(&ldir, &cdist, wdist, pos, ray, gx, gy, gz);
				SPHERE_CHECK(this_cell == '&')
				if(this_cell == '&') fog = FOGACC(fog, cdist, fogbeg);
				if(ldir == FYN || ldir == FYP)
				{
					*dist = cdist;
					WALL_HIT(gy > 0 ? COL_CEIL : COL_FLOOR);
				}
				STEP_XZ()
				if(gy > 0) wdist[1] -= iavel[1];
				cell = orc_get_cell(lv, cx, cz);
				if(cell == '"')
				{
					pos[1] -= 1.0f;
					if(gy > 0) wdist[1] += iavel[1];
					else wdist[1] -= iavel[1];
				}
				{
					int32_t xcell = cell;   /* peek through a portal, trace.h:404-413 */
					if(IS_PORTAL(xcell))
					{
						const orc_portal *pm = &lv->pm[PIDX(xcell)];
						if(pm->x1 == cx && pm->z1 == cz) xcell = pm->c2;
						else if(pm->x2 == cx && pm->z2 == cz) xcell = pm->c1;
					}
					if(pos[1] < 0.0f || pos[1] > 1.0f)
					switch(xcell)
					{
						case '#': case '&':
							break;
						case '"':   /* revert, trace.h:422-434 */
							pos[1] += 1.0f;
							if(gy > 0) wdist[1] -= iavel[1];
							else wdist[1] += iavel[1];
							*dist = cdist;
							WALL_HIT(COL_WALL);
						default:
							*dist = cdist;
							WALL_HIT(COL_WALL);
					}
				}
				break;

			case '>': case '<': case ',': case '^': {   /* trace.h:443-505 */
				const float ramp_delta = 0.5f;
				switch(this_cell)
				{
					case '>': ray[1] -= ramp_delta * ray[0]; break;
					case '<': ray[1] += ramp_delta * ray[0]; break;
					case ',': ray[1] -= ramp_delta * ray[2]; break;
					case '^': ray[1] += ramp_delta * ray[2]; break;
				}
				wdist[1] = pos[1];
				if(ray[1] >= 0.0f) wdist[1] = 1.0f - wdist[1];
				/* reciprocal then multiply, also in the shipped -ffast-math build (zbuf bit-exact);
				   SURVEY N9's single-divss folding does not happen in the inlined code */
				wdist[1] *= 1.0f / (ray[1] < 0.0f ? -ray[1] : ray[1]);
				if(aux_dist != -1.0f && cdist > aux_dist)   /* trace.h:463-468: no fog term */
				{
					*dist = aux_dist;
					ORC_REC(1, -1);
					orc_hit_bounce(lv, hitctr, seed, ray, aux_pos, aux_norm, aux_dir, aux_refl, dist, fog, aux_col, out, px);
					return;
				}
				orc_through(&ldir, &cdist, wdist, pos, ray, gy, gy, gz);   /* sic: gy as gx, trace.h:470 */
				if(ldir == FYN || ldir == FYP)
				{
					*dist = cdist;
					ldir = (ray[1] < 0.0f ? FYN : FYP);
					WALL_HIT(ray[1] >= 0.0f ? COL_CEIL : COL_FLOOR);
				}
				else if(ldir == FXN || ldir == FXP)
				{
					ldir = (ray[0] < 0.0f ? FXN : FXP);
					float t_ = wdist[0]; wdist[0] -= t_; wdist[1] -= t_; wdist[2] -= t_; wdist[3] -= t_;
					wdist[0] = iavel[0]; cx += gx;
				}
				else if(ldir == FZN || ldir == FZP)
				{
					ldir = (ray[2] < 0.0f ? FZN : FZP);
					float t_ = wdist[2]; wdist[0] -= t_; wdist[1] -= t_; wdist[2] -= t_; wdist[3] -= t_;
					wdist[2] = iavel[2]; cz += gz;
				}
				switch(this_cell)
				{
					case '>': ray[1] += ramp_delta * ray[0]; break;
					case '<': ray[1] -= ramp_delta * ray[0]; break;
					case ',': ray[1] += ramp_delta * ray[2]; break;
					case '^': ray[1] -= ramp_delta * ray[2]; break;
				}
				wdist[1] = pos[1];
				if(ray[1] >= 0.0f) wdist[1] = 1.0f - wdist[1];
				wdist[1] *= iavel[1];
				cell = orc_get_cell(lv, cx, cz);
			} break;

			default:
			if(IS_PORTAL(cell))   /* trace.h:508-650 */
			{
				const orc_portal *pm = &lv->pm[PIDX(cell)];
				if(pm->x2 == -1)
				{
					*dist = cdist;
					SPHERE_CHECK(0)
					WALL_HIT(COL_WALL);
				}

				int rot = 0;
				if(pm->x1 == cx && pm->z1 == cz)
				{
					cx = pm->x2; cz = pm->z2;
					pos[0] += (float)(pm->x2 - pm->x1);
					pos[2] += (float)(pm->z2 - pm->z1);
					rot = (-pm->rot12) & 3;
				}
				else if(pm->x2 == cx && pm->z2 == cz)
				{
					cx = pm->x1; cz = pm->z1;
					pos[0] -= (float)(pm->x2 - pm->x1);
					pos[2] -= (float)(pm->z2 - pm->z1);
					rot = (pm->rot12) & 3;
				}
				else
				{
					*dist = cdist;
					SPHERE_CHECK(0)
					WALL_HIT(COL_BADPORTAL);
				}

				float rx = pos[0], rz = pos[2];
				int tgx = gx, tgz = gz;
				float rvzx = ray[0], rvzz = ray[2];
				float trx = rx, trz = rz, trvzx = rvzx, trvzz = rvzz;
				float twdx = wdist[0], twdz = wdist[2], tiax = iavel[0], tiaz = iavel[2];

				ldir = (ldir - rot) & 3;
				nport++;

				switch(rot)   /* trace.h:578-617 */
				{
					case 0: break;
					case 1:
						rx = (cx + 0.5f) + (trz-(cz + 0.5f));
						rz = (cz + 0.5f) - (trx-(cx + 0.5f));
						rvzx =  trvzz; rvzz = -trvzx;
						gx =  tgz; gz = -tgx;
						wdist[0] = twdz; wdist[2] = twdx;
						iavel[0] = tiaz; iavel[2] = tiax;
						break;
					case 2:
						rx = (cx + 0.5f)*2.0f - rx;
						rz = (cz + 0.5f)*2.0f - rz;
						rvzx = -trvzx; rvzz = -trvzz;
						gx = -gx; gz = -gz;
						break;
					case 3:
						rx = (cx + 0.5f) - (trz-(cz + 0.5f));
						rz = (cz + 0.5f) + (trx-(cx + 0.5f));
						rvzx = -trvzz; rvzz =  trvzx;
						gx = -tgz; gz =  tgx;
						wdist[0] = twdz; wdist[2] = twdx;
						iavel[0] = tiaz; iavel[2] = tiax;
						break;
				}
				pos[0] = rx; pos[2] = rz;
				ray[0] = rvzx; ray[2] = rvzz;

				switch(ldir)   /* trace.h:624-647 */
				{
					default:
					case FXP: cx++; pos[0] += 1.0f; break;
					case FZP: cz++; pos[2] += 1.0f; break;
					case FXN: cx--; pos[0] -= 1.0f; break;
					case FZN: cz--; pos[2] -= 1.0f; break;
				}
				cell = orc_get_cell(lv, cx, cz);
				break;
			}
			else   /* solid, trace.h:651-664 */
			{
				*dist = cdist;
				SPHERE_CHECK(0)
				WALL_HIT(ldir == FYP ? COL_CEIL : COL_WALL);
			}
		}

		SPHERE_CHECK(0)   /* trace.h:668-673 */
	}

	ORC_REC(2, ldir);   /* trace.h:677-678: the ray itself comes back as the colour */
	memcpy(out, ray, sizeof(ray));
}

/* ---- level.h:1-39,64-81: bin spheres into cells ------------------------------------ */
static void orc_prepare_render(orc_level *lv)
{
	int i, x, z, n = lv->w*lv->h;
	free(lv->parts_off); free(lv->parts_idx);
	lv->parts_off = NULL; lv->parts_idx = NULL;
	if(lv->nsph == 0) return;
	int *cnt = calloc((size_t)n + 1, sizeof(int));
	int pass;
	for(pass = 0; pass < 2; pass++)
	{
		for(i = 0; i < lv->nsph; i++)
		{
			const orc_sphere *s = &lv->sph[i];
			int cx1 = cpu_cvt_tr(s->pos[0] - s->r), cz1 = cpu_cvt_tr(s->pos[2] - s->r);
			int cx2 = cpu_cvt_tr(s->pos[0] + s->r), cz2 = cpu_cvt_tr(s->pos[2] + s->r);
			for(z = cz1; z <= cz2; z++)
			for(x = cx1; x <= cx2; x++)
			{
				if(x < 0 || x >= lv->w || z < 0 || z >= lv->h) continue;   /* UB in the reference */
				if(pass == 0) cnt[z*lv->w + x + 1]++;
				else lv->parts_idx[cnt[z*lv->w + x]++] = i;
			}
		}
		if(pass == 0)
		{
			for(i = 0; i < n; i++) cnt[i+1] += cnt[i];
			lv->parts_off = malloc(((size_t)n + 1)*sizeof(int));
			memcpy(lv->parts_off, cnt, ((size_t)n + 1)*sizeof(int));
			lv->parts_idx = malloc((size_t)(cnt[n] > 0 ? cnt[n] : 1)*sizeof(int));
		}
	}
	free(cnt);
}

/* ---- exported test API -------------------------------------------------------------- */
int orc_level_set(int w, int h, const int32_t *cells, int nport, const int32_t *pm7)
{
	int i;
	if(w <= 0 || h <= 0) return -1;
	free(g_lv.cells); free(g_lv.pm);
	g_lv.w = w; g_lv.h = h;
	g_lv.cells = malloc((size_t)w*h*sizeof(int32_t));
	memcpy(g_lv.cells, cells, (size_t)w*h*sizeof(int32_t));
	g_lv.nport = nport;
	g_lv.pm = malloc((size_t)(nport > 0 ? nport : 1)*sizeof(orc_portal));
	for(i = 0; i < nport; i++)
	{
		g_lv.pm[i].x1 = pm7[i*7+0]; g_lv.pm[i].z1 = pm7[i*7+1];
		g_lv.pm[i].x2 = pm7[i*7+2]; g_lv.pm[i].z2 = pm7[i*7+3];
		g_lv.pm[i].rot12 = pm7[i*7+4]; g_lv.pm[i].c1 = pm7[i*7+5]; g_lv.pm[i].c2 = pm7[i*7+6];
	}
	/* every portal code in the grid must index the table */
	for(i = 0; i < w*h; i++)
		if(IS_PORTAL(cells[i]) && PIDX(cells[i]) >= nport) return -2;
	return 0;
}

/* sph: n x 10 floats: r refl x y z w b g r a */
int orc_spheres_set(int n, const float *sph)
{
	int i;
	free(g_lv.sph);
	g_lv.sph = NULL;
	g_lv.nsph = n;
	if(n > 0)
	{
		g_lv.sph = malloc((size_t)n*sizeof(orc_sphere));
		for(i = 0; i < n; i++)
		{
			g_lv.sph[i].r = sph[i*10+0]; g_lv.sph[i].refl = sph[i*10+1];
			memcpy(g_lv.sph[i].pos, &sph[i*10+2], 4*sizeof(float));
			memcpy(g_lv.sph[i].col, &sph[i*10+6], 4*sizeof(float));
		}
	}
	orc_prepare_render(&g_lv);
	return 0;
}

void orc_config(int reflect, int reflect_blur)
{
	g_lv.reflect = reflect;
	g_lv.reflect_blur = reflect_blur;
}

/* screen.h:31-57: per-frame ray basis.  out: rayb, rdx, rdy (4 floats each) */
void orc_ray_setup(const float *cam16, int dimx, int dimy, float *rayb, float *rdx, float *rdy)
{
	int i;
	float xrat = -1.0f;
	float yrat = -((float)dimy)/(float)dimx;
	float xsrat = 2.0f * xrat / (float)dimx;
	float ysrat = 2.0f * yrat / (float)dimy;
	const float *camx = cam16, *camy = cam16 + 4, *camz = cam16 + 8;
	for(i = 0; i < 4; i++)
	{
		rayb[i] = camz[i] + ((-xrat)*camx[i] + (-yrat)*camy[i]);
		rdx[i] = xsrat*camx[i];
		rdy[i] = ysrat*camy[i];
	}
}

/* screen.h:1-28 for one pixel */
static void orc_pixel(const orc_level *lv, const float *from, const float *rayb, const float *rdx, const float *rdy,
	int x, int y, int rwidth, uint32_t *sb, float *zb, orc_ids *ids)
{
	int i, k;
	float rayl[4], col[4];
	const float one[4] = {1.0f, 1.0f, 1.0f, 1.0f};
	int cx = x & ~31;
	for(i = 0; i < 4; i++)
	{
		rayl[i] = rayb[i] + ((float)y)*rdy[i];
		rayl[i] = rayl[i] + ((float)cx)*rdx[i];
	}
	for(k = 0; k <= (x & 31); k++)          /* rayl += rdx BEFORE each pixel, screen.h:18 */
		for(i = 0; i < 4; i++) rayl[i] = rayl[i] + rdx[i];
	uint32_t seed = (uint32_t)x + (uint32_t)y*(uint32_t)y*(uint32_t)(rwidth+1);   /* screen.h:19-21, wraps */
	seed *= seed * seed;
	seed *= seed * seed;
	orc_pix px;
	px.ids = ids;
	orc_trace_ray(lv, 0, &seed, zb, from, rayl, one, col, &px);
	*sb = orc_col_ftoint(col);
}

/*
 * Rows [y0,y1) of a dimx x dimy frame (the full-frame geometry decides the rays;
 * y0/y1 select the band a shard owns).  sbuf/zbuf/ids are FULL-frame arrays with
 * row stride rwidth (screen.h:14-15); untouched outside the band.
 */
int orc_trace_rows(const float *cam16, float sec, int dimx, int dimy, int rwidth, int y0, int y1,
	uint32_t *sbuf, float *zbuf, orc_ids *ids, int nthreads)
{
	float rayb[4], rdx[4], rdy[4];
	int y;
	if(g_lv.cells == NULL) return -1;
	g_lv.sec_current = sec;
	orc_ray_setup(cam16, dimx, dimy, rayb, rdx, rdy);
	const float *from = cam16 + 12;
#ifdef _OPENMP
	if(nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 4)
#endif
	for(y = y0; y < y1; y++)
	{
		int x;
		for(x = 0; x < dimx; x++)
		{
			size_t p = (size_t)y*rwidth + x;
			orc_pixel(&g_lv, from, rayb, rdx, rdy, x, y, rwidth, &sbuf[p], &zbuf[p], ids ? &ids[p*3] : NULL);
		}
	}
	return 0;
}

/* screen.h:69-123: focal blur, in-place on sbuf using zbuf; tsbuf is scratch (full frame) */
int orc_blur(int dimx, int dimy, int rwidth, int y0, int y1, uint32_t *sbuf, const float *zbuf, const uint32_t *tsbuf, int nthreads)
{
	int cy;
#ifdef _OPENMP
	if(nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for
#endif
	for(cy = y0; cy < y1; cy++)
	{
		int cx, i, j, b;
		uint32_t seed = (uint32_t)cy*(uint32_t)cy + 415135u;
		const float *dist = zbuf + (size_t)cy*rwidth;
		uint32_t *p = sbuf + (size_t)cy*rwidth;
		const float fstr = 0.002f * dimy;
		const float foffs = 1.0f;
		for(cx = 0; cx < dimx-3; cx += 4)
		{
			uint32_t vbuf[4][4];
			for(i = 0; i < 4; i++)
			for(j = 0; j < 4; j++)
			{
				float z = (dist[j] - foffs);
				int x = cpu_cvt_tr((float)(cx + j) + orc_randfs(&seed) * fstr * z);
				int y = cpu_cvt_tr((float)cy + orc_randfs(&seed) * fstr * z);
				if(x < 0) x = 0;
				if(y < 0) y = 0;
				if(x >= dimx) x = dimx-1;
				if(y >= dimy) y = dimy-1;
				vbuf[i][j] = tsbuf[x + (size_t)y*rwidth];
			}
			for(j = 0; j < 4; j++)
			{
				uint32_t o = 0;
				for(b = 0; b < 4; b++)   /* pavgb: (a+b+1)>>1 */
				{
					uint32_t a0 = (vbuf[0][j] >> (8*b)) & 255, a1 = (vbuf[1][j] >> (8*b)) & 255;
					uint32_t a2 = (vbuf[2][j] >> (8*b)) & 255, a3 = (vbuf[3][j] >> (8*b)) & 255;
					uint32_t l = (a0 + a1 + 1) >> 1, r = (a2 + a3 + 1) >> 1;
					o |= ((l + r + 1) >> 1) << (8*b);
				}
				p[j] = o;
			}
			dist += 4;
			p += 4;
		}
	}
	return 0;
}

/*
 * screen_upscale (screen.h:126-149), restated by destination index: the reference's single
 * running pointer puts sub-pixel (y, x) of source pixel (py, px) at word
 * py*(w*rscale + pitch*(rscale-1)) + px*rscale + y*pitch + x of the surface.
 */
int orc_upscale(const uint32_t *frame, int w, int h, int rwidth, int rscale, int pitch_words, uint32_t *surface)
{
	int px, py, x, y;
	const long long rowstep = (long long)w*rscale + (long long)pitch_words*(rscale - 1);
	if(w <= 0 || h <= 0 || rscale < 1 || pitch_words < w*rscale) return -1;
	for(py = 0; py < h; py++)
	for(px = 0; px < w; px++)
	{
		const uint32_t pd = frame[(size_t)py*rwidth + px];
		uint32_t *d = surface + py*rowstep + (long long)px*rscale;
		for(y = 0; y < rscale; y++)
		for(x = 0; x < rscale; x++)
			d[(long long)y*pitch_words + x] = pd;
	}
	return 0;
}

/* known-answer helpers */
void orc_kat_libm(const float *in, float *s, float *c, float *e, size_t n)
{
	size_t i;
	for(i = 0; i < n; i++) { s[i] = sinf(in[i]); c[i] = cosf(in[i]); e[i] = expf(in[i]); }
}
float orc_kat_randfs(uint32_t *seed) { return orc_randfs(seed); }
uint32_t orc_kat_col_ftoint(const float *v4) { return orc_col_ftoint(v4); }
float orc_kat_rcp(float x) { return cpu_rcp(x); }
float orc_kat_rsqrt(float x) { return cpu_rsqrt(x); }
int orc_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
