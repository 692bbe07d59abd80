/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Headless harness around the UNMODIFIED reference headers of fanzyflani/pwnfps.
 * It contains none of the reference's code: the five headers are #included from
 * where they lie (-I/root/reference, see oracle/Makefile) exactly the way the
 * reference's own main.c does it (main.c:1-39), and the result is a shared
 * object under oracle/_ref/ (git-ignored, travels to the GPU box prebuilt).
 *
 * What it replaces from main.c: SDL window + Lua VM + game loop (main.c:42-403).
 * What it keeps: the globals the hot path reads/writes (main.c:26-34), the
 * 16-byte-aligned sbuf/tsbuf/zbuf buffers (main.c:395-400), level_load
 * (main.c:51), level_prepare_render + trace_screen_centred per frame
 * (main.c:95,107).
 *
 * Build-time switches (all applied without touching the reference text):
 *   -DREF_NO_BLUR      #undef POSTPROC_BLUR after defs.h  (pre-blur frame)
 *   -DREF_REFLECT=n    redefine REFLECT (segment-split timing)
 *   -DREF_IDS          the including TU is the build-time *instrumented* copy
 *                      made by oracle/make_ref.py (ORC_* hooks, see below)
 *   -DREF_SHIM         the PRODUCT's drop-in shim (pwnfps_b200/host/pwn_screen.h) replaces
 *                      trace.h + screen.h: the reference host driving libpwntrace.so
 *   -DREF_WIDE         the including TU is the build-time *widened* copy
 *                      (LV_DIM x LV_DIM grid, PM_MAX portals)
 *   -DREF_TICK         also builds ref_player_tick(): main.c:188-379, extracted at build time by
 *                      oracle/make_ref.py into tick_body.inc (temporary directory)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the resulting library.
 */
#define _GNU_SOURCE
#include <string.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdio.h>
#include <errno.h>

#include <assert.h>
#include <math.h>
#include <sys/types.h>
#include <strings.h>

#include <mmintrin.h>
#include <xmmintrin.h>
#include <emmintrin.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* screen_upscale (screen.h:126-149) is the only SDL user in the included set */
typedef struct { int pitch; void *pixels; } SDL_Surface;

#ifdef REF_WIDE
#ifndef LV_DIM
#define LV_DIM 128
#endif
#ifndef PM_MAX
#define PM_MAX 4096
#endif
#endif

#include "defs.h"

#ifdef REF_NO_BLUR
#undef POSTPROC_BLUR
#endif
#ifdef REF_NO_REFLECT_BLUR
#undef REFLECT_BLUR
#endif
#ifdef REF_REFLECT
#undef REFLECT
#define REFLECT REF_REFLECT
#endif

/* globals as main.c:26-34 */
int rwidth = DEF_RWIDTH;
int rheight = DEF_RHEIGHT;
int rscale = DEF_SCALE;

SDL_Surface *screen = NULL;
uint32_t *sbuf = NULL;
uint32_t *tsbuf = NULL;
float *zbuf = NULL;
level *lvroot = NULL;

/* ---- hit-ID hooks used only by the instrumented copy (-DREF_IDS) ---------- */
typedef struct ref_ids_s
{
	int32_t cell;     /* cz*DIM + cx at termination, -1 = never terminated   */
	int16_t ldir;     /* FXP..FYN (defs.h:25-33); -1 = sphere (aux_dir)      */
	int16_t kind;     /* 0 wall/floor/ceiling via trace_hit_wall, 1 sphere via
	                     trace_hit_bounce, 2 walk exhausted (trace.h:677)     */
	int32_t nstep;    /* loop iterations entered (trace.h:250)               */
	int32_t nportal;  /* successful portal crossings (trace.h:576)           */
} ref_ids;

#ifdef REF_IDS
static ref_ids *orc_ids = NULL;      /* [rwidth*rheight][3] */
#define ORC_ENTER() int orc_nstep = 0, orc_nport = 0;
#define ORC_STEP() orc_nstep++;
#define ORC_PORTAL() orc_nport++;
#define ORC_REC(k, ld) do { \
	if(orc_ids != NULL && hitctr >= 0 && hitctr < 3) { \
		ref_ids *o_ = &orc_ids[(dist - zbuf)*3 + hitctr]; \
		o_->cell = cz*ORC_DIM + cx; o_->ldir = (int16_t)(ld); o_->kind = (k); \
		o_->nstep = orc_nstep; o_->nportal = orc_nport; } } while(0)
#define ORC_WALL() ORC_REC(0, ldir)
#define ORC_SPH() ORC_REC(1, -1)
#define ORC_EXHAUST() ORC_REC(2, ldir)
#ifdef REF_WIDE
#define ORC_DIM LV_DIM
#else
#define ORC_DIM 64
#endif
#endif

#include "util.h"
#ifdef REF_SHIM
/* the reference host with trace.h + screen.h swapped for the drop-in shim:
 * everything else -- defs.h, util.h, level.h and this harness -- is unchanged */
#include "pwn_screen.h"
#else
#include "trace.h"
#include "screen.h"
#endif
#include "level.h"

/* ---------------------------------------------------------------------------- */

static void *buf_base[3] = {NULL, NULL, NULL};

static int ref_alloc(int w, int h)
{
	int i;
	if(w <= 0 || h <= 0) return -1;
	if(w == rwidth && h == rheight && sbuf != NULL) return 0;
#ifdef REF_SHIM
	pwn_screen_release_buffers();   /* the shim page-locks sbuf/zbuf in place */
#endif
	for(i = 0; i < 3; i++) { free(buf_base[i]); buf_base[i] = NULL; }
	size_t n = (size_t)w*(size_t)h*4 + 64;
	for(i = 0; i < 3; i++)
		if(posix_memalign(&buf_base[i], 64, n) != 0) return -2;
	sbuf = buf_base[0]; tsbuf = buf_base[1]; zbuf = buf_base[2];
	rwidth = w; rheight = h;
	return 0;
}

int ref_dim(void)
{
#ifdef REF_WIDE
	return LV_DIM;
#else
	return 64;
#endif
}

int ref_pm_max(void)
{
#ifdef REF_WIDE
	return PM_MAX;
#else
	return 26;
#endif
}

int ref_flags(void)
{
	int f = 0;
#ifdef POSTPROC_BLUR
	f |= 1;
#endif
#ifdef REF_IDS
	f |= 2;
#endif
#ifdef REF_WIDE
	f |= 4;
#endif
#ifdef REFLECT_BLUR
	f |= 8;
#endif
	f |= (REFLECT << 8);
	return f;
}

/* level.txt path -> static lvbase via the reference loader (level.h:107) */
int ref_level_load(const char *fname)
{
	lvroot = level_load(fname);
	fflush(stdout);
	return lvroot == NULL ? -1 : 0;
}

/* In-memory level (wide or not): cells as int codes, portal table given by the
 * caller already derived (rot12/c1/c2 as level.h:194-221 would).  Used for the
 * levels that do not fit level.txt (SURVEY 8c "wide oracle"). */
int ref_level_set(int dim, const int32_t *cells, int nport, const int32_t *pm /* [n][7]: x1 z1 x2 z2 rot12 c1 c2 */, int sx, int sz)
{
	int x, z, i;
	if(dim != ref_dim() || nport > ref_pm_max()) return -1;
	level *lv = level_new();
	for(z = 0; z < dim; z++)
	for(x = 0; x < dim; x++)
	{
		lv->data[z][x] = cells[z*dim + x];
		lv->parts[z][x] = NULL;
		lv->parts_num[z][x] = 0;
		lv->parts_max[z][x] = 0;
	}
	for(i = 0; i < nport; i++)
	{
		lv->pmap[i].x1 = pm[i*7+0]; lv->pmap[i].z1 = pm[i*7+1];
		lv->pmap[i].x2 = pm[i*7+2]; lv->pmap[i].z2 = pm[i*7+3];
		lv->pmap[i].rot12 = pm[i*7+4];
		lv->pmap[i].c1 = pm[i*7+5]; lv->pmap[i].c2 = pm[i*7+6];
	}
	lv->sx = sx; lv->sz = sz;
	lvroot = lv;
	return 0;
}

int ref_level_spawn(int *sx, int *sz)
{
	if(lvroot == NULL) return -1;
	*sx = lvroot->sx; *sz = lvroot->sz;
	return 0;
}

/* loader KAT: cells as int32 [dim*dim], portals as int32 [pm_max][7] */
int ref_level_get(int32_t *cells, int32_t *pm)
{
	int x, z, i, dim = ref_dim();
	if(lvroot == NULL) return -1;
	for(z = 0; z < dim; z++)
	for(x = 0; x < dim; x++)
		cells[z*dim + x] = lvroot->data[z][x];
	for(i = 0; i < ref_pm_max(); i++)
	{
		/* level_new resets only x1/x2 (level.h:94-99): z1/z2/rot12 of an unused end are stale */
		pm[i*7+0] = lvroot->pmap[i].x1; pm[i*7+1] = lvroot->pmap[i].x1 == -1 ? 0 : lvroot->pmap[i].z1;
		pm[i*7+2] = lvroot->pmap[i].x2; pm[i*7+3] = lvroot->pmap[i].x2 == -1 ? 0 : lvroot->pmap[i].z2;
		pm[i*7+4] = lvroot->pmap[i].x2 == -1 ? 0 : lvroot->pmap[i].rot12;
		pm[i*7+5] = lvroot->pmap[i].c1; pm[i*7+6] = lvroot->pmap[i].c2;
	}
	return 0;
}

/* spheres: what script.h:10-40 (obj_set "sphere") does to a part */
int ref_objs_clear(void)
{
	if(lvroot == NULL) return -1;
	lvroot->objs_num = 0;
	return 0;
}

int ref_obj_sphere(float r, float refl, float x, float y, float z, float cb, float cg, float cr)
{
	if(lvroot == NULL) return -1;
	part *pt = level_obj_new(lvroot);
	if(pt == NULL) return -2;
	pt->typ = P_SPHERE;
	pt->sph.r = r;
	pt->sph.refl = refl;
	pt->sph.pos.v.x = x; pt->sph.pos.v.y = y; pt->sph.pos.v.z = z; pt->sph.pos.v.w = 1.0f;
	/* script.h:30-32 leaves col.a untouched; a fresh lvbase slot is zero */
	pt->sph.col.c.b = cb; pt->sph.col.c.g = cg; pt->sph.col.c.r = cr; pt->sph.col.c.a = 0.0f;
	return 0;
}

/*
 * One frame, exactly main.c:95 + main.c:107.  cam = 16 floats, rows x,y,z,w
 * (defs.h:46-52).  zbuf/sbuf are zeroed first so that a walk that exhausts
 * maxsteps (trace.h:677-678, leaves *dist untouched) is deterministic.
 * Returns wall-clock seconds of trace_screen_centred alone.
 */
double ref_trace_frame(const float *cam16, float sec, int w, int h, int nthreads,
	uint32_t *sbuf_out, float *zbuf_out, ref_ids *ids_out)
{
	mat4 cam;
	double t0 = 0.0, t1 = 0.0;
	if(lvroot == NULL) return -1.0;
	if(ref_alloc(w, h) != 0) return -2.0;
	memcpy(&cam, cam16, sizeof(cam));
	sec_current = sec;
	memset(sbuf, 0, (size_t)w*h*4);
	memset(tsbuf, 0, (size_t)w*h*4);
	memset(zbuf, 0, (size_t)w*h*4);
#ifdef REF_IDS
	orc_ids = ids_out;
	if(orc_ids != NULL)
	{
		size_t i, n = (size_t)w*h*3;
		for(i = 0; i < n; i++)
		{ orc_ids[i].cell = -1; orc_ids[i].ldir = -2; orc_ids[i].kind = -1; orc_ids[i].nstep = 0; orc_ids[i].nportal = 0; }
	}
#else
	(void)ids_out;
#endif
#ifdef _OPENMP
	if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
	level_prepare_render(lvroot);
#ifdef _OPENMP
	t0 = omp_get_wtime();
#endif
	trace_screen_centred(lvroot, 0, 0, rwidth, rheight, &cam);
#ifdef _OPENMP
	t1 = omp_get_wtime();
#endif
	if(sbuf_out != NULL) memcpy(sbuf_out, sbuf, (size_t)w*h*4);
	if(zbuf_out != NULL) memcpy(zbuf_out, zbuf, (size_t)w*h*4);
	return t1 - t0;
}

/*
 * The FP environment of the shipped executable.  The reference Makefile's -ffast-math makes gcc link crtfastmath.o into
 * an EXECUTABLE, whose constructor sets MXCSR.FTZ and MXCSR.DAZ; gcc does not link it into a shared object, so these
 * libraries, loaded into Python, run with denormals on.  ref_set_fpenv(1) sets FTZ|DAZ on the calling thread and on every
 * OpenMP worker (what the game's threads inherit from main), ref_set_fpenv(0) clears them: tests/test_oracle_ref.py traces
 * the same frames both ways.
 */
int ref_set_fpenv(int ftz_daz)
{
	const unsigned bits = 0x8040u;   /* FTZ (bit 15) | DAZ (bit 6) */
	_mm_setcsr(ftz_daz ? (_mm_getcsr() | bits) : (_mm_getcsr() & ~bits));
#ifdef _OPENMP
#pragma omp parallel
	{
		_mm_setcsr(ftz_daz ? (_mm_getcsr() | bits) : (_mm_getcsr() & ~bits));
	}
#endif
	return (int)(_mm_getcsr() & bits);
}

/* Repeat the same frame n times without copying out: the CPU-baseline timer. */
double ref_time_frames(const float *cam16, float sec, int w, int h, int nthreads, int nframes)
{
	int i;
	mat4 cam;
	double t0 = 0.0, t1 = 0.0;
	if(lvroot == NULL) return -1.0;
	if(ref_alloc(w, h) != 0) return -2.0;
	memcpy(&cam, cam16, sizeof(cam));
	sec_current = sec;
#ifdef _OPENMP
	if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
	level_prepare_render(lvroot);
#ifdef _OPENMP
	t0 = omp_get_wtime();
#endif
	for(i = 0; i < nframes; i++)
		trace_screen_centred(lvroot, 0, 0, rwidth, rheight, &cam);
#ifdef _OPENMP
	t1 = omp_get_wtime();
#endif
	return t1 - t0;
}

/*
 * screen_upscale (screen.h:126-149) of a caller-supplied frame into a caller-supplied surface:
 * the globals it reads (sbuf, rwidth, rheight, rscale, screen) are pointed at the arguments
 * for the call and restored.
 */
int ref_screen_upscale(const uint32_t *frame, int w, int h, int scale, int pitch_bytes, void *pixels)
{
	SDL_Surface surf;
	SDL_Surface *old_screen = screen;
	uint32_t *old_sbuf = sbuf;
	int ow = rwidth, oh = rheight, os = rscale;
	surf.pitch = pitch_bytes; surf.pixels = pixels;
	screen = &surf; sbuf = (uint32_t *)frame; rwidth = w; rheight = h; rscale = scale;
	screen_upscale();
	screen = old_screen; sbuf = old_sbuf; rwidth = ow; rheight = oh; rscale = os;
	return 0;
}

int ref_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* Known-answer helpers for the device restatement of util.h */
uint32_t ref_randi(uint32_t *seed) { return randi(seed); }
float ref_randfu(uint32_t *seed) { return randfu(seed); }
float ref_randfs(uint32_t *seed) { return randfs(seed); }
uint32_t ref_col_ftoint(const float *v4)
{
	return col_ftoint(_mm_loadu_ps(v4));
}
void ref_v_normalise(const float *in4, float *out4)
{
	_mm_storeu_ps(out4, v_normalise(_mm_loadu_ps(in4)));
}
float ref_v_dot(const float *a4, const float *b4)
{
	return v_dot(_mm_loadu_ps(a4), _mm_loadu_ps(b4));
}
void ref_mat4_roty(float *m16, float ang)
{
	mat4 m;
	memcpy(&m, m16, sizeof(m));
	mat4_roty(&m, ang);
	memcpy(m16, &m, sizeof(m));
}
/* the host CPU's rcpps / rsqrtps (call sites trace.h:231, util.h:155) */
void ref_rcpps(const float *in, float *out, size_t n)
{
	size_t i;
	for(i = 0; i < n; i++)
		out[i] = _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(in[i])));
}
void ref_rsqrtps(const float *in, float *out, size_t n)
{
	size_t i;
	for(i = 0; i < n; i++)
		out[i] = _mm_cvtss_f32(_mm_rsqrt_ps(_mm_set1_ps(in[i])));
}
void ref_libm(const float *in, float *s, float *c, float *e, size_t n)
{
	size_t i;
	for(i = 0; i < n; i++) { s[i] = sinf(in[i]); c[i] = cosf(in[i]); e[i] = expf(in[i]); }
}

#ifdef REF_TICK
/*
 * One game tick of player motion: the reference's own statements, main.c:188-379 (yaw, velocity, move,
 * wall push-back, gravity, raised-floor steps, portal traversal), extracted at build time into
 * tick_body.inc.  The block reads and writes only the names declared here -- the locals of mainloop()
 * it uses (main.c:46-48,70-77), with `lv` the level loaded into this library.
 *   cam16     mat4 cam, rows x y z w (in/out)        gravity4  vec4 gravity (in/out)
 *   keys      k_turnleft, k_turnright, k_turnup, k_turndown, k_moveforward, k_moveback, k_moveleft, k_moveright
 */
int ref_player_tick(float *cam16, float *gravity4, const int *keys, float tdiff)
{
	mat4 cam;
	vec4 gravity;
	level *lv = lvroot;
	if(lv == NULL) return -1;
	int k_turnleft = keys[0], k_turnright = keys[1], k_turnup = keys[2], k_turndown = keys[3];
	int k_moveforward = keys[4], k_moveback = keys[5], k_moveleft = keys[6], k_moveright = keys[7];
	(void)k_turnup; (void)k_turndown;
	memcpy(&cam, cam16, sizeof(cam));
	memcpy(&gravity, gravity4, sizeof(gravity));
#define printf(...) ((void)0)      /* the block announces every traversal on stdout (main.c:297) */
	{
#include "tick_body.inc"
	}
#undef printf
	memcpy(cam16, &cam, sizeof(cam));
	memcpy(gravity4, &gravity, sizeof(gravity));
	return 0;
}
#endif
