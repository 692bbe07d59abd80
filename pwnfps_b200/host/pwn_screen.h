/*
 * pwn_screen.h -- drop-in replacement for the reference's trace.h + screen.h.
 *
 * Include it in main.c where "trace.h" and "screen.h" were (main.c:37-38), after
 * defs.h / util.h and the globals of main.c:26-34.  It defines the same two entry
 * points the game loop calls --
 *
 *     void trace_screen_centred(level *lv, int x1, int y1, int x2, int y2, const mat4 *cam);   screen.h:31
 *     void screen_upscale(void);                                                                screen.h:126
 *
 * -- and forwards the first to libpwntrace.so (include/pwntrace.h).  Everything else in
 * the host (level.h loader, script.h Lua bindings, camera/player code) stays as it is.
 * What the shim reads, exactly what the reference's trace reads: lv->data, lv->pmap,
 * lv->objs (through the same filter level_prepare_render applies, level.h:77-79), the
 * globals rwidth, rheight, sec_current, and it fills the globals sbuf and zbuf.
 *
 * Errors: the reference's trace has none (void + assert); a library failure here prints
 * pwn_last_error() and aborts, like the reference's own unsupported-part path (level.h:34-37).
 */
#ifndef PWN_SCREEN_H
#define PWN_SCREEN_H

#include "pwntrace.h"

static pwn_ctx *pwn_screen_ctx = NULL;
static level *pwn_screen_level = NULL;
static uint64_t pwn_screen_level_sum = 0;
static int pwn_screen_w = 0, pwn_screen_h = 0;
static pwn_sphere *pwn_screen_sph = NULL;
static size_t pwn_screen_sph_cap = 0;

#define PWN_SCREEN_CHECK(call) do { int rc_ = (call); if(rc_ != PWN_OK) { \
	printf("pwntrace: %s failed (%d): %s\n", #call, rc_, pwn_last_error()); fflush(stdout); abort(); } } while(0)

#define PWN_SCREEN_DIM ((int)(sizeof(((level *)0)->data[0])/sizeof(((level *)0)->data[0][0])))
#define PWN_SCREEN_NPORTAL ((int)(sizeof(((level *)0)->pmap)/sizeof(((level *)0)->pmap[0])))

static uint32_t *pwn_screen_reg_s = NULL;
static float *pwn_screen_reg_z = NULL;

static void pwn_screen_release_buffers(void)
{
	if(pwn_screen_reg_s != NULL) pwn_host_unregister(pwn_screen_reg_s);
	if(pwn_screen_reg_z != NULL) pwn_host_unregister(pwn_screen_reg_z);
	pwn_screen_reg_s = NULL; pwn_screen_reg_z = NULL;
}

/*
 * What the trace reads of a level besides the parts: the grid and the portal table.  The reference's loader
 * always returns the one static `lvbase` (level.h:83,90), so the pointer says nothing about a level that was
 * loaded again: the shim keys the upload on the contents (FNV-1a over 4 KB + 26 portals, a few microseconds a frame).
 */
static uint64_t pwn_screen_level_hash(const level *lv)
{
	uint64_t h = 0xcbf29ce484222325ull;
	const unsigned char *p = (const unsigned char *)lv->data;
	size_t i;
	for(i = 0; i < sizeof(lv->data); i++) h = (h ^ p[i])*0x100000001b3ull;
	p = (const unsigned char *)lv->pmap;
	for(i = 0; i < sizeof(lv->pmap); i++) h = (h ^ p[i])*0x100000001b3ull;
	return h;
}

/* for hosts that edit lv->data or lv->pmap in place and do not want to wait for the next frame's check */
static void __attribute__((unused)) pwn_screen_invalidate_level(void)
{
	pwn_screen_level = NULL;
}

static void pwn_screen_upload_level(level *lv)
{
	int x, z, i;
	const int dim = PWN_SCREEN_DIM, np = PWN_SCREEN_NPORTAL;
	int32_t *cells = malloc(sizeof(int32_t)*dim*dim);
	pwn_portal *pm = malloc(sizeof(pwn_portal)*np);
	for(z = 0; z < dim; z++)
	for(x = 0; x < dim; x++)
		cells[z*dim + x] = lv->data[z][x];
	for(i = 0; i < np; i++)
	{
		pm[i].x1 = lv->pmap[i].x1; pm[i].z1 = lv->pmap[i].z1;
		pm[i].x2 = lv->pmap[i].x2; pm[i].z2 = lv->pmap[i].z2;
		pm[i].rot12 = lv->pmap[i].x2 == -1 ? 0 : lv->pmap[i].rot12;
		pm[i].c1 = lv->pmap[i].c1; pm[i].c2 = lv->pmap[i].c2;
	}
	PWN_SCREEN_CHECK(pwn_level_upload(pwn_screen_ctx, dim, dim, cells, np, pm));
	free(cells);
	free(pm);
	pwn_screen_level = lv;
	pwn_screen_level_sum = pwn_screen_level_hash(lv);
}

/* lv->objs[] -> spheres, in the order level_prepare_render adds them (level.h:77-79) */
static void pwn_screen_upload_parts(level *lv)
{
	ssize_t i;
	int n = 0;
	if((size_t)lv->objs_num > pwn_screen_sph_cap)
	{
		pwn_screen_sph_cap = (size_t)lv->objs_num*2;
		pwn_screen_sph = realloc(pwn_screen_sph, sizeof(pwn_sphere)*pwn_screen_sph_cap);
	}
	for(i = 0; i < lv->objs_num; i++)
	{
		const part *pt = &lv->objs[i];
		if(pt->typ == P_FREE) continue;
		if(pt->typ != P_SPHERE)
		{
			printf("unsupported part type %i\n", pt->typ);   /* level.h:34-37 */
			fflush(stdout);
			abort();
		}
		pwn_screen_sph[n].r = pt->sph.r;
		pwn_screen_sph[n].refl = pt->sph.refl;
		memcpy(pwn_screen_sph[n].pos, pt->sph.pos.a, sizeof(float)*4);
		memcpy(pwn_screen_sph[n].col, pt->sph.col.a, sizeof(float)*4);
		n++;
	}
	PWN_SCREEN_CHECK(pwn_parts_upload(pwn_screen_ctx, n, pwn_screen_sph));
}

/* [p1,p2) */
void trace_screen_centred(level *lv, int x1, int y1, int x2, int y2, const mat4 *cam)
{
	assert(x1 < x2);
	assert(y1 < y2);

	if(pwn_screen_ctx == NULL)
	{
		const char *dev = getenv("PWN_DEVICE");
		PWN_SCREEN_CHECK(pwn_create(&pwn_screen_ctx, dev != NULL ? atoi(dev) : 0));
		PWN_SCREEN_CHECK(pwn_math_tables_from_host(pwn_screen_ctx));
	}
	/* grid and portals never change after load (script.h:65-69), but a level can be loaded again into the same struct */
	if(lv != pwn_screen_level || pwn_screen_level_hash(lv) != pwn_screen_level_sum)
		pwn_screen_upload_level(lv);
	if(rwidth != pwn_screen_w || rheight != pwn_screen_h)
	{
		PWN_SCREEN_CHECK(pwn_frame_config(pwn_screen_ctx, rwidth, rheight, 0));
		pwn_screen_w = rwidth; pwn_screen_h = rheight;
	}
	/* the host's own malloc'd frame buffers (main.c:395-400), page-locked in place for the copies.
	 * In the reference they live as long as the process; a host that reallocates them calls
	 * pwn_screen_release_buffers() before free().  A failure only costs speed. */
	if(sbuf != pwn_screen_reg_s || zbuf != pwn_screen_reg_z)
	{
		pwn_screen_release_buffers();
		if(pwn_host_register(sbuf, sizeof(uint32_t)*(size_t)rwidth*rheight) == PWN_OK) pwn_screen_reg_s = sbuf;
		if(pwn_host_register(zbuf, sizeof(float)*(size_t)rwidth*rheight) == PWN_OK) pwn_screen_reg_z = zbuf;
	}
	pwn_screen_upload_parts(lv);

	PWN_SCREEN_CHECK(pwn_trace_config(pwn_screen_ctx, REFLECT));   /* defs.h:8 */
	unsigned flags = 0;
#ifdef POSTPROC_BLUR
	flags |= PWN_F_BLUR;
#endif
#ifndef REFLECT_BLUR
	flags |= PWN_F_NO_REFLECT_BLUR;
#endif
	/* zbuf is written by the trace and read by the blur pass of this same function; nothing else in the
	 * reference host reads it (main.c, script.h).  -DPWN_SCREEN_NO_ZBUF leaves the host's copy unfilled
	 * and halves the bytes that cross PCIe per frame. */
#ifdef PWN_SCREEN_NO_ZBUF
	PWN_SCREEN_CHECK(pwn_trace_screen(pwn_screen_ctx, (const float *)cam, x2 - x1, y2 - y1, sec_current,
		flags, sbuf, NULL));
#else
	PWN_SCREEN_CHECK(pwn_trace_screen(pwn_screen_ctx, (const float *)cam, x2 - x1, y2 - y1, sec_current,
		flags, sbuf, zbuf));
#endif
}

/*
 * nearest-neighbour blit of sbuf into the SDL surface, rscale x rscale per pixel (screen.h:126-149).
 * The reference walks one running pointer, so block row py starts at py*(w*rscale + pitch*(rscale-1))
 * words -- py*rscale*pitch only when the pitch is not padded; the same words are written here.
 * It stays on the host: sbuf has to cross PCIe anyway and is rscale^2 smaller than the surface
 * (pwn_screen_upscale is the device version for hosts that present from device memory).
 */
void screen_upscale(void)
{
	int px, py, y, x;
	const int pitch = screen->pitch/(int)sizeof(uint32_t);
	const size_t rowstep = (size_t)rwidth*rscale + (size_t)pitch*(rscale - 1);
	for(py = 0; py < rheight; py++)
	{
		const uint32_t *s = sbuf + (size_t)py*rwidth;
		uint32_t *row0 = (uint32_t *)screen->pixels + (size_t)py*rowstep;
		for(px = 0; px < rwidth; px++)
			for(x = 0; x < rscale; x++)
				row0[px*rscale + x] = s[px];
		for(y = 1; y < rscale; y++)
			memcpy(row0 + (size_t)y*pitch, row0, sizeof(uint32_t)*(size_t)rwidth*rscale);
	}
}

#endif /* PWN_SCREEN_H */
