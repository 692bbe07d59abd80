/* pwn_game.c -- see pwn_game.h.  Arithmetic in double like Lua's, narrowed to float where
 * obj_set stores into the part (script.h:22-32). */
#include <math.h>
#include <string.h>

#include "pwn_game.h"

static const double OPOS[PWN_GAME_NSPH][8] = {
	{ 0.0, 0.0,  0.0, 0.3,  0.8, 0.8, 0.8, 0.6},
	{ 0.0, 0.3,  0.0, 0.1,  0.4, 0.4, 0.4, 0.2},
	{ 0.3, 0.0,  0.0, 0.1,  0.7, 0.7, 1.0, 0.4},
	{ 0.0, 0.0,  0.3, 0.1,  0.7, 1.0, 0.7, 0.4},
	{-0.3, 0.0,  0.0, 0.1,  1.0, 0.7, 0.7, 0.4},
	{ 0.0, 0.0, -0.3, 0.1,  0.5, 1.0, 1.0, 0.4},
	{ 0.3, 0.0,  0.1, 0.03, 0.4, 0.4, 0.4, 0.2},
	{ 0.1, 0.0,  0.3, 0.03, 0.4, 0.4, 0.4, 0.2},
	{-0.3, 0.0,  0.1, 0.03, 0.4, 0.4, 0.4, 0.2},
	{ 0.1, 0.0, -0.3, 0.03, 0.4, 0.4, 0.4, 0.2},
	{ 0.3, 0.0, -0.1, 0.03, 0.4, 0.4, 0.4, 0.2},
	{-0.1, 0.0,  0.3, 0.03, 0.4, 0.4, 0.4, 0.2},
	{-0.3, 0.0, -0.1, 0.03, 0.4, 0.4, 0.4, 0.2},
	{-0.1, 0.0, -0.3, 0.03, 0.4, 0.4, 0.4, 0.2},
};

/* obj_set(o, "sphere", r, refl, x, y, z, cb, cg, cr), script.h:18-33 */
static void obj_set(pwn_sphere *s, double r, double refl, double x, double y, double z, double cb, double cg, double cr)
{
	s->r = (float)r; s->refl = (float)refl;
	s->pos[0] = (float)x; s->pos[1] = (float)y; s->pos[2] = (float)z; s->pos[3] = 1.0f;
	s->col[0] = (float)cb; s->col[1] = (float)cg; s->col[2] = (float)cr;
}

void pwn_game_init(pwn_game *g)
{
	int i;
	memset(g, 0, sizeof(*g));
	memcpy(g->opos, OPOS, sizeof(OPOS));
	g->obx = 9.5; g->oby = 0.3; g->obz = 5.5;
	g->obvx = 1; g->obvz = 0;
	for(i = 0; i < PWN_GAME_NSPH; i++)   /* game.lua:25-30 */
		obj_set(&g->sph[i], g->opos[i][3], g->opos[i][7], g->obx + g->opos[i][0], g->oby + g->opos[i][1], g->obz + g->opos[i][2],
			g->opos[i][4], g->opos[i][5], g->opos[i][6]);
}

/* level_get(x, z) with math.floor'ed arguments: lf_level_get -> get_cell (script.h:53-63) */
static int32_t level_get(const pwn_level *lv, double x, double z)
{
	return pwn_level_get_cell(lv, (int)floor(x), (int)floor(z));
}

static int blocked(int32_t c1, int32_t c2)   /* game.lua:70,75 */
{
	return c2 == '.' || ((c1 == '#' || c1 == '&') && c2 == '"');
}

void pwn_game_tick(pwn_game *g, const pwn_level *lv, double sec_current, double sec_delta)
{
	int i;
	const double PI = 3.14159265358979323846;
	if(fmod(sec_current, 0.5) < 0.15) { g->opos[1][4] = 0.3; g->opos[1][5] = 0.3; g->opos[1][6] = 1.3; }   /* game.lua:36-40 */
	else { g->opos[1][4] = 0.3; g->opos[1][5] = 0.3; g->opos[1][6] = 0.3; }

	const double rs = sin(sec_current*PI*2/2), rc = cos(sec_current*PI*2/2);   /* game.lua:48-49 */
	for(i = 0; i < PWN_GAME_NSPH; i++)
	{
		double rx = g->opos[i][0], ry = g->opos[i][1], rz = g->opos[i][2];
		const double nx = rc*rx + rs*rz, nz = rc*rz - rs*rx;   /* game.lua:51 */
		rx = nx; rz = nz;
		obj_set(&g->sph[i], g->opos[i][3], g->opos[i][7], g->obx + rx, g->oby + ry, g->obz + rz,
			g->opos[i][4], g->opos[i][5], g->opos[i][6]);
	}

	const int32_t c1 = level_get(lv, g->obx, g->obz);   /* game.lua:61 */
	const double spd = 2.0;
	double nobx = g->obx + g->obvx*sec_delta*spd, nobz = g->obz + g->obvz*sec_delta*spd;
	int32_t c2 = level_get(lv, nobx + g->obvx*0.5, nobz + g->obvz*0.5);
	if(c1 != c2 && blocked(c1, c2))   /* game.lua:69-83 */
	{
		double t = g->obvx; g->obvx = g->obvz; g->obvz = -t;
		nobx = g->obx + g->obvx*sec_delta*spd; nobz = g->obz + g->obvz*sec_delta*spd;
		c2 = level_get(lv, nobx + g->obvx*0.5, nobz + g->obvz*0.5);
		if(blocked(c1, c2))
		{
			g->obvx = -g->obvx; g->obvz = -g->obvz;
			nobx = g->obx + g->obvx*sec_delta*spd; nobz = g->obz + g->obvz*sec_delta*spd;
		}
	}
	g->obx = nobx; g->obz = nobz;
}
