/*
 * pwn_camera.c -- scripted camera for the headless harness (see pwn_camera.h).
 * Player motion follows main.c:188-379 (yaw 3 rad/s, move 5 units/s, wall push-back with
 * PLAYER_BBOX 0.2, defs.h:7, gravity, raised-floor steps, portal traversal), written around a
 * height-window table for the cells and a crossing record for the portals, and pinned bit for bit
 * to the reference's own block extracted at build time (oracle/make_ref.py -> libref_tick*.so).
 */
#include <math.h>
#include <string.h>

#include "pwn_camera.h"

#define PLAYER_BBOX 0.2f

void pwn_cam_spawn(int sx, int sz, float cam[16])
{
	int i, j;
	for(i = 0; i < 4; i++)
	for(j = 0; j < 4; j++)
		cam[i*4 + j] = (i == j ? 1.0f : 0.0f);
	cam[12] = 0.5f + sx;
	cam[13] = 0.5f;
	cam[14] = 0.5f + sz;
}

void pwn_mat4_roty(float cam[16], float ang)
{
	float vs = sinf(ang), vc = cosf(ang);
	float vxx = cam[0], vxz = cam[2], vzx = cam[8], vzz = cam[10];
	cam[0] = vc*vxx + vs*vxz;
	cam[2] = vc*vxz - vs*vxx;
	cam[8] = vc*vzx + vs*vzz;
	cam[10] = vc*vzz - vs*vzx;
}

/*
 * Can the player stand in a cell of code `c`, coming from a cell of code `from`, at height y?
 * (what celltype_is_solid answers, util.h:112-126, as a table of height windows: a cell is open for
 * lo <= y < hi)
 */
static int cell_blocks(const pwn_level *lv, int32_t c, int32_t from, float y)
{
	float lo = 0.0f, hi;
	switch(c)
	{
		case '#': case '&': hi = 2.0f; break;
		case '"': hi = 1.0f; if(from == '#' || from == '&') { lo = 1.0f; hi = 2.0f; } break;
		case ';': case '$': case '>': case '<': case '^': case ',': hi = 1.0f; break;
		default:
			if(PWN_IS_PORTAL(c))
			{
				const int pi = PWN_PORTAL_INDEX(c);
				return !(pi >= 0 && pi < lv->nportals && lv->portals[pi].x2 != -1);   /* a paired portal is open at any height */
			}
			return 1;
	}
	return y < lo || y >= hi;
}

void pwn_player_init(pwn_player *pl, const pwn_level *lv)
{
	memset(pl, 0, sizeof(*pl));
	pwn_cam_spawn(lv->sx, lv->sz, pl->cam);
}

/* the far side of a portal as the player crosses it: which cell he comes out of, how far he is moved, and by how
 * many quarter turns he and his view are rotated about that cell's centre (main.c:301-326) */
typedef struct crossing
{
	int ok;            /* the cell entered is one of the pair's two ends */
	int shiftx, shiftz;
	float cx, cz;      /* centre of the cell on the far side */
	int quarter;       /* 0..3 */
} crossing;

static crossing portal_crossing(const pwn_portal *pm, int x, int z)
{
	crossing k;
	memset(&k, 0, sizeof(k));
	if(pm->x2 == -1) return k;
	const int at1 = (pm->x1 == x && pm->z1 == z), at2 = (pm->x2 == x && pm->z2 == z);
	if(!at1 && !at2) return k;
	const int dir = at1 ? 1 : -1;            /* end 1 -> end 2 or back */
	k.ok = 1;
	k.shiftx = dir*(pm->x2 - pm->x1);
	k.shiftz = dir*(pm->z2 - pm->z1);
	k.cx = (float)(at1 ? pm->x2 : pm->x1) + 0.5f;
	k.cz = (float)(at1 ? pm->z2 : pm->z1) + 0.5f;
	k.quarter = (-dir*pm->rot12) & 3;
	return k;
}

/* (x, z) turned by `quarter` quarter turns: exact, only signs and an exchange */
static void turn_vec(int quarter, float *x, float *z)
{
	const float ox = *x, oz = *z;
	switch(quarter)
	{
		case 1: *x = oz; *z = -ox; break;
		case 2: *x = -ox; *z = -oz; break;
		case 3: *x = -oz; *z = ox; break;
		default: break;
	}
}

/*
 * One tick of the player (main.c:188-379).  Order of events as in the reference: yaw, move by the view's z and x
 * axes, push back out of walls, fall, step onto / off raised floors, walk through portals.  Every float
 * expression keeps the reference's operand order -- the tick is compared bit for bit with the reference's own
 * block (oracle/_ref/libref_tick_strict.so, tests/test_level.py).
 */
void pwn_player_tick(pwn_player *pl, const pwn_level *lv, const pwn_keys *k, float tdiff)
{
	float *cam = pl->cam, *pos = pl->cam + 12;
	const float *axx = cam, *axz = cam + 8;      /* view x and z axes */
	int i;

	pwn_mat4_roty(cam, tdiff*3.0f*(k->turnleft - k->turnright));

	const int ox = (int)pos[0], oz = (int)pos[2];       /* the cell the tick starts in */
	const int32_t here = pwn_level_get_cell(lv, ox, oz);

	/* move: all four lanes, as the SSE original does */
	const float ahead = tdiff*5.0f*(k->moveforward - k->moveback), aside = tdiff*5.0f*(k->moveleft - k->moveright);
	float step[4];
	for(i = 0; i < 4; i++) step[i] = axz[i]*ahead + axx[i]*aside;
	for(i = 0; i < 4; i++) pos[i] += step[i];

	/* push-back: the leading corner of the bounding box decides which walls are hit */
	{
		const int dx = (step[0] < 0.0f ? -1 : 1), dz = (step[2] < 0.0f ? -1 : 1);
		const float y = pos[1];
		const int nx = (int)(pos[0] + dx*PLAYER_BBOX), nz = (int)(pos[2] + dz*PLAYER_BBOX);
		const int leftx = (nx != ox), leftz = (nz != oz);
		int hitx = 0, hitz = 0;
		if(leftx && leftz)
		{
			const int wx = cell_blocks(lv, pwn_level_get_cell(lv, nx, oz), here, y);
			const int wz = cell_blocks(lv, pwn_level_get_cell(lv, ox, nz), here, y);
			const int wc = cell_blocks(lv, pwn_level_get_cell(lv, nx, nz), here, y);
			hitx = wx;
			hitz = wz || (!wx && wc);      /* a bare corner pushes back along z */
		}
		else if(leftx) hitx = cell_blocks(lv, pwn_level_get_cell(lv, nx, nz), here, y);
		else if(leftz) hitz = cell_blocks(lv, pwn_level_get_cell(lv, nx, nz), here, y);
		if(hitx) pos[0] = ox + 0.5f + (0.5f-PLAYER_BBOX)*dx;
		if(hitz) pos[2] = oz + 0.5f + (0.5f-PLAYER_BBOX)*dz;
	}

	/* fall */
	for(i = 0; i < 4; i++) pos[i] += pl->gravity[i];
	pl->gravity[1] -= 3.0f*tdiff*tdiff;
	if(pos[1] < 0.4f) { pos[1] = 0.4f; pl->gravity[1] = 0.0f; }

	/* a new cell: raised floors and portals */
	const int tx = (int)pos[0], tz = (int)pos[2];
	if(tx == ox && tz == oz) return;
	const int32_t there = pwn_level_get_cell(lv, tx, tz);
	const int two_here = (here == '#' || here == '&'), two_there = (there == '#' || there == '&');
	if(two_here && there == '"') pos[1] -= 1.0f;
	else if(here == '"' && two_there) pos[1] += 1.0f;
	else if(PWN_IS_PORTAL(there) && PWN_PORTAL_INDEX(there) < lv->nportals)
	{
		const crossing c = portal_crossing(&lv->portals[PWN_PORTAL_INDEX(there)], tx, tz);
		if(!c.ok) return;                  /* unpaired, or a third cell with the letter: nothing happens */
		float x = pos[0], z = pos[2];
		x += (tx - ox) + c.shiftx;         /* one cell further, then across */
		z += (tz - oz) + c.shiftz;
		/* about the centre of the far cell; s = +1 / -1 for one / three quarter turns (a product with +-1 is exact) */
		if(c.quarter == 2) { x = c.cx*2.0f - x; z = c.cz*2.0f - z; }
		else if(c.quarter != 0)
		{
			const float s = (c.quarter == 1 ? 1.0f : -1.0f), fx = x, fz = z;
			x = c.cx + s*(fz - c.cz);
			z = c.cz - s*(fx - c.cx);
		}
		pos[0] = x; pos[2] = z;
		turn_vec(c.quarter, &cam[0], &cam[2]);
		turn_vec(c.quarter, &cam[8], &cam[10]);
	}
}

/*
 * The fixed camera-path script: 60 Hz ticks from the spawn.  A 240-tick key cycle --
 * walk forward, arc left while walking, stand and pan right, strafe -- so that a
 * 600-frame run turns through every yaw, slides along walls and crosses portals.
 * Pose i is the camera BEFORE tick i, with sec_current = i/60 (main.c:107-114 order).
 */
int pwn_flythrough(const pwn_level *lv, int nframes, pwn_pose *out)
{
	pwn_player pl;
	int f;
	const float tdiff = 1.0f/60.0f;
	pwn_player_init(&pl, lv);
	float sec = 0.0f;
	for(f = 0; f < nframes; f++)
	{
		memcpy(out[f].cam, pl.cam, sizeof(pl.cam));
		out[f].sec_current = sec;
		pwn_keys k;
		memset(&k, 0, sizeof(k));
		int t = f % 240;
		if(t < 90) k.moveforward = 1;
		else if(t < 150) { k.moveforward = 1; k.turnleft = 1; }
		else if(t < 200) k.turnright = 1;
		else { k.moveleft = 1; k.moveforward = 1; }
		pwn_player_tick(&pl, lv, &k, tdiff);
		sec += tdiff;
	}
	return 0;
}
