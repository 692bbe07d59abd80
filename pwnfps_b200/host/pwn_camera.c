/*
 * pwn_camera.c -- scripted camera for the headless harness (see pwn_camera.h).
 * Player motion restates main.c:188-379: yaw 3 rad/s, move 5 units/s, wall push-back
 * with PLAYER_BBOX 0.2 (defs.h:7), gravity, raised-floor steps and portal traversal
 * with the same quarter-turn maths as the ray version (trace.h:578-617).
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

/* util.h:112-126 */
static int celltype_is_solid(const pwn_level *lv, int32_t c, int32_t oldcell, float y)
{
	if(c == '"' && (oldcell == '#' || oldcell == '&')) return y < 1.0f || y >= 2.0f;
	if(c == '#' || c == '&') return y < 0.0f || y >= 2.0f;
	if(c == ';' || c == '$' || c == '"') return y < 0.0f || y >= 1.0f;
	if(c == '>' || c == '<' || c == '^' || c == ',') return y < 0.0f || y >= 1.0f;
	if(PWN_IS_PORTAL(c))
	{
		int pi = PWN_PORTAL_INDEX(c);
		if(pi < 0 || pi >= lv->nportals) return 1;
		return lv->portals[pi].x2 != -1 ? 0 : 1;
	}
	return 1;
}

void pwn_player_init(pwn_player *pl, const pwn_level *lv)
{
	memset(pl, 0, sizeof(*pl));
	pwn_cam_spawn(lv->sx, lv->sz, pl->cam);
}

void pwn_player_tick(pwn_player *pl, const pwn_level *lv, const pwn_keys *k, float tdiff)
{
	float *cam = pl->cam, *w = pl->cam + 12;
	int i;

	pwn_mat4_roty(cam, tdiff*3.0f*(k->turnleft - k->turnright));   /* main.c:188 */

	int cx1 = (int)w[0], cz1 = (int)w[2];

	float vel[4];
	float fwd = tdiff*5.0f*(k->moveforward - k->moveback);
	float side = tdiff*5.0f*(k->moveleft - k->moveright);
	for(i = 0; i < 4; i++) vel[i] = cam[8+i]*fwd + cam[0+i]*side;
	for(i = 0; i < 4; i++) w[i] += vel[i];

	float px1 = w[0], py1 = w[1], pz1 = w[2];

	/* push-backs, main.c:215-264 */
	int gx1 = (vel[0] < 0.0f ? -1 : 1);
	int gz1 = (vel[2] < 0.0f ? -1 : 1);
	float bx1 = px1 + gx1*PLAYER_BBOX;
	float bz1 = pz1 + gz1*PLAYER_BBOX;
	int bcx = (int)bx1, bcz = (int)bz1;
	int32_t oldcell = pwn_level_get_cell(lv, cx1, cz1);

	if(cx1 != bcx && cz1 != bcz)
	{
		int solx = celltype_is_solid(lv, pwn_level_get_cell(lv, bcx, cz1), oldcell, py1);
		int solz = celltype_is_solid(lv, pwn_level_get_cell(lv, cx1, bcz), oldcell, py1);
		int solc = celltype_is_solid(lv, pwn_level_get_cell(lv, bcx, bcz), oldcell, py1);
		if(solx && solz)
		{
			w[0] = cx1 + 0.5f + (0.5f-PLAYER_BBOX)*gx1;
			w[2] = cz1 + 0.5f + (0.5f-PLAYER_BBOX)*gz1;
		}
		else if(solx) w[0] = cx1 + 0.5f + (0.5f-PLAYER_BBOX)*gx1;
		else if(solz) w[2] = cz1 + 0.5f + (0.5f-PLAYER_BBOX)*gz1;
		else if(solc) w[2] = cz1 + 0.5f + (0.5f-PLAYER_BBOX)*gz1;
	}
	else if(cx1 != bcx)
	{
		if(celltype_is_solid(lv, pwn_level_get_cell(lv, bcx, bcz), oldcell, py1))
			w[0] = cx1 + 0.5f + (0.5f-PLAYER_BBOX)*gx1;
	}
	else if(cz1 != bcz)
	{
		if(celltype_is_solid(lv, pwn_level_get_cell(lv, bcx, bcz), oldcell, py1))
			w[2] = cz1 + 0.5f + (0.5f-PLAYER_BBOX)*gz1;
	}

	/* gravity, main.c:266-274 */
	for(i = 0; i < 4; i++) w[i] += pl->gravity[i];
	pl->gravity[1] -= 3.0f*tdiff*tdiff;
	if(w[1] < 0.4f) { w[1] = 0.4f; pl->gravity[1] = 0.0f; }

	/* cell change, main.c:276-379 */
	int cx2 = (int)w[0], cz2 = (int)w[2];
	if(cx1 != cx2 || cz1 != cz2)
	{
		int32_t c1 = pwn_level_get_cell(lv, cx1, cz1);
		int32_t c2 = pwn_level_get_cell(lv, cx2, cz2);
		if((c1 == '#' || c1 == '&') && c2 == '"') w[1] -= 1.0f;
		else if(c1 == '"' && (c2 == '#' || c2 == '&')) w[1] += 1.0f;
		else if(PWN_IS_PORTAL(c2) && PWN_PORTAL_INDEX(c2) < lv->nportals)
		{
			const pwn_portal *pm = &lv->portals[PWN_PORTAL_INDEX(c2)];
			int rot = 0;
			float rx = w[0], rz = w[2];
			float rcx = cx2, rcz = cz2;
			float rvxx = cam[0], rvxz = cam[2], rvzx = cam[8], rvzz = cam[10];
			if(pm->x2 == -1) { }
			else if(pm->x1 == cx2 && pm->z1 == cz2)
			{
				rx += (cx2-cx1) + (pm->x2-pm->x1);
				rz += (cz2-cz1) + (pm->z2-pm->z1);
				rcx = pm->x2; rcz = pm->z2;
				rot = (-pm->rot12) & 3;
			}
			else if(pm->x2 == cx2 && pm->z2 == cz2)
			{
				rx += (cx2-cx1) - (pm->x2-pm->x1);
				rz += (cz2-cz1) - (pm->z2-pm->z1);
				rcx = pm->x1; rcz = pm->z1;
				rot = pm->rot12 & 3;
			}
			float trx = rx, trz = rz, txx = rvxx, txz = rvxz, tzx = rvzx, tzz = rvzz;
			switch(rot)
			{
				case 1:
					rx = (rcx + 0.5f) + (trz-(rcz + 0.5f));
					rz = (rcz + 0.5f) - (trx-(rcx + 0.5f));
					rvxx = txz; rvxz = -txx; rvzx = tzz; rvzz = -tzx;
					break;
				case 2:
					rx = (rcx + 0.5f)*2.0f - rx;
					rz = (rcz + 0.5f)*2.0f - rz;
					rvxx = -txx; rvxz = -txz; rvzx = -tzx; rvzz = -tzz;
					break;
				case 3:
					rx = (rcx + 0.5f) - (trz-(rcz + 0.5f));
					rz = (rcz + 0.5f) + (trx-(rcx + 0.5f));
					rvxx = -txz; rvxz = txx; rvzx = -tzz; rvzz = tzx;
					break;
				default: break;
			}
			w[0] = rx; w[2] = rz;
			cam[0] = rvxx; cam[2] = rvxz; cam[8] = rvzx; cam[10] = rvzz;
		}
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
