/*
 * pwn_camera.h -- camera state of the headless harness: the `mat4 cam` of main.c:47,60-64
 * and a deterministic restatement of one game tick of player motion (main.c:188-379),
 * driven by a fixed key schedule instead of SDL events (north_star item 4).
 */
#ifndef PWN_CAMERA_H
#define PWN_CAMERA_H

#include "pwn_level.h"

#ifdef __cplusplus
extern "C" {
#endif

/* key state of one tick: main.c:70-77 */
typedef struct pwn_keys
{
	int turnleft, turnright, moveforward, moveback, moveleft, moveright;
} pwn_keys;

typedef struct pwn_player
{
	float cam[16];      /* rows x y z w */
	float gravity[4];   /* main.c:48,60 */
} pwn_player;

/* mat4_iden + spawn position (main.c:60-64) */
void pwn_cam_spawn(int sx, int sz, float cam[16]);
/* mat4_roty (util.h:96-110) */
void pwn_mat4_roty(float cam[16], float ang);
void pwn_player_init(pwn_player *pl, const pwn_level *lv);
/* one tick of main.c:188-379 with time step tdiff */
void pwn_player_tick(pwn_player *pl, const pwn_level *lv, const pwn_keys *k, float tdiff);

/* the fixed script: nframes poses at 60 Hz from the spawn; returns 0 */
int pwn_flythrough(const pwn_level *lv, int nframes, pwn_pose *out);

#ifdef __cplusplus
}
#endif
#endif
