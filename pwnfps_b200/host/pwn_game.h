/*
 * pwn_game.h -- the sphere script of the reference restated natively.
 *
 * The reference animates 14 spheres from Lua (game.lua:1-88 through the script.h bindings
 * obj_new / obj_set / level_get); Lua 5.1 is not part of this image, so the headless host
 * carries the same logic in C: the sphere table (game.lua:2-20), their creation at
 * (9.5, 0.3, 5.5) (game.lua:22-30) and on_tick (game.lua:33-87): blink of sphere 2, rotation
 * of the cluster by pi*sec_current, and the walk of the cluster through the level, turning at
 * walls.  Unpinned: there is no Lua here to run game.lua against.
 */
#ifndef PWN_GAME_H
#define PWN_GAME_H

#include "pwn_level.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PWN_GAME_NSPH 14

typedef struct pwn_game
{
	double opos[PWN_GAME_NSPH][8];   /* x y z r cb cg cr refl (Lua numbers are doubles) */
	double obx, oby, obz, obvx, obvz;
	pwn_sphere sph[PWN_GAME_NSPH];   /* what obj_set wrote into lv->objs (script.h:18-33) */
} pwn_game;

void pwn_game_init(pwn_game *g);
/* on_tick(sec_current, sec_delta), game.lua:33-87 */
void pwn_game_tick(pwn_game *g, const pwn_level *lv, double sec_current, double sec_delta);

#ifdef __cplusplus
}
#endif
#endif
