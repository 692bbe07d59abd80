/*
 * pwn_level.h -- host-side level model of the B200 build: the part of the reference's
 * `level` (defs.h:96-111) the trace reads, without the 64x64 / 26-letter limits, plus
 * a loader for the reference's level.txt format (level.h:107-228).
 *
 * A host that keeps the reference's own level.h loader does not need this file: it
 * hands its `level` to the library through pwn_screen.h.  This loader exists for the
 * headless bench harness and for levels beyond the reference's limits.
 */
#ifndef PWN_LEVEL_H
#define PWN_LEVEL_H

#include <stdint.h>
#include "pwntrace.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pwn_level
{
	int32_t w, h;            /* grid, reference: 64 x 64 (defs.h:105) */
	int32_t *cells;          /* [h][w] codes, see pwntrace.h "Cell codes" */
	int32_t nportals;        /* reference: 26 (defs.h:103) */
	pwn_portal *portals;
	int32_t sx, sz;          /* spawn cell (defs.h:101) */
} pwn_level;

/* level_new (level.h:85-105): all '.', portals unpaired with c1 = c2 = ';' */
pwn_level *pwn_level_new(int w, int h, int nportals);
void pwn_level_free(pwn_level *lv);

/* level_load (level.h:107-228): 64x64 grid, 26 portals, same parsing quirks. NULL on I/O error. */
pwn_level *pwn_level_load(const char *fname);
/* same parser over a memory buffer */
pwn_level *pwn_level_parse(const char *text, size_t len);

/*
 * Portal pairing for levels built in memory: scans the grid row-major and registers
 * every portal cell in its slot the way the loader does (first two occurrences,
 * level.h:163-178), then derives rot12/c1/c2 (level.h:194-221).
 */
int pwn_level_link(pwn_level *lv);

/* level.h:194-221 alone: rot12, c1, c2 of every paired portal from the facing free cells */
void pwn_level_derive(pwn_level *lv);

/* util.h:151-158 */
int32_t pwn_level_get_cell(const pwn_level *lv, int cx, int cz);

/* writes the level in level.txt format; fails (-1) if it does not fit 64x64 / 26 letters */
int pwn_level_save_txt(const pwn_level *lv, const char *fname);

/*
 * Extended text format for levels beyond the reference's limits, read by
 * pwn_level_load / pwn_level_parse when the file starts with "!pwnlv":
 *   !pwnlv <w> <h> <nportals>          nportals >= 26
 *   <h rows of <= w characters>        the level.txt alphabet and rules (level.h:107-228: '*' spawn,
 *                                      letters, the lower-case double registration, short rows,
 *                                      blank lines); '@' = a cell of a portal pair with id >= 26
 *   @ <id> <x> <z>                     one line per '@' cell; the first two of an id are its ends
 *                                      (a line starting "@ " also ends a grid of fewer than h rows)
 * A level that fits 64x64 / 26 letters should be kept as plain level.txt (pwn_level_save_txt),
 * which the reference's own loader reads.  pwn_level_save_txtx writes this format for any level.
 */
int pwn_level_save_txtx(const pwn_level *lv, const char *fname);

/*
 * Binary level file for levels beyond the reference's limits: little endian,
 *   char magic[8] = "PWNLV001"; int32 w, h, nportals, sx, sz;
 *   int32 cells[h*w]; pwn_portal portals[nportals] (as derived: rot12, c1, c2 included).
 * A level that fits 64x64 / 26 letters should be kept as level.txt (pwn_level_save_txt), which
 * the reference's own loader reads.
 */
int pwn_level_save_bin(const pwn_level *lv, const char *fname);
pwn_level *pwn_level_load_bin(const char *fname);

#ifdef __cplusplus
}
#endif
#endif
