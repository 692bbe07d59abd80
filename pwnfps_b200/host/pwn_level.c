/*
 * pwn_level.c -- level.txt loader and portal derivation, restating level.h:83-228 and
 * util.h:128-158 for the pwn_level model (see pwn_level.h).  Host code, runs once per
 * level; checked against the reference's own loader in tests/test_level.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pwn_level.h"

pwn_level *pwn_level_new(int w, int h, int nportals)
{
	int i;
	if(w <= 0 || h <= 0 || nportals < 0) return NULL;
	pwn_level *lv = calloc(1, sizeof(pwn_level));
	if(lv == NULL) return NULL;
	lv->w = w; lv->h = h; lv->nportals = nportals;
	lv->cells = malloc(sizeof(int32_t)*(size_t)w*h);
	lv->portals = calloc((size_t)(nportals > 0 ? nportals : 1), sizeof(pwn_portal));
	if(lv->cells == NULL || lv->portals == NULL) { pwn_level_free(lv); return NULL; }
	for(i = 0; i < w*h; i++) lv->cells[i] = '.';
	for(i = 0; i < nportals; i++)
	{
		lv->portals[i].x1 = lv->portals[i].x2 = -1;
		lv->portals[i].z1 = lv->portals[i].z2 = 0;
		lv->portals[i].c1 = ';';
		lv->portals[i].c2 = ';';
	}
	return lv;
}

void pwn_level_free(pwn_level *lv)
{
	if(lv == NULL) return;
	free(lv->cells);
	free(lv->portals);
	free(lv);
}

int32_t pwn_level_get_cell(const pwn_level *lv, int cx, int cz)
{
	if(cx < 0 || cx >= lv->w) cx = 0;
	if(cz < 0 || cz >= lv->h) cz = 0;
	return lv->cells[cz*lv->w + cx];
}

/* util.h:128-138 */
static int celltype_is_free(int32_t c)
{
	if(c == ';' || c == '$' || c == '"') return 1;
	if(c == '#' || c == '&') return 1;
	if(c == '>' || c == '<' || c == '^' || c == ',') return 1;
	return 0;
}

/*
 * The reference indexes data[z][x+-1] / data[z+-1][x] unchecked (util.h:140-149); on its
 * contiguous 64x64 array an x overflow lands in the neighbouring row.  Same here through
 * the flat index; anything outside the array counts as not free.
 */
static int32_t cell_flat(const pwn_level *lv, int x, int z)
{
	long idx = (long)z*lv->w + x;
	if(idx < 0 || idx >= (long)lv->w*lv->h) return '.';
	return lv->cells[idx];
}

static int find_free_dir_2d(const pwn_level *lv, int x, int z)
{
	if(celltype_is_free(cell_flat(lv, x+1, z))) return PWN_FXP;
	if(celltype_is_free(cell_flat(lv, x, z+1))) return PWN_FZP;
	if(celltype_is_free(cell_flat(lv, x-1, z))) return PWN_FXN;
	if(celltype_is_free(cell_flat(lv, x, z-1))) return PWN_FZN;
	return PWN_FXP;   /* "stuff it" (util.h:148) */
}

static int32_t cell_in_dir(const pwn_level *lv, int x, int z, int d)
{
	switch(d)
	{
		case PWN_FXP: return cell_flat(lv, x+1, z);
		case PWN_FZP: return cell_flat(lv, x, z+1);
		case PWN_FXN: return cell_flat(lv, x-1, z);
		default: return cell_flat(lv, x, z-1);
	}
}

void pwn_level_derive(pwn_level *lv)
{
	int i;
	for(i = 0; i < lv->nportals; i++)
	{
		pwn_portal *pm = &lv->portals[i];
		if(pm->x2 == -1) continue;
		int d1 = find_free_dir_2d(lv, pm->x1, pm->z1);
		int d2 = find_free_dir_2d(lv, pm->x2, pm->z2);
		pm->rot12 = (d2 - d1 + 2) & 3;
		pm->c1 = cell_in_dir(lv, pm->x1, pm->z1, d1);
		pm->c2 = cell_in_dir(lv, pm->x2, pm->z2, d2);
	}
}

static void portal_register(pwn_portal *pm, int x, int z)
{
	if(pm->x1 == -1) { pm->x1 = x; pm->z1 = z; }
	else if(pm->x2 == -1) { pm->x2 = x; pm->z2 = z; }
}

int pwn_level_link(pwn_level *lv)
{
	int x, z, i;
	for(i = 0; i < lv->nportals; i++)
	{
		lv->portals[i].x1 = lv->portals[i].x2 = -1;
		lv->portals[i].c1 = lv->portals[i].c2 = ';';
	}
	for(z = 0; z < lv->h; z++)
	for(x = 0; x < lv->w; x++)
	{
		int32_t c = lv->cells[z*lv->w + x];
		if(PWN_IS_PORTAL(c))
		{
			int pi = PWN_PORTAL_INDEX(c);
			if(pi < 0 || pi >= lv->nportals) return -1;
			portal_register(&lv->portals[pi], x, z);
		}
	}
	pwn_level_derive(lv);
	return 0;
}

/* level.h:107-228 over a byte stream */
typedef struct { const unsigned char *p, *end; } rd_t;
static int rd_getc(rd_t *r) { return r->p < r->end ? *r->p++ : -1; }

static pwn_level *parse_extended(const char *text, size_t len);
static int32_t cell_as_written(const pwn_level *lv, int x, int z);

/*
 * What one character of the grid does, in both text formats (the rules of level.h:137-178): '*' is the spawn, on a plain
 * cell; a lower-case letter registers in its own pair and then stands for the NEXT letter; a capital registers in its pair.
 * The cell keeps the byte as a signed char, as `char data[64][64]` does (defs.h:105).
 */
static int32_t place_char(pwn_level *lv, int x, int z, int c)
{
	if(c == '*') { lv->sx = x; lv->sz = z; c = ';'; }
	if(c >= 'a' && c < 'z') { portal_register(&lv->portals[c - 'a'], x, z); c += 'A' - 'a' + 1; }
	if(c >= 'A' && c <= 'Z') portal_register(&lv->portals[c - 'A'], x, z);
	return (int32_t)(signed char)c;
}

/*
 * The rows of the grid from a byte stream, one reader for both formats.  A line break ends a row that has cells and is
 * skipped otherwise (blank lines, the LF of a CRLF); the stream may end anywhere; cells never reached keep their '.'.
 * level.txt: a row of w characters runs on into the next row without a line break (level.h:118-182 has no notion of
 * lines).  Extended: the rest of an over-long line is dropped, and a line beginning "@ " ends the grid.
 */
static void read_grid(rd_t *r, pwn_level *lv, int extended)
{
	int x = 0, z = 0;
	while(z < lv->h)
	{
		const int c = rd_getc(r);
		if(c == -1) return;
		if(c == '\r' || c == '\n') { if(x > 0) { x = 0; z++; } continue; }
		if(extended && x == 0 && c == '@' && r->p < r->end && *r->p == ' ') { r->p--; return; }
		lv->cells[z*lv->w + x] = place_char(lv, x, z, c);
		if(++x == lv->w)
		{
			x = 0; z++;
			if(extended) while(r->p < r->end && *r->p != '\n') r->p++;
		}
	}
}

pwn_level *pwn_level_parse(const char *text, size_t len)
{
	rd_t r;
	if(len >= 6 && memcmp(text, "!pwnlv", 6) == 0) return parse_extended(text, len);
	r.p = (const unsigned char *)text; r.end = r.p + len;
	pwn_level *lv = pwn_level_new(64, 64, 26);
	if(lv == NULL) return NULL;
	read_grid(&r, lv, 0);
	pwn_level_derive(lv);
	return lv;
}

/*
 * The extended text format (pwn_level.h): the same grid alphabet and the same per-character rules
 * as level.h:107-228, on a w x h grid, with '@' cells named by trailing "@ id x z" lines.
 */
static pwn_level *parse_extended(const char *text, size_t len)
{
	int w = 0, h = 0, np = 0, used = 0;
	char *copy = malloc(len + 1);
	if(copy == NULL) return NULL;
	memcpy(copy, text, len);
	copy[len] = 0;
	if(sscanf(copy, "!pwnlv %d %d %d%n", &w, &h, &np, &used) != 3 || w < 1 || h < 1 || w > 4096 || h > 4096 || np < 26 || np > (1 << 22))
	{ free(copy); return NULL; }
	pwn_level *lv = pwn_level_new(w, h, np);
	if(lv == NULL) { free(copy); return NULL; }
	rd_t r;
	r.p = (const unsigned char *)copy + used; r.end = (const unsigned char *)copy + len;
	while(r.p < r.end && *r.p != '\n') r.p++;      /* rest of the header line */
	if(r.p < r.end) r.p++;
	read_grid(&r, lv, 1);
	/* "@ id x z": the '@' cell at (x, z) belongs to portal pair id >= 26; registration in line order */
	{
		const char *q = (const char *)r.p;
		while(q < copy + len)
		{
			int id, px, pz;
			if(*q == '@' && sscanf(q, "@ %d %d %d", &id, &px, &pz) == 3)
			{
				if(id < 26 || id >= np || px < 0 || pz < 0 || px >= w || pz >= h || lv->cells[pz*w + px] != '@')
				{ pwn_level_free(lv); free(copy); return NULL; }
				lv->cells[pz*w + px] = PWN_PORTAL_CODE(id);
				portal_register(&lv->portals[id], px, pz);
			}
			while(q < copy + len && *q != '\n') q++;
			if(q < copy + len) q++;
		}
	}
	free(copy);
	pwn_level_derive(lv);
	return lv;
}

int pwn_level_save_txtx(const pwn_level *lv, const char *fname)
{
	int x, z;
	for(z = 0; z < lv->h; z++)
	for(x = 0; x < lv->w; x++)
	{
		int32_t c = lv->cells[z*lv->w + x];
		if(c >= 0x100) { if(PWN_PORTAL_INDEX(c) >= lv->nportals) return -1; continue; }
		if(c < 32 || c > 126 || c == '*' || c == '@' || (c >= 'a' && c <= 'z')) return -1;
	}
	FILE *fp = fopen(fname, "wb");
	if(fp == NULL) return -1;
	fprintf(fp, "!pwnlv %d %d %d\n", lv->w, lv->h, lv->nportals < 26 ? 26 : lv->nportals);
	for(z = 0; z < lv->h; z++)
	{
		for(x = 0; x < lv->w; x++)
		{
			int32_t c = lv->cells[z*lv->w + x];
			fputc(c >= 0x100 ? '@' : cell_as_written(lv, x, z), fp);
		}
		fputc('\n', fp);
	}
	/* row-major, so that reading registers the ends in the order pwn_level_link does */
	for(z = 0; z < lv->h; z++)
	for(x = 0; x < lv->w; x++)
	{
		int32_t c = lv->cells[z*lv->w + x];
		if(c >= 0x100) fprintf(fp, "@ %d %d %d\n", PWN_PORTAL_INDEX(c), x, z);
	}
	return fclose(fp) == 0 ? 0 : -1;
}

pwn_level *pwn_level_load(const char *fname)
{
	FILE *fp = fopen(fname, "rb");
	if(fp == NULL) return NULL;
	size_t cap = 1 << 16, len = 0;
	char *buf = malloc(cap);
	if(buf == NULL) { fclose(fp); return NULL; }
	for(;;)
	{
		size_t n = fread(buf + len, 1, cap - len, fp);
		len += n;
		if(n == 0) break;
		if(len == cap) { cap *= 2; buf = realloc(buf, cap); if(buf == NULL) { fclose(fp); return NULL; } }
	}
	fclose(fp);
	pwn_level *lv = pwn_level_parse(buf, len);
	free(buf);
	return lv;
}

/*
 * The character a cell was loaded from.  level.h:144-161 turns a lower-case letter into the NEXT
 * upper-case one after registering the cell as an end of its own pair: a cell 'L'+1 that the table
 * lists as an end of pair L was written in lower case.
 */
static int32_t cell_as_written(const pwn_level *lv, int x, int z)
{
	int32_t c = lv->cells[z*lv->w + x];
	if(c >= 'B' && c <= 'Z' && c - 'B' < lv->nportals)
	{
		const pwn_portal *pm = &lv->portals[c - 'B'];
		if((pm->x1 == x && pm->z1 == z) || (pm->x2 == x && pm->x2 != -1 && pm->z2 == z)) return (c - 'B') + 'a';
	}
	if(x == lv->sx && z == lv->sz && c == ';') return '*';
	return c;
}

int pwn_level_save_txt(const pwn_level *lv, const char *fname)
{
	int x, z;
	if(lv->w > 64 || lv->h > 64) return -1;
	for(z = 0; z < lv->h; z++)
	for(x = 0; x < lv->w; x++)
	{
		int32_t c = lv->cells[z*lv->w + x];
		if(c < 32 || c > 126 || c == '*' || (c >= 'a' && c <= 'z')) return -1;
	}
	FILE *fp = fopen(fname, "wb");
	if(fp == NULL) return -1;
	for(z = 0; z < lv->h; z++)
	{
		for(x = 0; x < lv->w; x++)
		{
			fputc(cell_as_written(lv, x, z), fp);
		}
		fputc('\n', fp);
	}
	fclose(fp);
	return 0;
}

int pwn_level_save_bin(const pwn_level *lv, const char *fname)
{
	FILE *fp = fopen(fname, "wb");
	if(fp == NULL) return -1;
	int32_t hdr[5] = {lv->w, lv->h, lv->nportals, lv->sx, lv->sz};
	int ok = fwrite("PWNLV001", 1, 8, fp) == 8
		&& fwrite(hdr, sizeof(int32_t), 5, fp) == 5
		&& fwrite(lv->cells, sizeof(int32_t), (size_t)lv->w*lv->h, fp) == (size_t)lv->w*lv->h
		&& fwrite(lv->portals, sizeof(pwn_portal), (size_t)lv->nportals, fp) == (size_t)lv->nportals;
	fclose(fp);
	return ok ? 0 : -1;
}

pwn_level *pwn_level_load_bin(const char *fname)
{
	FILE *fp = fopen(fname, "rb");
	char magic[8];
	int32_t hdr[5];
	if(fp == NULL) return NULL;
	if(fread(magic, 1, 8, fp) != 8 || memcmp(magic, "PWNLV001", 8) != 0 || fread(hdr, sizeof(int32_t), 5, fp) != 5
		|| hdr[0] <= 0 || hdr[1] <= 0 || hdr[0] > 16384 || hdr[1] > 16384 || hdr[2] < 0 || hdr[2] > (1 << 24))
	{ fclose(fp); return NULL; }
	pwn_level *lv = pwn_level_new(hdr[0], hdr[1], hdr[2]);
	if(lv == NULL) { fclose(fp); return NULL; }
	lv->sx = hdr[3]; lv->sz = hdr[4];
	if(fread(lv->cells, sizeof(int32_t), (size_t)lv->w*lv->h, fp) != (size_t)lv->w*lv->h
		|| fread(lv->portals, sizeof(pwn_portal), (size_t)lv->nportals, fp) != (size_t)lv->nportals)
	{ fclose(fp); pwn_level_free(lv); return NULL; }
	fclose(fp);
	return lv;
}
