/* pwn_image.c -- frame dumps for visual diffing. */
#include "pwn_image.h"

#include <stdio.h>
#include <stdlib.h>

int pwn_save_ppm(const char *path, const uint32_t *sbuf, int w, int h, int rwidth)
{
	int x, y;
	FILE *fp;
	unsigned char *row;
	if(path == NULL || sbuf == NULL || w <= 0 || h <= 0 || rwidth < w) return -1;
	fp = fopen(path, "wb");
	if(fp == NULL) return -2;
	row = malloc((size_t)w*3);
	if(row == NULL) { fclose(fp); return -3; }
	fprintf(fp, "P6\n%d %d\n255\n", w, h);
	for(y = 0; y < h; y++)
	{
		const uint32_t *s = sbuf + (size_t)y*rwidth;
		for(x = 0; x < w; x++)
		{
			row[3*x + 0] = (unsigned char)(s[x] >> 16);
			row[3*x + 1] = (unsigned char)(s[x] >> 8);
			row[3*x + 2] = (unsigned char)(s[x]);
		}
		if(fwrite(row, 3, (size_t)w, fp) != (size_t)w) { free(row); fclose(fp); return -4; }
	}
	free(row);
	return fclose(fp) == 0 ? 0 : -4;
}

int pwn_frame_diff(const uint32_t *a, const uint32_t *b, uint32_t *out, int w, int h, int rwidth, int gain)
{
	int x, y, c;
	if(a == NULL || b == NULL || out == NULL || w <= 0 || h <= 0 || rwidth < w || gain < 1) return -1;
	for(y = 0; y < h; y++)
	for(x = 0; x < w; x++)
	{
		const size_t i = (size_t)y*rwidth + x;
		uint32_t o = 0;
		for(c = 0; c < 3; c++)
		{
			int d = (int)((a[i] >> (8*c)) & 255) - (int)((b[i] >> (8*c)) & 255);
			if(d < 0) d = -d;
			d *= gain;
			if(d > 255) d = 255;
			o |= (uint32_t)d << (8*c);
		}
		out[i] = o;
	}
	return 0;
}
