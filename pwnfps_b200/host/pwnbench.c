/*
 * pwnbench.c -- the headless C host: what main.c's game loop (main.c:42-381) becomes when the
 * SDL window, the Lua VM and the keyboard are replaced by a fixed camera-path script
 * (north_star item 4).  Pure C against the C ABI (include/pwntrace.h) and the host library
 * (pwn_level.h, pwn_camera.h); no Python, no torch.
 *
 *   pwnbench level.txt|level.pwnlv [frames=600] [width=1920] [height=1080] [blur=1] [device=0] [spheres=0]
 *
 * Loads the level with the level.txt loader (level.h:107 restated in pwn_level.c), runs the
 * scripted fly-through (pwn_camera.c) through pwn_trace_poses into pinned host memory -- the
 * frame loop of main.c:93-109 -- and prints the throughput and one checksum per 100th frame.
 * With spheres=1 the loop is the reference's frame by frame: on_tick moves the 14 spheres of
 * game.lua (pwn_game.c), level_prepare_render becomes pwn_parts_upload, trace_screen_centred
 * becomes pwn_trace_screen (main.c:95,107,127-140).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pwntrace.h"
#include "pwn_level.h"
#include "pwn_camera.h"
#include "pwn_game.h"

#define CHECK(call) do { int rc_ = (call); if(rc_ != PWN_OK) { \
	fprintf(stderr, "pwnbench: %s failed (%d): %s\n", #call, rc_, pwn_last_error()); return 2; } } while(0)

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9*(double)ts.tv_nsec;
}

static uint64_t fnv1a64(const uint32_t *w, size_t n)
{
	uint64_t h = 0xcbf29ce484222325ull;
	size_t i;
	for(i = 0; i < n; i++) { h ^= w[i]; h *= 0x100000001b3ull; }
	return h;
}

int main(int argc, char **argv)
{
	if(argc < 2)
	{
		fprintf(stderr, "usage: pwnbench level.txt|level.pwnlv [frames] [width] [height] [blur] [device]\n");
		return 1;
	}
	const char *path = argv[1];
	int frames = argc > 2 ? atoi(argv[2]) : 600;
	int w = argc > 3 ? atoi(argv[3]) : 1920, h = argc > 4 ? atoi(argv[4]) : 1080;
	unsigned flags = (argc > 5 ? atoi(argv[5]) : 1) ? PWN_F_BLUR : 0u;
	int device = argc > 6 ? atoi(argv[6]) : 0;
	int spheres = argc > 7 ? atoi(argv[7]) : 0;
	size_t len = strlen(path);

	pwn_level *lv = (len > 6 && strcmp(path + len - 6, ".pwnlv") == 0) ? pwn_level_load_bin(path) : pwn_level_load(path);
	if(lv == NULL) { fprintf(stderr, "err loading level: %s\n", path); return 1; }
	printf("level %dx%d, %d portal slots, spawn %3i %3i\n", lv->w, lv->h, lv->nportals, lv->sx, lv->sz);

	pwn_pose *poses = malloc(sizeof(pwn_pose)*(size_t)frames);
	pwn_flythrough(lv, frames, poses);

	pwn_ctx *ctx = NULL;
	CHECK(pwn_create(&ctx, device));
	CHECK(pwn_math_tables_from_host(ctx));
	CHECK(pwn_level_upload(ctx, lv->w, lv->h, lv->cells, lv->nportals, lv->portals));
	CHECK(pwn_frame_config(ctx, w, h, 0));

	const int chunk = frames < 100 ? frames : 100;
	const size_t fpix = (size_t)w*h;
	uint32_t *host = NULL;
	CHECK(pwn_host_alloc((void **)&host, fpix*4*(size_t)chunk));

	CHECK(pwn_trace_poses(ctx, chunk < 8 ? chunk : 8, poses, w, h, flags, host, NULL));   /* warm-up */
	double t0 = now_s();
	int f;
	if(!spheres)
	{
		for(f = 0; f < frames; f += chunk)
		{
			int n = frames - f < chunk ? frames - f : chunk;
			CHECK(pwn_trace_poses(ctx, n, poses + f, w, h, flags, host, NULL));
			printf("frame %4d fnv1a64 %016llx\n", f, (unsigned long long)fnv1a64(host, fpix));
		}
	}
	else
	{
		pwn_game game;
		pwn_game_init(&game);
		for(f = 0; f < frames; f++)
		{
			CHECK(pwn_parts_upload(ctx, PWN_GAME_NSPH, game.sph));                       /* level_prepare_render, main.c:95 */
			CHECK(pwn_trace_screen(ctx, poses[f].cam, w, h, poses[f].sec_current, flags, host, NULL));   /* main.c:107 */
			if(f % 100 == 0) printf("frame %4d fnv1a64 %016llx\n", f, (unsigned long long)fnv1a64(host, fpix));
			pwn_game_tick(&game, lv, poses[f].sec_current + 1.0/60.0, 1.0/60.0);         /* on_tick, main.c:127-140 */
		}
	}
	double dt = now_s() - t0;
	uint64_t launches = 0;
	pwn_launch_count(ctx, &launches);
	printf("%d frames %dx%d blur=%d in %.3f s: %.1f frames/s, %.1f Mrays/s (host buffers, %llu kernel launches)\n",
		frames, w, h, flags ? 1 : 0, dt, frames/dt, frames*(double)fpix/dt/1e6, (unsigned long long)launches);

	pwn_host_free(host);
	pwn_destroy(ctx);
	free(poses);
	pwn_level_free(lv);
	return 0;
}
