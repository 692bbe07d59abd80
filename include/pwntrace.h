/*
 * pwntrace.h -- C ABI of libpwntrace.so: the B200 (sm_100a) implementation of the
 * pwnfps per-pixel portal ray cast.
 *
 * This is the drop-in boundary for ONE path of fanzyflani/pwnfps:
 *
 *     mainloop()                       main.c:93-109
 *       level_prepare_render(lv)       main.c:95    level.h:64-81
 *       trace_screen_centred(lv, 0, 0, rwidth, rheight, &cam)
 *                                      main.c:107   screen.h:31-124
 *         trace_ray_prelude(...)       screen.h:1-28      (per 32-row band)
 *           trace_ray(...)             trace.h:186-679    (per pixel, +2 bounces)
 *         POSTPROC_BLUR pass           screen.h:69-123
 *
 * The reference has no FFI/plugin layer: every function above is defined in a
 * header #included into the single TU main.c (main.c:36-40).  The seam is
 * therefore source level: the host includes pwnfps_b200/host/pwn_screen.h in
 * place of trace.h + screen.h; that shim defines the same
 * trace_screen_centred() and forwards to the entry points below
 * (INTEGRATION.md shows the two-line change to main.c).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no CUDA or torch types.
 *   - every call returns 0 on success or a negative PWN_E_* code; the message
 *     for the last failure on the calling thread is pwn_last_error().  The
 *     reference trace has no error channel (void + assert, screen.h:35-36);
 *     this library never calls exit()/abort().
 *   - one pwn_ctx drives one GPU.  The caller is single threaded per context, as
 *     the reference's main thread is (main.c:93-109); the library owns all
 *     parallelism (streams, kernels).
 *   - there is NO CPU fallback: without a usable sm_100 device pwn_create fails.
 */
#ifndef PWNTRACE_H
#define PWNTRACE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PWN_ABI_VERSION 1

enum
{
	PWN_OK = 0,
	PWN_E_ARG = -1,        /* bad argument (null pointer, size, range)           */
	PWN_E_CUDA = -2,       /* CUDA runtime failure; pwn_last_error has the text  */
	PWN_E_NODEVICE = -3,   /* no sm_100 device / device index out of range       */
	PWN_E_STATE = -4,      /* call order: level, tables or frame not set up yet  */
	PWN_E_LIMIT = -5,      /* level exceeds a device-layout limit                */
	PWN_E_HOSTMATH = -6,   /* host rcpps/rsqrtps is not a pure mantissa-bucket table */
	PWN_E_UNSUPPORTED = -7,/* input outside what the reference host can produce  */
	PWN_E_PEER = -8        /* sharded frame: a peer GPU missed the device barrier */
};

/* defs.h:25-33 */
enum { PWN_FXP = 0, PWN_FZP, PWN_FXN, PWN_FZN, PWN_FYP, PWN_FYN };

/*
 * Cell codes.  The reference stores one char per cell (defs.h:105) and names
 * portal pair p by the letter 'A'+p (defs.h:103).  Here a cell is an int32:
 * values < 0x100 are the reference's chars unchanged; PWN_PORTAL_CODE(p)
 * names portal pair p for p >= 26, which the reference cannot express.
 */
#define PWN_PORTAL_CODE(p) ((p) < 26 ? 'A' + (p) : 0x100 + (p))
#define PWN_IS_PORTAL(c) (((c) >= 'A' && (c) <= 'Z') || (c) >= 0x100)
#define PWN_PORTAL_INDEX(c) ((c) >= 0x100 ? (c) - 0x100 : (c) - 'A')

/* defs.h:87-94 `portal` (double1/double2 are never read by the trace) */
typedef struct pwn_portal
{
	int32_t x1, z1;
	int32_t x2, z2;     /* x2 == -1: unpaired, renders as wall (trace.h:512-526) */
	int32_t rot12;
	int32_t c1, c2;     /* cell code in front of end 1 / end 2 (level.h:204-218) */
} pwn_portal;

/* defs.h:73-79 `part.sph` */
typedef struct pwn_sphere
{
	float r, refl;
	float pos[4];       /* x y z w; w must equal the camera's w (1.0f, script.h:28) */
	float col[4];       /* b g r a (defs.h:40) */
} pwn_sphere;

/* per pixel, per segment (0 = primary, 1/2 = mirror bounces); test/diagnostic output */
typedef struct pwn_ids
{
	int32_t cell;       /* cz*w + cx when the segment terminated; -1 = never ran    */
	int16_t ldir;       /* PWN_FXP..PWN_FYN of the face hit; -1 = sphere            */
	int16_t kind;       /* 0 wall/floor/ceiling, 1 sphere, 2 walk exhausted (trace.h:677) */
	int32_t nstep;      /* walk-loop iterations entered (trace.h:250)               */
	int32_t nportal;    /* portal crossings (trace.h:576)                           */
} pwn_ids;

/* one camera pose: mat4 cam rows x,y,z,w (defs.h:46-52) + the global sec_current (defs.h:23) */
typedef struct pwn_pose
{
	float cam[16];
	float sec_current;
} pwn_pose;

typedef struct pwn_ctx pwn_ctx;

/* trace flags */
#define PWN_F_BLUR 1u        /* run the POSTPROC_BLUR pass (screen.h:69-123)         */
#define PWN_F_IDS 2u         /* also write pwn_ids (needs pwn_frame_config want_ids) */
#define PWN_F_NO_REFLECT_BLUR 4u /* build without REFLECT_BLUR (defs.h:6); diagnostics */
#define PWN_F_LOCAL_BANDS 8u /* pwn_trace_shard_blur: the blurred bands stay in this shard's own sbuf (for hosts that
                                read the bands of every GPU themselves, pwn_read_bands) instead of going to shard 0 */

/* ---- lifetime ------------------------------------------------------------------ */
int pwn_abi_version(void);
const char *pwn_last_error(void);

/* device: CUDA ordinal.  Fails with PWN_E_NODEVICE unless it is compute capability 10.x. */
int pwn_create(pwn_ctx **out, int device);
int pwn_destroy(pwn_ctx *ctx);

/*
 * rcpps / rsqrtps emulation tables (call sites trace.h:231, util.h:43).
 * Both instructions are, on the CPUs probed, pure functions of (exponent parity,)
 * the top `bits` mantissa bits; the device reproduces them from tables of the
 * results for inputs in [1,2) (rcp) and [1,4) (rsqrt).
 *   rcp_tab  : 1<<rcp_bits   entries, float bits of rcpps(1 + i*2^-rcp_bits)
 *   rsqrt_tab: 2<<rsqrt_bits entries, float bits of rsqrtps over [1,2) then [2,4)
 * pwn_math_tables_from_host() dumps them from the CPU it runs on and verifies the
 * bucket property exhaustively (all 2^23 mantissas, every exponent the trace can
 * produce); PWN_E_HOSTMATH if the property does not hold within 23 bits.
 */
int pwn_math_tables(pwn_ctx *ctx, const uint32_t *rcp_tab, int rcp_bits, const uint32_t *rsqrt_tab, int rsqrt_bits);
int pwn_math_tables_from_host(pwn_ctx *ctx);
/* same dump without a context (host-only, no GPU needed); tables may be NULL to query bits */
int pwn_host_math_probe(int *rcp_bits, int *rsqrt_bits, uint32_t *rcp_tab, size_t rcp_cap, uint32_t *rsqrt_tab, size_t rsqrt_cap);

/*
 * libm on the bounce-shading path: sinf/cosf/sincosf (trace.h:42-46) and expf (trace.h:97) come
 * from the host's glibc in the reference.  The device restates glibc 2.39's algorithms bit for
 * bit (csrc/pwn_libm.h) in both of the builds x86-64 glibc selects between at load time.
 * pwn_host_libm_probe asks THIS host's libm which one it runs (*fma_build 1 = "_fma", 0 =
 * "_sse2"; *exact = 0 if it is neither); pwn_math_tables_from_host applies the answer,
 * pwn_libm_config overrides it.
 */
int pwn_host_libm_probe(int *fma_build, int *exact);
int pwn_libm_config(pwn_ctx *ctx, int fma_build);
/*
 * What the context's emulation is pinned to: table widths, the libm variant in use, and whether
 * pwn_math_tables_from_host found this host's libm to be exactly that variant (1), neither of the two
 * (0: pwn_math_tables_from_host then also leaves a warning in pwn_last_error -- frames are bit-exact against
 * a glibc 2.39 host, not necessarily against this one), or -1 when tables / variant were handed in.
 * Any pointer may be NULL.
 */
int pwn_math_status(pwn_ctx *ctx, int *rcp_bits, int *rsqrt_bits, int *libm_fma, int *libm_exact);
/* test hook (host only, no GPU): the restatement itself over an array */
void pwn_host_libm_eval(const float *in, float *s, float *c, float *e, size_t n, int fma_build);
/* test hook: sin, cos, exp of n floats evaluated by the device functions K1 uses */
int pwn_debug_libm(pwn_ctx *ctx, const float *in, float *sin_out, float *cos_out, float *exp_out, int n);
/* test hook: the device's col_ftoint (util.h:48-59) on n BGRA float quadruples */
int pwn_debug_pack(pwn_ctx *ctx, const float *bgra, uint32_t *packed, int n);

/* ---- level: once per level (grid and portals are immutable after load, script.h:65-69) ---- */
/*
 * Flattens the reference's level (defs.h:96-111) into the device layout:
 * a w*h plane of cell classes and one transform record per portal cell.
 * cells: h rows of w int32 codes, row z then column x (lv->data[z][x]).
 * portals: nportals entries indexed by PWN_PORTAL_INDEX (lv->pmap[]).
 * Device-layout limits (PWN_E_LIMIT): at most 16384 cells per axis and 32768 portal CELLS
 * (the reference's own limits are 64x64 cells and 26 letters, defs.h:96-111).
 */
int pwn_level_upload(pwn_ctx *ctx, int w, int h, const int32_t *cells, int nportals, const pwn_portal *portals);

/* per frame, after level_prepare_render (main.c:95): spheres in objs[] order; bins them per cell (level.h:1-39) */
int pwn_parts_upload(pwn_ctx *ctx, int nsph, const pwn_sphere *sph);

/* ---- frame buffers (the globals sbuf/tsbuf/zbuf, main.c:31-33,395-400) ------------ */
/* (PWN_E_LIMIT for rwidth*rheight >= 2^31: the kernels index a frame with 32 bits) */
int pwn_frame_config(pwn_ctx *ctx, int rwidth, int rheight, int want_ids);
/* zero sbuf/zbuf (and ids) on the device: what a fresh malloc'd+memset frame is to the reference */
int pwn_frame_clear(pwn_ctx *ctx);

/*
 * Row-band sharding for multi-GPU: this context traces only the
 * 32-row bands b with b % nshards == shard (screen.h:63-64 bands).  Default 1/0.
 */
int pwn_shard_config(pwn_ctx *ctx, int nshards, int shard);
/*
 * How the bands are dealt: PWN_SHARD_INTERLEAVED (default; band b goes to shard b % nshards: even load whatever the
 * view) or PWN_SHARD_BLOCKS (shard s owns the contiguous bands [s*nbands/nshards, (s+1)*nbands/nshards)).  Blocks are
 * what makes a sharded frame WITH the blur pass cheap: the blur reaches only 0.002*dimy*(z-1) rows (screen.h:86,100-102),
 * so a shard needs a few halo bands of each neighbour instead of the whole pre-blur frame.  halo_bands: bands exchanged
 * with each neighbour, 0 = ceil(0.015*dimy/32); samples that reach further are fetched from the owning shard afterwards,
 * the frame is exact for any value.
 */
enum { PWN_SHARD_INTERLEAVED = 0, PWN_SHARD_BLOCKS = 1 };
int pwn_shard_layout(pwn_ctx *ctx, int layout, int halo_bands);
/*
 * Block layout: explicit boundaries.  band_start[s] is the first band of shard s, band_start[nshards] the number of
 * bands of the frame ((dimy + 31)/32; boundaries given for another height are ignored).  Rows differ in cost (the
 * floor below the horizon carries the wavy-normal bounce, trace.h:39-56): a harness times the blocks and moves the
 * boundaries until the shards finish together (pwnfps_b200/shard.py: balance_blocks).  Every shard must be given the
 * same boundaries.  NULL restores the even split.
 */
int pwn_shard_blocks(pwn_ctx *ctx, int nshards, const int32_t *band_start);
/*
 * The block layout's arithmetic, without a context (host only, no GPU): for a frame of dimy rows in nshards blocks
 * (band_start as for pwn_shard_blocks, or NULL) with halo_bands (0 = automatic), shard's
 *   out[0], out[1]  first band and number of bands       out[2], out[3]  halo bands it sends up / down
 *   out[4], out[5]  the bands [vlo, vhi) it will ever hold: its own and the halo bands its neighbours send
 */
int pwn_block_shape(int dimy, int nshards, int halo_bands, const int32_t *band_start, int shard, int32_t out[6]);

/* ---- the hot path ------------------------------------------------------------------ */
/*
 * trace_ray_prelude for all bands of this shard (screen.h:1-28, 63-67) and, with
 * PWN_F_BLUR, the blur pass.  from/rayb/rdx/rdy are the four vec4 that
 * trace_screen_centred derives from cam (screen.h:43-57); their w lanes must be
 * (1,0,0,0)-shaped as main.c:60-64 + mat4_roty produce (PWN_E_UNSUPPORTED
 * otherwise).  Results stay on the device; asynchronous on the context's stream.
 */
int pwn_trace_frame(pwn_ctx *ctx, const float from[4], const float rayb[4], const float rdx[4], const float rdy[4],
	int dimx, int dimy, float sec_current, unsigned flags);

/*
 * REFLECT of defs.h:8 -- mirror bounce levels per pixel -- for hosts built with another value
 * than the shipped 2.  0, 1 or 2; anything else is PWN_E_UNSUPPORTED.  (REFLECT_BLUR, defs.h:6,
 * is the flag PWN_F_NO_REFLECT_BLUR; POSTPROC_BLUR, defs.h:7, is PWN_F_BLUR.)
 */
int pwn_trace_config(pwn_ctx *ctx, int reflect);

/*
 * The blur pass alone (screen.h:69-123) over the frame currently in sbuf/zbuf: for
 * harnesses that trace row bands on several GPUs and gather them first.
 */
int pwn_blur_frame(pwn_ctx *ctx, int dimx, int dimy);

/* screen.h:31-57 on the host, bit-identical to the reference's arithmetic, then pwn_trace_frame */
int pwn_trace_cam(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags);

/* the ray basis alone (screen.h:43-57): exposed so a host can check it against its own build */
void pwn_ray_setup(const float cam[16], int dimx, int dimy, float rayb[4], float rdx[4], float rdy[4]);

/* wait for the context's streams */
int pwn_sync(pwn_ctx *ctx);
/*
 * Sharded frames of the block layout keep two frames in flight on two streams.  pwn_join makes everything the library
 * has queued so far a predecessor of whatever comes next on pwn_stream (an event recorded there afterwards marks the
 * completion of every frame submitted before it); no host wait.
 */
int pwn_join(pwn_ctx *ctx);

/* device -> host; any pointer may be NULL.  Full frames: rwidth*rheight elements (ids: x3). Synchronous. */
int pwn_read_frame(pwn_ctx *ctx, uint32_t *sbuf, float *zbuf, pwn_ids *ids);
/* host -> device, the inverse: a pre-blur frame assembled elsewhere, for pwn_blur_frame / pwn_screen_upscale */
int pwn_write_frame(pwn_ctx *ctx, const uint32_t *sbuf, const float *zbuf);

/*
 * The drop-in call: everything trace_screen_centred does for the host, with host
 * buffers.  Uploads nothing but the pose, traces (and blurs with PWN_F_BLUR),
 * copies sbuf (and zbuf if non-NULL) back, returns when they are complete -- the
 * host reads sbuf immediately afterwards (screen_upscale, main.c:108).
 */
int pwn_trace_screen(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags,
	uint32_t *sbuf_host, float *zbuf_host);
/*
 * pwn_trace_screen traces (and blurs) a frame in row chunks so that the copy of chunk k to the host runs under
 * the kernels of chunk k+1; only the traced window [0,dimx) x [0,dimy) of the host buffers is written, as in the
 * reference.  nchunks: 1..8, or 0 = automatic (4 for frames of half a megapixel and more, else 1).
 */
int pwn_screen_chunks(pwn_ctx *ctx, int nchunks);

/*
 * screen_upscale (screen.h:126-149) on the device: every pixel of the dimx x dimy frame in
 * sbuf becomes an rscale x rscale block of a display surface of pitch_bytes per row
 * (screen->pitch), at exactly the words the reference's pointer walk writes -- with a padded
 * pitch (pitch_bytes/4 > dimx*rscale) that walk drifts by the padding per block row, and so
 * does this.  The surface (dimy*rscale rows of pitch_bytes) stays on the device
 * (pwn_surface_device_ptr); with pixels_host != NULL it is also copied there, synchronously.
 * Words the reference leaves untouched are 0.
 */
int pwn_screen_upscale(pwn_ctx *ctx, int dimx, int dimy, int rscale, int pitch_bytes, void *pixels_host);
int pwn_surface_device_ptr(pwn_ctx *ctx, uint64_t *surface_dev, uint64_t *bytes);
/* device time in ms of the last pwn_screen_upscale kernel (CUDA events on the context stream; syncs) */
int pwn_last_upscale_ms(pwn_ctx *ctx, float *k3_ms);

/*
 * A scripted fly-through: n poses, frame i written to sbuf_host + i*rwidth*rheight
 * (pinned or pageable).  The batch is pipelined over three device frame sets and two
 * streams: with PWN_F_BLUR the launch that traces pose i also blurs frame i-1 (chained by
 * programmatic dependent launch), while frame i-2 is copied to the host.  sbuf_host may be
 * NULL: frames then stay on the device (pwn_read_frame afterwards returns frame set 0, not
 * necessarily the last pose).  sums, if non-NULL, receives one checksum per frame computed
 * on the device: sum over pixels i = y*dimx+x of mix64(i<<32 | sbuf[y][x]) mod 2^64,
 * mix64 = the splitmix64 finaliser (order independent, so it parallelises).
 */
int pwn_trace_poses(pwn_ctx *ctx, int n, const pwn_pose *poses, int dimx, int dimy, unsigned flags,
	uint32_t *sbuf_host, uint64_t *sums);

/* ---- device-resident access for harnesses that manage their own transfers ---------- */
/* raw device pointers of the current frame (CUdeviceptr values); for peer mapping and zero-copy consumers */
int pwn_frame_device_ptrs(pwn_ctx *ctx, uint64_t *sbuf_dev, uint64_t *zbuf_dev, uint64_t *tsbuf_dev);
/*
 * Redirect this context's K1/K2 output into another allocation (e.g. GPU 0's
 * frame, peer-mapped or opened through CUDA IPC): the fused "trace + gather"
 * path, bands land in the gathered frame as they are traced.  0 restores.
 */
int pwn_frame_redirect(pwn_ctx *ctx, uint64_t sbuf_dev, uint64_t zbuf_dev, uint64_t tsbuf_dev);
/* CUDA IPC handle (64 bytes) of this context's frame allocation, and its inverse */
int pwn_frame_ipc_export(pwn_ctx *ctx, void *handle64, uint64_t *sbuf_off, uint64_t *zbuf_off, uint64_t *tsbuf_off);
int pwn_frame_ipc_import(pwn_ctx *ctx, const void *handle64, uint64_t *base_dev);
int pwn_frame_ipc_close(pwn_ctx *ctx, uint64_t base_dev);

/*
 * A sharded frame WITH the blur pass, block layout (pwn_shard_layout PWN_SHARD_BLOCKS; what bench.py times at N > 1):
 * one K1 launch traces the shard's bands, halo bands first, storing those into the neighbour's pre-blur buffer too and
 * counting finished lanes per band; on a second stream the blur pass follows in a few groups of bands, each launch gated on the
 * device by the counters of the bands it reads, and stores straight into shard 0's frame -- the gather to GPU 0 runs under
 * the trace.  Samples beyond the exchanged rows are finished from the owning shard's rows after a device barrier.
 * Frames alternate between two groups of planes: the call for frame n+1 may be made at once, its trace runs while frame
 * n is still being blurred and gathered (pwn_sync / pwn_read_frame / pwn_join wait for or order after both).
 *
 * Interleaved layout: screen.h:69-123 gathers from anywhere in the pre-blur
 * frame, so every shard needs all of it: with pwn_frame_peers set, pwn_trace_shard_blur
 *   1. traces the own bands (pwn_shard_config) and stores every pixel into the pre-blur buffer
 *      of EVERY shard -- the all-gather is part of the trace kernel, whole 128-byte row
 *      segments over NVLink;
 *   2. meets the other shards at a barrier on the device (flags in each frame allocation,
 *      signalled and polled by a one-warp kernel in stream order; no host round trip);
 *   3. blurs the own bands from the local copy and stores them into shard 0's sbuf;
 *   4. meets them again: on return of a later pwn_sync, shard 0's frame is complete and the
 *      pre-blur buffers may be overwritten by the next frame.
 * base_dev[s] = device address of shard s's frame allocation in THIS process (pwn_frame_ipc_import
 * of the handle shard s exported; the own entry is ignored).  All shards must configure the same
 * frame size and call pwn_trace_shard_blur the same number of times.  nshards == 0 clears.
 * The call also restarts the frame count the device barriers and band counters work with: every shard calls it,
 * with no sharded frame in flight, and no shard starts a frame before all have returned from it (a host barrier).
 * A shard that never arrives ends the wait after 4 s: the next pwn_sync / pwn_read_frame / pwn_read_bands of
 * the waiting shard then fails with PWN_E_PEER (once per incident; the frame holds stale bands), and
 * pwn_frame_peers_status reports 1 until pwn_frame_peers is called again.  A shard whose own trace cannot be
 * launched still meets its peers at both barriers of that frame, so the others neither wait nor fall out of step,
 * and returns its error.
 */
int pwn_frame_peers(pwn_ctx *ctx, int nshards, const uint64_t *base_dev);
int pwn_trace_shard_blur(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags);
int pwn_frame_peers_status(pwn_ctx *ctx, int *timed_out);
/* test hook: the first nwords words behind the frame planes (barrier flags [0,16), error word 16, band counters from 64),
 * read on a stream of their own while kernels may be running; returns the number of words read */
int pwn_debug_flags(pwn_ctx *ctx, uint32_t *out, int nwords);
/*
 * Copy the bands this shard owns (pwn_shard_config) from its sbuf to the same rows of a HOST frame
 * of rwidth*rheight words -- typically one frame in shared memory that every shard's process has
 * page-locked (pwn_host_register): the frame is assembled in host memory over all the GPUs' PCIe
 * links at once and never gathered on one GPU.  Synchronous.
 */
int pwn_read_bands(pwn_ctx *ctx, int dimx, int dimy, uint32_t *sbuf_host_frame);

/* pinned host memory for frames (cudaHostAlloc): what pwn_trace_poses copies into at full PCIe rate */
int pwn_host_alloc(void **out, size_t bytes);
int pwn_host_free(void *p);
/*
 * Page-lock a buffer the HOST allocated (the reference mallocs sbuf/zbuf, main.c:395-400), in
 * place: pwn_trace_screen / pwn_read_frame / pwn_trace_poses then copy into it at PCIe rate
 * instead of through the driver's staging buffers.  Registering twice is harmless.  Unregister
 * before the host frees the buffer.
 */
int pwn_host_register(void *p, size_t bytes);
int pwn_host_unregister(void *p);

/* ---- measurement hooks (bench harness) ------------------------------------------------ */
/* kernels launched by this context since creation (K1, K2, checksum, clear) */
int pwn_launch_count(pwn_ctx *ctx, uint64_t *count);
/*
 * Algorithmic FP32 lane-ops of the last trace that ran with PWN_F_IDS: the work model of
 * DESIGN.md (per-event constants counted from the reference source) summed over all pixels.
 * This is the numerator of the FP32 roofline; it is counted, not timed.
 */
int pwn_last_opcount(pwn_ctx *ctx, uint64_t *ops);
/* device time in ms of the last pwn_trace_frame's K1 and K2 launches (CUDA events on the context stream; syncs) */
int pwn_last_kernel_ms(pwn_ctx *ctx, float *k1_ms, float *k2_ms);
/* run n frames of the same pose back to back on the device, return total K1(+K2) event time in ms */
int pwn_time_frames(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags, int n, float *total_ms);
/*
 * Measured FP32 issue peak of this GPU in 10^12 lane-ops/s: a kernel of independent, non-fused
 * FMUL/FADD chains (the trace may not use FMA: the reference build has none).  The denominator of
 * the FP32 roofline, measured under the clocks of the moment.
 */
int pwn_fp32_peak(pwn_ctx *ctx, double *tlaneops);
/* the CUDA stream the context launches on (cudaStream_t as an integer) */
int pwn_stream(pwn_ctx *ctx, uint64_t *stream);

#ifdef __cplusplus
}
#endif

#endif /* PWNTRACE_H */
