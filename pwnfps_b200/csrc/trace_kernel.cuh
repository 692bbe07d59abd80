// trace_kernel.cuh -- K1: ray generation + cell walk + shading + two mirror bounces + pack.
//
// B200-native restatement of trace_ray_prelude (screen.h:1-28), trace_ray
// (trace.h:186-679), trace_ray_through (trace.h:156-184), trace_hit_wall
// (trace.h:108-154), trace_hit_bounce (trace.h:3-106), randi/randfs, v_dot,
// v_normalise and col_ftoint (util.h:1-59).  One thread traces one pixel: its
// primary ray and both bounce segments, iteratively (the reference recurses,
// trace.h:90), with all per-ray state in registers.
//
// Numerics contract (DESIGN.md "Numerics"): every float operation is issued in
// the reference's source order with explicit round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn never contract into FMA); rcpps/rsqrtps come from the
// host-dumped tables; float->int uses the reference's truncation / RN-even.
//
// The kernel is instruction-issue bound (profiles/), so the walk is written to
// minimise instructions per cell step: the ray carries ONE linear index into a
// padded cell map (no per-step bounds checks or cz*w+cx), the face code of the
// last step is a select between two per-segment registers, and everything rare
// (raised floors, leaving the grid, portals) sits behind one predicate.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pwntrace.h"
#include "pwn_libm.cuh"
#include "blur_kernel.cuh"

namespace pwn {

// ---- device layout of a level (flattened by pwn_level_upload) -----------------------
enum : uint32_t
{
	CL_SOLID = 0,      // anything not listed below (trace.h:651-664)
	CL_PORTAL = 1,     // letter cell; bits 16.. of the map entry index PortalRec
	CL_RAMP_GT = 4,    // '>'  (trace.h:443-505)   ramps: (class & 12) == 4; bit 1: along z; bit 0: ray.y += tilt
	CL_RAMP_LT = 5,    // '<'
	CL_RAMP_COMMA = 6, // ','
	CL_RAMP_CARET = 7, // '^'
	CL_ONE = 8,        // ';'  1-high            (trace.h:302-352)   rooms: bit 3; bit 1: 2-high; bit 0: foggy
	CL_ONE_FOG = 9,    // '$'  1-high, foggy
	CL_TWO = 10,       // '#'  2-high            (trace.h:354-441)
	CL_TWO_FOG = 11,   // '&'  2-high, foggy
	CL_RAISED = 12,    // '"'  1-high on the upper floor of a 2-high room
	CL_OOB = 2,        // the one-cell border around the map: the ray is leaving the grid

	// Map entry = class | slow[4] << 4 | hwin[4] << 8 | rz[4] << 12 | portal record << 16.  slow[d] (d = FXP, FZP, FXN, FZN) is set
	// on a room cell when stepping out of it through face d can need more than the plain advance:
	// the neighbour is the border, a raised cell steps into a 2-high room (trace.h:343-351), or a
	// 2-high room steps into anything but another 2-high room (trace.h:395-440).  The flattening
	// (pwn_level_upload) derives the bits; the room step tests one of them instead of the classes.
	// hwin[d] (bits 8-11) marks the commonest slow case: a 2-high room stepping into a plain cell inside
	// the grid (solid, 1-high, ramp).  All the reference does there is end the ray on the wall above the
	// opening when it is outside the height window (trace.h:415-440); the slow path tests this bit first.
	// rz[d] (bits 12-15) is set on all four faces of a raised cell: with the three groups in the same
	// per-face layout, one shift of the entry by the face code (done before the next cell is loaded)
	// answers everything the slow path asks about the cell being left.
	CM_SLOW_SHIFT = 4, CM_HWIN_SHIFT = 8, CM_RZ_SHIFT = 12, CM_REC_SHIFT = 16,
	CM_TWO_SHIFT = 16,   // rooms have no portal record: on a 2-high room bits 16-19 repeat "2-high" once per face
	CM_REC_MASK = 0x7FFFu,   // a portal record index has 15 bits: bit 31 of every entry is
	CM_SPH_BIT = 1u << 31,   // "the cell's sphere list is not empty", set per frame in the copy of the map the sphere builds walk
	// the value of `oobm` while the ray is outside the grid: it forces the slow bit on and masks the
	// height-window bit off, each with the same single LOP3 that tests the bit
	OOB_MARK = (1u << CM_SLOW_SHIFT) | (1u << CM_HWIN_SHIFT)
};

enum : uint32_t { PK_WALL = 0, PK_PASS = 1, PK_RAISED = 2 };   // what a 2-high room sees through a cell (trace.h:404-440)
enum : uint32_t { PR_OK = 0, PR_UNPAIRED = 1, PR_STRAY = 2 };  // trace.h:512-526 / 547-559

// One record per portal CELL (not per letter): the crossing transform of
// trace.h:528-647 for a ray standing on this cell, precomputed.  It is the
// 3x4 affine map {R(rot*90deg) about (ccx, ccz), translate (dx,dz)} in factored
// form -- the reference evaluates it in a fixed association order
// (trace.h:583-616) that a generic matrix product would not reproduce bit for bit.
struct __align__(16) PortalRec
{
	int32_t dci;        // destination portal cell, index into the padded map
	float dx, dz;       // pos.xz += (dx,dz) before rotating        (trace.h:535-536,543-544)
	uint32_t flags;     // rot | kind<<8 | peek<<16
	float ccx, ccz;     // (float)dcx + 0.5f, (float)dcz + 0.5f     (trace.h:584-607)
	int32_t dcx, dcz;   // destination cell coordinates (ids only)
};

struct LevelDev
{
	const uint32_t *cellmap;   // (h+2)*(w+2) entries: class | slow<<4 | hwin<<8 | rz<<12 | rec<<16, border = CL_OOB
	const PortalRec *portals;
	int w, h, wp;              // wp = w + 2: row stride of the padded map
	const uint32_t *rcp_tab;   // 1<<rcp_bits
	const uint32_t *rsqrt_tab; // 2<<rsqrt_bits
	int rcp_shift, rsqrt_shift;   // 23 - bits
	uint32_t rsqrt_mask;          // (2 << rsqrt_bits) - 1
	// spheres (pwn_parts_upload); CSR over the padded map
	const float4 *sph_pos;     // x y z w
	const float4 *sph_col;     // b g r a
	const float2 *sph_rr;      // r, refl
	const uint32_t *parts_pk;  // (h+2)*(w+2): first list entry | min(count, 255) << 24 -- ONE load per cell step tells whether
	                           // the cell holds spheres at all; 255 means "count from parts_off"
	const uint32_t *parts_off; // (h+2)*(w+2)+1
	const uint32_t *parts_idx;
	int nsph;
};

struct TraceParams
{
	LevelDev lv;
	float from[4], rayb[4], rdx[4], rdy[4];
	float sec_current;
	int dimx, dimy, rwidth;
	int nshards, shard;        // 32-row bands b with b % nshards == shard
	int nbands_local;          // bands this shard owns
	int band0;                 // first of them this launch traces (pwn_trace_screen traces a frame in row chunks)
	uint32_t *sbuf;            // pre-blur or final frame (BGRA8)
	float *zbuf;
	pwn_ids *ids;              // [rwidth*rheight][3] or null
	unsigned long long *opcount; // IDS builds: sum of algorithmic FP32 lane-ops (DESIGN.md "Work model")
	int reflect_blur;
	int reflect;               // REFLECT (defs.h:8): bounce levels, 0..2
	uint32_t *peer[7];         // the same pre-blur buffer on the other shards' GPUs (pwn_frame_peers): every pixel is
	int npeers;                // stored there too -- the all-gather the sharded blur pass needs, fused into the trace
	int rowtile;               // 1: a warp covers 32x1 pixels (128-byte row stores, for frames reached over NVLink) instead of 8x4
	uint32_t zero;             // always 0; opaque to the compiler (see xmov)
	// Contiguous-block sharding (pwn_shard_layout PWN_SHARD_BLOCKS): this shard owns the bands [blk_b0, blk_b0 + nbands_local).
	// Its first blk_ht and last blk_hb bands are the halo the blur pass of the shard above / below reads: they are traced
	// FIRST and stored into that shard's pre-blur buffer as well.  Every group of lanes that finishes adds its lane count
	// to the band's counter (here, and for a halo band on the neighbour too), after a fence: the blur pass of a band
	// starts as soon as the counters of the bands it reads are full -- it runs under the trace of the later bands.
	int blk, blk_b0, blk_ht, blk_hb;
	uint32_t *halo_up, *halo_dn;          // pre-blur buffer of the shard above / below (null at the frame's edges)
	uint32_t *done_local, *done_up, *done_dn;   // band counters, indexed by the band's number in the frame
	uint32_t blk_lanes, blk_target;             // lanes per band (32 rows x the grid's width); a counter's value when this frame's band is complete
	BlurParams prev;           // BLURP builds: the blur pass of the PREVIOUS frame of a pose batch
};

// this group of lanes is done with band `band`: publish its stores, then count its lanes
__device__ __forceinline__ void band_arrive(const TraceParams &P, int band, int which /* 0 none, 1 up, 2 down */, bool wrote)
{
	// No fence: a __threadfence() here is a MEMBAR plus an invalidation of the SM's whole L1 (CCTL.IVALL), once per
	// warp -- it costs the walk its 98% L1 hit rate on the cell map (measured: K1 +28%).  The lanes meet at a warp
	// barrier (which orders their stores before what the leader does next) and the leader's add is a RELEASE: the
	// waiter's acquire load of the counter then sees every lane's pixels.
	// A halo band is signalled to the neighbour ONCE, by the group that completes it here: a system-scope release per
	// warp (MEMBAR.SYS: a round trip over NVLink with the warp's slot held) cost K1 +60% at 8 GPUs.  The completing group
	// has acquired every other group's release through the local counter, so its one system-scope fence publishes the
	// whole band's pixels on the neighbour before the neighbour's counter moves.
	const unsigned m = __activemask();
	__syncwarp(m);
	(void)wrote;
	if((threadIdx.x & 31) == (__ffs(m) - 1))
	{
		const unsigned n = __popc(m);
		if(which == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(P.done_local + band), "r"(n) : "memory");
		else
		{
			unsigned old;
			asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(P.done_local + band), "r"(n) : "memory");
			if(old + n == P.blk_target)
			{
				__threadfence_system();
				asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"((which == 1 ? P.done_up : P.done_dn) + band), "r"(P.blk_lanes) : "memory");
			}
		}
	}
}

// Algorithmic FP32 lane-ops per event (counted from the reference source: mul, add,
// cmp, abs, cvt, rcp, rsqrt = 1 each, live lanes only).  Only the IDS instantiation counts.
enum : uint32_t
{
	W_PIXEL = 20,     // ray generation + pack            (screen.h:12-24, util.h:48-59)
	W_SEGMENT = 40,   // normalise, signs, rcp, wdist     (trace.h:203-241)
	W_ONE = 14,       // 1-high room step                 (trace.h:156-184,302-352)
	W_TWO = 18,       // 2-high room step                 (trace.h:354-441)
	W_RAMP = 30,      // ramp step                        (trace.h:443-505)
	W_PORTAL = 12,    // portal crossing                  (trace.h:528-647)
	W_SPH_TEST = 14,  // one sphere test                  (trace.h:259-273)
	W_SPH_HIT = 20,   //   ... that becomes the nearest   (trace.h:276-292)
	W_WALL = 9,       // trace_hit_wall                   (trace.h:108-150)
	W_BOUNCE = 40,    // reflect + jitter + blend         (trace.h:12-37,77-93)
	W_FLOOR = 42,     // wavy floor normal: 38 + 4 libm   (trace.h:39-56)
	W_FOG = 13        // fog blend: 12 + expf             (trace.h:95-101)
};

// ---- host-CPU primitives, emulated ---------------------------------------------------
// rcpps(|x|), x normal: neither the table index nor the exponent field looks at the sign bit
// With the table widths of every x86 CPU seen so far (rcpps 11 bits, rsqrtps 10 + parity: 2048 entries each) the two
// tables sit in the 16 KB in front of the cell map (STD builds): one base pointer serves map and tables, the shifts and the
// mask are immediates, and nothing about the tables is read from the kernel parameters (five loads per segment less).
constexpr int STD_RCP_BITS = 11, STD_RSQRT_BITS = 10;
constexpr int STD_TAB_WORDS = 4096;      // rcp at cmap - 4096, rsqrt at cmap - 2048 (words)

template <bool STD>
__device__ __forceinline__ float rcp_emul(const LevelDev &L, const uint32_t *cmap, float x)
{
	uint32_t u = __float_as_uint(x);
	uint32_t t = STD ? __ldg(cmap - STD_TAB_WORDS + ((u & 0x7FFFFFu) >> (23 - STD_RCP_BITS))) : __ldg(&L.rcp_tab[(u & 0x7FFFFFu) >> L.rcp_shift]);
	return __uint_as_float(t - (u & 0x7F800000u) + 0x3F800000u);
}

// rsqrtps(x), x > 0.  The host's table is indexed by [parity of the unbiased exponent | top mantissa bits] and the
// result exponent is floor((E - 127)/2) = ((E + 1) >> 1) - 64, E the biased exponent.  pwn_math_tables stores the
// table re-indexed by [E & 1 | top mantissa bits] -- the bit pattern as it is -- with (E & 1) << 23 and the
// exponent bias already taken out of each entry, which leaves a shift, a mask and a subtract around the load.
template <bool STD>
__device__ __forceinline__ float rsqrt_emul(const LevelDev &L, const uint32_t *cmap, float x)
{
	const uint32_t u = __float_as_uint(x);
	const uint32_t t = STD ? __ldg(cmap - STD_TAB_WORDS/2 + ((u >> (23 - STD_RSQRT_BITS)) & ((2u << STD_RSQRT_BITS) - 1u)))
	                       : __ldg(&L.rsqrt_tab[(u >> L.rsqrt_shift) & L.rsqrt_mask]);
	return __uint_as_float(t - ((u >> 1) & 0x7F800000u));
}

// util.h:1-16 with the shipped -ffast-math folding of randfs: r*fl(2/3759) - 1
__device__ __forceinline__ float randfs(uint32_t &seed)
{
	seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
	uint32_t r = seed % 3759u;
	return __fsub_rn(__fmul_rn((float)r, __uint_as_float(0x3A0B79B3u)), 1.0f);
}

// col_ftoint (util.h:48-59): cvtps2dq + packssdw + packuswb, B | G<<8 | R<<16 | A<<24.  cvtps2dq rounds to
// nearest even and gives 0x80000000 for NaN and for |s| >= 2^31, which the two saturating packs turn into 0;
// everything else saturates to 0..255.  cvt.rni.sat.u8 does the same in one F2IP except for s >= 2^31 (255
// instead of 0): that one case is tested once for the four lanes and fixed behind a real call, which ptxas
// does not flatten into predicated instructions on every pixel's path.
__device__ __noinline__ uint32_t pack_huge(float s0, float s1, float s2, float s3, uint32_t word)
{
	if(s0 >= 2147483648.0f) word &= 0xFFFFFF00u;
	if(s1 >= 2147483648.0f) word &= 0xFFFF00FFu;
	if(s2 >= 2147483648.0f) word &= 0xFF00FFFFu;
	if(s3 >= 2147483648.0f) word &= 0x00FFFFFFu;
	return word;
}

__device__ __forceinline__ uint32_t pack_word(float b, float g, float r, float a)
{
	const float s0 = __fmul_rn(b, 255.0f), s1 = __fmul_rn(g, 255.0f), s2 = __fmul_rn(r, 255.0f), s3 = __fmul_rn(a, 255.0f);
	uint32_t q0, q1, q2, q3;
	asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q0) : "f"(s0));
	asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q1) : "f"(s1));
	asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q2) : "f"(s2));
	asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q3) : "f"(s3));
	uint32_t word = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
	if(fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)) >= 2147483648.0f) word = pack_huge(s0, s1, s2, s3, word);   // (fmaxf skips NaN lanes)
	return word;
}

// col_ftoint alone, for known-answer tests of the pack (pwn_debug_pack): the code K1 ends with
__global__ void k_pack_eval(const float4 *in, uint32_t *out, int n)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	const float4 v = in[i];
	out[i] = pack_word(v.x, v.y, v.z, v.w);
}

// The portal rotation exchanges x and z quantities.  Written as plain copies the exchange becomes a
// cyclic register copy that ptxas resolves with moves at the head of EVERY walk iteration.  The ray
// direction and the signs go through an xor with an opaque zero; the step state (wx/wz, iax/iaz) goes
// through shared memory (PWN_XCHG in k_trace), which removes the copies altogether.
__device__ __forceinline__ float xmov(float v, uint32_t zero) { return __uint_as_float(__float_as_uint(v) ^ zero); }

// a pointer ptxas cannot trace back to the kernel parameters (both halves xor an opaque zero): it stays in a
// register pair instead of being re-read from the constant bank at every use
__device__ __forceinline__ const uint32_t *opaque_ptr(const uint32_t *p, uint32_t zero)
{
	const unsigned long long v = (unsigned long long)(uintptr_t)p;
	const uint32_t lo = (uint32_t)v ^ zero, hi = (uint32_t)(v >> 32) ^ zero;
	return (const uint32_t *)(uintptr_t)(((unsigned long long)hi << 32) | lo);
}

// ---- packed FP32 (Blackwell FADD2 / FMUL2 / FFMA2: two lanes per issue slot) ------------------------------
// The x and z quantities of a ray are stepped by the same operations, so they travel as pairs.  Two rules keep the
// results those of the reference's separate roundings:
//   * ptxas contracts a packed multiply that feeds a packed add into ONE FFMA2 (a single rounding), even for
//     mul.rn/add.rn and under --fmad=false.  A product that feeds an add is therefore written as fma(a, b, -0)
//     with the pair {-0, -0} coming from a kernel parameter ptxas cannot see through: a*b + (-0) is the correctly
//     rounded product for every a*b (including +-0), and an FFMA2 is never contracted again.
//   * a packed multiply whose result feeds scalar adds is left alone (measured on the SASS: FMUL2 + FADD + FADD).
// Scalars ride along as broadcast operands (R.F32): no copies.
#ifndef PWN_PACKED
#define PWN_PACKED 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 u; asm("mov.b64 %0, {%1,%2};" : "=l"(u) : "f"(lo), "f"(hi)); return u; }
__device__ __forceinline__ void unpk2(f32x2 u, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(u)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// a*b, correctly rounded, in a form ptxas cannot fuse with a following add (nz = {-0, -0}, opaque)
__device__ __forceinline__ f32x2 mul2x(f32x2 a, f32x2 b, f32x2 nz) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(nz)); return d; }

// one bit of a map entry as a predicate.  Written in PTX so that the front end cannot rewrite the test
// into shift-and-compare chains: ptxas turns and + setp into a single LOP3 with a predicate result.
__device__ __forceinline__ bool bit_of(uint32_t v, uint32_t mask)
{
	uint32_t t;
	asm("and.b32 %0, %1, %2;" : "=r"(t) : "r"(v), "r"(mask));
	return t != 0u;
}

// surface colours, lanes b g r a (defs.h:17-19, trace.h:558)
// (the codes are the byte offsets of the colours in c_surface: the indexed load takes them as they are)
enum : int { SF_CEIL = 0, SF_FLOOR = 16, SF_WALL = 32, SF_STRAY = 48, SF_NONE = 64, SF_SPHERE = 80 };
#define PWN_SURFACE(H) c_surface[(H) >> 4]
__constant__ float4 c_surface[4] = {
	{30.0f, 30.0f, 0.0f, 0.0f},   // COL_CEIL
	{1.0f, 1.0f, 1.0f, 0.0f},     // COL_FLOOR
	{0.8f, 0.8f, 1.0f, 0.0f},     // COL_WALL
	{5.0f, 0.0f, 5.0f, 0.0f}      // a third cell with a portal's letter (trace.h:558)
};

// Nearest sphere so far (trace.h:196-200: aux_dist, aux_pos, aux_norm, aux_col, aux_refl).  The
// reference keeps thirteen floats; six registers carry the same information: the normal, the colour
// and the reflectance are pure functions of (pos, the sphere, diff) and are rebuilt with the same
// operations at the one place they are used, when the segment ends on the sphere.  diff is kept
// because it depends on the ray at the time of the test, which a later portal may rotate.
struct Aux
{
	float dist;
	float px, py, pz;
	float diff;       // 0.2 + 0.8*max(0, -ray.norm)  (trace.h:286-290)
	uint32_t si;      // which sphere
};

// v_normalise(hit - centre) (trace.h:283-284), shared by the test and by the rebuild at the hit
template <bool STD>
__device__ __forceinline__ void sphere_normal(const LevelDev &L, const uint32_t *cmap, float hx, float hy, float hz, const float4 &sp4, float &nx, float &ny, float &nz)
{
	nx = __fsub_rn(hx, sp4.x); ny = __fsub_rn(hy, sp4.y); nz = __fsub_rn(hz, sp4.z);
	const float nd = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(nz, nz)), __fmul_rn(ny, ny));
	const float ns = rsqrt_emul<STD>(L, cmap, nd);
	nx = __fmul_rn(nx, ns); ny = __fmul_rn(ny, ns); nz = __fmul_rn(nz, ns);
}

// trace.h:220-222: a ray component nearer to zero than EPSILON (defs.h:1) becomes +-EPSILON
__device__ __noinline__ float3 eps_clamp(float rx, float ry, float rz)
{
	const float EPS = 0.0000000000001f;
	if(fabsf(rx) < EPS) rx = (rx < 0.0f ? -EPS : EPS);
	if(fabsf(ry) < EPS) ry = (ry < 0.0f ? -EPS : EPS);
	if(fabsf(rz) < EPS) rz = (rz < 0.0f ? -EPS : EPS);
	return make_float3(rx, ry, rz);
}

// CTA = 32 x K1_ROWS pixels = four warps of 8x4 pixels (or of 32x1 when the frame is reached over NVLink)
constexpr int K1_ROWS = 4;
constexpr int K1_THREADS = 32*K1_ROWS;
#ifndef PWN_MIN_CTAS
#define PWN_MIN_CTAS 8
#endif
constexpr int K1_MIN_CTAS = PWN_MIN_CTAS;   // 8 CTAs x 4 warps per SM; the kernel needs 54-56 registers, and 9 or 10 CTAs measure the same
#ifndef PWN_SPH_MIN_CTAS
#define PWN_SPH_MIN_CTAS 8
#endif
#ifndef PWN_WALK_UNROLL
#define PWN_WALK_UNROLL 1
#endif
#ifndef PWN_SEG_UNROLL
#define PWN_SEG_UNROLL 1
#endif
constexpr int K1_SEG_UNROLL = PWN_SEG_UNROLL;     // 1: one copy of the segment body for the three segments
constexpr int K1_WALK_UNROLL = PWN_WALK_UNROLL;   // 1: the walk stays rolled (measured: see DESIGN.md)
constexpr int K1_MIN_CTAS_SPH = PWN_SPH_MIN_CTAS;   // with spheres the nearest-hit record adds 12 live registers

template <bool SPH, bool IDS, bool LFMA, bool BLURP, bool STD>
__global__ void __launch_bounds__(K1_THREADS, SPH ? K1_MIN_CTAS_SPH : K1_MIN_CTAS) k_trace(const __grid_constant__ TraceParams P)
{
	const LevelDev &L = P.lv;
	__shared__ float s_ray[3][K1_ROWS][32];
	// Per-thread scratch in shared memory, addressed from ONE register that holds this thread's 32-bit
	// shared address (rows of K1_THREADS floats, the thread's word in each): rows 0-5 and 6-11 are the blend
	// state of the two bounce levels, rows 12-15 the portal exchange.  The address is made opaque once, or
	// the compiler rebuilds it from the CTA's shared window at every use (six instructions a time).
	__shared__ float s_scr[16][K1_THREADS];
	uint32_t sthr = (uint32_t)__cvta_generic_to_shared(&s_scr[0][threadIdx.x]);
	asm volatile("" : "+r"(sthr));
#define PWN_STS(A, OFF, V) asm volatile("st.shared.f32 [%0+" #OFF "], %1;" :: "r"(A), "f"(V) : "memory")
#define PWN_LDS(V, A, OFF) asm volatile("ld.shared.f32 %0, [%1+" #OFF "];" : "=f"(V) : "r"(A) : "memory")
	// a quarter-turn portal exchanges the x and z step state through shared memory: a register exchange
	// makes ptxas keep copies of the exchanged values at the head of EVERY walk iteration
#define PWN_XCHG { PWN_STS(sthr, 6144, wx); PWN_STS(sthr, 6656, wz); PWN_STS(sthr, 7168, iax); PWN_STS(sthr, 7680, iaz); \
		PWN_LDS(wz, sthr, 6144); PWN_LDS(wx, sthr, 6656); PWN_LDS(iaz, sthr, 7168); PWN_LDS(iax, sthr, 7680); }
static_assert(K1_THREADS*4 == 512, "scratch row offsets");

	// the cell map's address in a register pair (without spheres there is room for it): read as a kernel
	// parameter it is re-read from the constant bank at every cell step
	const uint32_t *cmap = SPH ? L.cellmap : opaque_ptr(L.cellmap, P.zero);
	// -0.0f that ptxas cannot fold (see mul2x), in a register of the thread's own (mixed with the thread index: left
	// uniform it is re-read from the constant bank at every use)
	const float negz = __uint_as_float(0x80000000u ^ (SPH ? P.zero : (P.zero & threadIdx.x)));   // (the sphere builds have no register to spare)
	const f32x2 nz2 = pk2(negz, negz);

	// ---- pixel of this thread: CTA = 32 x K1_ROWS pixels, warp = 8x4 (or 32x1) ----------------------
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int tx = P.rowtile ? lane : (warp & 3)*8 + (lane & 7);
	const int ty = P.rowtile ? warp : (warp >> 2)*4 + (lane >> 3);
	const int x = blockIdx.x*32 + tx;
	const int lrow0 = (blockIdx.y + P.band0*(32/K1_ROWS))*K1_ROWS;
	// local row -> global row (8 | 32: a CTA never straddles bands): through the band interleave, or, for a contiguous
	// block, in the order top halo, bottom halo, interior
	int y0, halo = 0;
	if(P.blk)
	{
		const int lb = lrow0 >> 5;
		int b = lb - P.blk_hb;                                    // interior, after both halos
		if(lb < P.blk_ht) { b = lb; halo = (P.halo_up != nullptr) ? 1 : 0; }
		else if(lb < P.blk_ht + P.blk_hb) { b = P.nbands_local - P.blk_hb + (lb - P.blk_ht); halo = (P.halo_dn != nullptr) ? 2 : 0; }
		y0 = (P.blk_b0 + b)*32 + (lrow0 & 31);
	}
	else y0 = ((lrow0 >> 5)*P.nshards + P.shard)*32 + (lrow0 & 31);
	const int y = y0 + ty;

	// ---- screen.h:12-13,18: rayl = ((rayb + y*rdy) + cx*rdx) + rdx + ... + rdx --------
	// The x+1 adds are sequential by definition; warp 0 computes each of the CTA's rows once
	// (lane = 3*row + component, 32 chained adds) and shares them through smem.  The other warps
	// wait at the barrier, which costs no issue slots -- the resource this kernel is short of.
	if(warp == 0 && lane < 3*K1_ROWS)
	{
		const int row = lane/3, comp = lane - 3*row;
		const float fy = (float)(y0 + row), fcx = (float)(blockIdx.x*32);
		const float d = P.rdx[comp];
		float v = __fadd_rn(__fadd_rn(P.rayb[comp], __fmul_rn(fy, P.rdy[comp])), __fmul_rn(fcx, d));
		#pragma unroll
		for(int k = 0; k < 32; k++)
		{
			v = __fadd_rn(v, d);
			s_ray[comp][row][k] = v;
		}
	}
	__syncthreads();
	if(x >= P.dimx || y >= P.dimy)
	{
		if(P.done_local != nullptr) band_arrive(P, y0 >> 5, halo, false);   // (the counters count every lane of the grid)
		return;
	}
	const uint32_t pix = (uint32_t)y*(uint32_t)P.rwidth + (uint32_t)x;   // (pwn_frame_config keeps rwidth*rheight below 2^31)
	float irx = s_ray[0][ty][tx], iry = s_ray[1][ty][tx], irz = s_ray[2][ty][tx];

	// screen.h:19-21 (uint32 wrap)
	uint32_t ops = W_PIXEL;
	uint32_t seed = (uint32_t)x + (uint32_t)y*(uint32_t)y*(uint32_t)(P.rwidth + 1);
	seed *= seed*seed;
	seed *= seed*seed;

	float fx = P.from[0], fy = P.from[1], fz = P.from[2];
	float icb = 1.0f, icg = 1.0f, icr = 1.0f, ica = 1.0f;   // icol (screen.h:24)

	// ALPHA0: the alpha lane.  The surface colours have alpha 0 (defs.h:17-19, trace.h:558), so a wall, floor or
	// ceiling shades to alpha = icol.a*0*diffuse = +0 exactly for every finite ray (icol.a is 1 or an earlier
	// +0, diffuse is 0.1..1), and blending it in adds k*(+0) to refl*oa >= +0: nothing.  Only sphere colours
	// carry an alpha of their own, so the builds without spheres leave those operations out; the fog term,
	// which is what makes the alpha byte of a frame non-zero, stays.
	// per-level blend state of the two bounce levels (trace.h:89-101)
	// (colour, refl, fog) of a bounce level wait in shared memory from the bounce to the unwind: twelve
	// values per pixel that would otherwise sit in registers through the walks of the later segments
	float ob, og, orr, oa;   // colour coming back from the deepest segment
	bool exhausted = false;
	int depth = 0;           // bounce levels to unwind

	const int W = L.w, H = L.h, WP = L.wp;

	#pragma unroll K1_SEG_UNROLL
	for(int seg = 0; ; seg++)
	{
		// ================= trace.h:203-248: segment setup =================
		if(IDS) ops += W_SEGMENT;
		float rx, ry, rz;
#if PWN_PACKED
		{
			const f32x2 irxz = pk2(irx, irz);
			float sx, sz;
			unpk2(mul2x(irxz, irxz, nz2), sx, sz);
			const float d = __fadd_rn(__fadd_rn(sx, sz), __fmul_rn(iry, iry));
			const float s = rsqrt_emul<STD>(L, cmap, d);
			unpk2(mul2x(irxz, pk2(s, s), nz2), rx, rz);
			ry = __fmul_rn(iry, s);
		}
#else
		{
			const float d = __fadd_rn(__fadd_rn(__fmul_rn(irx, irx), __fmul_rn(irz, irz)), __fmul_rn(iry, iry));
			const float s = rsqrt_emul<STD>(L, cmap, d);
			rx = __fmul_rn(irx, s); ry = __fmul_rn(iry, s); rz = __fmul_rn(irz, s);
		}
#endif
		int cx = __float2int_rz(fx), cz = __float2int_rz(fz);   // real cell coordinates (IDS builds keep them current)
		const float EPS = 0.0000000000001f;
		// The reference tests the signs twice: gx/gy/gz from the un-normalised ray (trace.h:225-227) and "ray >= 0"
		// on the normalised, clamped one (trace.h:233-238).  A component the clamp leaves alone is non-zero and has
		// the sign of its un-normalised value (the scale is positive), so the second test is the negation of the
		// first; only behind the clamp can the two disagree (a negative component that underflowed to -0 becomes
		// +EPSILON), and there the second test is made for real.  Everything that reads the signs sits in front of
		// the clamp's call, so each is tested once and no predicate has to live across the call.
		const bool xneg = irx < 0.0f, yneg = iry < 0.0f, zneg = irz < 0.0f;
		int gx = (xneg ? -1 : 1), gz = (zneg ? -1 : 1);
		const bool up = !yneg;                                  // gy > 0
		uint32_t upmask = up ? 2u : 0u;
		asm volatile("" : "+r"(upmask));   // a value of its own: ptxas would otherwise rebuild it from the ray at every step
#if PWN_PACKED
		float w0x, w0z, omx, omz;
		unpk2(sub2(pk2(fx, fz), pk2((float)cx, (float)cz)), w0x, w0z);
		unpk2(sub2(pk2(1.0f, 1.0f), pk2(w0x, w0z)), omx, omz);
		float wx = xneg ? w0x : omx, wy = yneg ? fy : __fsub_rn(1.0f, fy), wz = zneg ? w0z : omz;
#else
		const float w0x = __fsub_rn(fx, (float)cx), w0z = __fsub_rn(fz, (float)cz);
		float wx = xneg ? w0x : __fsub_rn(1.0f, w0x), wy = yneg ? fy : __fsub_rn(1.0f, fy), wz = zneg ? w0z : __fsub_rn(1.0f, w0z);
#endif
		// linear index into the padded map.  While a coordinate is outside the grid the lookup
		// folds to index 0 on that axis (util.h:151-158): such rays (oobm != 0) take the slow path
		// of every step, which tracks the real (cx, cz) and recomputes the index from them
		uint32_t oobm = 0u;
		int ci = (cz + 1)*WP + cx + 1;
		const bool outside = ((unsigned)cx >= (unsigned)W) | ((unsigned)cz >= (unsigned)H);
		// the two rare cases of the setup share ONE branch (the clamp is the identity on a ray that does not need it)
		if((fminf(fminf(fabsf(rx), fabsf(ry)), fabsf(rz)) < EPS) | outside)   // trace.h:220-222, almost never taken: a real
		{                                                         // call keeps the clamp out of the instruction stream of every segment
			const float3 c3 = eps_clamp(rx, ry, rz);
			rx = c3.x; ry = c3.y; rz = c3.z;
			wx = (rx >= 0.0f) ? __fsub_rn(1.0f, w0x) : w0x;
			wy = (ry >= 0.0f) ? __fsub_rn(1.0f, fy) : fy;
			wz = (rz >= 0.0f) ? __fsub_rn(1.0f, w0z) : w0z;
			if(outside)
			{
				oobm = OOB_MARK;
				ci = (((unsigned)cz >= (unsigned)H ? 0 : cz) + 1)*WP + ((unsigned)cx >= (unsigned)W ? 0 : cx) + 1;
			}
		}
		float iax = rcp_emul<STD>(L, cmap, rx), iaz = rcp_emul<STD>(L, cmap, rz);   // iavel = rcpps(abs(ray)), trace.h:229-231
		const float iay = rcp_emul<STD>(L, cmap, ry);
		float px = fx, py = fy, pz = fz;
#if PWN_PACKED
		unpk2(mul2x(pk2(wx, wz), pk2(iax, iaz), nz2), wx, wz);
		wy = __fmul_rn(wy, iay);
#else
		wx = __fmul_rn(wx, iax); wy = __fmul_rn(wy, iay); wz = __fmul_rn(wz, iaz);
#endif

		int stepz = gz*WP;   // index step of a z move; an x move steps by gx
		// the faces the ray crosses: FXP/FXN = 0/2 and FZP/FZN = 1/3 for a sign of +1/-1, as arithmetic (the FMA
		// pipe has room, the ALU pipe that would do the selects has not)
		static_assert(PWN_FXP == 0 && PWN_FZP == 1 && PWN_FXN == 2 && PWN_FZN == 3, "face codes");
		int ldx = 1 - gx, ldz = 2 - gz;
		uint32_t cell = __ldg(&cmap[ci]);
		int ldir = PWN_FYN;
		float cdist = 0.0f, fog = 0.0f;
		int nstep = 0, nport = 0;
		Aux aux;
		// "no sphere yet" is +inf here, not the reference's -1 (trace.h:196): then "a sphere, and nearer than where the
		// walk has got to" is ONE compare at every step, and "none yet, or nearer than the one so far" one at every
		// test.  A hit at exactly -1, which the reference then takes for "none", is stored as +inf for the same effect.
		aux.dist = __int_as_float(0x7F800000);
		aux.px = aux.py = aux.pz = 0.0f; aux.diff = 0.0f; aux.si = 0u;

		// how the walk ended: a surface (SF_CEIL..SF_STRAY), a sphere, or not at all (maxsteps, trace.h:677).
		// What the reference stores to *dist is cdist as the walk left it, or the sphere's distance.
		int hsurf = SF_NONE;
#define hkind (hsurf < SF_NONE ? 0 : (hsurf == SF_SPHERE ? 1 : 2))
#define hdist ((SPH && hsurf == SF_SPHERE) ? aux.dist : cdist)

// the "sphere nearer than where we got to" exits (trace.h:313-319 and friends)
// How a lane leaves the walk.  A `break` out of the divergent middle of the loop costs a BREAK (the lane takes itself
// off the loop's convergence barrier) beside the branch; with PWN_EXIT_AT_TAIL the lane instead runs the step counter
// down and jumps to the loop's tail, where the one exit test of the loop lets it go: one instruction fewer on the
// room step.  Whether the walk ended on a surface or ran out of steps is read from hsurf, never from the counter.
#ifndef PWN_EXIT_AT_TAIL
#define PWN_EXIT_AT_TAIL 0
#endif
#if PWN_EXIT_AT_TAIL
#define PWN_LEAVE { maxsteps = 1; continue; }
#else
#define PWN_LEAVE break;
#endif
#define PWN_SPHERE_MARK hsurf = SF_SPHERE; ldir = -1;
#define PWN_SPHERE_EXIT(FOGC, FOGBEG) \
		if(SPH && cdist > aux.dist) { \
			PWN_SPHERE_MARK \
			if((FOGC) && aux.dist > (FOGBEG)) fog = __fadd_rn(fog, __fsub_rn(aux.dist, (FOGBEG))); \
			PWN_LEAVE }
#define PWN_WALL_EXIT(SURF) { hsurf = (SURF); PWN_LEAVE }
// The step just taken (along x if XSTEP) put the ray on the border of the padded map, or the ray
// already was outside the grid: bring the real cell coordinates up to date and look the cell up
// the way get_cell does (util.h:151-158).  Builds with IDS keep cx/cz current themselves.
#define PWN_OOB_FIX(XSTEP) do { \
		if(!oobm) { const int row_ = ci / WP; if(!IDS) { cx = ci - row_*WP - 1; cz = row_ - 1; } } \
		else if(!IDS) { if(XSTEP) cx += gx; else cz += gz; } \
		oobm = (((unsigned)cx >= (unsigned)W) | ((unsigned)cz >= (unsigned)H)) ? (uint32_t)OOB_MARK : 0u; \
		ci = (((unsigned)cz >= (unsigned)H ? 0 : cz) + 1)*WP + ((unsigned)cx >= (unsigned)W ? 0 : cx) + 1; \
		cell = __ldg(&cmap[ci]); } while(0)

// the same, entered by either reason, with the two reasons tested once each
#define PWN_OOB_RELOOK_ \
		oobm = (((unsigned)cx >= (unsigned)W) | ((unsigned)cz >= (unsigned)H)) ? (uint32_t)OOB_MARK : 0u; \
		ci = (((unsigned)cz >= (unsigned)H ? 0 : cz) + 1)*WP + ((unsigned)cx >= (unsigned)W ? 0 : cx) + 1; \
		cell = __ldg(&cmap[ci]);
#define PWN_OOB_CHECK(XSTEP) \
		if(oobm != 0u) { if(!IDS) { if(XSTEP) cx += gx; else cz += gz; } PWN_OOB_RELOOK_ } \
		else if((cell & 15u) == CL_OOB) { const int row_ = ci / WP; if(!IDS) { cx = ci - row_*WP - 1; cz = row_ - 1; } PWN_OOB_RELOOK_ }

		// ================= trace.h:250-675: the walk =================
		#pragma unroll K1_WALK_UNROLL
		for(int maxsteps = 1000; maxsteps > 0; maxsteps--)
		{
			if(IDS) nstep++;

			if(SPH)   // trace.h:252-296
			{
				if((int32_t)cell < 0 && !oobm)   // CM_SPH_BIT (outside the grid the reference looks at no list, trace.h:252)
				{
					const uint32_t pk = __ldg(&L.parts_pk[ci]);
					uint32_t i = pk & 0xFFFFFFu, e = i + (pk >> 24);
					if((pk >> 24) == 255u) e = __ldg(&L.parts_off[ci + 1]);
					for(; i < e; i++)
					{
						const uint32_t si = __ldg(&L.parts_idx[i]);
						const float4 sp4 = __ldg(&L.sph_pos[si]);
						const float2 rr = __ldg(&L.sph_rr[si]);
						const float rad2 = __fmul_rn(rr.x, rr.x);
						const float sx = __fsub_rn(sp4.x, px), sy = __fsub_rn(sp4.y, py), sz = __fsub_rn(sp4.z, pz);
						// v_dot: (x + z) + (y + w), w lanes are 0 (pwn_parts_upload checks)
						const float dist2 = __fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sz, sz)), __fmul_rn(sy, sy));
						const float dot = __fadd_rn(__fadd_rn(__fmul_rn(sx, rx), __fmul_rn(sz, rz)), __fmul_rn(sy, ry));
						if(IDS) ops += W_SPH_TEST;
						if(dot > 0.0f)
						{
							const float calcrad2 = __fsub_rn(dist2, __fmul_rn(dot, dot));
							if(calcrad2 < rad2)
							{
								const float sdist2 = __fsub_rn(1.0f, __fdiv_rn(calcrad2, rad2));
								const float sd = __fsub_rn(__fsqrt_rn(dist2), __fsqrt_rn(sdist2));
								const float ad = __fadd_rn(sd, cdist);
								if(ad < aux.dist)
								{
									if(IDS) ops += W_SPH_HIT;
									aux.dist = (ad == -1.0f) ? __int_as_float(0x7F800000) : ad;
									aux.px = __fadd_rn(px, __fmul_rn(sd, rx));
									aux.py = __fadd_rn(py, __fmul_rn(sd, ry));
									aux.pz = __fadd_rn(pz, __fmul_rn(sd, rz));
									float nx, ny, nz;
									sphere_normal<STD>(L, cmap, aux.px, aux.py, aux.pz, sp4, nx, ny, nz);
									float diff = -__fadd_rn(__fadd_rn(__fmul_rn(rx, nx), __fmul_rn(rz, nz)), __fmul_rn(ry, ny));
									if(diff < 0.0f) diff = 0.0f;
									aux.diff = __fadd_rn(0.2f, __fmul_rn(__fsub_rn(1.0f, 0.2f), diff));
									aux.si = si;
								}
							}
						}
					}
				}
			}

			if(cell & 8u)
			{
				// 1-high (trace.h:302-352) and 2-high (trace.h:354-441) rooms share the step
				const bool fogc = bit_of(cell, 1u);
				const bool up2 = bit_of(cell, upmask);   // 2-high and the ray goes up, in one bit test
				const bool two = IDS ? bit_of(cell, 2u) : false;
				if(IDS) ops += two ? W_TWO : W_ONE;
				if(up2) wy = __fadd_rn(wy, iay);   // ceiling one unit further (trace.h:357)
				const float fogbeg = cdist;        // trace.h:309,359 (read only when fogc)

				// trace_ray_through (trace.h:156-184)
				// the three-way pick as two minima: with no NaN among the distances (iavel is finite, trace.h:220-231)
				// min(wx, wz) is the x/z pick, "y strictly first" is wy < that, and t is the smaller of the two
				const bool xfirst = wx < wz;
				const float mxz = fminf(wx, wz);
				const bool yfirst = wy < mxz;
				const float t = fminf(wy, mxz);
#if PWN_PACKED
				float ax, az;
				{
					const f32x2 tt = pk2(t, t);
					unpk2(add2(mul2x(tt, pk2(rx, rz), nz2), pk2(px, pz)), px, pz);     // pos.xz += t*ray.xz
					unpk2(sub2(pk2(wx, wz), tt), ax, az);
				}
				cdist = __fadd_rn(cdist, t);
				py = __fadd_rn(__fmul_rn(t, ry), py);
#else
				cdist = __fadd_rn(cdist, t);
				px = __fadd_rn(__fmul_rn(t, rx), px);
				py = __fadd_rn(__fmul_rn(t, ry), py);
				pz = __fadd_rn(__fmul_rn(t, rz), pz);
#endif

				PWN_SPHERE_EXIT(fogc, fogbeg)
				if(fogc) fog = __fadd_rn(fog, __fsub_rn(cdist, fogbeg));
				if(yfirst)
				{
					ldir = up ? PWN_FYP : PWN_FYN;
					PWN_WALL_EXIT(up ? SF_CEIL : SF_FLOOR)
				}
				// X/Z cell advance (trace.h:331-340): all lanes minus t, the stepped lane reloads
				{
#if !PWN_PACKED
					const float ax = __fsub_rn(wx, t), az = __fsub_rn(wz, t);
#endif
					wy = __fsub_rn(wy, t);
					wx = xfirst ? iax : ax;
					wz = xfirst ? az : iaz;
					ci += xfirst ? gx : stepz;
					ldir = xfirst ? ldx : ldz;
					if(IDS) { if(xfirst) cx += gx; else cz += gz; }
				}
				if(up2) wy = __fsub_rn(wy, iay);   // trace.h:392
				const uint32_t faceb = cell >> ldir;   // the bits of the face just crossed, in the positions of face 0
				cell = __ldg(&cmap[ci]);

				// slow transitions only (the cell's slow bit for the face just crossed, or a ray
				// outside the grid): leaving the grid, raised floors, 2-high room into anything else
				if((faceb | oobm) & (1u << CM_SLOW_SHIFT))
				{
					if(bit_of(faceb & ~oobm, 1u << CM_HWIN_SHIFT))
					{
						if(py < 0.0f || py > 1.0f) PWN_WALL_EXIT(SF_WALL)   // trace.h:415,432-439
					}
					else
					{
					PWN_OOB_CHECK(xfirst)
					const uint32_t ncls = cell & 15u;
					if(!bit_of(faceb, 1u << CM_TWO_SHIFT))
					{
						if((faceb & (1u << CM_RZ_SHIFT)) != 0u && (ncls == CL_TWO || ncls == CL_TWO_FOG))   // leaving '"': trace.h:343-351
						{
							py = __fadd_rn(py, 1.0f);
							wy = up ? __fsub_rn(wy, iay) : __fadd_rn(wy, iay);
						}
					}
					else
					{
						if(ncls == CL_RAISED)   // trace.h:395-402
						{
							py = __fsub_rn(py, 1.0f);
							wy = up ? __fadd_rn(wy, iay) : __fsub_rn(wy, iay);
						}
						if(py < 0.0f || py > 1.0f)   // trace.h:404-440
						{
							uint32_t pk;
							if(ncls == CL_PORTAL) pk = oobm ? PK_WALL : ((__ldg(&L.portals[(cell >> CM_REC_SHIFT) & (SPH ? CM_REC_MASK : 0xFFFFu)].flags) >> 16) & 255u);
							else pk = (ncls == CL_TWO || ncls == CL_TWO_FOG) ? PK_PASS : (ncls == CL_RAISED ? PK_RAISED : PK_WALL);
							if(pk != PK_PASS)
							{
								if(pk == PK_RAISED)   // revert (trace.h:422-431)
								{
									py = __fadd_rn(py, 1.0f);
									wy = up ? __fsub_rn(wy, iay) : __fadd_rn(wy, iay);
								}
								PWN_WALL_EXIT(SF_WALL)
							}
						}
					}
					}
				}
			}
			else if(cell & 4u)
			{
				// trace.h:443-505
				if(IDS) ops += W_RAMP;
				const float tilt = __fmul_rn(0.5f, bit_of(cell, 2u) ? rz : rx);   // '>' '<' tilt along x, ',' '^' along z
				const bool minus = !bit_of(cell, 1u);                            // '>' and ','
				ry = minus ? __fsub_rn(ry, tilt) : __fadd_rn(ry, tilt);
				wy = py;
				if(ry >= 0.0f) wy = __fsub_rn(1.0f, wy);
				wy = __fmul_rn(wy, __fdiv_rn(1.0f, (ry < 0.0f ? -ry : ry)));
				PWN_SPHERE_EXIT(false, 0.0f)
				// trace_ray_through(..., gy, gy, gz): gy stands in for gx (trace.h:470); the face is
				// then re-derived from the ray's own signs (trace.h:474,481,487)
				const bool yfirst = (wy < wx) & (wy < wz);
				const bool xfirst = wx < wz;
				const float t = yfirst ? wy : (xfirst ? wx : wz);
				cdist = __fadd_rn(cdist, t);
				px = __fadd_rn(__fmul_rn(t, rx), px);
				py = __fadd_rn(__fmul_rn(t, ry), py);
				pz = __fadd_rn(__fmul_rn(t, rz), pz);
				if(yfirst)
				{
					ldir = (ry < 0.0f ? PWN_FYN : PWN_FYP);
					PWN_WALL_EXIT(ry >= 0.0f ? SF_CEIL : SF_FLOOR)
				}
				{
					const float ax = __fsub_rn(wx, t), az = __fsub_rn(wz, t);
					wx = xfirst ? iax : ax;
					wz = xfirst ? az : iaz;
					ci += xfirst ? gx : stepz;
					ldir = xfirst ? (rx < 0.0f ? PWN_FXN : PWN_FXP) : (rz < 0.0f ? PWN_FZN : PWN_FZP);
					if(IDS) { if(xfirst) cx += gx; else cz += gz; }
				}
				// un-tilt: the product is recomputed from the (unchanged) rx/rz, as the source does
				ry = minus ? __fadd_rn(ry, tilt) : __fsub_rn(ry, tilt);
				wy = py;
				if(ry >= 0.0f) wy = __fsub_rn(1.0f, wy);
				wy = __fmul_rn(wy, iay);
				cell = __ldg(&cmap[ci]);
				PWN_OOB_CHECK(xfirst)
				PWN_SPHERE_EXIT(false, 0.0f)   // trace.h:668-673: the ramp looked before it stepped
			}
			else if((cell & 15u) == CL_PORTAL)
			{
				// trace.h:508-650
				const uint4 pa = __ldg(reinterpret_cast<const uint4 *>(&L.portals[(cell >> CM_REC_SHIFT) & (SPH ? CM_REC_MASK : 0xFFFFu)]));
				uint32_t pkind = (pa.w >> 8) & 255u;
				if(oobm && pkind == PR_OK) pkind = PR_STRAY;   // outside the grid no end matches (trace.h:530,539)
				if(pkind != PR_OK)
				{
					PWN_SPHERE_EXIT(false, 0.0f)
					PWN_WALL_EXIT(pkind == PR_UNPAIRED ? SF_WALL : SF_STRAY)
				}
				const uint4 pb = __ldg(reinterpret_cast<const uint4 *>(&L.portals[(cell >> CM_REC_SHIFT) & (SPH ? CM_REC_MASK : 0xFFFFu)]) + 1);
				ci = (int)pa.x;
				if(IDS) { cx = (int)pb.z; cz = (int)pb.w; ops += W_PORTAL; nport++; }
				px = __fadd_rn(px, __uint_as_float(pa.y));
				pz = __fadd_rn(pz, __uint_as_float(pa.z));
				const int rot = (int)(pa.w & 255u);
				ldir = (ldir - rot) & 3;
				const float ccx = __uint_as_float(pb.x), ccz = __uint_as_float(pb.y);
				if(rot == 1)
				{
					const float trx = px, trz = pz, trvx = rx;
					px = __fadd_rn(ccx, __fsub_rn(trz, ccz));
					pz = __fsub_rn(ccz, __fsub_rn(trx, ccx));
					rx = xmov(rz, P.zero); rz = -trvx;
					const int tg = gx; gx = gz ^ (int)P.zero; gz = -tg;
					PWN_XCHG
				}
				else if(rot == 2)
				{
					px = __fsub_rn(__fmul_rn(ccx, 2.0f), px);
					pz = __fsub_rn(__fmul_rn(ccz, 2.0f), pz);
					rx = -rx; rz = -rz;
					gx = -gx; gz = -gz;
				}
				else if(rot == 3)
				{
					const float trx = px, trz = pz, trvx = rx;
					px = __fsub_rn(ccx, __fsub_rn(trz, ccz));
					pz = __fadd_rn(ccz, __fsub_rn(trx, ccx));
					rx = -rz; rz = xmov(trvx, P.zero);
					const int tg = gx; gx = -gz; gz = tg ^ (int)P.zero;
					PWN_XCHG
				}
				stepz = gz*WP;
				ldx = 1 - gx; ldz = 2 - gz;
				// step out of the destination portal cell (trace.h:624-647)
				{
					// FXP FZP FXN FZN = 0 1 2 3: bit 0 = along z, bit 1 = backwards (a - 1.0f and a + (-1.0f) are
					// the same operation); selects instead of a four-way branch
					const bool onz = (ldir & 1) != 0, neg = (ldir & 2) != 0;
					const float one = neg ? -1.0f : 1.0f;
					const int sgn = neg ? -1 : 1;
					ci += onz ? sgn*WP : sgn;
					px = onz ? px : __fadd_rn(px, one);
					pz = onz ? __fadd_rn(pz, one) : pz;
					if(IDS) { if(onz) cz += sgn; else cx += sgn; }
				}
				cell = __ldg(&cmap[ci]);
				if((cell & 15u) == CL_OOB) PWN_OOB_FIX(false);   // a portal on the rim: never outside before the step
				PWN_SPHERE_EXIT(false, 0.0f)   // trace.h:668-673
			}
			else
			{
				// solid (trace.h:651-664)
				PWN_SPHERE_EXIT(false, 0.0f)
				PWN_WALL_EXIT(ldir == PWN_FYP ? SF_CEIL : SF_WALL)
			}

			// (trace.h:668-673, the look at the end of every iteration, sits at the end of the ramp and portal
			// branches: after a room step it repeats the look the step has just taken, with nothing changed)
		}
#undef PWN_SPHERE_EXIT
#undef PWN_SPHERE_MARK
#undef PWN_WALL_EXIT
#undef PWN_LEAVE
#undef PWN_OOB_FIX
#undef PWN_OOB_CHECK
#undef PWN_OOB_RELOOK_

		if(IDS && P.ids != nullptr)
		{
			pwn_ids rec;
			rec.cell = cz*W + cx; rec.ldir = (int16_t)ldir; rec.kind = (int16_t)hkind;
			rec.nstep = nstep; rec.nportal = nport;
			P.ids[(size_t)pix*3 + seg] = rec;
		}

		// zbuf keeps the PRIMARY distance: bounces save/zero/restore *dist (trace.h:86-87,103);
		// an exhausted walk leaves it untouched (trace.h:677-678)
		if(seg == 0 && hkind != 2) P.zbuf[pix] = hdist;

		if(hkind == 2)
		{
			// the ray comes back as the colour (trace.h:677-678).  It goes through the scratch words: written as
			// register copies, ptxas hoists the copies above this branch, into every segment's path
			PWN_STS(sthr, 6144, rx); PWN_STS(sthr, 6656, ry); PWN_STS(sthr, 7168, rz);
			exhausted = true;
			break;
		}

		// ================= shading =================
		// the shaded colour is the child's icol (trace.h:90): it is built in place
		float &cb = icb, &cg = icg, &cr = icr, &ca = ica;
		float refl;
		if(!SPH || hkind == 0)
		{
			// trace_hit_wall, trace.h:108-154
			if(IDS) ops += W_WALL;
			// one indexed load from the constant bank instead of a four-way branch: the lanes of a warp hit
			// ceilings, floors and walls side by side
			const float4 sc = PWN_SURFACE(hsurf);
			const float sb = sc.x, sg = sc.y, sr = sc.z;
#if PWN_PACKED
			unpk2(mul2x(pk2(icb, icg), pk2(sb, sg), nz2), cb, cg);
			cr = __fmul_rn(icr, sr);
#else
			cb = __fmul_rn(icb, sb); cg = __fmul_rn(icg, sg); cr = __fmul_rn(icr, sr);
#endif
			if(SPH) ca = __fmul_rn(ica, 0.0f);   // (without spheres the alpha lane of a shaded colour is +0 exactly, see ALPHA0)
			// diffuse = +-ray[axis(ldir)]: FXP FZP FXN FZN FYP FYN = 0..5
			float diffuse = (ldir >= PWN_FYP) ? ry : ((ldir & 1) ? rz : rx);
			// FXN, FZN, FYN (2, 3, 5) face against their axis: bit ldir of 0b101100, moved up to the sign bit
			diffuse = __uint_as_float(__float_as_uint(diffuse) ^ ((0x2Cu << (31 - ldir)) & 0x80000000u));
			if(diffuse < 0.0f) diffuse = 0.0f;
			diffuse = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, 0.1f), diffuse), 0.1f);
#if PWN_PACKED
			unpk2(mul2x(pk2(cb, cg), pk2(diffuse, diffuse), nz2), cb, cg);
			cr = __fmul_rn(cr, diffuse);
#else
			cb = __fmul_rn(cb, diffuse); cg = __fmul_rn(cg, diffuse); cr = __fmul_rn(cr, diffuse);
#endif
			if(SPH) ca = __fmul_rn(ca, diffuse);
			refl = (ldir == PWN_FYN ? 0.7f : 0.25f);
		}
		else
		{
			// aux_col = diff * sphere colour, aux_refl (trace.h:288-292), rebuilt from the record
			const float4 c4 = __ldg(&L.sph_col[aux.si]);
			cb = __fmul_rn(aux.diff, c4.x); cg = __fmul_rn(aux.diff, c4.y);
			cr = __fmul_rn(aux.diff, c4.z); ca = __fmul_rn(aux.diff, c4.w);
			refl = __ldg(&L.sph_rr[aux.si]).y;
		}

		// ================= trace_hit_bounce, trace.h:3-106 =================
		if(seg >= P.reflect || refl == 0.0f)   // hitctr >= REFLECT (trace.h:6-7; REFLECT == 2 as shipped, defs.h:8)
		{
			break;
		}

		if(IDS) ops += W_BOUNCE + (ldir == PWN_FYN ? W_FLOOR : 0u) + (fog != 0.0f ? W_FOG : 0u);
		float nrx = rx, nry = ry, nrz = rz;
		float npx, npy, npz;
		if(SPH && hkind == 1) { npx = aux.px; npy = aux.py; npz = aux.pz; }
		else { npx = px; npy = py; npz = pz; }

		// the five flat mirrors (trace.h:12-37): flip the component of the face's axis and step 0.001 back
		// off the face (a - 0.001f and a + (-0.001f) are the same IEEE operation).  One predicated
		// sequence for all five: the lanes of a warp hit different faces, and the compiler's jump table
		// for an if-chain would run once per face present.
		if((unsigned)ldir <= (unsigned)PWN_FYP)
		{
			const float off = (ldir & 2) ? 0.001f : -0.001f;          // FXN, FZN: +; FXP, FZP, FYP: -
			const bool onx = (ldir == PWN_FXP) | (ldir == PWN_FXN), onz = (ldir == PWN_FZP) | (ldir == PWN_FZN);
			const bool ony = ldir == PWN_FYP;
			nrx = onx ? -nrx : nrx; npx = onx ? __fadd_rn(npx, off) : npx;
			nrz = onz ? -nrz : nrz; npz = onz ? __fadd_rn(npz, off) : npz;
			nry = ony ? -nry : nry; npy = ony ? __fadd_rn(npy, off) : npy;
		}
		else
		{
			float nx, ny, nz;
			if(!SPH || ldir == PWN_FYN)
			{
				// wavy floor normal (trace.h:39-46)
				npy = __fsub_rn(npy, 0.001f);
				const float HPI = __uint_as_float(0x3FC90FDBu);   // (float)M_PI*0.5f
				const float TPI = __uint_as_float(0x40C90FDBu);   // (float)M_PI*2.0f
				const float a0 = libm_sinf<LFMA>(__fmul_rn(HPI, npx));
				const float a1 = libm_cosf<LFMA>(__fmul_rn(HPI, npz));
				const float ang = __fmul_rn(TPI, __fadd_rn(__fadd_rn(a0, a1), P.sec_current));
				float s, c;
				libm_sincosf<LFMA>(ang, &s, &c);
				// v_normalise(sin, 38, cos, 0): (x^2 + z^2) + (y^2 + 0)
				const float nd = __fadd_rn(__fadd_rn(__fmul_rn(s, s), __fmul_rn(c, c)), __fmul_rn(38.0f, 38.0f));
				const float ns = rsqrt_emul<STD>(L, cmap, nd);
				nx = __fmul_rn(s, ns); ny = __fmul_rn(38.0f, ns); nz = __fmul_rn(c, ns);
			}
			else
			{
				// sphere (trace.h:58-64)
				npx = __fsub_rn(npx, __fmul_rn(0.001f, nrx));
				npy = __fsub_rn(npy, __fmul_rn(0.001f, nry));
				npz = __fsub_rn(npz, __fmul_rn(0.001f, nrz));
				sphere_normal<STD>(L, cmap, aux.px, aux.py, aux.pz, __ldg(&L.sph_pos[aux.si]), nx, ny, nz);   // aux_norm, rebuilt
			}
			// rmul = -2*(((0 + rx*nx) + ry*ny) + rz*nz); ray = normalise(rmul*norm + ray)
#if PWN_PACKED
			{
				const f32x2 nxz = pk2(nx, nz);
				float dx, dz, qx, qz;
				unpk2(mul2x(pk2(nrx, nrz), nxz, nz2), dx, dz);
				const float rmul = __fmul_rn(-2.0f, __fadd_rn(__fadd_rn(dx, __fmul_rn(nry, ny)), dz));
				unpk2(add2(mul2x(pk2(rmul, rmul), nxz, nz2), pk2(nrx, nrz)), nrx, nrz);
				nry = __fadd_rn(__fmul_rn(rmul, ny), nry);
				const f32x2 rxz2 = pk2(nrx, nrz);
				unpk2(mul2x(rxz2, rxz2, nz2), qx, qz);
				const float d = __fadd_rn(__fadd_rn(qx, qz), __fmul_rn(nry, nry));
				const float s = rsqrt_emul<STD>(L, cmap, d);
				unpk2(mul2x(rxz2, pk2(s, s), nz2), nrx, nrz);
				nry = __fmul_rn(nry, s);
			}
#else
			const float rmul = __fmul_rn(-2.0f, __fadd_rn(__fadd_rn(__fmul_rn(nrx, nx), __fmul_rn(nry, ny)), __fmul_rn(nrz, nz)));
			nrx = __fadd_rn(__fmul_rn(rmul, nx), nrx);
			nry = __fadd_rn(__fmul_rn(rmul, ny), nry);
			nrz = __fadd_rn(__fmul_rn(rmul, nz), nrz);
			const float d = __fadd_rn(__fadd_rn(__fmul_rn(nrx, nrx), __fmul_rn(nrz, nrz)), __fmul_rn(nry, nry));
			const float s = rsqrt_emul<STD>(L, cmap, d);
			nrx = __fmul_rn(nrx, s); nry = __fmul_rn(nry, s); nrz = __fmul_rn(nrz, s);
#endif
		}

		if(P.reflect_blur)   // trace.h:77-84: draws x, y, (discard), z, (discard)
		{
#if PWN_PACKED
			{
				// two draws, then both lanes together: randfs = r*fl(2/3759) - 1, ray += randfs*0.03
				seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
				const float r0 = (float)(seed % 3759u);
				seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
				const float r1 = (float)(seed % 3759u);
				const float c = __uint_as_float(0x3A0B79B3u);
				const f32x2 v = sub2(mul2x(pk2(r0, r1), pk2(c, c), nz2), pk2(1.0f, 1.0f));
				unpk2(add2(pk2(nrx, nry), mul2x(v, pk2(0.03f, 0.03f), nz2)), nrx, nry);
			}
#else
			nrx = __fadd_rn(nrx, __fmul_rn(randfs(seed), 0.03f));
			nry = __fadd_rn(nry, __fmul_rn(randfs(seed), 0.03f));
#endif
			seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
			nrz = __fadd_rn(nrz, __fmul_rn(randfs(seed), 0.03f));
			seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
		}

		// (indexed directly: through a pointer variable the compiler addresses shared memory generically
		// and rebuilds the window base from the cluster CTA id at every bounce)
		{
			const uint32_t sl = sthr + (uint32_t)seg*3072u;
			PWN_STS(sl, 0, cb); PWN_STS(sl, 512, cg); PWN_STS(sl, 1024, cr);
			if(SPH) PWN_STS(sl, 1536, ca);
			PWN_STS(sl, 2048, refl); PWN_STS(sl, 2560, fog);
		}
		depth = seg + 1;

		// child: trace_ray(hitctr+1, seed, lv, dist, &pos, &ray, col)  (trace.h:90)
		fx = npx; fy = npy; fz = npz;
		irx = nrx; iry = nry; irz = nrz;
	}
#undef hkind
#undef hdist

	// the colour the deepest segment returns: its shaded hit (built in place in icol), or its ray if the walk
	// ran out of steps (trace.h:677-678)
	if(exhausted) { PWN_LDS(ob, sthr, 6144); PWN_LDS(og, sthr, 6656); PWN_LDS(orr, sthr, 7168); oa = 0.0f; }
	else { ob = icb; og = icg; orr = icr; oa = SPH ? ica : 0.0f; }
	// ---- unwind the blends, deepest first (trace.h:91-101) ------------------------------
	#pragma unroll
	for(int lvl = 1; lvl >= 0; lvl--)
	if(lvl < depth)
	{
		const uint32_t sl = sthr + (uint32_t)lvl*3072u;
		float refl, fogl, lb, lg, lr, la = 0.0f;
		PWN_LDS(refl, sl, 2048); PWN_LDS(fogl, sl, 2560);
		PWN_LDS(lb, sl, 0); PWN_LDS(lg, sl, 512); PWN_LDS(lr, sl, 1024);
		if(SPH) PWN_LDS(la, sl, 1536);
		const float k = __fsub_rn(1.0f, refl);
#if PWN_PACKED
		{
			// (b, g) and (r, a) as pairs.  Without spheres the parked alpha is +0 (ALPHA0): refl*oa + k*(+0) is refl*oa,
			// which is never -0, so the pair form gives the same bits as leaving the term out.
			const f32x2 rr2 = pk2(refl, refl), kk2 = pk2(k, k);
			unpk2(add2(mul2x(rr2, pk2(ob, og), nz2), mul2x(kk2, pk2(lb, lg), nz2)), ob, og);
			unpk2(add2(mul2x(rr2, pk2(orr, oa), nz2), mul2x(kk2, pk2(lr, la), nz2)), orr, oa);
		}
#else
		ob = __fadd_rn(__fmul_rn(refl, ob), __fmul_rn(k, lb));
		og = __fadd_rn(__fmul_rn(refl, og), __fmul_rn(k, lg));
		orr = __fadd_rn(__fmul_rn(refl, orr), __fmul_rn(k, lr));
		if(SPH) oa = __fadd_rn(__fmul_rn(refl, oa), __fmul_rn(k, la));
		else oa = __fmul_rn(refl, oa);   // ALPHA0: + k*(+0) changes nothing, refl*oa is never -0
#endif
		if(fogl != 0.0f)
		{
			const float f = libm_expf<LFMA>(__fmul_rn(-0.6f, fogl)), g = __fsub_rn(1.0f, f);
#if PWN_PACKED
			const f32x2 ff2 = pk2(f, f), gg2 = pk2(g, g);
			unpk2(add2(mul2x(ff2, pk2(ob, og), nz2), gg2), ob, og);
			unpk2(add2(mul2x(ff2, pk2(orr, oa), nz2), gg2), orr, oa);
#else
			ob = __fadd_rn(__fmul_rn(f, ob), g); og = __fadd_rn(__fmul_rn(f, og), g);
			orr = __fadd_rn(__fmul_rn(f, orr), g); oa = __fadd_rn(__fmul_rn(f, oa), g);
#endif
		}
	}

	if(IDS && P.opcount != nullptr)
	{
		// (a sum over whatever lanes are here: partial warps at the frame's edges have holes in the mask, which a
		// shuffle butterfly would mis-add; a lane's count is far below 2^32/32)
		const unsigned m = __activemask();
		const unsigned tot = __reduce_add_sync(m, ops);
		if((threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(P.opcount, (unsigned long long)tot);
	}

	// ---- col_ftoint (util.h:48-59): B | G<<8 | R<<16 | A<<24 ------------------------------
	const uint32_t word = pack_word(ob, og, orr, oa);
	P.sbuf[pix] = word;
	for(int p = 0; p < P.npeers; p++) P.peer[p][pix] = word;
	if(P.done_local != nullptr)
	{
		if(halo == 1) P.halo_up[pix] = word;
		if(halo == 2) P.halo_dn[pix] = word;
		band_arrive(P, y0 >> 5, halo, true);
	}

	// ---- pose batches: this thread also blurs its pixel of the previous frame (screen.h:69-123),
	// whose trace ran in the previous launch.  The launches are chained with programmatic dependent
	// launch: the trace above needs nothing from the previous grid, so this grid may start while
	// that one drains; only here does it wait for it (griddepcontrol.wait = previous grid complete
	// and flushed), and only after that does it let the next grid in, so at most two grids overlap.
	if(BLURP)
	{
		asm volatile("griddepcontrol.wait;" ::: "memory");
		asm volatile("griddepcontrol.launch_dependents;");
		blur_pixel<false>(P.prev, x, y);
	}
}

} // namespace pwn
