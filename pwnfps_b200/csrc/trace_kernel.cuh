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
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pwntrace.h"
#include "pwn_libm.cuh"

namespace pwn {

// ---- device layout of a level (flattened by pwn_level_upload) -----------------------
enum : uint32_t
{
	CL_SOLID = 0,      // anything not listed below (trace.h:651-664)
	CL_ONE = 1,        // ';'  1-high            (trace.h:302-352)
	CL_ONE_FOG = 2,    // '$'  1-high, foggy
	CL_RAISED = 3,     // '"'  1-high on the upper floor of a 2-high room
	CL_TWO = 4,        // '#'  2-high            (trace.h:354-441)
	CL_TWO_FOG = 5,    // '&'  2-high, foggy
	CL_RAMP_GT = 6,    // '>'  (trace.h:443-505)
	CL_RAMP_LT = 7,    // '<'
	CL_RAMP_COMMA = 8, // ','
	CL_RAMP_CARET = 9, // '^'
	CL_PORTAL = 10     // letter cell; bits 4..15 of the map entry index PortalRec
};

enum : uint32_t { PK_WALL = 0, PK_PASS = 1, PK_RAISED = 2 };   // what a 2-high room sees through a cell (trace.h:404-440)
enum : uint32_t { PR_OK = 0, PR_UNPAIRED = 1, PR_STRAY = 2 };  // trace.h:512-526 / 547-559

// One record per portal CELL (not per letter): the crossing transform of
// trace.h:528-647 for a ray standing on this cell, precomputed.  It is the
// 3x4 affine map {R(rot*90deg) about (dcx+.5, dcz+.5), translate (dx,dz)} in
// factored form -- the reference evaluates it in a fixed association order
// (trace.h:583-616) that a generic matrix product would not reproduce bit for bit.
struct __align__(16) PortalRec
{
	int16_t dcx, dcz;   // destination portal cell
	int16_t dx, dz;     // pos.xz += (dx,dz) before rotating
	uint8_t rot;        // quarter turns 0..3
	uint8_t kind;       // PR_*
	uint8_t peek;       // PK_* of the cell in front of the other end (pm->c2 / c1)
	uint8_t pad;
	uint32_t pad2;
};

struct LevelDev
{
	const uint16_t *cellmap;   // h*w entries: class | rec<<4
	const PortalRec *portals;
	int w, h;
	const uint32_t *rcp_tab;   // 1<<rcp_bits
	const uint32_t *rsqrt_tab; // 2<<rsqrt_bits
	int rcp_shift, rsqrt_shift;   // 23 - bits
	// spheres (pwn_parts_upload)
	const float4 *sph_pos;     // x y z w
	const float4 *sph_col;     // b g r a
	const float2 *sph_rr;      // r, refl
	const uint32_t *parts_off; // h*w+1
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
	uint32_t *sbuf;            // pre-blur or final frame (BGRA8)
	float *zbuf;
	pwn_ids *ids;              // [rwidth*rheight][3] or null
	unsigned long long *opcount; // IDS builds: sum of algorithmic FP32 lane-ops (DESIGN.md "Work model")
	int reflect_blur;
};

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
// rcpps(x), x > 0 normal
__device__ __forceinline__ float rcp_emul(const LevelDev &L, float x)
{
	uint32_t u = __float_as_uint(x);
	uint32_t t = __ldg(&L.rcp_tab[(u & 0x7FFFFFu) >> L.rcp_shift]);
	return __uint_as_float(t - (((u >> 23) - 127u) << 23));
}

// rsqrtps(x), x > 0 normal
__device__ __forceinline__ float rsqrt_emul(const LevelDev &L, float x)
{
	uint32_t u = __float_as_uint(x);
	int e = (int)(u >> 23) - 127;
	int p = e & 1;
	uint32_t idx = ((u & 0x7FFFFFu) >> L.rsqrt_shift) + ((uint32_t)p << (23 - L.rsqrt_shift));
	uint32_t t = __ldg(&L.rsqrt_tab[idx]);
	return __uint_as_float(t - ((uint32_t)((e - p) >> 1) << 23));
}

// util.h:1-16 with the shipped -ffast-math folding of randfs: r*fl(2/3759) - 1
__device__ __forceinline__ float randfs(uint32_t &seed)
{
	seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
	uint32_t r = seed % 3759u;
	return __fsub_rn(__fmul_rn((float)r, __uint_as_float(0x3A0B79B3u)), 1.0f);
}

// cvtps2dq + packssdw + packuswb on one lane (util.h:48-59)
__device__ __forceinline__ uint32_t pack_lane(float v)
{
	float s = __fmul_rn(v, 255.0f);
	// cvtps2dq: RN-even; NaN and |s| >= 2^31 give 0x80000000, which the packs clamp to 0
	if(!(s < 2147483648.0f)) return 0u;
	int q = __float2int_rn(s);
	return (uint32_t)min(max(q, 0), 255);
}

__device__ __forceinline__ uint16_t get_cell(const LevelDev &L, int cx, int cz)
{
	// util.h:151-158: out-of-range coordinates fold to index 0 on that axis
	if((unsigned)cx >= (unsigned)L.w) cx = 0;
	if((unsigned)cz >= (unsigned)L.h) cz = 0;
	return __ldg(&L.cellmap[cz*L.w + cx]);
}

// surface colours, lanes b g r a (defs.h:17-19, trace.h:558)
enum : int { SF_CEIL = 0, SF_FLOOR = 1, SF_WALL = 2, SF_STRAY = 3 };

struct Hit
{
	int kind;        // 0 wall/floor/ceiling, 1 sphere, 2 exhausted
	int ldir;
	int surf;
	float dist;      // value the reference stores to *dist
};

struct Aux   // nearest sphere so far (trace.h:196-200)
{
	float dist, refl;
	float px, py, pz;
	float nx, ny, nz;
	float cb, cg, cr, ca;
};

template <bool SPH, bool IDS>
__global__ void __launch_bounds__(256) k_trace(const __grid_constant__ TraceParams P)
{
	const LevelDev &L = P.lv;

	// ---- pixel of this thread: CTA = 32x8 pixels, warp = 8x4 ----------------------
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int x = blockIdx.x*32 + (warp & 3)*8 + (lane & 7);
	const int lrow = blockIdx.y*8 + (warp >> 2)*4 + (lane >> 3);
	// local row -> global row through the band interleave
	const int y = ((lrow >> 5)*P.nshards + P.shard)*32 + (lrow & 31);
	if(x >= P.dimx || y >= P.dimy) return;
	const size_t pix = (size_t)y*P.rwidth + x;

	// ---- screen.h:12-13,18: rayl = ((rayb + y*rdy) + cx*rdx) + rdx + ... + rdx ----
	float irx, iry, irz;
	{
		const float fy = (float)y, fcx = (float)(x & ~31);
		irx = __fadd_rn(__fadd_rn(P.rayb[0], __fmul_rn(fy, P.rdy[0])), __fmul_rn(fcx, P.rdx[0]));
		iry = __fadd_rn(__fadd_rn(P.rayb[1], __fmul_rn(fy, P.rdy[1])), __fmul_rn(fcx, P.rdx[1]));
		irz = __fadd_rn(__fadd_rn(P.rayb[2], __fmul_rn(fy, P.rdy[2])), __fmul_rn(fcx, P.rdx[2]));
		const int nadd = (x & 31) + 1;
		for(int k = 0; k < nadd; k++)
		{
			irx = __fadd_rn(irx, P.rdx[0]);
			iry = __fadd_rn(iry, P.rdx[1]);
			irz = __fadd_rn(irz, P.rdx[2]);
		}
	}

	// screen.h:19-21 (uint32 wrap)
	uint32_t ops = W_PIXEL;
	uint32_t seed = (uint32_t)x + (uint32_t)y*(uint32_t)y*(uint32_t)(P.rwidth + 1);
	seed *= seed*seed;
	seed *= seed*seed;

	float fx = P.from[0], fy = P.from[1], fz = P.from[2];
	float icb = 1.0f, icg = 1.0f, icr = 1.0f, ica = 1.0f;   // icol (screen.h:24)

	// per-level blend state of the two bounce levels (trace.h:89-101)
	float b0b = 0, b0g = 0, b0r = 0, b0a = 0, refl0 = 0, fog0 = 0;
	float b1b = 0, b1g = 0, b1r = 0, b1a = 0, refl1 = 0, fog1 = 0;
	float ob, og, orr, oa;   // colour coming back from the deepest segment
	int depth = 0;           // bounce levels to unwind

	#pragma unroll 1
	for(int seg = 0; ; seg++)
	{
		// ================= trace.h:203-248: segment setup =================
		if(IDS) ops += W_SEGMENT;
		float rx, ry, rz;
		{
			const float d = __fadd_rn(__fadd_rn(__fmul_rn(irx, irx), __fmul_rn(irz, irz)), __fmul_rn(iry, iry));
			const float s = rsqrt_emul(L, d);
			rx = __fmul_rn(irx, s); ry = __fmul_rn(iry, s); rz = __fmul_rn(irz, s);
		}
		int cx = __float2int_rz(fx), cz = __float2int_rz(fz);
		const float EPS = 0.0000000000001f;
		if(rx > -EPS && rx < EPS) rx = (rx < 0.0f ? -EPS : EPS);
		if(ry > -EPS && ry < EPS) ry = (ry < 0.0f ? -EPS : EPS);
		if(rz > -EPS && rz < EPS) rz = (rz < 0.0f ? -EPS : EPS);
		int gx = (irx < 0.0f ? -1 : 1), gz = (irz < 0.0f ? -1 : 1);
		const int gy = (iry < 0.0f ? -1 : 1);
		float iax = rcp_emul(L, fabsf(rx)), iaz = rcp_emul(L, fabsf(rz));
		const float iay = rcp_emul(L, fabsf(ry));
		float px = fx, py = fy, pz = fz;
		float wx = __fsub_rn(px, (float)cx), wy = py, wz = __fsub_rn(pz, (float)cz);
		if(rx >= 0.0f) wx = __fsub_rn(1.0f, wx);
		if(ry >= 0.0f) wy = __fsub_rn(1.0f, wy);
		if(rz >= 0.0f) wz = __fsub_rn(1.0f, wz);
		wx = __fmul_rn(wx, iax); wy = __fmul_rn(wy, iay); wz = __fmul_rn(wz, iaz);

		uint32_t cell = get_cell(L, cx, cz);
		int ldir = PWN_FYN;
		float cdist = 0.0f, fog = 0.0f, fogbeg = 0.0f;
		int nstep = 0, nport = 0;
		Aux aux;
		aux.dist = -1.0f; aux.refl = 0.25f;
		aux.px = aux.py = aux.pz = 0.0f; aux.nx = aux.ny = aux.nz = 0.0f;
		aux.cb = aux.cg = aux.cr = aux.ca = 1.0f;

		Hit hit;
		hit.kind = 2; hit.ldir = PWN_FYN; hit.surf = SF_WALL; hit.dist = 0.0f;

// trace.h:156-184
#define PWN_THROUGH(GX, GY, GZ) do { \
		float t_; \
		if(wy < wx && wy < wz) { t_ = wy; ldir = ((GY) < 0 ? PWN_FYN : PWN_FYP); } \
		else if(wx < wz) { t_ = wx; ldir = ((GX) < 0 ? PWN_FXN : PWN_FXP); } \
		else { t_ = wz; ldir = ((GZ) < 0 ? PWN_FZN : PWN_FZP); } \
		cdist = __fadd_rn(cdist, t_); \
		px = __fadd_rn(__fmul_rn(t_, rx), px); \
		py = __fadd_rn(__fmul_rn(t_, ry), py); \
		pz = __fadd_rn(__fmul_rn(t_, rz), pz); } while(0)
// the "sphere nearer than where we got to" exits (trace.h:313-319 and friends)
#define PWN_SPHERE_EXIT(FOGC) \
		if(SPH && aux.dist != -1.0f && cdist > aux.dist) { \
			hit.kind = 1; hit.ldir = -1; hit.dist = aux.dist; \
			if((FOGC) && aux.dist > fogbeg) fog = __fadd_rn(fog, __fsub_rn(aux.dist, fogbeg)); \
			break; }
#define PWN_WALL_EXIT(SURF) { hit.kind = 0; hit.ldir = ldir; hit.surf = (SURF); hit.dist = cdist; break; }
// X/Z cell advance (trace.h:331-340)
#define PWN_STEP_XZ() do { \
		if(ldir == PWN_FXN || ldir == PWN_FXP) { \
			wy = __fsub_rn(wy, wx); wz = __fsub_rn(wz, wx); wx = iax; cx += gx; \
		} else { \
			wy = __fsub_rn(wy, wz); wx = __fsub_rn(wx, wz); wz = iaz; cz += gz; } } while(0)

		// ================= trace.h:250-675: the walk =================
		#pragma unroll 1
		for(int maxsteps = 1000; maxsteps > 0; maxsteps--)
		{
			nstep++;

			if(SPH)   // trace.h:252-296
			{
				if((unsigned)cx < (unsigned)L.w && (unsigned)cz < (unsigned)L.h)
				{
					const int ci = cz*L.w + cx;
					const uint32_t e = __ldg(&L.parts_off[ci + 1]);
					for(uint32_t i = __ldg(&L.parts_off[ci]); i < e; i++)
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
								if(aux.dist == -1.0f || ad < aux.dist)
								{
									if(IDS) ops += W_SPH_HIT;
									aux.dist = ad;
									aux.px = __fadd_rn(px, __fmul_rn(sd, rx));
									aux.py = __fadd_rn(py, __fmul_rn(sd, ry));
									aux.pz = __fadd_rn(pz, __fmul_rn(sd, rz));
									float nx = __fsub_rn(aux.px, sp4.x), ny = __fsub_rn(aux.py, sp4.y), nz = __fsub_rn(aux.pz, sp4.z);
									const float nd = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(nz, nz)), __fmul_rn(ny, ny));
									const float ns = rsqrt_emul(L, nd);
									nx = __fmul_rn(nx, ns); ny = __fmul_rn(ny, ns); nz = __fmul_rn(nz, ns);
									aux.nx = nx; aux.ny = ny; aux.nz = nz;
									float diff = -__fadd_rn(__fadd_rn(__fmul_rn(rx, nx), __fmul_rn(rz, nz)), __fmul_rn(ry, ny));
									if(diff < 0.0f) diff = 0.0f;
									aux.refl = rr.y;
									diff = __fadd_rn(0.2f, __fmul_rn(__fsub_rn(1.0f, 0.2f), diff));
									const float4 c4 = __ldg(&L.sph_col[si]);
									aux.cb = __fmul_rn(diff, c4.x); aux.cg = __fmul_rn(diff, c4.y);
									aux.cr = __fmul_rn(diff, c4.z); aux.ca = __fmul_rn(diff, c4.w);
								}
							}
						}
					}
				}
			}

			const uint32_t cls = cell & 15u;

			if(cls >= CL_ONE && cls <= CL_TWO_FOG)
			{
				// 1-high (trace.h:302-352) and 2-high (trace.h:354-441) rooms share the step
				const bool two = cls >= CL_TWO;
				if(IDS) ops += two ? W_TWO : W_ONE;
				const bool fogc = (cls == CL_ONE_FOG) || (cls == CL_TWO_FOG);
				const bool up2 = two && gy > 0;
				if(up2) wy = __fadd_rn(wy, iay);   // ceiling one unit further (trace.h:357)
				if(fogc) fogbeg = cdist;
				PWN_THROUGH(gx, gy, gz);
				PWN_SPHERE_EXIT(fogc)
				if(fogc) fog = __fadd_rn(fog, __fsub_rn(cdist, fogbeg));
				if(ldir >= PWN_FYP) PWN_WALL_EXIT(gy > 0 ? SF_CEIL : SF_FLOOR)
				PWN_STEP_XZ();
				if(up2) wy = __fsub_rn(wy, iay);   // trace.h:392
				cell = get_cell(L, cx, cz);
				const uint32_t ncls = cell & 15u;
				if(!two)
				{
					if(cls == CL_RAISED && (ncls == CL_TWO || ncls == CL_TWO_FOG))   // trace.h:343-351
					{
						py = __fadd_rn(py, 1.0f);
						wy = (gy < 0) ? __fadd_rn(wy, iay) : __fsub_rn(wy, iay);
					}
				}
				else
				{
					if(ncls == CL_RAISED)   // trace.h:395-402
					{
						py = __fsub_rn(py, 1.0f);
						wy = (gy > 0) ? __fadd_rn(wy, iay) : __fsub_rn(wy, iay);
					}
					if(py < 0.0f || py > 1.0f)   // trace.h:404-440
					{
						uint32_t pk;
						if(ncls == CL_PORTAL) pk = (__ldg(&reinterpret_cast<const uint32_t *>(L.portals)[(cell >> 4)*4 + 2]) >> 16) & 255u;
						else pk = (ncls == CL_TWO || ncls == CL_TWO_FOG) ? PK_PASS : (ncls == CL_RAISED ? PK_RAISED : PK_WALL);
						if(pk != PK_PASS)
						{
							if(pk == PK_RAISED)   // revert (trace.h:422-431)
							{
								py = __fadd_rn(py, 1.0f);
								wy = (gy > 0) ? __fsub_rn(wy, iay) : __fadd_rn(wy, iay);
							}
							PWN_WALL_EXIT(SF_WALL)
						}
					}
				}
			}
			else if(cls >= CL_RAMP_GT && cls <= CL_RAMP_CARET)
			{
				// trace.h:443-505
				if(IDS) ops += W_RAMP;
				const float tilt = __fmul_rn(0.5f, (cls <= CL_RAMP_LT) ? rx : rz);
				const bool minus = (cls == CL_RAMP_GT) || (cls == CL_RAMP_COMMA);
				ry = minus ? __fsub_rn(ry, tilt) : __fadd_rn(ry, tilt);
				wy = py;
				if(ry >= 0.0f) wy = __fsub_rn(1.0f, wy);
				wy = __fmul_rn(wy, __fdiv_rn(1.0f, (ry < 0.0f ? -ry : ry)));
				PWN_SPHERE_EXIT(false)
				PWN_THROUGH(gy, gy, gz);   // sic: gy stands in for gx (trace.h:470)
				if(ldir >= PWN_FYP)
				{
					ldir = (ry < 0.0f ? PWN_FYN : PWN_FYP);
					PWN_WALL_EXIT(ry >= 0.0f ? SF_CEIL : SF_FLOOR)
				}
				else if(ldir == PWN_FXN || ldir == PWN_FXP)
				{
					ldir = (rx < 0.0f ? PWN_FXN : PWN_FXP);
					wy = __fsub_rn(wy, wx); wz = __fsub_rn(wz, wx); wx = iax; cx += gx;
				}
				else
				{
					ldir = (rz < 0.0f ? PWN_FZN : PWN_FZP);
					wy = __fsub_rn(wy, wz); wx = __fsub_rn(wx, wz); wz = iaz; cz += gz;
				}
				// un-tilt: the product is recomputed from the (unchanged) rx/rz, as the source does
				ry = minus ? __fadd_rn(ry, tilt) : __fsub_rn(ry, tilt);
				wy = py;
				if(ry >= 0.0f) wy = __fsub_rn(1.0f, wy);
				wy = __fmul_rn(wy, iay);
				cell = get_cell(L, cx, cz);
			}
			else if(cls == CL_PORTAL)
			{
				// trace.h:508-650
				const uint4 pr = __ldg(reinterpret_cast<const uint4 *>(&L.portals[cell >> 4]));
				const uint32_t pkind = (pr.z >> 8) & 255u;
				if(pkind != PR_OK)
				{
					PWN_SPHERE_EXIT(false)
					PWN_WALL_EXIT(pkind == PR_UNPAIRED ? SF_WALL : SF_STRAY)
				}
				cx = (int)(int16_t)(pr.x & 0xFFFFu); cz = (int)(int16_t)(pr.x >> 16);
				px = __fadd_rn(px, (float)(int)(int16_t)(pr.y & 0xFFFFu));
				pz = __fadd_rn(pz, (float)(int)(int16_t)(pr.y >> 16));
				const int rot = (int)(pr.z & 255u);
				ldir = (ldir - rot) & 3;
				nport++;
				if(IDS) ops += W_PORTAL;
				const float ccx = __fadd_rn((float)cx, 0.5f), ccz = __fadd_rn((float)cz, 0.5f);
				if(rot == 1)
				{
					const float trx = px, trz = pz, trvx = rx;
					px = __fadd_rn(ccx, __fsub_rn(trz, ccz));
					pz = __fsub_rn(ccz, __fsub_rn(trx, ccx));
					rx = rz; rz = -trvx;
					const int tg = gx; gx = gz; gz = -tg;
					float t = wx; wx = wz; wz = t;
					t = iax; iax = iaz; iaz = t;
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
					rx = -rz; rz = trvx;
					const int tg = gx; gx = -gz; gz = tg;
					float t = wx; wx = wz; wz = t;
					t = iax; iax = iaz; iaz = t;
				}
				// step out of the destination portal cell (trace.h:624-647)
				if(ldir == PWN_FXP) { cx++; px = __fadd_rn(px, 1.0f); }
				else if(ldir == PWN_FZP) { cz++; pz = __fadd_rn(pz, 1.0f); }
				else if(ldir == PWN_FXN) { cx--; px = __fsub_rn(px, 1.0f); }
				else { cz--; pz = __fsub_rn(pz, 1.0f); }
				cell = get_cell(L, cx, cz);
			}
			else
			{
				// solid (trace.h:651-664)
				PWN_SPHERE_EXIT(false)
				PWN_WALL_EXIT(ldir == PWN_FYP ? SF_CEIL : SF_WALL)
			}

			PWN_SPHERE_EXIT(false)   // trace.h:668-673
		}
#undef PWN_THROUGH
#undef PWN_SPHERE_EXIT
#undef PWN_WALL_EXIT
#undef PWN_STEP_XZ

		if(IDS && P.ids != nullptr)
		{
			pwn_ids rec;
			rec.cell = cz*L.w + cx; rec.ldir = (int16_t)hit.ldir; rec.kind = (int16_t)hit.kind;
			rec.nstep = nstep; rec.nportal = nport;
			P.ids[pix*3 + seg] = rec;
		}

		// zbuf keeps the PRIMARY distance: bounces save/zero/restore *dist (trace.h:86-87,103);
		// an exhausted walk leaves it untouched (trace.h:677-678)
		if(seg == 0 && hit.kind != 2) P.zbuf[pix] = hit.dist;

		if(hit.kind == 2)
		{
			ob = rx; og = ry; orr = rz; oa = 0.0f;   // the ray comes back as the colour
			break;
		}

		// ================= shading =================
		float cb, cg, cr, ca, refl;
		if(hit.kind == 0)
		{
			// trace_hit_wall, trace.h:108-154
			if(IDS) ops += W_WALL;
			float sb, sg, sr;
			if(hit.surf == SF_CEIL) { sb = 30.0f; sg = 30.0f; sr = 0.0f; }
			else if(hit.surf == SF_FLOOR) { sb = 1.0f; sg = 1.0f; sr = 1.0f; }
			else if(hit.surf == SF_WALL) { sb = 0.8f; sg = 0.8f; sr = 1.0f; }
			else { sb = 5.0f; sg = 0.0f; sr = 5.0f; }
			cb = __fmul_rn(icb, sb); cg = __fmul_rn(icg, sg); cr = __fmul_rn(icr, sr); ca = __fmul_rn(ica, 0.0f);
			float diffuse;
			switch(hit.ldir)
			{
				default:
				case PWN_FXP: diffuse = rx; break;
				case PWN_FYP: diffuse = ry; break;
				case PWN_FZP: diffuse = rz; break;
				case PWN_FXN: diffuse = -rx; break;
				case PWN_FYN: diffuse = -ry; break;
				case PWN_FZN: diffuse = -rz; break;
			}
			if(diffuse < 0.0f) diffuse = 0.0f;
			diffuse = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, 0.1f), diffuse), 0.1f);
			cb = __fmul_rn(cb, diffuse); cg = __fmul_rn(cg, diffuse); cr = __fmul_rn(cr, diffuse); ca = __fmul_rn(ca, diffuse);
			refl = (hit.ldir == PWN_FYN ? 0.7f : 0.25f);
		}
		else
		{
			cb = aux.cb; cg = aux.cg; cr = aux.cr; ca = aux.ca;
			refl = aux.refl;
		}

		// ================= trace_hit_bounce, trace.h:3-106 =================
		if(seg >= 2 || refl == 0.0f)   // REFLECT == 2 (defs.h:8)
		{
			ob = cb; og = cg; orr = cr; oa = ca;
			break;
		}

		if(IDS) ops += W_BOUNCE + (hit.ldir == PWN_FYN ? W_FLOOR : 0u) + (fog != 0.0f ? W_FOG : 0u);
		float nrx = rx, nry = ry, nrz = rz;
		float npx, npy, npz;
		if(hit.kind == 1) { npx = aux.px; npy = aux.py; npz = aux.pz; }
		else { npx = px; npy = py; npz = pz; }

		if(hit.ldir == PWN_FXP) { nrx = -nrx; npx = __fsub_rn(npx, 0.001f); }
		else if(hit.ldir == PWN_FXN) { nrx = -nrx; npx = __fadd_rn(npx, 0.001f); }
		else if(hit.ldir == PWN_FZP) { nrz = -nrz; npz = __fsub_rn(npz, 0.001f); }
		else if(hit.ldir == PWN_FZN) { nrz = -nrz; npz = __fadd_rn(npz, 0.001f); }
		else if(hit.ldir == PWN_FYP) { nry = -nry; npy = __fsub_rn(npy, 0.001f); }
		else
		{
			float nx, ny, nz;
			if(hit.ldir == PWN_FYN)
			{
				// wavy floor normal (trace.h:39-46)
				npy = __fsub_rn(npy, 0.001f);
				const float HPI = __uint_as_float(0x3FC90FDBu);   // (float)M_PI*0.5f
				const float TPI = __uint_as_float(0x40C90FDBu);   // (float)M_PI*2.0f
				const float a0 = libm_sinf(__fmul_rn(HPI, npx));
				const float a1 = libm_cosf(__fmul_rn(HPI, npz));
				const float ang = __fmul_rn(TPI, __fadd_rn(__fadd_rn(a0, a1), P.sec_current));
				float s, c;
				libm_sincosf(ang, &s, &c);
				// v_normalise(sin, 38, cos, 0): (x^2 + z^2) + (y^2 + 0)
				const float nd = __fadd_rn(__fadd_rn(__fmul_rn(s, s), __fmul_rn(c, c)), __fmul_rn(38.0f, 38.0f));
				const float ns = rsqrt_emul(L, nd);
				nx = __fmul_rn(s, ns); ny = __fmul_rn(38.0f, ns); nz = __fmul_rn(c, ns);
			}
			else
			{
				// sphere (trace.h:58-64)
				npx = __fsub_rn(npx, __fmul_rn(0.001f, nrx));
				npy = __fsub_rn(npy, __fmul_rn(0.001f, nry));
				npz = __fsub_rn(npz, __fmul_rn(0.001f, nrz));
				nx = aux.nx; ny = aux.ny; nz = aux.nz;
			}
			// rmul = -2*(((0 + rx*nx) + ry*ny) + rz*nz); ray = normalise(rmul*norm + ray)
			const float rmul = __fmul_rn(-2.0f, __fadd_rn(__fadd_rn(__fmul_rn(nrx, nx), __fmul_rn(nry, ny)), __fmul_rn(nrz, nz)));
			nrx = __fadd_rn(__fmul_rn(rmul, nx), nrx);
			nry = __fadd_rn(__fmul_rn(rmul, ny), nry);
			nrz = __fadd_rn(__fmul_rn(rmul, nz), nrz);
			const float d = __fadd_rn(__fadd_rn(__fmul_rn(nrx, nrx), __fmul_rn(nrz, nrz)), __fmul_rn(nry, nry));
			const float s = rsqrt_emul(L, d);
			nrx = __fmul_rn(nrx, s); nry = __fmul_rn(nry, s); nrz = __fmul_rn(nrz, s);
		}

		if(P.reflect_blur)   // trace.h:77-84: draws x, y, (discard), z, (discard)
		{
			nrx = __fadd_rn(nrx, __fmul_rn(randfs(seed), 0.03f));
			nry = __fadd_rn(nry, __fmul_rn(randfs(seed), 0.03f));
			randfs(seed);
			nrz = __fadd_rn(nrz, __fmul_rn(randfs(seed), 0.03f));
			randfs(seed);
		}

		if(seg == 0) { b0b = cb; b0g = cg; b0r = cr; b0a = ca; refl0 = refl; fog0 = fog; }
		else { b1b = cb; b1g = cg; b1r = cr; b1a = ca; refl1 = refl; fog1 = fog; }
		depth = seg + 1;

		// child: trace_ray(hitctr+1, seed, lv, dist, &pos, &ray, col)  (trace.h:90)
		fx = npx; fy = npy; fz = npz;
		irx = nrx; iry = nry; irz = nrz;
		icb = cb; icg = cg; icr = cr; ica = ca;
	}

	// ---- unwind the blends, deepest first (trace.h:91-101) ------------------------------
	if(depth >= 2)
	{
		const float k = __fsub_rn(1.0f, refl1);
		ob = __fadd_rn(__fmul_rn(refl1, ob), __fmul_rn(k, b1b));
		og = __fadd_rn(__fmul_rn(refl1, og), __fmul_rn(k, b1g));
		orr = __fadd_rn(__fmul_rn(refl1, orr), __fmul_rn(k, b1r));
		oa = __fadd_rn(__fmul_rn(refl1, oa), __fmul_rn(k, b1a));
		if(fog1 != 0.0f)
		{
			const float f = libm_expf(__fmul_rn(-0.6f, fog1)), g = __fsub_rn(1.0f, f);
			ob = __fadd_rn(__fmul_rn(f, ob), g); og = __fadd_rn(__fmul_rn(f, og), g);
			orr = __fadd_rn(__fmul_rn(f, orr), g); oa = __fadd_rn(__fmul_rn(f, oa), g);
		}
	}
	if(depth >= 1)
	{
		const float k = __fsub_rn(1.0f, refl0);
		ob = __fadd_rn(__fmul_rn(refl0, ob), __fmul_rn(k, b0b));
		og = __fadd_rn(__fmul_rn(refl0, og), __fmul_rn(k, b0g));
		orr = __fadd_rn(__fmul_rn(refl0, orr), __fmul_rn(k, b0r));
		oa = __fadd_rn(__fmul_rn(refl0, oa), __fmul_rn(k, b0a));
		if(fog0 != 0.0f)
		{
			const float f = libm_expf(__fmul_rn(-0.6f, fog0)), g = __fsub_rn(1.0f, f);
			ob = __fadd_rn(__fmul_rn(f, ob), g); og = __fadd_rn(__fmul_rn(f, og), g);
			orr = __fadd_rn(__fmul_rn(f, orr), g); oa = __fadd_rn(__fmul_rn(f, oa), g);
		}
	}

	if(IDS && P.opcount != nullptr)
	{
		unsigned long long tot = ops;
		const unsigned m = __activemask();
		for(int o = 16; o > 0; o >>= 1)
		{
			const unsigned long long v = __shfl_xor_sync(m, tot, o);
			if(m & (1u << ((threadIdx.x & 31) ^ o))) tot += v;
		}
		if((threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(P.opcount, tot);
	}

	// ---- col_ftoint (util.h:48-59): B | G<<8 | R<<16 | A<<24 ------------------------------
	P.sbuf[pix] = pack_lane(ob) | (pack_lane(og) << 8) | (pack_lane(orr) << 16) | (pack_lane(oa) << 24);
}

} // namespace pwn
