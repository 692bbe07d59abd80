// blur_kernel.cuh -- K2: the POSTPROC_BLUR focal blur (screen.h:69-123) and the frame checksum.
//
// The reference walks each row left to right with ONE sequential LCG stream
// (seed = cy*cy + 415135, 32 draws per 4-pixel group, screen.h:82,95-109).  The LCG
// is affine mod 2^31, so a thread reaches its own draws through a per-frame table of
// the composed map for each 4-pixel group (built by pwn_frame_config); one thread =
// one output pixel.  In pose batches the blur of frame i-1 rides in the trace launch
// of frame i (k_trace<..., BLURP>): its scattered gathers hide under the trace's ALU work.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pwn {

struct BlurParams
{
	const uint32_t *src;   // pre-blur frame (the reference's tsbuf copy, screen.h:75)
	const float *zbuf;
	uint32_t *dst;         // sbuf
	const uint2 *jump;     // per 4-pixel group g: the affine map (A, C) that advances a row's stream by 32*g draws
	int dimx, dimy, rwidth;
	int y0, y1;            // rows to produce
	int nshards, shard;    // rows y0 + the 32-row bands b with b % nshards == shard (1, 0: all rows)
	float fstr;            // 0.002f * dimy (screen.h:86), computed on the host in float
	// pwn_trace_screen blurs a frame in row chunks while later chunks are still being traced: rows [0, ylimit) of src
	// are final.  A pixel with a sample below them is not written but queued (defer_list[atomicAdd(defer_count)] =
	// pixel index) for k_blur_fix, which runs when the whole frame has been traced.
	int ylo, ylimit;       // rows [ylo, ylimit) of src are final
	uint32_t *defer_count, *defer_list;
	// k_blur_fix on a frame sharded in contiguous blocks: row y of the pre-blur frame is read from the shard that traced it
	const uint32_t *src_sh[8];
	int band_start[9];     // shard s owns the bands [band_start[s], band_start[s+1])
	int nsh;               // 0: src holds every row
	uint32_t zero;         // always 0; opaque to the compiler (the -0 of the packed products)
};

// s -> A*s + C (mod 2^31) composed n times for n = 0, 2, 4, 6 (a pixel's offset inside its group)
// and for n = 6 (from one sample of a pixel to its next)
__constant__ uint32_t c_jumpA[8];
__constant__ uint32_t c_jumpC[8];

// ---- packed FP32 (FADD2 / FFMA2): the x and y offsets of a sample go through the same operations ----
// (see trace_kernel.cuh for why a product that feeds an add is written fma(a, b, -0) with an opaque -0)
#ifndef PWN_PACKED
#define PWN_PACKED 1
#endif
typedef unsigned long long bf32x2;
__device__ __forceinline__ bf32x2 bpk2(float lo, float hi) { bf32x2 u; asm("mov.b64 %0, {%1,%2};" : "=l"(u) : "f"(lo), "f"(hi)); return u; }
__device__ __forceinline__ void bunpk2(bf32x2 u, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(u)); }
__device__ __forceinline__ bf32x2 badd2(bf32x2 a, bf32x2 b) { bf32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ bf32x2 bsub2(bf32x2 a, bf32x2 b) { bf32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ bf32x2 bmul2x(bf32x2 a, bf32x2 b, bf32x2 nz) { bf32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(nz)); return d; }

// one LCG draw reduced mod 3759, as a float (util.h:1-12 up to the division)
__device__ __forceinline__ float blur_draw(uint32_t &seed)
{
	seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
	return (float)(seed % 3759u);
}

// the clamped sample coordinates (sx, sy) of one sample: base + randfs*fstr*z per lane, randfs = r*fl(2/3759) - 1
__device__ __forceinline__ void blur_sample_xy(uint32_t &seed, bf32x2 base, float fstr, float z, bf32x2 nz, int xmax, int ymax, int &sx, int &sy);

__device__ __forceinline__ float blur_randfs(uint32_t &seed)
{
	seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
	uint32_t r = seed % 3759u;
	return __fsub_rn(__fmul_rn((float)r, __uint_as_float(0x3A0B79B3u)), 1.0f);
}

// (int)float as cvttss2si does it, FOLLOWED by the clamp to [0, hi] of screen.h:103-106.
// cvttss2si gives INT_MIN for NaN and for anything outside int range, which the clamp turns into 0.
// The unsigned conversion F2I.U32.TRUNC gives 0 for NaN and negatives, the value below 2^32 and
// 0xFFFFFFFF above: read as a signed number, everything from 2^31 up is negative, so max(., 0) sends
// exactly the values cvttss2si cannot represent to 0 as well -- three instructions, no compare.
__device__ __forceinline__ int cvtt_clamp(float v, int hi)
{
	const int i = (int)__float2uint_rz(v);
	return min(max(i, 0), hi);
}

__device__ __forceinline__ void blur_sample_xy(uint32_t &seed, bf32x2 base, float fstr, float z, bf32x2 nz, int xmax, int ymax, int &sx, int &sy)
{
#if PWN_PACKED
	const float r0 = blur_draw(seed), r1 = blur_draw(seed);
	const float c = __uint_as_float(0x3A0B79B3u);
	bf32x2 v = bsub2(bmul2x(bpk2(r0, r1), bpk2(c, c), nz), bpk2(1.0f, 1.0f));          // randfs, both lanes
	v = bmul2x(bmul2x(v, bpk2(fstr, fstr), nz), bpk2(z, z), nz);                         // (randfs*fstr)*z
	float fx, fy;
	bunpk2(badd2(base, v), fx, fy);
	sx = cvtt_clamp(fx, xmax);
	sy = cvtt_clamp(fy, ymax);
#else
	float fx, fy;
	bunpk2(base, fx, fy);
	sx = cvtt_clamp(__fadd_rn(fx, __fmul_rn(__fmul_rn(blur_randfs(seed), fstr), z)), xmax);
	sy = cvtt_clamp(__fadd_rn(fy, __fmul_rn(__fmul_rn(blur_randfs(seed), fstr), z)), ymax);
#endif
}

// pavgb on four bytes at once: (a + b + 1) >> 1 per byte = (a | b) - (((a ^ b) >> 1) & 0x7f7f7f7f)
__device__ __forceinline__ uint32_t pavgb4(uint32_t a, uint32_t b)
{
	return (a | b) - (((a ^ b) >> 1) & 0x7F7F7F7Fu);
}

// one output pixel of the blur pass (screen.h:91-117)
template <bool DEFER, bool PEER = false>
__device__ __forceinline__ void blur_pixel(const BlurParams &P, int x, int y)
{
	const uint32_t pix = (uint32_t)y*(uint32_t)P.rwidth + (uint32_t)x;   // (a frame has fewer than 2^31 pixels: pwn_frame_config)
	const int g = x >> 2, j = x & 3;
	if((x & ~3) >= P.dimx - 3)
	{
		// the last dimx%4 pixels are not blurred (screen.h:88-91)
		P.dst[pix] = P.src[pix];
		return;
	}
	// the row's stream (seed = cy*cy + 415135, screen.h:82) advanced to draw 32*g + 2*j
	const uint2 jg = __ldg(&P.jump[g]);
	uint32_t seed = jg.x*((uint32_t)y*(uint32_t)y + 415135u) + jg.y;
	// (the jumps are left unmasked: bit 31 of a seed reaches only bit 31 of the next product, and every draw
	// masks before it takes the remainder)
	seed = c_jumpA[j]*seed + c_jumpC[j];

	const float z = __fsub_rn(P.zbuf[pix], 1.0f);
	const bf32x2 base = bpk2((float)x, (float)y);
	const float negz = __uint_as_float(0x80000000u ^ (P.zero & (uint32_t)x));   // -0.0f ptxas cannot fold
	const bf32x2 nz = bpk2(negz, negz);
	// six draws separate this pixel's pair in sample i from its pair in sample i+1
	const uint32_t a6 = c_jumpA[3], c6 = c_jumpC[3];
	uint32_t s[4];
	if(DEFER)
	{
		uint32_t at[4];
		int ymax = 0, ymin = 0x7FFFFFFF;
		#pragma unroll
		for(int i = 0; i < 4; i++)
		{
			int sx, sy;
			blur_sample_xy(seed, base, P.fstr, z, nz, P.dimx - 1, P.dimy - 1, sx, sy);
			at[i] = (uint32_t)sy*(uint32_t)P.rwidth + (uint32_t)sx;
			ymax = max(ymax, sy);
			ymin = min(ymin, sy);
			seed = a6*seed + c6;
		}
		if(ymax >= P.ylimit || ymin < P.ylo)   // a sample in a row that is not there (yet): k_blur_fix finishes this pixel
		{
			P.defer_list[atomicAdd(P.defer_count, 1u)] = pix;
			return;
		}
		// (plain loads, not the non-coherent path: other rows of src are being written by a trace that runs beside this
		// kernel.  The rows read here were released by their writers before the stream-ordered wait that precedes this
		// launch acquired their counters, so they happen-before every load of this kernel.)
		#pragma unroll
		for(int i = 0; i < 4; i++) s[i] = P.src[at[i]];
	}
	else if(PEER)
	{
		#pragma unroll
		for(int i = 0; i < 4; i++)
		{
			int sx, sy;
			blur_sample_xy(seed, base, P.fstr, z, nz, P.dimx - 1, P.dimy - 1, sx, sy);
			const uint32_t *src = P.src;
			if(P.nsh > 0)
			{
				const int sb = sy >> 5;
				int o = 0;
				while(o + 1 < P.nsh && sb >= P.band_start[o + 1]) o++;
				src = P.src_sh[o];
			}
			s[i] = src[(uint32_t)sy*(uint32_t)P.rwidth + (uint32_t)sx];
			seed = a6*seed + c6;
		}
	}
	else
	{
	#pragma unroll
	for(int i = 0; i < 4; i++)
	{
		int sx, sy;
		blur_sample_xy(seed, base, P.fstr, z, nz, P.dimx - 1, P.dimy - 1, sx, sy);
		s[i] = __ldg(&P.src[(uint32_t)sy*(uint32_t)P.rwidth + (uint32_t)sx]);
		seed = a6*seed + c6;
	}
	}
	// pavgb: per byte (a+b+1)>>1 (screen.h:112-115)
	P.dst[pix] = pavgb4(pavgb4(s[0], s[1]), pavgb4(s[2], s[3]));
}

template <bool DEFER>
__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ BlurParams P)
{
	const int x = blockIdx.x*blockDim.x + threadIdx.x;
	const int lrow = blockIdx.y;
	const int y = P.y0 + ((lrow >> 5)*P.nshards + P.shard)*32 + (lrow & 31);
	if(x >= P.dimx || y >= P.y1) return;
	blur_pixel<DEFER>(P, x, y);
}

// the pixels k_blur<true> queued, once the whole pre-blur frame is there
__global__ void __launch_bounds__(256) k_blur_fix(const __grid_constant__ BlurParams P)
{
	const uint32_t n = *P.defer_count;
	for(uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x)
	{
		const uint32_t pix = P.defer_list[i];
		const uint32_t y = pix / (uint32_t)P.rwidth;
		blur_pixel<false, true>(P, (int)(pix - y*(uint32_t)P.rwidth), (int)y);
	}
}

// stream-ordered wait for band counters (contiguous-block sharding): lane t polls band b0 + t until it holds
// `target` lanes or more; a band that never fills ends the wait after timeout_ns with the error word set
struct WaitParams
{
	const uint32_t *done;
	int b0, b1;
	uint32_t target;
	unsigned long long timeout_ns;
	uint32_t *err;
};

__global__ void k_wait_bands(const WaitParams W)
{
	unsigned long long t0, t1;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	for(int b = W.b0 + (int)threadIdx.x; b < W.b1; b += (int)blockDim.x)
	{
		for(;;)
		{
			uint32_t v;
			asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(W.done + b) : "memory");
			if((int32_t)(v - W.target) >= 0) break;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			if(t1 - t0 > W.timeout_ns) { *W.err = 1u; break; }
			__nanosleep(100);
		}
	}
}

// ---- barrier between the GPUs of a sharded frame, on the device ---------------------------------
// Lane t signals shard t (stores this launch's epoch into its own slot of t's flag array, over
// NVLink) and then waits for t's signal in the local array.  Launched in stream order after the
// kernel whose stores it publishes; a peer that never arrives ends the wait after `timeout_ns`
// with the error word set instead of hanging the GPU.
struct BarrierParams
{
	uint32_t *peer_flags[8];   // flag array of every shard (own entry = local)
	uint32_t *local;           // [0..7] barrier A, [8..15] barrier B, [16] error
	int nshards, shard, slot;
	uint32_t epoch;
	unsigned long long timeout_ns;
};

__global__ void k_peer_barrier(const BarrierParams B)
{
	const int t = threadIdx.x;
	if(t >= B.nshards || t == B.shard) return;
	__threadfence_system();
	uint32_t *remote = B.peer_flags[t] + B.slot*8 + B.shard;
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(remote), "r"(B.epoch) : "memory");
	const uint32_t *mine = B.local + B.slot*8 + t;
	unsigned long long t0, t1;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	for(;;)
	{
		uint32_t v;
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
		if((int32_t)(v - B.epoch) >= 0) break;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if(t1 - t0 > B.timeout_ns) { B.local[16] = 1u; break; }
		__nanosleep(200);
	}
}

// ---- order-independent frame checksum: sum_i mix64(i<<32 | word_i) mod 2^64 ----------
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_checksum(const uint32_t *buf, int dimx, int dimy, int rwidth, unsigned long long *out)
{
	uint64_t acc = 0;
	const size_t n = (size_t)dimx*dimy;
	for(size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x*blockDim.x)
	{
		const size_t yy = i / dimx, xx = i - yy*dimx;
		acc += mix64(((uint64_t)i << 32) | buf[yy*rwidth + xx]);
	}
	#pragma unroll
	for(int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
	if((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)acc);
}

// ---- measured FP32 issue peak: 8 independent chains of non-fused FMUL/FADD per thread ----------
// (the trace may not use FMA, so the relevant ceiling is one FP32 lane-op per lane per clock)
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int iters, float a, float b)
{
	float v0 = threadIdx.x, v1 = v0 + 1.0f, v2 = v0 + 2.0f, v3 = v0 + 3.0f, v4 = v0 + 4.0f, v5 = v0 + 5.0f, v6 = v0 + 6.0f, v7 = v0 + 7.0f;
	#pragma unroll 1
	for(int i = 0; i < iters; i++)
	{
		#pragma unroll
		for(int k = 0; k < 8; k++)
		{
			v0 = __fadd_rn(__fmul_rn(v0, a), b); v1 = __fadd_rn(__fmul_rn(v1, a), b);
			v2 = __fadd_rn(__fmul_rn(v2, a), b); v3 = __fadd_rn(__fmul_rn(v3, a), b);
			v4 = __fadd_rn(__fmul_rn(v4, a), b); v5 = __fadd_rn(__fmul_rn(v5, a), b);
			v6 = __fadd_rn(__fmul_rn(v6, a), b); v7 = __fadd_rn(__fmul_rn(v7, a), b);
		}
	}
	out[blockIdx.x*blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

} // namespace pwn
