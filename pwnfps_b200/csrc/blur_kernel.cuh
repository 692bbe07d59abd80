// blur_kernel.cuh -- K2: the POSTPROC_BLUR focal blur (screen.h:69-123) and the frame checksum.
//
// The reference walks each row left to right with ONE sequential LCG stream
// (seed = cy*cy + 415135, 32 draws per 4-pixel group, screen.h:82,95-109).  The LCG
// is affine mod 2^31, so a thread reaches its own draws by composing precomputed
// power-of-two jumps; one thread = one output pixel, gathers go through L2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pwn {

struct BlurParams
{
	const uint32_t *src;   // pre-blur frame (the reference's tsbuf copy, screen.h:75)
	const float *zbuf;
	uint32_t *dst;         // sbuf
	int dimx, dimy, rwidth;
	int y0, y1;            // rows to produce
	float fstr;            // 0.002f * dimy (screen.h:86), computed on the host in float
};

// s -> A*s + C (mod 2^31) composed 2^k times: jumpA[k], jumpC[k]
__constant__ uint32_t c_jumpA[32];
__constant__ uint32_t c_jumpC[32];

__device__ __forceinline__ float blur_randfs(uint32_t &seed)
{
	seed = (seed * 25739u + 4u) & 0x7FFFFFFFu;
	uint32_t r = seed % 3759u;
	return __fsub_rn(__fmul_rn((float)r, __uint_as_float(0x3A0B79B3u)), 1.0f);
}

// (int)float as cvttss2si does it: out of range and NaN give INT_MIN
__device__ __forceinline__ int cvtt(float v)
{
	if(!(v < 2147483648.0f && v >= -2147483648.0f)) return (int)0x80000000;
	return __float2int_rz(v);
}

__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ BlurParams P)
{
	const int x = blockIdx.x*blockDim.x + threadIdx.x;
	const int y = P.y0 + blockIdx.y;
	if(x >= P.dimx || y >= P.y1) return;
	const size_t pix = (size_t)y*P.rwidth + x;
	const int cx = x & ~3, j = x & 3;
	if(cx >= P.dimx - 3)
	{
		// the last dimx%4 pixels are not blurred (screen.h:88-91)
		P.dst[pix] = P.src[pix];
		return;
	}

	// jump the row's stream to draw number 32*(cx/4) + 2*j
	uint32_t n = 8u*(uint32_t)cx + 2u*(uint32_t)j;
	uint32_t a = 1u, c = 0u;
	for(int k = 0; n != 0u; k++, n >>= 1)
		if(n & 1u) { a = c_jumpA[k]*a; c = c_jumpA[k]*c + c_jumpC[k]; }
	uint32_t seed = (a*((uint32_t)y*(uint32_t)y + 415135u) + c) & 0x7FFFFFFFu;

	const float z = __fsub_rn(P.zbuf[pix], 1.0f);
	const float fbx = (float)x, fby = (float)y;
	// six draws separate this pixel's pair in sample i from its pair in sample i+1
	const uint32_t a6 = c_jumpA[1]*c_jumpA[2], c6 = c_jumpA[2]*c_jumpC[1] + c_jumpC[2];
	uint32_t s[4];
	#pragma unroll
	for(int i = 0; i < 4; i++)
	{
		int sx = cvtt(__fadd_rn(fbx, __fmul_rn(__fmul_rn(blur_randfs(seed), P.fstr), z)));
		int sy = cvtt(__fadd_rn(fby, __fmul_rn(__fmul_rn(blur_randfs(seed), P.fstr), z)));
		if(sx < 0) sx = 0;
		if(sy < 0) sy = 0;
		if(sx >= P.dimx) sx = P.dimx - 1;
		if(sy >= P.dimy) sy = P.dimy - 1;
		s[i] = __ldg(&P.src[(size_t)sy*P.rwidth + sx]);
		seed = (a6*seed + c6) & 0x7FFFFFFFu;
	}
	// pavgb: per byte (a+b+1)>>1 (screen.h:112-115)
	P.dst[pix] = __vavgu4(__vavgu4(s[0], s[1]), __vavgu4(s[2], s[3]));
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

} // namespace pwn
