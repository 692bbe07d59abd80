// upscale_kernel.cuh -- K3: screen_upscale (screen.h:126-149), the display tail of the frame.
//
// Every frame pixel becomes an rscale x rscale block of the display surface.  The reference
// walks ONE destination pointer: per source pixel +rscale, per source row +pitch*(rscale-1)
// on top of the w*rscale already walked.  Surface word of (source row py, source pixel px,
// sub-row y, sub-column x):
//
//     py*(w*rscale + pitch*(rscale-1)) + px*rscale + y*pitch + x            (words)
//
// which is the usual py*rscale*pitch + ... only when pitch == w*rscale; with a padded pitch
// the rows drift.  Both kernels below write exactly those words.
//
// This kernel is HBM-bound (4 B read, 4*rscale^2 B written per frame pixel, nothing computed):
// the packed variant stores 16 bytes per thread per sub-row, fully coalesced, and reads each
// frame pixel once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pwn {

struct UpscaleParams
{
	const uint32_t *src;   // sbuf
	uint32_t *dst;         // surface
	int w, h, rwidth;      // frame size and stride (pixels)
	int rscale;
	int pitch;             // surface stride in 32-bit words (screen->pitch / 4)
	long long rowstep;     // w*rscale + pitch*(rscale-1)
};

// pitch == w*rscale and (w*rscale) % 4 == 0: thread = four consecutive surface pixels of a source
// row; it reads their (at most four, usually one or two) source pixels once and stores rscale uint4s
template <int RS>
__global__ void __launch_bounds__(256) k_upscale_packed(const UpscaleParams P)
{
	const int rs = RS > 0 ? RS : P.rscale;
	const int q = blockIdx.x*blockDim.x + threadIdx.x;   // uint4 column of the surface
	const int py = blockIdx.y;
	if(q*4 >= P.pitch) return;
	const uint32_t *s = P.src + (size_t)py*P.rwidth;
	uint4 v;
	if(RS == 1)
		v = __ldg(reinterpret_cast<const uint4 *>(s) + q);
	else if(RS == 2)
	{
		const uint2 a = __ldg(reinterpret_cast<const uint2 *>(s) + q);
		v = make_uint4(a.x, a.x, a.y, a.y);
	}
	else if(RS == 4)
	{
		const uint32_t a = __ldg(s + q);
		v = make_uint4(a, a, a, a);
	}
	else
	{
		const int o = q*4;
		v.x = __ldg(s + (o    )/rs);
		v.y = __ldg(s + (o + 1)/rs);
		v.z = __ldg(s + (o + 2)/rs);
		v.w = __ldg(s + (o + 3)/rs);
	}
	uint4 *d = reinterpret_cast<uint4 *>(P.dst + (size_t)py*rs*P.pitch) + q;
	const size_t p4 = (size_t)(P.pitch >> 2);
	#pragma unroll
	for(int y = 0; y < rs; y++)
		__stcs(d + (size_t)y*p4, v);   // streaming: the surface is written once and read by the copy engine
}

// any pitch >= w*rscale: thread = one surface word of the block row (x along the surface row)
__global__ void __launch_bounds__(256) k_upscale_any(const UpscaleParams P)
{
	const int ox = blockIdx.x*blockDim.x + threadIdx.x;   // 0 .. w*rscale-1
	const int py = blockIdx.y;
	if(ox >= P.w*P.rscale) return;
	const uint32_t pd = __ldg(P.src + (size_t)py*P.rwidth + ox/P.rscale);
	uint32_t *d = P.dst + (long long)py*P.rowstep + ox;
	for(int y = 0; y < P.rscale; y++)
		d[(size_t)y*P.pitch] = pd;
}

} // namespace pwn
