// pwn_libm.cuh -- device side of pwn_libm.h: the reference host's glibc sinf / cosf / sincosf /
// expf, bit for bit, in FP64 on the B200's DFMA pipe (which the rest of K1 leaves idle).
// LFMA selects glibc's "_fma" build (one rounding per a*b+c) or its "_sse2" build; the host
// probe (pwn_host_libm_probe) decides which one the reference would be running.
#pragma once
#include <cuda_runtime.h>

#include "pwn_libm.h"

namespace pwn {

// __exp2f_data.tab: 256 bytes, read with a per-lane index -> global memory through L1
__device__ uint64_t d_exp2f_tab[32] = PWN_EXP2F_TAB;

template <bool LFMA>
__device__ __forceinline__ float libm_sinf(float x) { return pwn_sinf(x, LFMA ? 1 : 0); }
template <bool LFMA>
__device__ __forceinline__ float libm_cosf(float x) { return pwn_cosf(x, LFMA ? 1 : 0); }
template <bool LFMA>
__device__ __forceinline__ void libm_sincosf(float x, float *s, float *c) { pwn_sincosf(x, s, c, LFMA ? 1 : 0); }
template <bool LFMA>
__device__ __forceinline__ float libm_expf(float x) { return pwn_expf(x, d_exp2f_tab, LFMA ? 1 : 0); }

// test hook: evaluates the four functions over an array (pwn_debug_libm)
template <bool LFMA>
__global__ void k_libm_eval(const float *in, float *s, float *c, float *e, int n)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	float a, b;
	libm_sincosf<LFMA>(in[i], &a, &b);
	s[i] = a;
	c[i] = b;
	e[i] = libm_expf<LFMA>(in[i]);
}

} // namespace pwn
