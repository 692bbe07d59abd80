// pwn_libm.cuh -- the libm calls of the bounce shading: sinf/cosf of the wavy floor
// normal (trace.h:42-46) and expf of the fog blend (trace.h:97).  The reference
// gets them from the host's glibc; see DESIGN.md "Numerics / libm".
#pragma once
#include <cuda_runtime.h>

namespace pwn {

__device__ __forceinline__ float libm_sinf(float x) { return sinf(x); }
__device__ __forceinline__ float libm_cosf(float x) { return cosf(x); }
__device__ __forceinline__ void libm_sincosf(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ float libm_expf(float x) { return expf(x); }

} // namespace pwn
