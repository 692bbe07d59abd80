// pwntrace.cu -- libpwntrace.so: the C ABI of include/pwntrace.h over the sm_100a kernels.
//
// Host side of the boundary described in pwntrace.h: level flattening (the
// reference walks its AoS `level` in place, defs.h:96-111; the device gets a
// class plane + per-portal-cell transform records), per-frame sphere binning
// (level.h:1-39,64-81), frame buffers (main.c:31-33,395-400), launches, copies.
// No CPU implementation of the trace exists in this library.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pwntrace.h"
#include "trace_kernel.cuh"
#include "upscale_kernel.cuh"

using namespace pwn;

// ---- errors ----------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

#define CU(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
	return fail(PWN_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while(0)

// scoped device memory / events for the helpers that allocate per call: an early error return frees them
struct DevMem
{
	void *p = nullptr;
	~DevMem() { if(p) cudaFree(p); }
	template <typename T> T *as() const { return static_cast<T *>(p); }
};
struct ScopedEvent
{
	cudaEvent_t e = nullptr;
	~ScopedEvent() { if(e) cudaEventDestroy(e); }
};

extern "C" const char *pwn_last_error(void) { return g_err; }
extern "C" int pwn_abi_version(void) { return PWN_ABI_VERSION; }

// ---- context -----------------------------------------------------------------------------
struct FrameSet
{
	void *base = nullptr;      // one allocation: sbuf | tsbuf | zbuf  (so one IPC handle covers the frame)
	uint32_t *sbuf = nullptr, *tsbuf = nullptr;
	float *zbuf = nullptr;
	uint32_t *flags = nullptr; // 64 words after the planes: inter-GPU barrier flags (k_peer_barrier); then one counter per band
	uint32_t *band_done = nullptr;   // (contiguous-block sharding: lanes that have finished the band, summed over the frames so far)
	size_t off_sbuf = 0, off_tsbuf = 0, off_zbuf = 0, off_flags = 0, off_band_done = 0, bytes = 0;
	// frame set 0 only: a second group sbuf | tsbuf | zbuf | band counters at off_alt (same offsets inside the group), so that
	// two sharded frames can be in flight (block layout: the trace of frame n+1 runs under the blur and gather of frame n)
	size_t off_alt = 0, group_bytes = 0;
};

struct pwn_ctx
{
	int device = -1;
	cudaStream_t stream = nullptr, copy_stream = nullptr;
	cudaEvent_t ev_k1a = nullptr, ev_k1b = nullptr, ev_k2b = nullptr;
	cudaEvent_t ev_done[3] = {nullptr, nullptr, nullptr}, ev_copied[3] = {nullptr, nullptr, nullptr};
	cudaEvent_t ev_chunk[16] = {};   // pwn_trace_screen: "chunk c traced" (0-7) and "chunk c blurred" (8-15)
	uint32_t *h_defer = nullptr;     // pinned: how many pixels the chunked blur had to leave to k_blur_fix
	cudaStream_t dbg_stream = nullptr;   // pwn_debug_flags
	uint32_t *h_dbg = nullptr;

	// level
	uint32_t *d_cellmap = nullptr;
	std::vector<uint32_t> h_cellmap;   // host copy: pwn_parts_upload stages a copy with the "cell holds spheres" bits
	uint32_t *d_cellmap_sph = nullptr; // that copy on the device (inside d_parts_blob); what the sphere builds of K1 walk
	PortalRec *d_portals = nullptr;
	int w = 0, h = 0;
	bool have_level = false;

	// math tables
	uint32_t *d_rcp = nullptr, *d_rsqrt = nullptr;
	std::vector<uint32_t> h_rcp, h_rsqrt;   // the tables as the device reads them: copied in front of every cell map (STD builds of K1)
	bool std_tables = false;                // the usual widths (rcpps 11 bits, rsqrtps 10 + parity): 2048 entries each
	uint32_t *d_cellmap_alloc = nullptr;    // d_cellmap's allocation: 16 KB of tables, then the map
	int rcp_bits = 0, rsqrt_bits = 0;
	bool have_tables = false;

	// spheres
	float4 *d_sph_pos = nullptr, *d_sph_col = nullptr;
	float2 *d_sph_rr = nullptr;
	uint32_t *d_parts_off = nullptr, *d_parts_idx = nullptr, *d_parts_pk = nullptr;   // all six point into d_parts_blob
	void *d_parts_blob = nullptr, *h_parts_stage = nullptr;
	size_t parts_cap = 0;
	cudaEvent_t ev_parts = nullptr;
	int nsph = 0;
	float sph_w = 1.0f;

	// frame
	int rwidth = 0, rheight = 0;
	FrameSet fs[3];            // three sets: the pose pipeline keeps trace(i), blur(i-1) and copy(i-2) apart
	pwn_ids *d_ids = nullptr;
	uint2 *d_jump = nullptr;   // per 4-pixel group: the LCG map composed 32*g times (K2)
	uint32_t *r_sbuf = nullptr, *r_tsbuf = nullptr;   // redirect targets (peer frame) or null
	float *r_zbuf = nullptr;
	unsigned long long *d_sums = nullptr;
	size_t sums_cap = 0;
	unsigned long long *d_opcount = nullptr;
	uint32_t *d_surface = nullptr;   // K3: the display surface (screen->pixels)
	size_t surface_words = 0;
	cudaEvent_t ev_k3a = nullptr, ev_k3b = nullptr;
	bool timed_k3 = false;

	int nshards = 1, shard = 0;
	int layout = PWN_SHARD_INTERLEAVED;
	int halo_bands = 0;        // contiguous blocks: bands exchanged with each neighbour, 0 = from the frame height
	std::vector<int> blk_bounds;   // contiguous blocks: first band of every shard + the band count (pwn_shard_blocks), or empty
	uint32_t blk_frames = 0;   // sharded frames traced in the block layout since pwn_frame_peers (the band counters' epoch)
	cudaEvent_t ev_blk_k1[2] = {nullptr, nullptr}, ev_blk_done[2] = {nullptr, nullptr};
	cudaEvent_t ev_join = nullptr;
	int blk_par = 0;           // which group of frame set 0 holds the latest sharded frame of the block layout
	int reflect = 2;           // REFLECT as shipped (defs.h:8)
	int screen_chunks = 0;     // pwn_trace_screen: row chunks per frame, 0 = automatic
	// sharded frame with the blur pass: the other shards' frame allocations (pwn_frame_peers)
	char *peer_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	int npeer_shards = 0;
	uint32_t epoch = 0;
	bool barrier_pending = false;  // a sharded frame's barriers are in the stream and their error word has not been looked at
	unsigned peer_timeouts = 0;    // sharded frames that ended in a barrier time-out since pwn_frame_peers
	bool allgather_now = false;   // launch_trace: trace into the pre-blur buffer here and on every peer, no blur
	int libm_fma = 1;          // glibc build the reference host would run: 1 "_fma", 0 "_sse2"
	int libm_exact = -1;       // pwn_host_libm_probe's verdict: 1 the host runs exactly that build, 0 neither, -1 tables were handed in
	unsigned long long launches = 0;
	bool timed = false;        // events of the last trace are valid
	bool timed_blur = false;
};

static int set_device(pwn_ctx *ctx)
{
	CU(cudaSetDevice(ctx->device));
	return PWN_OK;
}

#define CTX_ENTER(ctx) do { if((ctx) == nullptr) return fail(PWN_E_ARG, "null context"); \
	int rc_ = set_device(ctx); if(rc_ != PWN_OK) return rc_; } while(0)

// Under CUDA's lazy module loading (the default since 12.2) a kernel's code reaches the device at its first launch,
// and that load waits for the device to go idle.  A sharded frame parks kernels on the device that spin until other work
// arrives (k_peer_barrier, k_wait_bands): a first launch behind one of them stalls until its time-out.  Every kernel of the
// library is therefore loaded when the first context is created (cudaFuncGetAttributes loads the function).
// Every kernel also asks for the SAME shared-memory carve-out (the 100 KB step: K1's nine CTAs of 9.5 KB + 1 KB each).  An SM
// changes its L1 / shared-memory split only when it is empty: left to the driver's per-kernel choice, a blur launch whose
// waiting CTAs sit on every SM keeps the trace kernel they wait for from starting anywhere (measured: each frame that
// started its blur launch before its trace launch ran into the 4 s time-out).
template <typename F> static cudaError_t preload(F *fn)
{
	cudaFuncAttributes a;
	cudaError_t e = cudaFuncGetAttributes(&a, (const void *)fn);
	if(e != cudaSuccess) return e;
	return cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout, 43);
}

static int preload_kernels()
{
#define PWN_PRELOAD(K) do { cudaError_t e_ = preload(K); if(e_ != cudaSuccess) \
	return fail(PWN_E_CUDA, "loading %s: %s", #K, cudaGetErrorString(e_)); } while(0)
#define PWN_PRELOAD_TRACE_(S, L, T) PWN_PRELOAD((k_trace<S, false, L, false, T>)); PWN_PRELOAD((k_trace<S, false, L, true, T>)); PWN_PRELOAD((k_trace<S, true, L, false, T>))
#define PWN_PRELOAD_TRACE(S, L) PWN_PRELOAD_TRACE_(S, L, false); PWN_PRELOAD_TRACE_(S, L, true)
	PWN_PRELOAD_TRACE(false, false); PWN_PRELOAD_TRACE(false, true); PWN_PRELOAD_TRACE(true, false); PWN_PRELOAD_TRACE(true, true);
	PWN_PRELOAD(k_blur<false>); PWN_PRELOAD(k_blur<true>); PWN_PRELOAD(k_wait_bands); PWN_PRELOAD(k_blur_fix); PWN_PRELOAD(k_peer_barrier);
	PWN_PRELOAD(k_checksum); PWN_PRELOAD(k_fp32_peak); PWN_PRELOAD(k_pack_eval); PWN_PRELOAD(k_libm_eval<false>); PWN_PRELOAD(k_libm_eval<true>);
	PWN_PRELOAD(k_upscale_packed<0>); PWN_PRELOAD(k_upscale_packed<1>); PWN_PRELOAD(k_upscale_packed<2>); PWN_PRELOAD(k_upscale_packed<4>);
	PWN_PRELOAD(k_upscale_any);
#undef PWN_PRELOAD_TRACE
#undef PWN_PRELOAD_TRACE_
#undef PWN_PRELOAD
	return PWN_OK;
}

extern "C" int pwn_create(pwn_ctx **out, int device)
{
	if(out == nullptr) return fail(PWN_E_ARG, "pwn_create: null out");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if(e != cudaSuccess || n <= 0)
		return fail(PWN_E_NODEVICE, "no CUDA device (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
	if(device < 0 || device >= n) return fail(PWN_E_NODEVICE, "device %d out of range (%d devices)", device, n);
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	if(prop.major != 10)
		return fail(PWN_E_NODEVICE, "device %d is sm_%d%d; libpwntrace is built for sm_100a only", device, prop.major, prop.minor);

	pwn_ctx *ctx = new pwn_ctx();
	struct Guard { pwn_ctx *c; ~Guard() { if(c) pwn_destroy(c); } } guard{ctx};   // a failed step below undoes the ones before it
	ctx->device = device;
	CU(cudaSetDevice(device));
	{ int rc = preload_kernels(); if(rc != PWN_OK) return rc; }
	CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	{
		// the second stream carries copies and, for sharded frames, the blur launches that run under the trace: with the
		// higher priority their few CTAs are placed ahead of the trace's pending ones whenever an SM has room
		int lo = 0, hi = 0;
		CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		if(getenv("PWN_NO_STREAM_PRIORITY") != nullptr) hi = lo;
		CU(cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, hi));
	}
	CU(cudaEventCreate(&ctx->ev_k1a));
	CU(cudaEventCreate(&ctx->ev_k1b));
	CU(cudaEventCreate(&ctx->ev_k2b));
	CU(cudaEventCreate(&ctx->ev_k3a));
	CU(cudaEventCreate(&ctx->ev_k3b));
	for(int i = 0; i < 3; i++)
	{
		CU(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
	}
	for(int i = 0; i < 16; i++) CU(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
	for(int i = 0; i < 2; i++)
	{
		CU(cudaEventCreateWithFlags(&ctx->ev_blk_k1[i], cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&ctx->ev_blk_done[i], cudaEventDisableTiming));
	}
	CU(cudaHostAlloc((void **)&ctx->h_defer, 64, cudaHostAllocDefault));
	ctx->h_defer[0] = 0;

	// LCG jumps for K2 (util.h:1-6: s -> 25739 s + 4): entries 0..3 = the map composed 0, 2, 4, 6 times
	uint32_t ja[8] = {1u, 0, 0, 0, 0, 0, 0, 0}, jc[8] = {0u, 0, 0, 0, 0, 0, 0, 0};
	for(int k = 1; k < 4; k++)
	{
		uint32_t a = ja[k-1], c = jc[k-1];
		for(int r = 0; r < 2; r++) { c = 25739u*c + 4u; a = 25739u*a; }
		ja[k] = a; jc[k] = c;
	}
	CU(cudaMemcpyToSymbol(c_jumpA, ja, sizeof(ja)));
	CU(cudaMemcpyToSymbol(c_jumpC, jc, sizeof(jc)));

	guard.c = nullptr;
	*out = ctx;
	return PWN_OK;
}

static void free_frames(pwn_ctx *ctx)
{
	for(int i = 0; i < 3; i++) { cudaFree(ctx->fs[i].base); ctx->fs[i] = FrameSet(); }
	cudaFree(ctx->d_ids); ctx->d_ids = nullptr;
	cudaFree(ctx->d_jump); ctx->d_jump = nullptr;
	ctx->rwidth = ctx->rheight = 0;
}

extern "C" int pwn_destroy(pwn_ctx *ctx)
{
	if(ctx == nullptr) return PWN_OK;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	free_frames(ctx);
	cudaFree(ctx->d_cellmap_alloc); cudaFree(ctx->d_portals);
	cudaFree(ctx->d_rcp); cudaFree(ctx->d_rsqrt);
	cudaFree(ctx->d_parts_blob); cudaFreeHost(ctx->h_parts_stage);
	if(ctx->ev_parts) cudaEventDestroy(ctx->ev_parts);
	cudaFree(ctx->d_sums); cudaFree(ctx->d_opcount); cudaFree(ctx->d_surface);
	cudaEvent_t evs[] = {ctx->ev_k3a, ctx->ev_k3b, ctx->ev_k1a, ctx->ev_k1b, ctx->ev_k2b, ctx->ev_done[0], ctx->ev_done[1], ctx->ev_done[2],
		ctx->ev_copied[0], ctx->ev_copied[1], ctx->ev_copied[2]};
	for(cudaEvent_t e : evs) if(e) cudaEventDestroy(e);
	for(cudaEvent_t e : ctx->ev_chunk) if(e) cudaEventDestroy(e);
	for(int i = 0; i < 2; i++) { if(ctx->ev_blk_k1[i]) cudaEventDestroy(ctx->ev_blk_k1[i]); if(ctx->ev_blk_done[i]) cudaEventDestroy(ctx->ev_blk_done[i]); }
	if(ctx->ev_join) cudaEventDestroy(ctx->ev_join);
	if(ctx->h_defer) cudaFreeHost(ctx->h_defer);
	if(ctx->h_dbg) cudaFreeHost(ctx->h_dbg);
	if(ctx->dbg_stream) cudaStreamDestroy(ctx->dbg_stream);
	if(ctx->stream) cudaStreamDestroy(ctx->stream);
	if(ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
	cudaGetLastError();
	delete ctx;
	return PWN_OK;
}

// ---- math tables ---------------------------------------------------------------------------
extern "C" int pwn_math_tables(pwn_ctx *ctx, const uint32_t *rcp_tab, int rcp_bits, const uint32_t *rsqrt_tab, int rsqrt_bits)
{
	CTX_ENTER(ctx);
	if(rcp_tab == nullptr || rsqrt_tab == nullptr || rcp_bits < 1 || rcp_bits > 23 || rsqrt_bits < 1 || rsqrt_bits > 23)
		return fail(PWN_E_ARG, "pwn_math_tables: bad tables (bits %d, %d)", rcp_bits, rsqrt_bits);
	CU(cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_rcp); cudaFree(ctx->d_rsqrt);
	ctx->d_rcp = ctx->d_rsqrt = nullptr;
	ctx->have_tables = false;
	size_t nr = (size_t)1 << rcp_bits, ns = (size_t)2 << rsqrt_bits;
	CU(cudaMalloc(&ctx->d_rcp, nr*sizeof(uint32_t)));
	CU(cudaMalloc(&ctx->d_rsqrt, ns*sizeof(uint32_t)));
	CU(cudaMemcpy(ctx->d_rcp, rcp_tab, nr*sizeof(uint32_t), cudaMemcpyHostToDevice));
	// The device's copy of the rsqrt table is re-indexed and re-biased so that rsqrt_emul needs neither the flip of
	// the exponent's low bit nor the "+1 before halving" (trace_kernel.cuh): entry i2 = [E & 1 | top mantissa bits]
	// holds the host entry of the other parity, minus (E & 1) << 23, plus the 0x20000000 of the result exponent.
	// (E + 1) >> 1 == (E >> 1) + (E & 1) for every E, so the result is the same word for every positive input.
	{
		std::vector<uint32_t> st2(ns);
		for(size_t i2 = 0; i2 < ns; i2++)
			st2[i2] = rsqrt_tab[i2 ^ ((size_t)1 << rsqrt_bits)] - ((uint32_t)((i2 >> rsqrt_bits) & 1) << 23) + 0x20000000u;
		CU(cudaMemcpy(ctx->d_rsqrt, st2.data(), ns*sizeof(uint32_t), cudaMemcpyHostToDevice));
		ctx->h_rsqrt = st2;
	}
	ctx->h_rcp.assign(rcp_tab, rcp_tab + nr);
	ctx->rcp_bits = rcp_bits; ctx->rsqrt_bits = rsqrt_bits;
	ctx->std_tables = (rcp_bits == STD_RCP_BITS && rsqrt_bits == STD_RSQRT_BITS);
	if(ctx->std_tables && ctx->d_cellmap_alloc != nullptr)      // a level is there already: its copy of the tables
	{
		CU(cudaMemcpy(ctx->d_cellmap_alloc, ctx->h_rcp.data(), STD_TAB_WORDS/2*4, cudaMemcpyHostToDevice));
		CU(cudaMemcpy(ctx->d_cellmap_alloc + STD_TAB_WORDS/2, ctx->h_rsqrt.data(), STD_TAB_WORDS/2*4, cudaMemcpyHostToDevice));
	}
	ctx->nsph = 0;      // (the sphere builds' copy of the map carries the tables too: pwn_parts_upload rebuilds it)
	ctx->have_tables = true;
	return PWN_OK;
}

extern "C" int pwn_math_tables_from_host(pwn_ctx *ctx)
{
	CTX_ENTER(ctx);
	int rb = 0, sb = 0;
	int rc = pwn_host_math_probe(&rb, &sb, nullptr, 0, nullptr, 0);
	if(rc != PWN_OK) return fail(rc, "host rcpps/rsqrtps is not a mantissa-bucket table on this CPU");
	std::vector<uint32_t> rt((size_t)1 << rb), st((size_t)2 << sb);
	rc = pwn_host_math_probe(&rb, &sb, rt.data(), rt.size(), st.data(), st.size());
	if(rc != PWN_OK) return fail(rc, "host math probe failed");
	int fma_build = 1, exact = 0;
	pwn_host_libm_probe(&fma_build, &exact);
	ctx->libm_fma = fma_build;
	ctx->libm_exact = exact;
	rc = pwn_math_tables(ctx, rt.data(), rb, st.data(), sb);
	if(rc != PWN_OK) return rc;
	// neither of the two glibc 2.39 builds the device restates: the floor and fog shading can then differ from this
	// host's reference build in the last bit.  Not an error (the frame is still a valid frame of the reference on a
	// glibc 2.39 host); the caller can ask (pwn_math_status) and refuse.
	if(!exact) snprintf(g_err, sizeof(g_err), "warning: this host's libm is neither glibc 2.39's _fma nor its _sse2 build of sinf/cosf/expf; "
		"the device uses the %s variant, bounce shading may differ from this host's reference build in the last bit", fma_build ? "_fma" : "_sse2");
	return PWN_OK;
}

extern "C" int pwn_math_status(pwn_ctx *ctx, int *rcp_bits, int *rsqrt_bits, int *libm_fma, int *libm_exact)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	if(rcp_bits) *rcp_bits = ctx->rcp_bits;
	if(rsqrt_bits) *rsqrt_bits = ctx->rsqrt_bits;
	if(libm_fma) *libm_fma = ctx->libm_fma;
	if(libm_exact) *libm_exact = ctx->libm_exact;
	return PWN_OK;
}

extern "C" int pwn_libm_config(pwn_ctx *ctx, int fma_build)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	if(fma_build != 0 && fma_build != 1) return fail(PWN_E_ARG, "pwn_libm_config: variant %d", fma_build);
	ctx->libm_fma = fma_build;
	ctx->libm_exact = -1;
	return PWN_OK;
}

extern "C" int pwn_debug_libm(pwn_ctx *ctx, const float *in, float *sin_out, float *cos_out, float *exp_out, int n)
{
	CTX_ENTER(ctx);
	if(in == nullptr || sin_out == nullptr || cos_out == nullptr || exp_out == nullptr || n <= 0) return fail(PWN_E_ARG, "pwn_debug_libm: bad arguments");
	DevMem mem;
	CU(cudaMalloc(&mem.p, sizeof(float)*4*(size_t)n));
	float *d = mem.as<float>();
	CU(cudaMemcpyAsync(d, in, sizeof(float)*n, cudaMemcpyHostToDevice, ctx->stream));
	if(ctx->libm_fma) k_libm_eval<true><<<(n + 255)/256, 256, 0, ctx->stream>>>(d, d + n, d + 2*(size_t)n, d + 3*(size_t)n, n);
	else k_libm_eval<false><<<(n + 255)/256, 256, 0, ctx->stream>>>(d, d + n, d + 2*(size_t)n, d + 3*(size_t)n, n);
	CU(cudaGetLastError());
	ctx->launches++;
	CU(cudaMemcpyAsync(sin_out, d + n, sizeof(float)*n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(cos_out, d + 2*(size_t)n, sizeof(float)*n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(exp_out, d + 3*(size_t)n, sizeof(float)*n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PWN_OK;
}

extern "C" int pwn_debug_pack(pwn_ctx *ctx, const float *bgra, uint32_t *packed, int n)
{
	CTX_ENTER(ctx);
	if(bgra == nullptr || packed == nullptr || n <= 0) return fail(PWN_E_ARG, "pwn_debug_pack: bad arguments");
	DevMem md, mo;
	CU(cudaMalloc(&md.p, sizeof(float4)*(size_t)n));
	CU(cudaMalloc(&mo.p, sizeof(uint32_t)*(size_t)n));
	float4 *d = md.as<float4>();
	uint32_t *o = mo.as<uint32_t>();
	CU(cudaMemcpyAsync(d, bgra, sizeof(float4)*(size_t)n, cudaMemcpyHostToDevice, ctx->stream));
	k_pack_eval<<<(n + 255)/256, 256, 0, ctx->stream>>>(d, o, n);
	CU(cudaGetLastError());
	ctx->launches++;
	CU(cudaMemcpyAsync(packed, o, sizeof(uint32_t)*(size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PWN_OK;
}

// ---- level ---------------------------------------------------------------------------------
static uint32_t class_of(int32_t c)
{
	switch(c)
	{
		case ';': return CL_ONE;
		case '$': return CL_ONE_FOG;
		case '"': return CL_RAISED;
		case '#': return CL_TWO;
		case '&': return CL_TWO_FOG;
		case '>': return CL_RAMP_GT;
		case '<': return CL_RAMP_LT;
		case ',': return CL_RAMP_COMMA;
		case '^': return CL_RAMP_CARET;
		default: return PWN_IS_PORTAL(c) ? CL_PORTAL : CL_SOLID;
	}
}

// what `switch(xcell)` at trace.h:416-440 does with a cell code
static uint8_t peek_of(int32_t xcell)
{
	if(xcell == '#' || xcell == '&') return PK_PASS;
	if(xcell == '"') return PK_RAISED;
	return PK_WALL;
}

extern "C" int pwn_level_upload(pwn_ctx *ctx, int w, int h, const int32_t *cells, int nportals, const pwn_portal *portals)
{
	CTX_ENTER(ctx);
	if(cells == nullptr || w <= 0 || h <= 0 || nportals < 0 || (nportals > 0 && portals == nullptr))
		return fail(PWN_E_ARG, "pwn_level_upload: bad arguments");
	if(w > 16384 || h > 16384) return fail(PWN_E_LIMIT, "level %dx%d exceeds 16384 cells per axis", w, h);

	// padded map: one border cell of CL_OOB all round, so the walk needs no bounds checks
	const int wp = w + 2, hp = h + 2;
	std::vector<uint32_t> map((size_t)wp*hp, (uint32_t)CL_OOB);
	std::vector<PortalRec> recs;
	for(int z = 0; z < h; z++)
	for(int x = 0; x < w; x++)
	{
		const int32_t c = cells[(size_t)z*w + x];
		uint32_t cls = class_of(c);
		uint32_t entry = cls;
		if(cls == CL_PORTAL)
		{
			const int pi = PWN_PORTAL_INDEX(c);
			if(pi < 0 || pi >= nportals)
				return fail(PWN_E_ARG, "cell (%d,%d) names portal %d but the table has %d entries", x, z, pi, nportals);
			const pwn_portal &pm = portals[pi];
			PortalRec r;
			memset(&r, 0, sizeof(r));
			// peek through the portal from a 2-high room (trace.h:404-413)
			int32_t xcell = c;
			if(pm.x1 == x && pm.z1 == z) xcell = pm.c2;
			else if(pm.x2 == x && pm.z2 == z) xcell = pm.c1;
			uint32_t kind, rot = 0;
			int dcx = 0, dcz = 0, dx = 0, dz = 0;
			if(pm.x2 == -1) kind = PR_UNPAIRED;                      // trace.h:512-526
			else if(pm.x1 == x && pm.z1 == z)                        // 1 -> 2, trace.h:530-537
			{
				kind = PR_OK; dcx = pm.x2; dcz = pm.z2;
				dx = pm.x2 - pm.x1; dz = pm.z2 - pm.z1;
				rot = (uint32_t)((-pm.rot12) & 3);
			}
			else if(pm.x2 == x && pm.z2 == z)                        // 2 -> 1, trace.h:539-545
			{
				kind = PR_OK; dcx = pm.x1; dcz = pm.z1;
				dx = -(pm.x2 - pm.x1); dz = -(pm.z2 - pm.z1);
				rot = (uint32_t)(pm.rot12 & 3);
			}
			else kind = PR_STRAY;                                    // trace.h:547-559
			if(kind == PR_OK && (dcx < 0 || dcx >= w || dcz < 0 || dcz >= h))
				return fail(PWN_E_ARG, "portal %d: end (%d,%d) lies outside the %dx%d grid", pi, dcx, dcz, w, h);
			r.dci = (dcz + 1)*wp + dcx + 1;
			r.dx = (float)dx; r.dz = (float)dz;
			r.flags = rot | (kind << 8) | ((uint32_t)peek_of(xcell) << 16);
			r.ccx = (float)dcx + 0.5f; r.ccz = (float)dcz + 0.5f;   // (cx + 0.5f), trace.h:584
			r.dcx = dcx; r.dcz = dcz;
			if(recs.size() >= (1u << 15)) return fail(PWN_E_LIMIT, "more than 32768 portal cells");   // bit 31 of an entry is CM_SPH_BIT
			entry |= (uint32_t)recs.size() << CM_REC_SHIFT;
			recs.push_back(r);
		}
		map[(size_t)(z + 1)*wp + x + 1] = entry;
	}
	// slow-transition bits of the room cells (see CM_SLOW_SHIFT in trace_kernel.cuh)
	for(int z = 0; z < h; z++)
	for(int x = 0; x < w; x++)
	{
		const size_t ci = (size_t)(z + 1)*wp + x + 1;
		const uint32_t c = map[ci] & 15u;
		if(!(c & 8u)) continue;   // rooms only
		if(c == CL_RAISED) map[ci] |= 15u << CM_RZ_SHIFT;
		if(c == CL_TWO || c == CL_TWO_FOG) map[ci] |= 15u << CM_TWO_SHIFT;
		const long nb[4] = {1, wp, -1, -wp};   // FXP FZP FXN FZN (defs.h:25-33)
		for(int d = 0; d < 4; d++)
		{
			const uint32_t n = map[ci + nb[d]] & 15u;
			const bool ntwo = (n == CL_TWO) || (n == CL_TWO_FOG);
			bool slow;
			if(c == CL_TWO || c == CL_TWO_FOG) slow = !ntwo;
			else if(c == CL_RAISED) slow = ntwo || n == CL_OOB;
			else slow = n == CL_OOB;
			if(slow) map[ci] |= 1u << (CM_SLOW_SHIFT + d);
			// a 2-high room next to a plain cell inside the grid: only the height window matters there
			if((c == CL_TWO || c == CL_TWO_FOG) && (n == CL_SOLID || n == CL_ONE || n == CL_ONE_FOG || (n >= CL_RAMP_GT && n <= CL_RAMP_CARET)))
				map[ci] |= 1u << (CM_HWIN_SHIFT + d);
		}
	}
	if(recs.empty()) { PortalRec r; memset(&r, 0, sizeof(r)); recs.push_back(r); }

	CU(cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_cellmap_alloc); cudaFree(ctx->d_portals);
	ctx->d_cellmap_alloc = ctx->d_cellmap = nullptr; ctx->d_portals = nullptr;
	ctx->have_level = false;
	CU(cudaMalloc(&ctx->d_cellmap_alloc, (STD_TAB_WORDS + map.size())*sizeof(uint32_t)));
	ctx->d_cellmap = ctx->d_cellmap_alloc + STD_TAB_WORDS;
	CU(cudaMemset(ctx->d_cellmap_alloc, 0, STD_TAB_WORDS*sizeof(uint32_t)));
	if(ctx->std_tables)
	{
		CU(cudaMemcpy(ctx->d_cellmap_alloc, ctx->h_rcp.data(), STD_TAB_WORDS/2*4, cudaMemcpyHostToDevice));
		CU(cudaMemcpy(ctx->d_cellmap_alloc + STD_TAB_WORDS/2, ctx->h_rsqrt.data(), STD_TAB_WORDS/2*4, cudaMemcpyHostToDevice));
	}
	CU(cudaMalloc(&ctx->d_portals, recs.size()*sizeof(PortalRec)));
	CU(cudaMemcpy(ctx->d_cellmap, map.data(), map.size()*sizeof(uint32_t), cudaMemcpyHostToDevice));
	CU(cudaMemcpy(ctx->d_portals, recs.data(), recs.size()*sizeof(PortalRec), cudaMemcpyHostToDevice));
	ctx->w = w; ctx->h = h;
	ctx->h_cellmap = map;
	ctx->d_cellmap_sph = nullptr;
	ctx->have_level = true;
	ctx->nsph = 0;   // the per-cell sphere lists belong to the previous grid: pwn_parts_upload rebuilds them
	return PWN_OK;
}

// (int)float the way the reference's x86 build converts (cvttss2si)
static int cvtt_host(float v)
{
	if(!(v < 2147483648.0f && v >= -2147483648.0f)) return (int)0x80000000;
	return (int)v;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static size_t align16(size_t v) { return align_up(v, 16); }

extern "C" int pwn_parts_upload(pwn_ctx *ctx, int nsph, const pwn_sphere *sph)
{
	CTX_ENTER(ctx);
	if(!ctx->have_level) return fail(PWN_E_STATE, "pwn_parts_upload before pwn_level_upload");
	if(nsph < 0 || (nsph > 0 && sph == nullptr)) return fail(PWN_E_ARG, "pwn_parts_upload: bad arguments");
	if(nsph == 0) { ctx->nsph = 0; return PWN_OK; }
	for(int i = 1; i < nsph; i++)
		if(sph[i].pos[3] != sph[0].pos[3]) return fail(PWN_E_UNSUPPORTED, "sphere %d: pos.w differs (script.h:28 sets 1.0f)", i);
	ctx->sph_w = sph[0].pos[3];

	// level_prepare_render + level_part_add_bbox (level.h:1-39,64-81): objs[] order per cell
	const int w = ctx->w, h = ctx->h, wp = w + 2;
	const size_t ncell = (size_t)wp*(h + 2);   // CSR over the padded map, indexed like the cell map
	std::vector<uint32_t> off(ncell + 1, 0), idx, cur;
	for(int pass = 0; pass < 2; pass++)
	{
		if(pass == 1) { idx.assign(off[ncell] ? off[ncell] : 1, 0); cur.assign(off.begin(), off.end() - 1); }
		for(int i = 0; i < nsph; i++)
		{
			int cx1 = cvtt_host(sph[i].pos[0] - sph[i].r), cz1 = cvtt_host(sph[i].pos[2] - sph[i].r);
			int cx2 = cvtt_host(sph[i].pos[0] + sph[i].r), cz2 = cvtt_host(sph[i].pos[2] + sph[i].r);
			// the reference writes out of bounds for cells outside the grid (level.h:1-39); only the cells inside are
			// ever looked at (trace.h:252), so the box is cut to the grid first -- a far-away or huge sphere would
			// otherwise walk up to 2^64 cells here
			if(cx1 < 0) cx1 = 0;
			if(cz1 < 0) cz1 = 0;
			if(cx2 > w - 1) cx2 = w - 1;
			if(cz2 > h - 1) cz2 = h - 1;
			for(int z = cz1; z <= cz2; z++)
			for(int x = cx1; x <= cx2; x++)
			{
				const size_t ci = (size_t)(z + 1)*wp + x + 1;
				if(pass == 0) off[ci + 1]++;
				else idx[cur[ci]++] = (uint32_t)i;
			}
		}
		if(pass == 0) for(size_t c = 0; c < ncell; c++) off[c+1] += off[c];
		else
		{
			// one pinned staging block [pos | col | rr | off | idx] and ONE stream-ordered copy into one
			// device block: the upload sits in the frame loop (level_prepare_render runs every frame,
			// main.c:95), five synchronous pageable copies cost more than the binning itself
			const size_t b_pos = 0, b_col = b_pos + align16((size_t)nsph*sizeof(float4)), b_rr = b_col + align16((size_t)nsph*sizeof(float4));
			const size_t b_off = b_rr + align16((size_t)nsph*sizeof(float2)), b_idx = b_off + align16((ncell + 1)*sizeof(uint32_t));
			const size_t b_pk = b_idx + align16(idx.size()*sizeof(uint32_t));
			const size_t b_tab = b_pk + align16(ncell*sizeof(uint32_t));      // 16 KB of tables in front of the map, as in d_cellmap_alloc
			const size_t b_map = b_tab + STD_TAB_WORDS*sizeof(uint32_t);
			const size_t total = b_map + align16(ncell*sizeof(uint32_t));
			if(ctx->h_cellmap.size() != ncell) return fail(PWN_E_STATE, "pwn_parts_upload: no host copy of the cell map");
			if(idx.size() >= (1u << 24)) return fail(PWN_E_LIMIT, "more than 2^24 sphere list entries");
			if(ctx->ev_parts == nullptr) CU(cudaEventCreateWithFlags(&ctx->ev_parts, cudaEventDisableTiming));
			else CU(cudaEventSynchronize(ctx->ev_parts));      // the previous upload has left the staging block
			if(total > ctx->parts_cap)
			{
				CU(cudaStreamSynchronize(ctx->stream));
				cudaFree(ctx->d_parts_blob); cudaFreeHost(ctx->h_parts_stage);
				ctx->d_parts_blob = nullptr; ctx->h_parts_stage = nullptr; ctx->parts_cap = 0;
				CU(cudaMalloc(&ctx->d_parts_blob, total*2));
				CU(cudaHostAlloc(&ctx->h_parts_stage, total*2, cudaHostAllocDefault));
				ctx->parts_cap = total*2;
			}
			char *hs = (char *)ctx->h_parts_stage, *ds = (char *)ctx->d_parts_blob;
			float4 *pos = (float4 *)(hs + b_pos), *col = (float4 *)(hs + b_col);
			float2 *rr = (float2 *)(hs + b_rr);
			for(int i = 0; i < nsph; i++)
			{
				pos[i] = make_float4(sph[i].pos[0], sph[i].pos[1], sph[i].pos[2], sph[i].pos[3]);
				col[i] = make_float4(sph[i].col[0], sph[i].col[1], sph[i].col[2], sph[i].col[3]);
				rr[i] = make_float2(sph[i].r, sph[i].refl);
			}
			memcpy(hs + b_off, off.data(), (ncell + 1)*sizeof(uint32_t));
			memcpy(hs + b_idx, idx.data(), idx.size()*sizeof(uint32_t));
			// per cell: the packed list handle, and the map entry again with CM_SPH_BIT where the list is not empty --
			// the walk then learns from the entry it has loaded anyway whether the cell's list is worth a look
			uint32_t *pk = (uint32_t *)(hs + b_pk), *mp = (uint32_t *)(hs + b_map);
			if(ctx->std_tables)
			{
				memcpy(hs + b_tab, ctx->h_rcp.data(), STD_TAB_WORDS/2*4);
				memcpy(hs + b_tab + STD_TAB_WORDS/2*4, ctx->h_rsqrt.data(), STD_TAB_WORDS/2*4);
			}
			for(size_t c = 0; c < ncell; c++)
			{
				const uint32_t cnt = off[c + 1] - off[c];
				pk[c] = off[c] | ((cnt < 255u ? cnt : 255u) << 24);
				mp[c] = ctx->h_cellmap[c] | (cnt ? (uint32_t)CM_SPH_BIT : 0u);
			}
			CU(cudaMemcpyAsync(ds, hs, total, cudaMemcpyHostToDevice, ctx->stream));
			CU(cudaEventRecord(ctx->ev_parts, ctx->stream));
			ctx->d_sph_pos = (float4 *)(ds + b_pos); ctx->d_sph_col = (float4 *)(ds + b_col); ctx->d_sph_rr = (float2 *)(ds + b_rr);
			ctx->d_parts_off = (uint32_t *)(ds + b_off); ctx->d_parts_idx = (uint32_t *)(ds + b_idx);
			ctx->d_parts_pk = (uint32_t *)(ds + b_pk);
			ctx->d_cellmap_sph = (uint32_t *)(ds + b_map);
		}
	}
	ctx->nsph = nsph;
	return PWN_OK;
}

// ---- frame buffers ----------------------------------------------------------------------------

extern "C" int pwn_frame_config(pwn_ctx *ctx, int rwidth, int rheight, int want_ids)
{
	CTX_ENTER(ctx);
	if(rwidth <= 0 || rheight <= 0) return fail(PWN_E_ARG, "pwn_frame_config: %dx%d", rwidth, rheight);
	// (K1 indexes the frame with 32 bits; three planes of 2^31 pixels would not fit a B200 anyway)
	if((unsigned long long)rwidth*(unsigned long long)rheight >= (1ull << 31))
		return fail(PWN_E_LIMIT, "pwn_frame_config: %dx%d has 2^31 pixels or more", rwidth, rheight);
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaStreamSynchronize(ctx->copy_stream));
	if(rwidth != ctx->rwidth || rheight != ctx->rheight)
	{
		free_frames(ctx);
		const size_t plane = align_up((size_t)rwidth*rheight*4, 512);
		for(int i = 0; i < 3; i++)
		{
			FrameSet &f = ctx->fs[i];
			const size_t nbands = ((size_t)rheight + 31)/32;
			f.group_bytes = 3*plane + 256 + align_up(4*(nbands + 1), 256);
			f.bytes = f.group_bytes*(i == 0 ? 2 : 1);
			f.off_alt = (i == 0) ? f.group_bytes : 0;
			CU(cudaMalloc(&f.base, f.bytes));
			f.off_sbuf = 0; f.off_tsbuf = plane; f.off_zbuf = 2*plane; f.off_flags = 3*plane; f.off_band_done = 3*plane + 256;
			f.flags = (uint32_t *)((char *)f.base + f.off_flags);
			f.band_done = (uint32_t *)((char *)f.base + f.off_band_done);
			f.sbuf = (uint32_t *)((char *)f.base + f.off_sbuf);
			f.tsbuf = (uint32_t *)((char *)f.base + f.off_tsbuf);
			f.zbuf = (float *)((char *)f.base + f.off_zbuf);
			CU(cudaMemset(f.base, 0, f.bytes));
		}
		// the blur's stream position of every 4-pixel group: 32 draws per group (screen.h:95-109)
		const int ngroups = rwidth/4 + 1;
		std::vector<uint2> jt(ngroups);
		uint32_t a32 = 1u, c32 = 0u;
		for(int r = 0; r < 32; r++) { c32 = 25739u*c32 + 4u; a32 = 25739u*a32; }
		uint32_t a = 1u, c = 0u;
		for(int g = 0; g < ngroups; g++) { jt[g] = make_uint2(a, c); c = a32*c + c32; a = a32*a; }
		CU(cudaMalloc(&ctx->d_jump, ngroups*sizeof(uint2)));
		CU(cudaMemcpy(ctx->d_jump, jt.data(), ngroups*sizeof(uint2), cudaMemcpyHostToDevice));
		ctx->rwidth = rwidth; ctx->rheight = rheight;
		for(int t = 0; t < 8; t++) ctx->peer_base[t] = nullptr;
		ctx->npeer_shards = 0;
		ctx->epoch = 0;
	}
	if(want_ids && ctx->d_ids == nullptr)
	{
		CU(cudaMalloc(&ctx->d_ids, (size_t)rwidth*rheight*3*sizeof(pwn_ids)));
		CU(cudaMemset(ctx->d_ids, 0xFF, (size_t)rwidth*rheight*3*sizeof(pwn_ids)));
	}
	ctx->r_sbuf = ctx->r_tsbuf = nullptr; ctx->r_zbuf = nullptr;
	return PWN_OK;
}

extern "C" int pwn_frame_clear(pwn_ctx *ctx)
{
	CTX_ENTER(ctx);
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "pwn_frame_clear before pwn_frame_config");
	// the planes only: the barrier flags behind them carry the epoch of the last sharded frame
	for(int i = 0; i < 3; i++) CU(cudaMemsetAsync(ctx->fs[i].base, 0, ctx->fs[i].off_flags, ctx->stream));
	CU(cudaMemsetAsync((char *)ctx->fs[0].base + ctx->fs[0].off_alt, 0, ctx->fs[0].off_flags, ctx->stream));
	if(ctx->d_ids) CU(cudaMemsetAsync(ctx->d_ids, 0xFF, (size_t)ctx->rwidth*ctx->rheight*3*sizeof(pwn_ids), ctx->stream));
	return PWN_OK;
}

extern "C" int pwn_shard_layout(pwn_ctx *ctx, int layout, int halo_bands)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	if(layout != PWN_SHARD_INTERLEAVED && layout != PWN_SHARD_BLOCKS) return fail(PWN_E_ARG, "pwn_shard_layout: layout %d", layout);
	if(halo_bands < 0 || halo_bands > 4096) return fail(PWN_E_ARG, "pwn_shard_layout: %d halo bands", halo_bands);
	ctx->layout = layout; ctx->halo_bands = halo_bands;
	return PWN_OK;
}

extern "C" int pwn_shard_blocks(pwn_ctx *ctx, int nshards, const int32_t *band_start)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	if(band_start == nullptr || nshards == 0) { ctx->blk_bounds.clear(); return PWN_OK; }
	if(nshards < 1 || nshards > 8) return fail(PWN_E_ARG, "pwn_shard_blocks: %d shards", nshards);
	if(band_start[0] != 0) return fail(PWN_E_ARG, "pwn_shard_blocks: the first block starts at band %d", band_start[0]);
	for(int t = 0; t < nshards; t++)
		if(band_start[t + 1] < band_start[t]) return fail(PWN_E_ARG, "pwn_shard_blocks: boundaries must not decrease");
	ctx->blk_bounds.assign(band_start, band_start + nshards + 1);
	return PWN_OK;
}

extern "C" int pwn_shard_config(pwn_ctx *ctx, int nshards, int shard)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	if(nshards < 1 || shard < 0 || shard >= nshards) return fail(PWN_E_ARG, "pwn_shard_config: shard %d of %d", shard, nshards);
	ctx->nshards = nshards; ctx->shard = shard;
	return PWN_OK;
}

// ---- contiguous-block sharding (PWN_SHARD_BLOCKS): who owns which bands, and which of them are halo -------------
// Pure host arithmetic (no context): exported as pwn_block_shape so that the CPU tests can check it for every shape.
static int blk_start_pure(int nbands, int nshards, const int32_t *bounds, int s)
{
	// explicit boundaries (pwn_shard_blocks), valid for the frame height they were given for; else an even split
	if(bounds != nullptr && bounds[nshards] == nbands) return bounds[s];
	return (int)((long long)s*nbands/nshards);
}

static int blk_halo_pure(int halo_bands, int dimy)
{
	if(halo_bands > 0) return halo_bands;
	// the blur reaches 0.002*dimy*(z-1) rows (screen.h:86,100-102): bands enough for surfaces up to ~8.5 units away;
	// farther ones are finished by k_blur_fix from the owning shard's rows
	int h = (int)((0.015*dimy + 31.0)/32.0);
	return h < 1 ? 1 : h;
}

// shard s: bands [b0, b0 + nb); its first ht bands are read by the shard above, its last hb by the shard below
static void blk_shape_pure(int dimy, int nshards, int halo_bands, const int32_t *bounds, int s, int &b0, int &nb, int &ht, int &hb)
{
	const int nbands = (dimy + 31)/32, h = blk_halo_pure(halo_bands, dimy);
	b0 = blk_start_pure(nbands, nshards, bounds, s);
	nb = blk_start_pure(nbands, nshards, bounds, s + 1) - b0;
	ht = (s > 0) ? (h < nb ? h : nb) : 0;
	hb = (s + 1 < nshards) ? (h < nb - ht ? h : nb - ht) : 0;
}

// ... and the bands [vlo, vhi) whose rows shard s will ever hold: its own and the halo bands its neighbours send
static void blk_window_pure(int dimy, int nshards, int halo_bands, const int32_t *bounds, int s, int &vlo, int &vhi)
{
	int b0, nb, ht, hb, ub0 = 0, unb = 0, uht = 0, uhb = 0, db0 = 0, dnb = 0, dht = 0, dhb = 0;
	blk_shape_pure(dimy, nshards, halo_bands, bounds, s, b0, nb, ht, hb);
	if(s > 0) blk_shape_pure(dimy, nshards, halo_bands, bounds, s - 1, ub0, unb, uht, uhb);
	if(s + 1 < nshards) blk_shape_pure(dimy, nshards, halo_bands, bounds, s + 1, db0, dnb, dht, dhb);
	vlo = b0 - uhb; vhi = b0 + nb + dht;
}

extern "C" int pwn_block_shape(int dimy, int nshards, int halo_bands, const int32_t *band_start, int shard, int32_t out[6])
{
	if(out == nullptr || dimy < 1 || nshards < 1 || nshards > 8 || shard < 0 || shard >= nshards || halo_bands < 0)
		return fail(PWN_E_ARG, "pwn_block_shape: bad arguments");
	int b0, nb, ht, hb, vlo, vhi;
	blk_shape_pure(dimy, nshards, halo_bands, band_start, shard, b0, nb, ht, hb);
	blk_window_pure(dimy, nshards, halo_bands, band_start, shard, vlo, vhi);
	out[0] = b0; out[1] = nb; out[2] = ht; out[3] = hb; out[4] = vlo; out[5] = vhi;
	return PWN_OK;
}

static const int32_t *blk_bounds_of(const pwn_ctx *ctx)
{
	return ((int)ctx->blk_bounds.size() == ctx->nshards + 1) ? ctx->blk_bounds.data() : nullptr;
}

static int blk_start(const pwn_ctx *ctx, int nbands, int s) { return blk_start_pure(nbands, ctx->nshards, blk_bounds_of(ctx), s); }
static int blk_halo(const pwn_ctx *ctx, int dimy) { return blk_halo_pure(ctx->halo_bands, dimy); }
static void blk_shape(const pwn_ctx *ctx, int dimy, int s, int &b0, int &nb, int &ht, int &hb)
{
	blk_shape_pure(dimy, ctx->nshards, ctx->halo_bands, blk_bounds_of(ctx), s, b0, nb, ht, hb);
}

struct BlkTrace   // what a K1 launch of pwn_trace_shard_blur adds in the block layout
{
	uint32_t *halo_up = nullptr, *halo_dn = nullptr, *done_local = nullptr, *done_up = nullptr, *done_dn = nullptr;
	int ht = 0, hb = 0;
	uint32_t lanes = 0, target = 0;
};

// ---- the hot path ----------------------------------------------------------------------------------
static BlurParams blur_params(pwn_ctx *ctx, const uint32_t *src, const float *z, uint32_t *dst, int dimx, int dimy)
{
	BlurParams B;
	B.src = src; B.zbuf = z; B.dst = dst; B.jump = ctx->d_jump;
	B.dimx = dimx; B.dimy = dimy; B.rwidth = ctx->rwidth; B.y0 = 0; B.y1 = dimy;
	B.nshards = 1; B.shard = 0;
	B.fstr = 0.002f * (float)dimy;   // screen.h:86
	B.ylo = 0; B.ylimit = dimy; B.defer_count = nullptr; B.defer_list = nullptr; B.zero = 0u;
	B.nsh = 0;
	for(int t = 0; t < 8; t++) B.src_sh[t] = nullptr;
	for(int t = 0; t < 9; t++) B.band_start[t] = 0;
	return B;
}

static int launch_blur(pwn_ctx *ctx, const BlurParams &B, int nrows = -1, cudaStream_t st = nullptr)
{
	dim3 grid((B.dimx + 255)/256, nrows < 0 ? B.dimy : nrows), block(256);
	if(B.defer_count != nullptr && (B.ylimit < B.dimy || B.ylo > 0)) k_blur<true><<<grid, block, 0, st ? st : ctx->stream>>>(B);
	else k_blur<false><<<grid, block, 0, st ? st : ctx->stream>>>(B);
	CU(cudaGetLastError());
	ctx->launches++;
	return PWN_OK;
}

template <bool SPH, bool LFMA, bool STD>
static void launch_k1(const TraceParams &P, dim3 grid, cudaStream_t st, bool ids, bool blurp)
{
	if(ids) k_trace<SPH, true, LFMA, false, STD><<<grid, K1_THREADS, 0, st>>>(P);
	else if(blurp)
	{
		// programmatic dependent launch: this grid may begin before the previous one has drained
		// (see the end of k_trace)
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = grid; cfg.blockDim = dim3(K1_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		cudaLaunchKernelEx(&cfg, k_trace<SPH, false, LFMA, true, STD>, P);
	}
	else k_trace<SPH, false, LFMA, false, STD><<<grid, K1_THREADS, 0, st>>>(P);
}

// K1 for one frame into frame set f (+ the blur of `prev`'s frame riding along, pose batches),
// then, for a single blurred frame, K2
static int launch_trace(pwn_ctx *ctx, const FrameSet &f, const float from[4], const float rayb[4], const float rdx[4], const float rdy[4],
	int dimx, int dimy, float sec, unsigned flags, bool timing, const FrameSet *prev = nullptr, bool defer_blur = false,
	int band0 = 0, int nbands_launch = -1, const BlkTrace *blk = nullptr)
{
	if(!ctx->have_level) return fail(PWN_E_STATE, "trace before pwn_level_upload");
	if(!ctx->have_tables) return fail(PWN_E_STATE, "trace before pwn_math_tables");
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "trace before pwn_frame_config");
	if(dimx <= 0 || dimy <= 0 || dimx > ctx->rwidth || dimy > ctx->rheight)
		return fail(PWN_E_ARG, "trace %dx%d does not fit the %dx%d frame", dimx, dimy, ctx->rwidth, ctx->rheight);
	// w lanes: cam rows x,y,z carry w == 0 and the position w == 1 through everything the
	// reference host does to cam (main.c:60-64,188,201-208,267,372-377)
	if(rayb[3] != 0.0f || rdx[3] != 0.0f || rdy[3] != 0.0f)
		return fail(PWN_E_UNSUPPORTED, "ray basis with non-zero w lanes");
	if(ctx->nsph > 0 && from[3] != ctx->sph_w)
		return fail(PWN_E_UNSUPPORTED, "camera w (%g) differs from sphere w (%g)", from[3], ctx->sph_w);
	if((flags & PWN_F_IDS) && ctx->d_ids == nullptr) return fail(PWN_E_STATE, "PWN_F_IDS without pwn_frame_config(want_ids)");
	if((flags & PWN_F_BLUR) && ctx->nshards > 1 && !ctx->allgather_now && blk == nullptr)
		return fail(PWN_E_UNSUPPORTED, "blur needs the whole pre-blur frame: gather the bands, then pwn_blur_frame");

	TraceParams P;
	memset(&P, 0, sizeof(P));
	P.lv.cellmap = (ctx->nsph > 0 && ctx->d_cellmap_sph) ? ctx->d_cellmap_sph : ctx->d_cellmap; P.lv.portals = ctx->d_portals; P.lv.w = ctx->w; P.lv.h = ctx->h; P.lv.wp = ctx->w + 2;
	P.lv.rcp_tab = ctx->d_rcp; P.lv.rsqrt_tab = ctx->d_rsqrt;
	P.lv.rcp_shift = 23 - ctx->rcp_bits; P.lv.rsqrt_shift = 23 - ctx->rsqrt_bits;
	P.lv.rsqrt_mask = (2u << ctx->rsqrt_bits) - 1u;
	P.lv.sph_pos = ctx->d_sph_pos; P.lv.sph_col = ctx->d_sph_col; P.lv.sph_rr = ctx->d_sph_rr;
	P.lv.parts_off = ctx->d_parts_off; P.lv.parts_idx = ctx->d_parts_idx; P.lv.parts_pk = ctx->d_parts_pk; P.lv.nsph = ctx->nsph;
	for(int i = 0; i < 4; i++) { P.from[i] = from[i]; P.rayb[i] = rayb[i]; P.rdx[i] = rdx[i]; P.rdy[i] = rdy[i]; }
	P.sec_current = sec;
	P.dimx = dimx; P.dimy = dimy; P.rwidth = ctx->rwidth;
	P.nshards = ctx->nshards; P.shard = ctx->shard;
	const int nbands = (dimy + 31)/32;
	P.nbands_local = nbands > ctx->shard ? (nbands - ctx->shard + ctx->nshards - 1)/ctx->nshards : 0;
	if(ctx->layout == PWN_SHARD_BLOCKS && ctx->nshards > 1)
	{
		int b0, nb, ht, hb;
		blk_shape(ctx, dimy, ctx->shard, b0, nb, ht, hb);
		P.blk = 1; P.blk_b0 = b0; P.nbands_local = nb;
		if(blk != nullptr)
		{
			P.blk_ht = blk->ht; P.blk_hb = blk->hb;
			P.halo_up = blk->halo_up; P.halo_dn = blk->halo_dn;
			P.done_local = blk->done_local; P.done_up = blk->done_up; P.done_dn = blk->done_dn;
			P.blk_lanes = blk->lanes; P.blk_target = blk->target;
		}
	}
	const bool blur = (flags & PWN_F_BLUR) != 0;
	uint32_t *sb = ctx->r_sbuf ? ctx->r_sbuf : f.sbuf, *tsb = ctx->r_tsbuf ? ctx->r_tsbuf : f.tsbuf;
	P.sbuf = blur ? tsb : sb;
	P.zbuf = ctx->r_zbuf ? ctx->r_zbuf : f.zbuf;
	P.ids = (flags & PWN_F_IDS) ? ctx->d_ids : nullptr;
	P.opcount = nullptr;
	if(flags & PWN_F_IDS)
	{
		if(ctx->d_opcount == nullptr) CU(cudaMalloc(&ctx->d_opcount, sizeof(unsigned long long)));
		CU(cudaMemsetAsync(ctx->d_opcount, 0, sizeof(unsigned long long), ctx->stream));
		P.opcount = ctx->d_opcount;
	}
	P.reflect_blur = (flags & PWN_F_NO_REFLECT_BLUR) ? 0 : 1;
	P.reflect = ctx->reflect;
	// a redirected frame lives on another GPU: whole 128-byte row segments per warp store cross NVLink
	// at about twice the rate of the four 32-byte pieces an 8x4 warp tile would send
	P.rowtile = (ctx->r_sbuf != nullptr || (blk != nullptr && (blk->halo_up != nullptr || blk->halo_dn != nullptr))) ? 1 : 0;
	if(ctx->allgather_now)
	{
		P.sbuf = f.tsbuf; P.zbuf = f.zbuf;
		for(int t = 0; t < ctx->nshards; t++)
			if(t != ctx->shard) P.peer[P.npeers++] = (uint32_t *)(ctx->peer_base[t] + f.off_tsbuf);
		P.rowtile = 1;
	}
	const bool blurp = blur && prev != nullptr && P.ids == nullptr;
	if(blurp) P.prev = blur_params(ctx, prev->tsbuf, prev->zbuf, prev->sbuf, dimx, dimy);

	ctx->timed = ctx->timed_blur = false;
	// (a launch may cover only bands [band0, band0 + nbands_launch) of the shard's: pwn_trace_screen's row chunks)
	P.band0 = band0;
	int nb_launch = P.nbands_local - band0;
	if(nbands_launch >= 0 && nbands_launch < nb_launch) nb_launch = nbands_launch;
	if(nb_launch > 0)
	{
		dim3 grid((dimx + 31)/32, nb_launch*(32/K1_ROWS));
		if(timing) CU(cudaEventRecord(ctx->ev_k1a, ctx->stream));
		const bool ids = P.ids != nullptr;
		const int variant = (ctx->nsph > 0 ? 4 : 0) | (ctx->libm_fma ? 2 : 0) | (ctx->std_tables ? 1 : 0);
		switch(variant)
		{
			case 0: launch_k1<false, false, false>(P, grid, ctx->stream, ids, blurp); break;
			case 1: launch_k1<false, false, true>(P, grid, ctx->stream, ids, blurp); break;
			case 2: launch_k1<false, true, false>(P, grid, ctx->stream, ids, blurp); break;
			case 3: launch_k1<false, true, true>(P, grid, ctx->stream, ids, blurp); break;
			case 4: launch_k1<true, false, false>(P, grid, ctx->stream, ids, blurp); break;
			case 5: launch_k1<true, false, true>(P, grid, ctx->stream, ids, blurp); break;
			case 6: launch_k1<true, true, false>(P, grid, ctx->stream, ids, blurp); break;
			default: launch_k1<true, true, true>(P, grid, ctx->stream, ids, blurp); break;
		}
		CU(cudaGetLastError());
		ctx->launches++;
		if(timing) { CU(cudaEventRecord(ctx->ev_k1b, ctx->stream)); ctx->timed = true; }
	}
	if(blur && !defer_blur && !ctx->allgather_now)
	{
		int rc = launch_blur(ctx, blur_params(ctx, tsb, P.zbuf, sb, dimx, dimy));
		if(rc != PWN_OK) return rc;
		if(timing) { CU(cudaEventRecord(ctx->ev_k2b, ctx->stream)); ctx->timed_blur = true; }
	}
	return PWN_OK;
}

extern "C" int pwn_trace_frame(pwn_ctx *ctx, const float from[4], const float rayb[4], const float rdx[4], const float rdy[4],
	int dimx, int dimy, float sec_current, unsigned flags)
{
	CTX_ENTER(ctx);
	if(from == nullptr || rayb == nullptr || rdx == nullptr || rdy == nullptr) return fail(PWN_E_ARG, "pwn_trace_frame: null vector");
	return launch_trace(ctx, ctx->fs[0], from, rayb, rdx, rdy, dimx, dimy, sec_current, flags, true);
}

// (the band counters of the block layout count frames since this call on EVERY shard: call it on all of them, with
// no sharded frame in flight)
static int reset_band_counters(pwn_ctx *ctx)
{
	ctx->blk_frames = 0;
	ctx->epoch = 0;          // the barrier flags count sharded frames since this call, like the band counters
	if(ctx->rwidth == 0) return PWN_OK;
	const FrameSet &f = ctx->fs[0];
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaStreamSynchronize(ctx->copy_stream));
	CU(cudaMemset(f.flags, 0, f.group_bytes - f.off_flags));
	CU(cudaMemset((char *)f.base + f.off_alt + f.off_flags, 0, f.group_bytes - f.off_flags));
	ctx->blk_par = 0;
	return PWN_OK;
}

// ---- sharded frame WITH the blur pass: all-gather fused into K1, device-side barriers, sharded K2 -----
extern "C" int pwn_frame_peers(pwn_ctx *ctx, int nshards, const uint64_t *base_dev)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	for(int t = 0; t < 8; t++) ctx->peer_base[t] = nullptr;
	ctx->npeer_shards = 0;
	ctx->barrier_pending = false;
	ctx->peer_timeouts = 0;
	{ int rc0 = set_device(ctx); if(rc0 != PWN_OK) return rc0; rc0 = reset_band_counters(ctx); if(rc0 != PWN_OK) return rc0; }
	if(nshards == 0) return PWN_OK;
	if(nshards < 2 || nshards > 8 || base_dev == nullptr) return fail(PWN_E_ARG, "pwn_frame_peers: %d shards (2..8)", nshards);
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "pwn_frame_peers before pwn_frame_config");
	if(nshards != ctx->nshards) return fail(PWN_E_STATE, "pwn_frame_peers: %d shards but pwn_shard_config says %d", nshards, ctx->nshards);
	for(int t = 0; t < nshards; t++)
	{
		ctx->peer_base[t] = (t == ctx->shard) ? (char *)ctx->fs[0].base : (char *)(uintptr_t)base_dev[t];
		if(ctx->peer_base[t] == nullptr) return fail(PWN_E_ARG, "pwn_frame_peers: null frame for shard %d", t);
	}
	ctx->npeer_shards = nshards;
	return PWN_OK;
}

static int launch_barrier(pwn_ctx *ctx, int slot, cudaStream_t st = nullptr)
{
	BarrierParams B;
	memset(&B, 0, sizeof(B));
	for(int t = 0; t < ctx->nshards; t++) B.peer_flags[t] = (uint32_t *)(ctx->peer_base[t] + ctx->fs[0].off_flags);
	B.local = ctx->fs[0].flags;
	B.nshards = ctx->nshards; B.shard = ctx->shard; B.slot = slot; B.epoch = ctx->epoch;
	B.timeout_ns = 4000000000ull;
	k_peer_barrier<<<1, 32, 0, st ? st : ctx->stream>>>(B);
	CU(cudaGetLastError());
	ctx->launches++;
	return PWN_OK;
}

// A sharded frame WITH the blur pass in the contiguous-block layout.  Stream 1: one K1 launch per frame traces the shard's
// bands, halo bands first, storing those into the neighbour's pre-blur buffer as well and counting finished lanes per band.
// Stream 2: the blur pass follows band group by band group, each launch gated by a one-warp kernel that waits for the
// counters of the bands it reads, and stores straight into shard 0's frame; then barrier A (every shard's trace of this
// frame is complete), k_blur_fix (samples that reached beyond the exchanged rows, read from the owning shard), barrier B
// (the frame on shard 0 is complete).  Frames alternate between two groups of planes, so stream 1 goes on with the trace of
// frame n+1 while stream 2 still blurs and gathers frame n: per frame the GPU does the trace and the blur, and GPU 0's
// NVLink ingest (the frame's 4 bytes per pixel) runs under both.  K1 of frame n+2 waits for barrier B of frame n.
static int trace_shard_blur_blocks(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags)
{
	float rayb[4], rdx[4], rdy[4];
	pwn_ray_setup(cam, dimx, dimy, rayb, rdx, rdy);
	const FrameSet &f0 = ctx->fs[0];
	const int nbands = (dimy + 31)/32, S = ctx->nshards, r = ctx->shard;
	int b0, nb, ht, hb;
	blk_shape(ctx, dimy, r, b0, nb, ht, hb);
	// what reaches this shard from its neighbours: the last hb bands of the shard above, the first ht of the one below
	int vlo, vhi;                                            // bands whose rows this shard will hold
	blk_window_pure(dimy, S, ctx->halo_bands, blk_bounds_of(ctx), r, vlo, vhi);
	ctx->epoch++;
	ctx->blk_frames++;
	ctx->barrier_pending = true;
	const int par = (int)((ctx->blk_frames - 1) & 1u);
	const size_t alt = par ? f0.off_alt : 0;      // this frame's group of planes (the same offset on every shard)
	ctx->blk_par = par;
	FrameSet f = f0;                               // a view of that group
	f.sbuf = (uint32_t *)((char *)f0.base + alt + f0.off_sbuf);
	f.tsbuf = (uint32_t *)((char *)f0.base + alt + f0.off_tsbuf);
	f.zbuf = (float *)((char *)f0.base + alt + f0.off_zbuf);
	f.band_done = (uint32_t *)((char *)f0.base + alt + f0.off_band_done);
	const uint32_t lanes_per_band = 32u*32u*(uint32_t)((dimx + 31)/32);
	const uint32_t target = ((ctx->blk_frames + 1)/2)*lanes_per_band;      // frames that have used this group, this one included
	uint32_t *defer_count = f0.flags + 32;
	cudaStream_t s1 = ctx->stream, s2 = ctx->copy_stream;
	const unsigned long long tmo = 4000000000ull;

	// ---- stream 1: the trace ----
	CU(cudaStreamWaitEvent(s1, ctx->ev_blk_done[par], 0));     // the frame before last used these planes: its barrier B
	CU(cudaEventRecord(ctx->ev_k1a, s1));
	BlkTrace bt;
	bt.ht = ht; bt.hb = hb;
	bt.done_local = f.band_done;
	bt.lanes = lanes_per_band; bt.target = target;
	if(r > 0) { bt.halo_up = (uint32_t *)(ctx->peer_base[r - 1] + alt + f0.off_tsbuf); bt.done_up = (uint32_t *)(ctx->peer_base[r - 1] + alt + f0.off_band_done); }
	if(r + 1 < S) { bt.halo_dn = (uint32_t *)(ctx->peer_base[r + 1] + alt + f0.off_tsbuf); bt.done_dn = (uint32_t *)(ctx->peer_base[r + 1] + alt + f0.off_band_done); }
	const int rc_trace = launch_trace(ctx, f, cam + 12, rayb, rdx, rdy, dimx, dimy, sec_current, (flags | PWN_F_BLUR) & ~PWN_F_IDS, false, nullptr, true, 0, -1, &bt);
	char trace_err[sizeof(g_err)];
	if(rc_trace != PWN_OK) memcpy(trace_err, g_err, sizeof(g_err));
	CU(cudaEventRecord(ctx->ev_k1b, s1));
	CU(cudaEventRecord(ctx->ev_blk_k1[par], s1));
	ctx->timed = true;

	// ---- stream 2: the blur pass in groups of bands, the barriers, the fix ----
	uint32_t *dst = (flags & PWN_F_LOCAL_BANDS) ? f.sbuf : (uint32_t *)(ctx->peer_base[0] + alt + f0.off_sbuf);
	CU(cudaMemsetAsync(defer_count, 0, 4, s2));      // (the queue is filled and drained on this stream, frame after frame)
	if(nb > 0 && rc_trace == PWN_OK)
	{
		// The block's bands in a few groups; each group is one ordinary blur launch behind a one-warp kernel that waits for
		// the counters of the bands the group reads.  (Nothing but those single warps ever waits on the device: a blur
		// launch whose own CTAs did the waiting, parked on every SM, kept the trace kernel of a frame that had not started
		// yet from starting at all.)
		const int h = blk_halo(ctx, dimy);
		static int ngroups = -1;
		if(ngroups < 0) { const char *e = getenv("PWN_BLK_BLUR_GROUPS"); ngroups = e ? atoi(e) : 4; if(ngroups < 1) ngroups = 1; }
		int group = (nb + ngroups - 1)/ngroups;
		if(group < 1) group = 1;
		for(int g0 = b0; g0 < b0 + nb; g0 += group)
		{
			const int g1 = (g0 + group < b0 + nb) ? g0 + group : b0 + nb;
			int w0 = g0 - h, w1 = g1 + h;          // bands this group may read without queueing
			if(w0 < vlo) w0 = vlo;
			if(w1 > vhi) w1 = vhi;
			if(w0 < 0) w0 = 0;
			if(w1 > nbands) w1 = nbands;
			WaitParams W;
			W.done = f.band_done; W.b0 = w0; W.b1 = w1; W.target = target; W.timeout_ns = tmo; W.err = f0.flags + 16;
			k_wait_bands<<<1, 32, 0, s2>>>(W);
			CU(cudaGetLastError());
			BlurParams B = blur_params(ctx, f.tsbuf, f.zbuf, dst, dimx, dimy);
			B.y0 = g0*32; B.y1 = (g1*32 < dimy) ? g1*32 : dimy;
			B.ylo = w0*32; B.ylimit = (w1*32 < dimy) ? w1*32 : dimy;
			B.defer_count = defer_count; B.defer_list = ctx->fs[1].sbuf;
			if(B.ylo == 0 && B.ylimit >= dimy) B.ylimit = dimy + 1;      // (the whole frame is here: still the queueing build, which reads with plain loads)
			k_blur<true><<<dim3((dimx + 255)/256, B.y1 - B.y0), 256, 0, s2>>>(B);
			CU(cudaGetLastError());
			ctx->launches += 2;
		}
	}
	CU(cudaStreamWaitEvent(s2, ctx->ev_blk_k1[par], 0));      // the own trace of this frame is complete
	int rc = launch_barrier(ctx, 0, s2);          // A: every shard's trace and blur launches of this frame are complete
	if(rc != PWN_OK) return rc;
	if(nb > 0 && rc_trace == PWN_OK)
	{
		BlurParams B = blur_params(ctx, f.tsbuf, f.zbuf, dst, dimx, dimy);
		B.defer_count = defer_count; B.defer_list = ctx->fs[1].sbuf;
		B.nsh = S;
		for(int t = 0; t < S; t++) B.src_sh[t] = (const uint32_t *)(ctx->peer_base[t] + alt + f0.off_tsbuf);
		for(int t = 0; t <= S; t++) B.band_start[t] = blk_start(ctx, nbands, t);
		k_blur_fix<<<296, 256, 0, s2>>>(B);
		CU(cudaGetLastError());
		ctx->launches++;
	}
	rc = launch_barrier(ctx, 1, s2);              // B: the frame on shard 0 is complete; these planes may be rewritten
	if(rc != PWN_OK) return rc;
	CU(cudaEventRecord(ctx->ev_k2b, s2));
	CU(cudaEventRecord(ctx->ev_blk_done[par], s2));
	ctx->timed_blur = true;
	if(rc_trace != PWN_OK) { memcpy(g_err, trace_err, sizeof(g_err)); return rc_trace; }
	return PWN_OK;
}

// Work the library has put on its second stream (the blur pass and the barriers of sharded frames in the block layout)
// becomes part of the first: what the caller records or launches on pwn_stream afterwards comes after it.
extern "C" int pwn_join(pwn_ctx *ctx)
{
	CTX_ENTER(ctx);
	if(ctx->ev_join == nullptr) CU(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
	CU(cudaEventRecord(ctx->ev_join, ctx->copy_stream));
	CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
	return PWN_OK;
}

extern "C" int pwn_trace_shard_blur(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags)
{
	CTX_ENTER(ctx);
	if(cam == nullptr) return fail(PWN_E_ARG, "pwn_trace_shard_blur: null cam");
	if(ctx->npeer_shards != ctx->nshards || ctx->nshards < 2) return fail(PWN_E_STATE, "pwn_trace_shard_blur before pwn_frame_peers");
	if(ctx->r_sbuf || ctx->r_zbuf || ctx->r_tsbuf) return fail(PWN_E_STATE, "pwn_trace_shard_blur with a redirected frame");
	if(dimx <= 0 || dimy <= 0 || dimx > ctx->rwidth || dimy > ctx->rheight) return fail(PWN_E_ARG, "pwn_trace_shard_blur: bad size");
	if(ctx->layout == PWN_SHARD_BLOCKS) return trace_shard_blur_blocks(ctx, cam, dimx, dimy, sec_current, flags);
	float rayb[4], rdx[4], rdy[4];
	pwn_ray_setup(cam, dimx, dimy, rayb, rdx, rdy);
	const FrameSet &f = ctx->fs[0];
	// The epoch counts the calls, whatever becomes of them: a shard whose trace cannot be launched still meets its
	// peers at both barriers of this frame (they would otherwise spin into the time-out, and every later frame would
	// be one epoch out of step) and reports the failure to its own caller.
	ctx->epoch++;
	ctx->barrier_pending = true;
	// K1: own bands into the pre-blur buffer of every shard (and zbuf here)
	ctx->allgather_now = true;
	const int rc_trace = launch_trace(ctx, f, cam + 12, rayb, rdx, rdy, dimx, dimy, sec_current, (flags | PWN_F_BLUR) & ~PWN_F_IDS, true);
	ctx->allgather_now = false;
	char trace_err[sizeof(g_err)];
	if(rc_trace != PWN_OK) memcpy(trace_err, g_err, sizeof(g_err));
	if(!ctx->timed) { CU(cudaEventRecord(ctx->ev_k1a, ctx->stream)); CU(cudaEventRecord(ctx->ev_k1b, ctx->stream)); ctx->timed = true; }
	// A: every shard's bands have landed everywhere
	int rc = launch_barrier(ctx, 0);
	if(rc != PWN_OK) return rc;
	// K2 on the own bands, straight into shard 0's frame
	const int nbands = (dimy + 31)/32;
	const int nb_local = nbands > ctx->shard ? (nbands - ctx->shard + ctx->nshards - 1)/ctx->nshards : 0;
	if(nb_local > 0 && rc_trace == PWN_OK)
	{
		uint32_t *dst = (flags & PWN_F_LOCAL_BANDS) ? f.sbuf : (uint32_t *)(ctx->peer_base[0] + f.off_sbuf);
		BlurParams B = blur_params(ctx, f.tsbuf, f.zbuf, dst, dimx, dimy);
		B.nshards = ctx->nshards; B.shard = ctx->shard;
		rc = launch_blur(ctx, B, nb_local*32);
		if(rc != PWN_OK) return rc;
	}
	// B: every shard's blurred bands are in shard 0's frame, and nobody still reads a pre-blur
	// buffer that the next frame's K1 will overwrite
	rc = launch_barrier(ctx, 1);
	if(rc != PWN_OK) return rc;
	CU(cudaEventRecord(ctx->ev_k2b, ctx->stream));
	ctx->timed_blur = true;
	if(rc_trace != PWN_OK) { memcpy(g_err, trace_err, sizeof(g_err)); return rc_trace; }
	return PWN_OK;
}

// After a sharded frame the stream has run two device-side barriers.  A peer that never arrived left the error word
// set (k_peer_barrier): the calls that hand the frame to the host look at it once the stream has drained, clear it, and
// fail -- the frame then holds stale bands of the late shard.
static int check_peer_barriers(pwn_ctx *ctx)
{
	if(!ctx->barrier_pending) return PWN_OK;
	ctx->barrier_pending = false;
	uint32_t v = 0;
	CU(cudaMemcpyAsync(&v, ctx->fs[0].flags + 16, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	if(v == 0u) return PWN_OK;
	ctx->peer_timeouts++;
	CU(cudaMemsetAsync(ctx->fs[0].flags + 16, 0, 4, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return fail(PWN_E_PEER, "sharded frame %u: a peer shard did not reach the device barrier within 4 s; the frame is incomplete", ctx->epoch);
}

extern "C" int pwn_read_bands(pwn_ctx *ctx, int dimx, int dimy, uint32_t *sbuf_host_frame)
{
	CTX_ENTER(ctx);
	if(sbuf_host_frame == nullptr) return fail(PWN_E_ARG, "pwn_read_bands: null frame");
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "pwn_read_bands before pwn_frame_config");
	if(dimx <= 0 || dimy <= 0 || dimx > ctx->rwidth || dimy > ctx->rheight) return fail(PWN_E_ARG, "pwn_read_bands: bad size");
	const FrameSet &f = ctx->fs[0];
	const int nbands = (dimy + 31)/32;
	if(ctx->layout == PWN_SHARD_BLOCKS && ctx->nshards > 1)
	{
		int b0, nb, ht, hb;
		blk_shape(ctx, dimy, ctx->shard, b0, nb, ht, hb);
		const int y0 = b0*32, y1 = ((b0 + nb)*32 < dimy) ? (b0 + nb)*32 : dimy;
		const size_t alt = (ctx->blk_par && ctx->blk_frames > 0) ? f.off_alt : 0;
		CU(cudaStreamSynchronize(ctx->copy_stream));
		if(y1 > y0)
			CU(cudaMemcpyAsync(sbuf_host_frame + (size_t)y0*ctx->rwidth, (char *)(f.sbuf + (size_t)y0*ctx->rwidth) + alt, (size_t)(y1 - y0)*ctx->rwidth*4,
				cudaMemcpyDeviceToHost, ctx->stream));
	}
	else
	for(int b = ctx->shard; b < nbands; b += ctx->nshards)
	{
		const int y0 = b*32, rows = (dimy - y0 < 32) ? dimy - y0 : 32;
		const size_t off = (size_t)y0*ctx->rwidth;
		CU(cudaMemcpyAsync(sbuf_host_frame + off, f.sbuf + off, (size_t)rows*ctx->rwidth*4, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CU(cudaStreamSynchronize(ctx->stream));
	return check_peer_barriers(ctx);
}

// test hook: the barrier flags, the error word and the band counters of frame set 0, read while kernels may be running
extern "C" int pwn_debug_flags(pwn_ctx *ctx, uint32_t *out, int nwords)
{
	CTX_ENTER(ctx);
	if(out == nullptr || nwords == 0) return fail(PWN_E_ARG, "pwn_debug_flags: bad arguments");
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "no frame configured");
	const FrameSet &f = ctx->fs[0];
	const bool alt = nwords < 0;          // (negative: the same words of the second group of planes)
	if(alt) nwords = -nwords;
	const size_t have = (f.group_bytes - f.off_flags)/4;
	if((size_t)nwords > have) nwords = (int)have;
	if(nwords > 1024) nwords = 1024;
	if(ctx->dbg_stream == nullptr)
	{
		CU(cudaStreamCreateWithFlags(&ctx->dbg_stream, cudaStreamNonBlocking));
		CU(cudaHostAlloc((void **)&ctx->h_dbg, 4096, cudaHostAllocDefault));
	}
	CU(cudaMemcpyAsync(ctx->h_dbg, (char *)f.flags + (alt ? f.off_alt : 0), (size_t)nwords*4, cudaMemcpyDeviceToHost, ctx->dbg_stream));
	CU(cudaStreamSynchronize(ctx->dbg_stream));
	memcpy(out, ctx->h_dbg, (size_t)nwords*4);
	return nwords;
}

extern "C" int pwn_frame_peers_status(pwn_ctx *ctx, int *timed_out)
{
	CTX_ENTER(ctx);
	if(timed_out == nullptr) return fail(PWN_E_ARG, "null pointer");
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "no frame configured");
	uint32_t v = 0;
	CU(cudaMemcpyAsync(&v, ctx->fs[0].flags + 16, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	*timed_out = (v != 0u || ctx->peer_timeouts != 0) ? 1 : 0;
	return PWN_OK;
}

extern "C" int pwn_trace_config(pwn_ctx *ctx, int reflect)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	// the kernel keeps the blend state of two bounce levels in registers
	if(reflect < 0 || reflect > 2) return fail(PWN_E_UNSUPPORTED, "REFLECT %d: this build traces 0..2 bounce levels", reflect);
	ctx->reflect = reflect;
	return PWN_OK;
}

extern "C" int pwn_blur_frame(pwn_ctx *ctx, int dimx, int dimy)
{
	CTX_ENTER(ctx);
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "pwn_blur_frame before pwn_frame_config");
	if(dimx <= 0 || dimy <= 0 || dimx > ctx->rwidth || dimy > ctx->rheight) return fail(PWN_E_ARG, "pwn_blur_frame: bad size");
	const FrameSet &f = ctx->fs[0];
	// the traced (pre-blur) frame is in sbuf: it becomes the gather source, as memcpy(tsbuf, sbuf) does (screen.h:75)
	CU(cudaMemcpyAsync(f.tsbuf, f.sbuf, (size_t)ctx->rwidth*ctx->rheight*4, cudaMemcpyDeviceToDevice, ctx->stream));
	int rc = launch_blur(ctx, blur_params(ctx, f.tsbuf, f.zbuf, f.sbuf, dimx, dimy));
	if(rc != PWN_OK) return rc;
	return PWN_OK;
}

extern "C" int pwn_trace_cam(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags)
{
	CTX_ENTER(ctx);
	if(cam == nullptr) return fail(PWN_E_ARG, "pwn_trace_cam: null cam");
	if(dimx <= 0 || dimy <= 0) return fail(PWN_E_ARG, "pwn_trace_cam: bad size");
	float rayb[4], rdx[4], rdy[4];
	pwn_ray_setup(cam, dimx, dimy, rayb, rdx, rdy);
	return launch_trace(ctx, ctx->fs[0], cam + 12, rayb, rdx, rdy, dimx, dimy, sec_current, flags, true);
}

extern "C" int pwn_sync(pwn_ctx *ctx)
{
	CTX_ENTER(ctx);
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaStreamSynchronize(ctx->copy_stream));
	return check_peer_barriers(ctx);
}

extern "C" int pwn_read_frame(pwn_ctx *ctx, uint32_t *sbuf, float *zbuf, pwn_ids *ids)
{
	CTX_ENTER(ctx);
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "pwn_read_frame before pwn_frame_config");
	const size_t n = (size_t)ctx->rwidth*ctx->rheight;
	const FrameSet &f = ctx->fs[0];
	// (after sharded frames of the block layout the latest frame is in the group of planes it was traced into)
	const size_t alt = (ctx->blk_par && ctx->blk_frames > 0) ? f.off_alt : 0;
	CU(cudaStreamSynchronize(ctx->copy_stream));
	if(sbuf) CU(cudaMemcpyAsync(sbuf, (char *)f.sbuf + alt, n*4, cudaMemcpyDeviceToHost, ctx->stream));
	if(zbuf) CU(cudaMemcpyAsync(zbuf, (char *)f.zbuf + alt, n*4, cudaMemcpyDeviceToHost, ctx->stream));
	if(ids)
	{
		if(ctx->d_ids == nullptr) return fail(PWN_E_STATE, "no ids buffer configured");
		CU(cudaMemcpyAsync(ids, ctx->d_ids, n*3*sizeof(pwn_ids), cudaMemcpyDeviceToHost, ctx->stream));
	}
	CU(cudaStreamSynchronize(ctx->stream));
	return check_peer_barriers(ctx);
}

extern "C" int pwn_write_frame(pwn_ctx *ctx, const uint32_t *sbuf, const float *zbuf)
{
	CTX_ENTER(ctx);
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "pwn_write_frame before pwn_frame_config");
	const size_t n = (size_t)ctx->rwidth*ctx->rheight;
	const FrameSet &f = ctx->fs[0];
	if(sbuf) CU(cudaMemcpyAsync(f.sbuf, sbuf, n*4, cudaMemcpyHostToDevice, ctx->stream));
	if(zbuf) CU(cudaMemcpyAsync(f.zbuf, zbuf, n*4, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PWN_OK;
}

// device -> host copy of rows [y0, y1) of a plane, the traced window only: the reference's trace leaves everything
// outside [0,dimx) x [0,dimy) of the host's buffers untouched
static int copy_rows_out(pwn_ctx *ctx, void *host, const void *dev, int dimx, int y0, int y1, cudaStream_t st)
{
	if(y1 <= y0) return PWN_OK;
	const size_t pitch = (size_t)ctx->rwidth*4, off = (size_t)y0*pitch;
	if(dimx == ctx->rwidth) CU(cudaMemcpyAsync((char *)host + off, (const char *)dev + off, (size_t)(y1 - y0)*pitch, cudaMemcpyDeviceToHost, st));
	else CU(cudaMemcpy2DAsync((char *)host + off, pitch, (const char *)dev + off, pitch, (size_t)dimx*4, (size_t)(y1 - y0), cudaMemcpyDeviceToHost, st));
	return PWN_OK;
}

// The frame is traced in up to PWN_SCREEN_CHUNKS row chunks so that the copies to the host run under the kernels:
// chunk c's zbuf rows leave as soon as K1(c) is done; with the blur pass, chunk c is blurred after K1(c+1) -- its
// gathers reach rows below it by 0.002*dimy*(z-1) (screen.h:86,100-102), which the next chunk covers unless a surface
// is very far away; a pixel whose samples reach further still is queued and finished by k_blur_fix once the whole
// frame is traced, and (only then) the frame is copied again.  One kernel stream, one copy stream.
#ifndef PWN_SCREEN_CHUNKS
#define PWN_SCREEN_CHUNKS 4
#endif
extern "C" int pwn_screen_chunks(pwn_ctx *ctx, int nchunks)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	if(nchunks < 0 || nchunks > 8) return fail(PWN_E_ARG, "pwn_screen_chunks: %d (0 = automatic, 1..8)", nchunks);
	ctx->screen_chunks = nchunks;
	return PWN_OK;
}

extern "C" int pwn_trace_screen(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags,
	uint32_t *sbuf_host, float *zbuf_host)
{
	CTX_ENTER(ctx);
	if(cam == nullptr || sbuf_host == nullptr) return fail(PWN_E_ARG, "pwn_trace_screen: null pointer");
	if(dimx <= 0 || dimy <= 0) return fail(PWN_E_ARG, "pwn_trace_screen: bad size");
	if(ctx->nshards != 1) return fail(PWN_E_STATE, "pwn_trace_screen on a sharded context");
	flags &= ~PWN_F_IDS;
	const bool blur = (flags & PWN_F_BLUR) != 0;
	float rayb[4], rdx[4], rdy[4];
	pwn_ray_setup(cam, dimx, dimy, rayb, rdx, rdy);
	const FrameSet &f = ctx->fs[0];
	const uint32_t *sb_dev = ctx->r_sbuf ? ctx->r_sbuf : f.sbuf, *tsb_dev = ctx->r_tsbuf ? ctx->r_tsbuf : f.tsbuf;
	const float *zb_dev = ctx->r_zbuf ? ctx->r_zbuf : f.zbuf;
	// chunks of whole 32-row bands; small frames are not worth the extra launches
	const int nbands = (dimy + 31)/32;
	int nchunk = PWN_SCREEN_CHUNKS;
	if((long long)dimx*dimy < 1024ll*512ll) nchunk = 1;
	if(ctx->screen_chunks > 0) nchunk = ctx->screen_chunks;
	if(nchunk > nbands/2) nchunk = nbands/2 > 0 ? nbands/2 : 1;
	if(nchunk > 8) nchunk = 8;
	const int bpc = (nbands + nchunk - 1)/nchunk;    // bands per chunk
	nchunk = (nbands + bpc - 1)/bpc;
	uint32_t *defer_count = f.flags + 32;
	if(blur && nchunk > 1) CU(cudaMemsetAsync(defer_count, 0, 4, ctx->stream));
	auto rows_of = [&](int c, int &y0, int &y1) { y0 = c*bpc*32; y1 = (c + 1)*bpc*32; if(y1 > dimy) y1 = dimy; };
	auto blur_chunk = [&](int c) -> int {
		int y0, y1, ny0, ny1;
		rows_of(c, y0, y1);
		BlurParams B = blur_params(ctx, tsb_dev, zb_dev, const_cast<uint32_t *>(sb_dev), dimx, dimy);
		B.y0 = y0; B.y1 = y1;
		if(c + 1 < nchunk) { rows_of(c + 1, ny0, ny1); B.ylimit = ny1; B.defer_count = defer_count; B.defer_list = ctx->fs[1].sbuf; }
		int rc = launch_blur(ctx, B, y1 - y0);
		if(rc != PWN_OK) return rc;
		CU(cudaEventRecord(ctx->ev_chunk[8 + c], ctx->stream));
		CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[8 + c], 0));
		return copy_rows_out(ctx, sbuf_host, sb_dev, dimx, y0, y1, ctx->copy_stream);
	};
	CU(cudaEventRecord(ctx->ev_k1a, ctx->stream));
	for(int c = 0; c < nchunk; c++)
	{
		int y0, y1;
		rows_of(c, y0, y1);
		// K1 alone (with the blur flag it writes the pre-blur buffer and leaves the blur to this function)
		int rc = launch_trace(ctx, f, cam + 12, rayb, rdx, rdy, dimx, dimy, sec_current, flags, false, nullptr, true, c*bpc, bpc);
		if(rc != PWN_OK) return rc;
		CU(cudaEventRecord(ctx->ev_chunk[c], ctx->stream));
		CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[c], 0));
		// zbuf rows are final when their K1 ends; without the blur pass so are the sbuf rows
		if(!blur) { rc = copy_rows_out(ctx, sbuf_host, sb_dev, dimx, y0, y1, ctx->copy_stream); if(rc != PWN_OK) return rc; }
		else if(c > 0) { rc = blur_chunk(c - 1); if(rc != PWN_OK) return rc; }
		if(zbuf_host) { rc = copy_rows_out(ctx, zbuf_host, zb_dev, dimx, y0, y1, ctx->copy_stream); if(rc != PWN_OK) return rc; }
	}
	CU(cudaEventRecord(ctx->ev_k1b, ctx->stream));
	ctx->timed = true;
	ctx->h_defer[0] = 0;
	if(blur)
	{
		int rc = blur_chunk(nchunk - 1);
		if(rc != PWN_OK) return rc;
		if(nchunk > 1)
		{
			BlurParams B = blur_params(ctx, tsb_dev, zb_dev, const_cast<uint32_t *>(sb_dev), dimx, dimy);
			B.defer_count = defer_count; B.defer_list = ctx->fs[1].sbuf;
			k_blur_fix<<<296, 256, 0, ctx->stream>>>(B);
			CU(cudaGetLastError());
			ctx->launches++;
			CU(cudaMemcpyAsync(ctx->h_defer, defer_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
		}
		CU(cudaEventRecord(ctx->ev_k2b, ctx->stream));
		ctx->timed_blur = true;
	}
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaStreamSynchronize(ctx->copy_stream));
	if(ctx->h_defer[0] != 0)
	{
		// (rare: a surface so far away that its blur samples reach beyond the next chunk) the chunks went out before
		// k_blur_fix had finished those pixels: the blurred frame goes out again
		int rc = copy_rows_out(ctx, sbuf_host, sb_dev, dimx, 0, dimy, ctx->stream);
		if(rc != PWN_OK) return rc;
		CU(cudaStreamSynchronize(ctx->stream));
	}
	return PWN_OK;
}

// ---- K3: screen_upscale (screen.h:126-149) ----------------------------------------------------------
extern "C" int pwn_screen_upscale(pwn_ctx *ctx, int dimx, int dimy, int rscale, int pitch_bytes, void *pixels_host)
{
	CTX_ENTER(ctx);
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "pwn_screen_upscale before pwn_frame_config");
	if(dimx <= 0 || dimy <= 0 || dimx > ctx->rwidth || dimy > ctx->rheight) return fail(PWN_E_ARG, "pwn_screen_upscale: bad size");
	if(rscale < 1 || rscale > 64) return fail(PWN_E_ARG, "pwn_screen_upscale: rscale %d", rscale);
	if(pitch_bytes <= 0 || (pitch_bytes & 3) || pitch_bytes/4 < dimx*rscale)
		return fail(PWN_E_ARG, "pwn_screen_upscale: pitch %d bytes for %d pixels x %d", pitch_bytes, dimx, rscale);
	UpscaleParams U;
	U.src = ctx->fs[0].sbuf;
	U.w = dimx; U.h = dimy; U.rwidth = ctx->rwidth; U.rscale = rscale;
	U.pitch = pitch_bytes/4;
	U.rowstep = (long long)dimx*rscale + (long long)U.pitch*(rscale - 1);
	// the words the reference's walk can touch end at (h-1)*rowstep + (rscale-1)*pitch + w*rscale <= h*rscale*pitch
	const size_t words = (size_t)dimy*rscale*U.pitch;
	if(words > ctx->surface_words)
	{
		CU(cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_surface); ctx->d_surface = nullptr; ctx->surface_words = 0;
		CU(cudaMalloc(&ctx->d_surface, words*4));
		CU(cudaMemsetAsync(ctx->d_surface, 0, words*4, ctx->stream));
		ctx->surface_words = words;
	}
	U.dst = ctx->d_surface;
	CU(cudaEventRecord(ctx->ev_k3a, ctx->stream));
	if(U.pitch == dimx*rscale && (U.pitch & 3) == 0 && (ctx->rwidth & 3) == 0)
	{
		dim3 grid((U.pitch/4 + 255)/256, dimy);
		if(rscale == 1) k_upscale_packed<1><<<grid, 256, 0, ctx->stream>>>(U);
		else if(rscale == 2) k_upscale_packed<2><<<grid, 256, 0, ctx->stream>>>(U);
		else if(rscale == 4) k_upscale_packed<4><<<grid, 256, 0, ctx->stream>>>(U);
		else k_upscale_packed<0><<<grid, 256, 0, ctx->stream>>>(U);
	}
	else
	{
		// padded pitch: the walk leaves words untouched; they read as 0 whatever the surface held before
		CU(cudaMemsetAsync(ctx->d_surface, 0, words*4, ctx->stream));
		dim3 grid((dimx*rscale + 255)/256, dimy);
		k_upscale_any<<<grid, 256, 0, ctx->stream>>>(U);
	}
	CU(cudaGetLastError());
	CU(cudaEventRecord(ctx->ev_k3b, ctx->stream));
	ctx->timed_k3 = true;
	ctx->launches++;
	if(pixels_host != nullptr)
	{
		CU(cudaMemcpyAsync(pixels_host, ctx->d_surface, words*4, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
	}
	return PWN_OK;
}

extern "C" int pwn_surface_device_ptr(pwn_ctx *ctx, uint64_t *surface_dev, uint64_t *bytes)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	if(surface_dev) *surface_dev = (uint64_t)(uintptr_t)ctx->d_surface;
	if(bytes) *bytes = (uint64_t)ctx->surface_words*4;
	return PWN_OK;
}

extern "C" int pwn_last_upscale_ms(pwn_ctx *ctx, float *k3_ms)
{
	CTX_ENTER(ctx);
	if(k3_ms == nullptr) return fail(PWN_E_ARG, "null pointer");
	CU(cudaStreamSynchronize(ctx->stream));
	*k3_ms = 0.0f;
	if(ctx->timed_k3) CU(cudaEventElapsedTime(k3_ms, ctx->ev_k3a, ctx->ev_k3b));
	return PWN_OK;
}

extern "C" int pwn_trace_poses(pwn_ctx *ctx, int n, const pwn_pose *poses, int dimx, int dimy, unsigned flags,
	uint32_t *sbuf_host, uint64_t *sums)
{
	CTX_ENTER(ctx);
	if(n < 0 || (n > 0 && poses == nullptr)) return fail(PWN_E_ARG, "pwn_trace_poses: bad arguments");
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "pwn_trace_poses before pwn_frame_config");
	if(ctx->r_sbuf) return fail(PWN_E_STATE, "pwn_trace_poses with a redirected frame");
	const size_t fpix = (size_t)ctx->rwidth*ctx->rheight;
	flags &= ~PWN_F_IDS;
	if(sums != nullptr)
	{
		if((size_t)n > ctx->sums_cap)
		{
			cudaFree(ctx->d_sums);
			ctx->sums_cap = (size_t)n;
			CU(cudaMalloc(&ctx->d_sums, ctx->sums_cap*sizeof(unsigned long long)));
		}
		CU(cudaMemsetAsync(ctx->d_sums, 0, (size_t)n*sizeof(unsigned long long), ctx->stream));
	}
	// Blurred batches are software-pipelined across frames: launch i traces pose i into set i%3
	// and, in the same kernel, blurs frame i-1 in set (i-1)%3; a plain K2 finishes the last frame.
	// Frame j is final (and is checksummed / copied out) after launch j+1.
	const bool pipelined = (flags & PWN_F_BLUR) != 0 && ctx->nshards == 1 && n > 1;
	auto finish = [&](int j) -> int {
		const int s = j % 3;
		if(sums != nullptr)
		{
			k_checksum<<<296, 256, 0, ctx->stream>>>(ctx->fs[s].sbuf, dimx, dimy, ctx->rwidth, ctx->d_sums + j);
			CU(cudaGetLastError());
			ctx->launches++;
		}
		if(sbuf_host != nullptr)
		{
			CU(cudaEventRecord(ctx->ev_done[s], ctx->stream));
			CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[s], 0));
			CU(cudaMemcpyAsync(sbuf_host + (size_t)j*fpix, ctx->fs[s].sbuf, fpix*4, cudaMemcpyDeviceToHost, ctx->copy_stream));
			CU(cudaEventRecord(ctx->ev_copied[s], ctx->copy_stream));
		}
		return PWN_OK;
	};
	for(int i = 0; i < n; i++)
	{
		const int s = i % 3;
		float rayb[4], rdx[4], rdy[4];
		pwn_ray_setup(poses[i].cam, dimx, dimy, rayb, rdx, rdy);
		// Before a frame set's sbuf is rewritten, the copy of the frame it still holds must have
		// drained.  Unpipelined: this launch rewrites sbuf of set s, which holds frame i-3.
		// Pipelined: the blur riding in this launch rewrites sbuf of set (i-1)%3, which holds frame i-4.
		if(sbuf_host != nullptr)
		{
			if(!pipelined && i >= 3) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[s], 0));
			if(pipelined && i >= 4) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[(i - 1) % 3], 0));
		}
		int rc = launch_trace(ctx, ctx->fs[s], poses[i].cam + 12, rayb, rdx, rdy, dimx, dimy, poses[i].sec_current, flags, false,
			(pipelined && i > 0) ? &ctx->fs[(i - 1) % 3] : nullptr, pipelined);
		if(rc != PWN_OK) return rc;
		if(pipelined) { if(i > 0) { rc = finish(i - 1); if(rc != PWN_OK) return rc; } }
		else { rc = finish(i); if(rc != PWN_OK) return rc; }
	}
	if(pipelined)
	{
		const FrameSet &f = ctx->fs[(n - 1) % 3];
		if(sbuf_host != nullptr && n >= 4) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[(n - 1) % 3], 0));
		int rc = launch_blur(ctx, blur_params(ctx, f.tsbuf, f.zbuf, f.sbuf, dimx, dimy));
		if(rc != PWN_OK) return rc;
		rc = finish(n - 1);
		if(rc != PWN_OK) return rc;
	}
	if(sums != nullptr && n > 0)
		CU(cudaMemcpyAsync(sums, ctx->d_sums, (size_t)n*sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaStreamSynchronize(ctx->copy_stream));
	return PWN_OK;
}

// ---- device-resident access ----------------------------------------------------------------------------
extern "C" int pwn_frame_device_ptrs(pwn_ctx *ctx, uint64_t *sbuf_dev, uint64_t *zbuf_dev, uint64_t *tsbuf_dev)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "no frame configured");
	if(sbuf_dev) *sbuf_dev = (uint64_t)(uintptr_t)ctx->fs[0].sbuf;
	if(zbuf_dev) *zbuf_dev = (uint64_t)(uintptr_t)ctx->fs[0].zbuf;
	if(tsbuf_dev) *tsbuf_dev = (uint64_t)(uintptr_t)ctx->fs[0].tsbuf;
	return PWN_OK;
}

extern "C" int pwn_frame_redirect(pwn_ctx *ctx, uint64_t sbuf_dev, uint64_t zbuf_dev, uint64_t tsbuf_dev)
{
	if(ctx == nullptr) return fail(PWN_E_ARG, "null context");
	ctx->r_sbuf = (uint32_t *)(uintptr_t)sbuf_dev;
	ctx->r_zbuf = (float *)(uintptr_t)zbuf_dev;
	ctx->r_tsbuf = (uint32_t *)(uintptr_t)tsbuf_dev;
	return PWN_OK;
}

extern "C" int pwn_frame_ipc_export(pwn_ctx *ctx, void *handle64, uint64_t *sbuf_off, uint64_t *zbuf_off, uint64_t *tsbuf_off)
{
	CTX_ENTER(ctx);
	if(handle64 == nullptr) return fail(PWN_E_ARG, "null handle");
	if(ctx->rwidth == 0) return fail(PWN_E_STATE, "no frame configured");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	cudaIpcMemHandle_t hnd;
	CU(cudaIpcGetMemHandle(&hnd, ctx->fs[0].base));
	memcpy(handle64, &hnd, 64);
	if(sbuf_off) *sbuf_off = ctx->fs[0].off_sbuf;
	if(zbuf_off) *zbuf_off = ctx->fs[0].off_zbuf;
	if(tsbuf_off) *tsbuf_off = ctx->fs[0].off_tsbuf;
	return PWN_OK;
}

extern "C" int pwn_frame_ipc_import(pwn_ctx *ctx, const void *handle64, uint64_t *base_dev)
{
	CTX_ENTER(ctx);
	if(handle64 == nullptr || base_dev == nullptr) return fail(PWN_E_ARG, "null pointer");
	cudaIpcMemHandle_t hnd;
	memcpy(&hnd, handle64, 64);
	void *p = nullptr;
	CU(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
	*base_dev = (uint64_t)(uintptr_t)p;
	return PWN_OK;
}

extern "C" int pwn_frame_ipc_close(pwn_ctx *ctx, uint64_t base_dev)
{
	CTX_ENTER(ctx);
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaIpcCloseMemHandle((void *)(uintptr_t)base_dev));
	return PWN_OK;
}

extern "C" int pwn_host_alloc(void **out, size_t bytes)
{
	if(out == nullptr) return fail(PWN_E_ARG, "null out");
	CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
	return PWN_OK;
}

extern "C" int pwn_host_free(void *p)
{
	CU(cudaFreeHost(p));
	return PWN_OK;
}

// The reference host mallocs sbuf/zbuf (main.c:395-400): pageable memory, which the copy engine
// reaches only through the driver's staging buffers.  Page-locking the host's own buffers in place
// lets pwn_trace_screen copy at PCIe rate without the host changing its allocation.
extern "C" int pwn_host_register(void *p, size_t bytes)
{
	if(p == nullptr || bytes == 0) return fail(PWN_E_ARG, "pwn_host_register: null buffer");
	cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
	if(e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return PWN_OK; }
	if(e != cudaSuccess) { cudaGetLastError(); return fail(PWN_E_CUDA, "cudaHostRegister(%p, %zu): %s", p, bytes, cudaGetErrorString(e)); }
	return PWN_OK;
}

extern "C" int pwn_host_unregister(void *p)
{
	if(p == nullptr) return fail(PWN_E_ARG, "pwn_host_unregister: null buffer");
	cudaError_t e = cudaHostUnregister(p);
	if(e != cudaSuccess) { cudaGetLastError(); return fail(PWN_E_CUDA, "cudaHostUnregister(%p): %s", p, cudaGetErrorString(e)); }
	return PWN_OK;
}

// ---- measurement hooks ------------------------------------------------------------------------------------
extern "C" int pwn_launch_count(pwn_ctx *ctx, uint64_t *count)
{
	if(ctx == nullptr || count == nullptr) return fail(PWN_E_ARG, "null pointer");
	*count = ctx->launches;
	return PWN_OK;
}

extern "C" int pwn_last_opcount(pwn_ctx *ctx, uint64_t *ops)
{
	CTX_ENTER(ctx);
	if(ops == nullptr) return fail(PWN_E_ARG, "null pointer");
	if(ctx->d_opcount == nullptr) return fail(PWN_E_STATE, "no PWN_F_IDS trace has run");
	unsigned long long v = 0;
	CU(cudaMemcpyAsync(&v, ctx->d_opcount, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	*ops = v;
	return PWN_OK;
}

extern "C" int pwn_last_kernel_ms(pwn_ctx *ctx, float *k1_ms, float *k2_ms)
{
	CTX_ENTER(ctx);
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaStreamSynchronize(ctx->copy_stream));
	if(k1_ms) { *k1_ms = 0.0f; if(ctx->timed) CU(cudaEventElapsedTime(k1_ms, ctx->ev_k1a, ctx->ev_k1b)); }
	if(k2_ms) { *k2_ms = 0.0f; if(ctx->timed && ctx->timed_blur) CU(cudaEventElapsedTime(k2_ms, ctx->ev_k1b, ctx->ev_k2b)); }
	return PWN_OK;
}

extern "C" int pwn_time_frames(pwn_ctx *ctx, const float cam[16], int dimx, int dimy, float sec_current, unsigned flags, int n, float *total_ms)
{
	CTX_ENTER(ctx);
	if(cam == nullptr || total_ms == nullptr || n <= 0) return fail(PWN_E_ARG, "pwn_time_frames: bad arguments");
	float rayb[4], rdx[4], rdy[4];
	pwn_ray_setup(cam, dimx, dimy, rayb, rdx, rdy);
	ScopedEvent ea, eb;
	CU(cudaEventCreate(&ea.e)); CU(cudaEventCreate(&eb.e));
	cudaEvent_t a = ea.e, b = eb.e;
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaEventRecord(a, ctx->stream));
	for(int i = 0; i < n; i++)
	{
		int rc = launch_trace(ctx, ctx->fs[0], cam + 12, rayb, rdx, rdy, dimx, dimy, sec_current, flags & ~PWN_F_IDS, false);
		if(rc != PWN_OK) return rc;
	}
	CU(cudaEventRecord(b, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaEventElapsedTime(total_ms, a, b));
	return PWN_OK;
}

extern "C" int pwn_fp32_peak(pwn_ctx *ctx, double *tlaneops)
{
	CTX_ENTER(ctx);
	if(tlaneops == nullptr) return fail(PWN_E_ARG, "null pointer");
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, ctx->device));
	const int blocks = prop.multiProcessorCount*8, threads = 256, iters = 4096;
	DevMem mem;
	CU(cudaMalloc(&mem.p, sizeof(float)*(size_t)blocks*threads));
	float *d = mem.as<float>();
	ScopedEvent ea, eb;
	CU(cudaEventCreate(&ea.e)); CU(cudaEventCreate(&eb.e));
	cudaEvent_t a = ea.e, b = eb.e;
	float best = 1e30f;
	for(int rep = 0; rep < 5; rep++)
	{
		CU(cudaEventRecord(a, ctx->stream));
		k_fp32_peak<<<blocks, threads, 0, ctx->stream>>>(d, iters, 1.0000001f, 1e-7f);
		CU(cudaGetLastError());
		CU(cudaEventRecord(b, ctx->stream));
		CU(cudaEventSynchronize(b));
		float ms = 0.0f;
		CU(cudaEventElapsedTime(&ms, a, b));
		if(rep > 0 && ms < best) best = ms;
		ctx->launches++;
	}
	const double ops = (double)blocks*threads*(double)iters*8.0*8.0*2.0;   // 8 unrolled x 8 chains x (mul + add)
	*tlaneops = ops/((double)best*1e-3)/1e12;
	return PWN_OK;
}

extern "C" int pwn_stream(pwn_ctx *ctx, uint64_t *stream)
{
	if(ctx == nullptr || stream == nullptr) return fail(PWN_E_ARG, "null pointer");
	*stream = (uint64_t)(uintptr_t)ctx->stream;
	return PWN_OK;
}
