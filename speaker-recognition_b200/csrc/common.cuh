// common.cuh -- shared host/device helpers for libsrb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

namespace srb {

// ---- error state (thread-local message behind srb200_last_error) ----
void set_error(const char *fmt, ...);
void clear_error();

#define SRB_CUDA(call)                                                                             \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess) {                                                                   \
			srb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			return 1;                                                                              \
		}                                                                                          \
	} while (0)

#define SRB_REQUIRE(cond, ...)          \
	do {                                \
		if (!(cond)) {                  \
			srb::set_error(__VA_ARGS__); \
			return 1;                   \
		}                               \
	} while (0)

// ---- process-wide context: one device, one library stream, grow-only workspaces ----
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes);   // grow-only cudaMalloc
	void release();
	template <class T> T *as() { return (T *)p; }
};
struct PinBuf {                 // pinned host staging
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes);
	void release();
	template <class T> T *as() { return (T *)p; }
};

struct Context {
	std::recursive_mutex mu;
	int device = 0;
	bool inited = false;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	int compat_clamp = 1;
	int score_path = -1;        // 0: fp32 FMA kernel, 1: tcgen05 3xTF32 kernel; -1: default (tensor-core route unless SRB200_SCORE_TC=0)
	int last_score_route = -1;  // route the most recent scoring launch took (0 / 1)
	int em_path = -1;           // -1 auto, 0 FMA kernels only, 1 tensor cores wherever the shape allows (tests)
	int last_em_route = -1;     // route the most recent E-step took (0 = FMA kernels, 1 = tensor cores)
	long long score_launches = 0;   // scoring launches so far (tests: "one launch for all speakers and the UBM")
	// workspaces for host entry points (guarded by mu)
	DevBuf d_x, d_out, d_tiles, d_partial, d_sums, d_misc, d_pcm, d_ceps, d_feat, d_stats, d_frame;
	PinBuf h_x, h_out, h_misc, h_pcm;
	int init();                 // lazily binds the device, creates the stream; fails loudly without a GPU
};
Context &ctx();

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// ---- device-side PTX wrappers (mbarrier + 1-D bulk async copy = TMA engine, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	while (!mbar_try_wait(bar, parity)) {
	}
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, both 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
	float y;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
	float y;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
#endif

}  // namespace srb
