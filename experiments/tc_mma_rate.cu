// Probe: tcgen05.mma kind::tf32 issue/execute rate per SM as a function of the operand source (A from shared memory vs A
// from TMEM) and of N (64/128/256), with no-swizzle K-major operands like the scorer uses.  Also checks the numerics of
// the A-from-TMEM form (A written with tcgen05.st, lane = row, column = k).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_mma_rate tc_mma_rate.cu && ./tc_mma_rate
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

constexpr int M = 128, K = 80, NMAX = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int A_TMEM>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
	if (A_TMEM) {
		asm volatile(
		    "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
		    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}" ::"r"(d),
		    "r"(a_lo), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
		    : "memory");
	} else {
		asm volatile(
		    "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
		    "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d),
		    "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
		    : "memory");
	}
}

// gA: row-major [128][K] fp32 (tf32-exact values); gB: canonical no-swizzle image of [N][K]; out: D[128][N] of the LAST repetition
template <int A_TMEM>
__global__ void __launch_bounds__(128) rate_kernel(const float *gA, const float *gB, float *gD, int N, int reps, long long *cycles) {
	extern __shared__ __align__(1024) unsigned char smem[];
	float *sA = reinterpret_cast<float *>(smem);                 // canonical image of A (smem variant)
	float *sB = sA + M * K;
	uint64_t *bar_mma = reinterpret_cast<uint64_t *>(sB + NMAX * K);
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
	const int tid = threadIdx.x, warp = tid >> 5;
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_mma)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	// operands -> smem with plain stores (then made visible to the async proxy)
	for (int i = tid; i < M * K; i += 128) {
		const int r = i / K, k = i % K;
		sA[(size_t)(k / 4) * (M / 8) * 32 + (r / 8) * 32 + (r % 8) * 4 + (k % 4)] = gA[i];
	}
	for (int i = tid; i < N * K; i += 128) sB[i] = gB[i];
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = *tmem_slot;
	const uint32_t tmem_a = tmem_base + 256;
	if (A_TMEM) {
		// thread = row (lane 32*warp + lane), 8 columns per store
		for (int k0 = 0; k0 < K; k0 += 8) {
			uint32_t v[8];
			for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(gA[(size_t)tid * K + k0 + j]);
			const uint32_t taddr = tmem_a + ((uint32_t)(warp * 32) << 16) + (uint32_t)k0;
			asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
			             "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
			             : "memory");
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
		asm volatile("tcgen05.fence::before_thread_sync;");
		__syncthreads();
		asm volatile("tcgen05.fence::after_thread_sync;");
	}
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
	long long t0 = 0;
	if (tid == 0) {
		const uint32_t lbo_a = (M / 8) * 128, lbo_b = (uint32_t)(N / 8) * 128;
		const uint32_t hi = (128u >> 4) | (1u << 14);
		t0 = clock64();
		for (int r = 0; r < reps; ++r) {
			uint32_t da = A_TMEM ? tmem_a : (((smem_u32(sA) >> 4) & 0x3FFF) | ((lbo_a >> 4) << 16));
			uint32_t db = ((smem_u32(sB) >> 4) & 0x3FFF) | ((lbo_b >> 4) << 16);
#pragma unroll 2
			for (int k = 0; k < K / 8; ++k) {
				mma<A_TMEM>(tmem_base, da, hi, db, hi, idesc, k > 0 ? 1u : 0u);
				da += A_TMEM ? 8u : (2 * lbo_a) >> 4;
				db += (2 * lbo_b) >> 4;
			}
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma)) : "memory");
	}
	{
		uint32_t ok = 0;
		long long spins = 0;
		while (!ok) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(ok)
			             : "r"(smem_u32(bar_mma)), "r"(0u)
			             : "memory");
			if (++spins > 200000000LL) __trap();
		}
	}
	if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
	asm volatile("tcgen05.fence::after_thread_sync;");
	if (blockIdx.x == 0) {
		for (int c0 = 0; c0 < N; c0 += 8) {
			uint32_t r[8];
			const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
			asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
			             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
			             : "r"(taddr));
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			for (int j = 0; j < 8; ++j) gD[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

static float to_tf32(float x) {
	uint32_t u;
	memcpy(&u, &x, 4);
	u &= ~0x1FFFu;
	memcpy(&x, &u, 4);
	return x;
}

int main() {
	std::vector<float> A((size_t)M * K), B((size_t)NMAX * K);
	srand(1);
	for (auto &v : A) v = to_tf32((float)rand() / RAND_MAX - 0.5f);
	for (auto &v : B) v = to_tf32((float)rand() / RAND_MAX - 0.5f);
	float *dA, *dB, *dD;
	long long *dC;
	cudaMalloc(&dA, A.size() * 4);
	cudaMalloc(&dB, B.size() * 4);
	cudaMalloc(&dD, (size_t)M * NMAX * 4);
	cudaMalloc(&dC, 148 * 8);
	cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
	const size_t smem = (size_t)(M * K + NMAX * K) * 4 + 64;
	cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	const int reps = 400;
	for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
		for (int N : {64, 128, 256}) {
			std::vector<float> pB((size_t)N * K, 0.f), out((size_t)M * N);
			for (int r = 0; r < N; ++r)
				for (int k = 0; k < K; ++k) pB[(size_t)(k / 4) * (N / 8) * 32 + (r / 8) * 32 + (r % 8) * 4 + (k % 4)] = B[(size_t)r * K + k];
			cudaMemcpy(dB, pB.data(), pB.size() * 4, cudaMemcpyHostToDevice);
			cudaMemset(dD, 0, (size_t)M * NMAX * 4);
			for (int it = 0; it < 2; ++it) {
				if (a_tmem) rate_kernel<1><<<148, 128, smem>>>(dA, dB, dD, N, reps, dC);
				else rate_kernel<0><<<148, 128, smem>>>(dA, dB, dD, N, reps, dC);
				cudaError_t e = cudaDeviceSynchronize();
				if (e != cudaSuccess) {
					printf("A_%s N=%d CUDA error: %s\n", a_tmem ? "tmem" : "smem", N, cudaGetErrorString(e));
					return 1;
				}
			}
			long long cyc[148];
			cudaMemcpy(cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost);
			cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
			double avg = 0;
			for (int i = 0; i < 148; ++i) avg += (double)cyc[i] / 148;
			double maxerr = 0;
			for (int i = 0; i < M; ++i)
				for (int j = 0; j < N; ++j) {
					double s = 0;
					for (int k = 0; k < K; ++k) s += (double)A[(size_t)i * K + k] * B[(size_t)j * K + k];
					maxerr = fmax(maxerr, fabs(s - out[(size_t)i * N + j]));
				}
			const double per = avg / (reps * (K / 8));
			printf("A from %s  N=%3d : %.1f cycles/MMA  (tensor-bound would be %d)  -> %.0f%% of tf32 peak   max|err| %.2e\n", a_tmem ? "TMEM" : "smem",
			       N, per, N / 2, 100.0 * (N / 2) / per, maxerr);
		}
	return 0;
}
