// Probe for the tensor-core EM accumulation: D[128 x 80] = sum_f A[m][f] * B[n][f] with BOTH operands MN-MAJOR in shared memory
// (the natural production layout when a thread owns a frame f and writes its row of gamma / of [x, x^2, 1]):
//   core matrix = 8 K-rows (frames) x 16 bytes (8 consecutive MN elements); SBO = bytes between MN-adjacent core matrices,
//   LBO = bytes between K-adjacent core matrices (8 frames).  tcgen05.mma cta_group::1 kind::f16, M 128, N 80, K 16 per MMA.
// Prints the error against a CPU reference for the candidate (LBO, SBO) assignments.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int M = 128, N = 80, KF = 128;       // KF frames = reduction length

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) mn_test(const __half *gA, const __half *gB, float *gD, uint32_t lboA, uint32_t sboA, uint32_t lboB, uint32_t sboB,
                                              uint32_t kstepA, uint32_t kstepB, uint32_t idesc) {
	extern __shared__ __align__(1024) unsigned char smem[];
	__half *sA = reinterpret_cast<__half *>(smem);                 // KF * M halfs
	__half *sB = sA + KF * M;                                      // KF * N halfs
	uint64_t *bar_mma = reinterpret_cast<uint64_t *>(sB + KF * N);
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
	const int tid = threadIdx.x, warp = tid >> 5;
	for (int i = tid; i < KF * M; i += 128) sA[i] = gA[i];
	for (int i = tid; i < KF * N; i += 128) sB[i] = gB[i];
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_mma)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of the operands -> visible to the tensor core
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = *tmem_slot;
	if (tid == 0) {
		for (int k = 0; k < KF / 16; ++k) {
			uint64_t da = 0, db = 0;
			da |= (uint64_t)(((smem_u32(sA) + k * kstepA) >> 4) & 0x3FFF) | ((uint64_t)((lboA >> 4) & 0x3FFF) << 16) | ((uint64_t)((sboA >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
			db |= (uint64_t)(((smem_u32(sB) + k * kstepB) >> 4) & 0x3FFF) | ((uint64_t)((lboB >> 4) & 0x3FFF) << 16) | ((uint64_t)((sboB >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
			const uint32_t acc = k > 0 ? 1u : 0u;
			asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base), "l"(da),
			             "l"(db), "r"(idesc), "r"(acc)
			             : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma)) : "memory");
	}
	{
		uint32_t ok = 0;
		long long spins = 0;
		while (!ok) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar_mma)), "r"(0u) : "memory");
			if (++spins > 20000000LL) __trap();
		}
	}
	asm volatile("tcgen05.fence::after_thread_sync;");
	for (int c0 = 0; c0 < N; c0 += 16) {
		uint32_t r[16];
		const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
		asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
		               "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
		             : "r"(taddr));
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		for (int j = 0; j < 16; ++j) gD[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
}

// MN-major canonical image of a [rows = MN][KF] operand: element (mn, f) at kg * lbo + mg * sbo + (f & 7) * 16 + (mn & 7) * 2 bytes
static void pack_mn(const std::vector<__half> &src, int rows, uint32_t lbo, uint32_t sbo, std::vector<__half> &dst) {
	dst.assign((size_t)rows * KF, __float2half(0.f));
	for (int mn = 0; mn < rows; ++mn)
		for (int f = 0; f < KF; ++f) {
			const size_t byte = (size_t)(f >> 3) * lbo + (size_t)(mn >> 3) * sbo + (size_t)(f & 7) * 16 + (size_t)(mn & 7) * 2;
			dst[byte / 2] = src[(size_t)mn * KF + f];
		}
}

int main() {
	std::vector<__half> A((size_t)M * KF), B((size_t)N * KF);
	std::vector<float> Af(A.size()), Bf(B.size());
	srand(3);
	for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half((float)rand() / RAND_MAX - 0.5f); Af[i] = __half2float(A[i]); }
	for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half((float)rand() / RAND_MAX - 0.5f); Bf[i] = __half2float(B[i]); }
	std::vector<double> ref((size_t)M * N);
	for (int m = 0; m < M; ++m)
		for (int n = 0; n < N; ++n) {
			double s = 0;
			for (int f = 0; f < KF; ++f) s += (double)Af[(size_t)m * KF + f] * Bf[(size_t)n * KF + f];
			ref[(size_t)m * N + n] = s;
		}
	__half *dA, *dB;
	float *dD;
	cudaMalloc(&dA, A.size() * 2);
	cudaMalloc(&dB, B.size() * 2);
	cudaMalloc(&dD, ref.size() * 4);
	const size_t smem = (size_t)(KF * M + KF * N) * 2 + 64;
	cudaFuncSetAttribute(mn_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	// idesc: D = F32, A/B = F16, BOTH MN-major (bits 15 and 16), N >> 3 at bit 17, M >> 4 at bit 24
	const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
	// storage: K-groups (8 frames) outermost: per K-group, MN-adjacent core matrices are contiguous (128 B apart)
	const uint32_t stoLboA = (M / 8) * 128, stoLboB = (N / 8) * 128, stoSbo = 128;
	struct Try { uint32_t lboA, sboA, lboB, sboB; const char *name; } tries[] = {
	    {stoLboA, stoSbo, stoLboB, stoSbo, "descriptor LBO = K-group stride, SBO = MN-group stride (128)"},
	    {stoSbo, stoLboA, stoSbo, stoLboB, "descriptor LBO / SBO swapped"},
	};
	for (const Try &t : tries) {
		std::vector<__half> pA, pB;
		pack_mn(A, M, stoLboA, stoSbo, pA);
		pack_mn(B, N, stoLboB, stoSbo, pB);
		cudaMemcpy(dA, pA.data(), pA.size() * 2, cudaMemcpyHostToDevice);
		cudaMemcpy(dB, pB.data(), pB.size() * 2, cudaMemcpyHostToDevice);
		cudaMemset(dD, 0, ref.size() * 4);
		mn_test<<<1, 128, smem>>>(dA, dB, dD, t.lboA, t.sboA, t.lboB, t.sboB, 2 * stoLboA, 2 * stoLboB, idesc);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) {
			printf("%-70s CUDA error: %s\n", t.name, cudaGetErrorString(e));
			return 1;
		}
		std::vector<float> out(ref.size());
		cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
		double maxerr = 0, maxref = 0;
		for (size_t i = 0; i < ref.size(); ++i) {
			maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
			maxref = fmax(maxref, fabs(ref[i]));
		}
		printf("%-70s max|err| = %.3e (max|ref| = %.3f)\n", t.name, maxerr, maxref);
	}
	return 0;
}
