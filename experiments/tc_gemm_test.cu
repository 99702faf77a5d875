// Stage-1 probe for a tcgen05 scorer: one CTA computes D[128x128] = A[128xK] * B[128xK]^T with kind::tf32,
// operands K-major in the NO-SWIZZLE canonical layout (8-row x 16-byte core matrices), loaded by 1-D bulk copies,
// accumulator in TMEM read back with tcgen05.ld.32x32b.  Prints the max error vs a CPU reference for several
// descriptor interpretations so the right one can be identified in a single run.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int M = 128, N = 128, K = 80;     // K multiple of 8 (tf32 UMMA_K = 8)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3FFF);
	d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
	d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
	d |= (uint64_t)1 << 46;          // descriptor version for sm_100
	// base_offset = 0, lbo_mode = 0, layout_type (bits 61-63) = 0: SWIZZLE_NONE
	return d;
}

__global__ void __launch_bounds__(128) tc_test(const float *gA, const float *gB, float *gD, uint32_t lboA, uint32_t sboA, uint32_t lboB,
                                              uint32_t sboB, uint32_t kstepA, uint32_t kstepB, uint32_t idesc) {
	extern __shared__ __align__(1024) unsigned char smem[];
	float *sA = reinterpret_cast<float *>(smem);                 // M*K floats
	float *sB = sA + M * K;                                      // N*K floats
	uint64_t *bar_load = reinterpret_cast<uint64_t *>(sB + N * K);
	uint64_t *bar_mma = bar_load + 1;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
	const int tid = threadIdx.x, warp = tid >> 5;

	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_load)));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_mma)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = *tmem_slot;

	if (tid == 0) {
		const uint32_t bytes = (M * K + N * K) * 4;
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar_load)), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sA)), "l"(gA),
		             "r"((uint32_t)(M * K * 4)), "r"(smem_u32(bar_load))
		             : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)), "l"(gB),
		             "r"((uint32_t)(N * K * 4)), "r"(smem_u32(bar_load))
		             : "memory");
		// wait for the operands
		uint32_t ok = 0;
		long long spins = 0;
		while (!ok) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(ok)
			             : "r"(smem_u32(bar_load)), "r"(0u)
			             : "memory");
			if (++spins > 20000000LL) __trap();
		}
		asm volatile("tcgen05.fence::after_thread_sync;");
		for (int k = 0; k < K / 8; ++k) {
			const uint64_t da = make_desc(smem_u32(sA) + k * kstepA, lboA, sboA);
			const uint64_t db = make_desc(smem_u32(sB) + k * kstepB, lboB, sboB);
			const uint32_t acc = k > 0 ? 1u : 0u;
			asm volatile(
			    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
			    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base),
			    "l"(da), "l"(db), "r"(idesc), "r"(acc)
			    : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma)) : "memory");
	}
	// everyone waits for the MMA to finish
	{
		uint32_t ok = 0;
		long long spins = 0;
		while (!ok) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(ok)
			             : "r"(smem_u32(bar_mma)), "r"(0u)
			             : "memory");
			if (++spins > 20000000LL) __trap();
		}
	}
	asm volatile("tcgen05.fence::after_thread_sync;");
	// epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows), 32 columns at a time
	for (int c0 = 0; c0 < N; c0 += 32) {
		uint32_t r[32];
		const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
		asm volatile(
		    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
		      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
		      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
		      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		    : "r"(taddr));
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		const int row = tid;
		for (int j = 0; j < 32; ++j) gD[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
}

static float to_tf32(float x) {
	uint32_t u;
	memcpy(&u, &x, 4);
	u &= ~0x1FFFu;
	memcpy(&x, &u, 4);
	return x;
}

// canonical no-swizzle K-major image: element (r, k): core matrix (ri = r/8, ki = k/4), inside: rr = r%8 rows of 16 bytes
// variant 0: address = ki*(rows/8*128) + ri*128 + rr*16 + kk*4   (K-groups outermost: LBO = rows*16, SBO = 128)
// variant 1: address = ri*(K/4*128) + ki*128 + rr*16 + kk*4      (row-groups outermost: LBO = 128, SBO = K*32)
static void pack(const std::vector<float> &src, int rows, int variant, std::vector<float> &dst) {
	dst.assign((size_t)rows * K, 0.f);
	for (int r = 0; r < rows; ++r)
		for (int k = 0; k < K; ++k) {
			const int ri = r / 8, rr = r % 8, ki = k / 4, kk = k % 4;
			size_t off = variant == 0 ? (size_t)ki * (rows / 8) * 32 + (size_t)ri * 32 + rr * 4 + kk
			                          : (size_t)ri * (K / 4) * 32 + (size_t)ki * 32 + rr * 4 + kk;
			dst[off] = src[(size_t)r * K + k];
		}
}

int main() {
	std::vector<float> A((size_t)M * K), B((size_t)N * K), ref((size_t)M * N);
	srand(1);
	for (auto &v : A) v = to_tf32((float)rand() / RAND_MAX - 0.5f);
	for (auto &v : B) v = to_tf32((float)rand() / RAND_MAX - 0.5f);
	for (int i = 0; i < M; ++i)
		for (int j = 0; j < N; ++j) {
			double s = 0;
			for (int k = 0; k < K; ++k) s += (double)A[(size_t)i * K + k] * B[(size_t)j * K + k];
			ref[(size_t)i * N + j] = (float)s;
		}
	float *dA, *dB, *dD;
	cudaMalloc(&dA, A.size() * 4);
	cudaMalloc(&dB, B.size() * 4);
	cudaMalloc(&dD, ref.size() * 4);
	const size_t smem = (size_t)(M * K + N * K) * 4 + 64;
	cudaFuncSetAttribute(tc_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	// idesc: c_format F32 (1) at bits 4-5, a/b format TF32 (2) at bits 7-9 / 10-12, K-major both, N>>3 at bit 17, M>>4 at bit 24
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
	struct Try { int variant; uint32_t lbo, sbo, kstep; const char *name; };
	const Try tries[] = {
	    {0, (uint32_t)(M / 8) * 128, 128, 2u * (M / 8) * 128, "v0: K-groups outer  LBO=rows*16 SBO=128"},
	    {0, 128, (uint32_t)(M / 8) * 128, 2u * (M / 8) * 128, "v0 with LBO/SBO swapped"},
	    {1, 128, (uint32_t)(K / 4) * 128, 256, "v1: row-groups outer LBO=128 SBO=K*32"},
	    {1, (uint32_t)(K / 4) * 128, 128, 256, "v1 with LBO/SBO swapped"},
	};
	for (const Try &t : tries) {
		std::vector<float> pA, pB, out(ref.size());
		pack(A, M, t.variant, pA);
		pack(B, N, t.variant, pB);
		cudaMemcpy(dA, pA.data(), pA.size() * 4, cudaMemcpyHostToDevice);
		cudaMemcpy(dB, pB.data(), pB.size() * 4, cudaMemcpyHostToDevice);
		cudaMemset(dD, 0, ref.size() * 4);
		tc_test<<<1, 128, smem>>>(dA, dB, dD, t.lbo, t.sbo, t.lbo, t.sbo, t.kstep, t.kstep, idesc);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) {
			printf("%-45s CUDA error: %s\n", t.name, cudaGetErrorString(e));
			return 1;
		}
		cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
		double maxerr = 0, maxref = 0;
		for (size_t i = 0; i < ref.size(); ++i) {
			maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
			maxref = fmax(maxref, fabs((double)ref[i]));
		}
		printf("%-45s max|err| = %.3e (max|ref| = %.3f)  D[0][0..3] = %.4f %.4f %.4f %.4f  ref %.4f %.4f %.4f %.4f\n", t.name, maxerr,
		       maxref, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
	}
	return 0;
}
