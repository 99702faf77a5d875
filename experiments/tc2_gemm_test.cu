// Round-2 probe for the fp16x3 scorer: a CTA PAIR (cluster of 2) computes D[256 x 128] = A[256 x K] * B[128 x K]^T with
// tcgen05.mma.cta_group::2.kind::f16 (fp16 operands, fp32 accumulate):
//   * A lives in TENSOR MEMORY (each CTA writes its 128 rows with tcgen05.st, two fp16 per 32-bit column),
//   * B is K-major in the no-swizzle canonical layout (8-row x 16-byte core matrices); each CTA bulk-copies HALF of the rows
//     (64) into its own shared memory and the peer's arrival is relayed to the leader with a remote mbarrier arrive,
//   * one thread of the leader CTA issues the MMAs, tcgen05.commit multicasts completion to both CTAs.
// Prints the error against a CPU reference under the candidate interpretations (which CTA supplies which half of N; which
// half of a TMEM column is the lower k) and the measured cycles per MMA.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

constexpr int M2 = 256, N = 128, K = 80;      // K multiple of 16 (fp16 UMMA_K = 16)
constexpr int NH = N / 2;                     // B rows per CTA

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
	uint32_t r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
	return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
	asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) {
	uint32_t ok = 0;
	long long spins = 0;
	while (!ok) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok)
		             : "r"(smem_u32(bar)), "r"(parity)
		             : "memory");
		if (++spins > 50000000LL) __trap();
	}
}

struct Bars {
	uint64_t load, peer, a_ready, mma;
	uint32_t tmem_slot, pad;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192) tc2_test(const __half *gA, const __half *gB, float *gD, int swap_halves,
                                                                         int reps, long long *cycles_out, float *probe_out) {
	extern __shared__ __align__(1024) unsigned char smem[];
	__half *sB = reinterpret_cast<__half *>(smem);                     // NH * K halfs, canonical layout with rows = NH
	Bars *bars = reinterpret_cast<Bars *>(sB + NH * K);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t rank = cluster_rank();

	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars->load)));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars->peer)));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&bars->a_ready)));      // 4 local + 4 remote warps
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars->mma)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 5) {
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(256u));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	cluster_sync_all();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = bars->tmem_slot;
	const uint32_t TM_A = tmem_base + 128;          // A: columns [128, 128 + K/2)

	if (warp < 4) {
		// ---- A rows -> tensor memory: thread = row = TMEM lane; column c holds k = 2c (low half) and 2c + 1 (high half)
		const int row = (int)rank * 128 + tid;
		const uint32_t lane_base = TM_A + ((uint32_t)(warp * 32) << 16);
		const uint32_t *src = reinterpret_cast<const uint32_t *>(gA + (size_t)row * K);
		for (int c0 = 0; c0 < K / 2; c0 += 8) {
			uint32_t v[8];
			for (int j = 0; j < 8; ++j) v[j] = src[c0 + j];
			asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + c0), "r"(v[0]),
			             "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
			             : "memory");
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
		asm volatile("tcgen05.fence::before_thread_sync;");
		__syncwarp();
		if (lane == 0) remote_arrive(mapa(smem_u32(&bars->a_ready), 0));
	} else if (warp == 4) {
		if (lane == 0) {
			const uint32_t bytes = NH * K * 2;
			const int half = (int)rank ^ swap_halves;
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars->load)), "r"(bytes) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)),
			             "l"(gB + (size_t)half * NH * K), "r"(bytes), "r"(smem_u32(&bars->load))
			             : "memory");
			if (rank == 1) {                 // relay: my half has landed -> tell the leader
				wait_bar(&bars->load, 0);
				remote_arrive(mapa(smem_u32(&bars->peer), 0));
			}
		}
	} else if (warp == 5 && rank == 0) {
		wait_bar(&bars->load, 0);
		wait_bar(&bars->peer, 0);
		wait_bar(&bars->a_ready, 0);
		asm volatile("tcgen05.fence::after_thread_sync;");
		// idesc: D = F32 (bits 4-5 = 1), A/B = F16 (0), K-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 for the pair)
		const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M2 >> 4) << 24);
		constexpr uint32_t lbo = (NH / 8) * 128, sbo = 128;
		const uint32_t d_hi = (uint32_t)(sbo >> 4) | (1u << 14);
		const uint32_t dB0 = ((smem_u32(sB) >> 4) & 0x3FFF) | ((lbo >> 4) << 16);
		constexpr uint32_t stepB = (2 * lbo) >> 4;
		long long t0 = clock64();
		for (int r = 0; r < reps; ++r) {
#pragma unroll
			for (int k = 0; k < K / 16; ++k) {
				const uint32_t acc = (r > 0 || k > 0) ? 1u : 0u;
				asm volatile(
				    "{\n\t.reg .b64 db;\n\t.reg .pred p, pe;\n\t"
				    "mov.b64 db, {%2, %3};\n\t"
				    "setp.ne.b32 p, %5, 0;\n\t"
				    "elect.sync _|pe, 0xffffffff;\n\t"
				    "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_base),
				    "r"(TM_A + 8 * k), "r"(dB0 + stepB * k), "r"(d_hi), "r"(idesc), "r"(acc)
				    : "memory");
			}
		}
		asm volatile(
		    "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
		    "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
		        smem_u32(&bars->mma)),
		    "h"((uint16_t)3)
		    : "memory");
		wait_bar(&bars->mma, 0);
		long long t1 = clock64();
		if (lane == 0 && cycles_out) *cycles_out = t1 - t0;
	}
	if (warp < 4) {
		wait_bar(&bars->mma, 0);
		asm volatile("tcgen05.fence::after_thread_sync;");
		const int row = (int)rank * 128 + tid;
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
			for (int j = 0; j < 32; ++j) gD[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
		}
		// three-input max and packed f32x2 ops compile check (results land in probe_out so nothing is optimised away)
		if (probe_out && tid == 0 && rank == 0) {
			float a = gD[0], b = gD[1], c = gD[2], m3;
			asm("max.f32 %0, %1, %2, %3;" : "=f"(m3) : "f"(a), "f"(b), "f"(c));
			probe_out[0] = m3;
			probe_out[1] = fmaxf(a, fmaxf(b, c));
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	cluster_sync_all();
	if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
}

// fp32 accumulation behaviour of the tensor core: D = sum_k a_k * b_k with a = [2^12, 1 + 2^-10, ...]: does the accumulator round to
// nearest, truncate, or keep extra bits?  (reported, used to judge the bias of the expanded-square form)
static void pack_half(const std::vector<__half> &src, int row0, int rows, std::vector<__half> &dst, size_t dst_off) {
	for (int r = 0; r < rows; ++r)
		for (int k = 0; k < K; ++k) {
			const int ri = r / 8, rr = r % 8, ki = k / 8, kk = k % 8;
			const size_t off = (size_t)ki * (rows / 8) * 64 + (size_t)ri * 64 + rr * 8 + kk;      // K-groups outermost, core matrix = 8 rows x 8 halfs
			dst[dst_off + off] = src[(size_t)(row0 + r) * K + k];
		}
}

int main() {
	std::vector<__half> A((size_t)M2 * K), B((size_t)N * K);
	std::vector<float> Af(A.size()), Bf(B.size());
	srand(1);
	for (size_t i = 0; i < A.size(); ++i) {
		A[i] = __float2half((float)rand() / RAND_MAX - 0.5f);
		Af[i] = __half2float(A[i]);
	}
	for (size_t i = 0; i < B.size(); ++i) {
		B[i] = __float2half((float)rand() / RAND_MAX - 0.5f);
		Bf[i] = __half2float(B[i]);
	}
	std::vector<double> ref((size_t)M2 * N);
	for (int i = 0; i < M2; ++i)
		for (int j = 0; j < N; ++j) {
			double s = 0;
			for (int k = 0; k < K; ++k) s += (double)Af[(size_t)i * K + k] * Bf[(size_t)j * K + k];
			ref[(size_t)i * N + j] = s;
		}
	std::vector<__half> pB(B.size());
	pack_half(B, 0, NH, pB, 0);
	pack_half(B, NH, NH, pB, (size_t)NH * K);
	__half *dA, *dB;
	float *dD, *dProbe;
	long long *dCyc;
	cudaMalloc(&dA, A.size() * 2);
	cudaMalloc(&dB, pB.size() * 2);
	cudaMalloc(&dD, ref.size() * 4);
	cudaMalloc(&dCyc, 8);
	cudaMalloc(&dProbe, 16);
	cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
	cudaMemcpy(dB, pB.data(), pB.size() * 2, cudaMemcpyHostToDevice);
	const size_t smem = (size_t)NH * K * 2 + sizeof(Bars) + 64;
	cudaFuncSetAttribute(tc2_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	for (int swap = 0; swap < 2; ++swap) {
		cudaMemset(dD, 0, ref.size() * 4);
		tc2_test<<<2, 192, smem>>>(dA, dB, dD, swap, 1, dCyc, dProbe);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) {
			printf("swap=%d CUDA error: %s\n", swap, cudaGetErrorString(e));
			return 1;
		}
		std::vector<float> out(ref.size());
		cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
		double e_id = 0, e_sw = 0, mr = 0;
		for (int i = 0; i < M2; ++i)
			for (int j = 0; j < N; ++j) {
				e_id = fmax(e_id, fabs(out[(size_t)i * N + j] - ref[(size_t)i * N + j]));
				e_sw = fmax(e_sw, fabs(out[(size_t)i * N + j] - ref[(size_t)i * N + ((j + NH) % N)]));
				mr = fmax(mr, fabs(ref[(size_t)i * N + j]));
			}
		printf("CTA r loads B half (r ^ %d): max|err| columns as-is %.3e, columns halves swapped %.3e (max|ref| %.3f)\n", swap, e_id, e_sw, mr);
	}
	float probe[2];
	cudaMemcpy(probe, dProbe, 8, cudaMemcpyDeviceToHost);
	printf("max3 probe: %.6f vs %.6f\n", probe[0], probe[1]);
	for (int reps : {1, 20, 200}) {
		tc2_test<<<2, 192, smem>>>(dA, dB, dD, 0, reps, dCyc, nullptr);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) {
			printf("reps=%d CUDA error: %s\n", reps, cudaGetErrorString(e));
			return 1;
		}
		long long cyc = 0;
		cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost);
		printf("reps %4d: %lld cycles for %d MMAs (M256 N128 K16, cta_group::2, A in TMEM) = %.1f cycles per MMA\n", reps, cyc, reps * (K / 16),
		       (double)cyc / (reps * (K / 16)));
	}
	return 0;
}
