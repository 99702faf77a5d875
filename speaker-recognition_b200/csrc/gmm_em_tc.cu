// gmm_em_tc.cu -- the EM E-step of large models on the tensor cores (sm_100a, tcgen05).
//
// gmm_em.cu recomputes the exponents and accumulates the sufficient statistics on the fp32 FMA pipe (0.43 of its peak: 7.9 ms
// per iteration at BASELINE config 3, K = 512, N = 1e6, D = 39).  Both halves of that work are GEMMs:
//   pass A   l2[t][k] = PHI[t] . W[k]      the expanded-square scores of gmm_score_tc2.cuh  -> per-frame log2 p(x_t), total LL
//   pass B   gamma[t][k] = 2^(l2[t][k] - lse2[t]);   S[k][:] = sum_t gamma[t][k] * [x''_t^2, x''_t, 1]      (GEMM over the FRAMES)
// Here the fp16 head/tail operand images of the CURRENT parameters are built on the device from the float64 master
// (em_tc_pack_kernel: no host round trip inside an EM iteration), pass A is the tcgen05 scorer reading those images, and pass B is
// em_accumulate_tc_kernel: per (128-mixture block, frame range) a CTA pipelines, per tile of 64 frames,
//     GEMM 1  L^T[128 mixtures x 64 frames] = W (tensor memory, written once per CTA) . PHI^T (shared-memory ring, K-major)
//     gamma   thread = mixture row: gamma = 2^(L^T - lse2), fp16 head + tail, written back to tensor memory
//     GEMM 2  ACC[128 mixtures x Kp] += gamma (tensor memory) . Y,  Y = the same shared-memory tile read MN-major   -- 3 passes each
// with ACC in tensor memory, folded into float64 running sums (registers) every 1024 frames; em_tc_reduce_kernel sums the splits and maps
// the centred / scaled moments back to sum gamma x, sum gamma x^2.  Exactly corrected entries (gmm_score_tc2's add-back) are NOT
// used here: the packer computes max_k sum_d nb''^2 and raises a flag when the expanded form would not be accurate enough
// (collapsed mixtures far from the centre) -- the caller then runs the FMA kernels of gmm_em.cu for that iteration.
#include <cuda_fp16.h>
#include <algorithm>
#include "gmm_score_tc_layout.h"
#include "kernels.h"

namespace srb {

constexpr double ET_SAFE_NB2 = TC_HARD_BUDGET;      // max over rows of sum nb''^2 the uncorrected form is trusted with: the scorer's own budget
constexpr int ET_FT = 64;                   // frames per tile of pass B
constexpr int ET_NS = 4;                    // PHI stages in shared memory
constexpr int ET_NX = 4;                    // raw frame-tile buffers (bulk-copy targets), loaded three tiles ahead
constexpr int ET_THREADS = 352;             // warp 0: MMA issue; warps 1-10 build PHI; warps 1-8 also turn L^T into gamma (TMEM lane quarter = warp % 4, frame half = (warp - 1) / 4)
constexpr int ET_WARP_PROD0 = 9;            // warps 9, 10: PHI only
constexpr int ET_BUILDERS = 320;            // threads of warps 1-10: one (frame, 16-column chunk) task each at Kp = 80
constexpr int ET_FLUSH_TILES = 16;          // tiles (of 64 frames) between float64 flushes of the accumulator: 1024 frames
// tensor memory: W (A operand of GEMM 1: head Kp/2 | tail Kp/2 columns) | L^T accumulators [2][64] | gamma (A operand of GEMM 2) [2][head 32 | tail 32] | ACC [Kp]
constexpr int ET_W_COL = 0, ET_L_COL = 128, ET_G_COL = 256, ET_ACC_COL = 384;

// Workspace of the route: operand images of the model being trained (rebuilt from its float64 master every iteration) and the
// float64 partials.  One per process, grow-only (DevBuf), shared by every model that trains: an E-step call builds, uses and is
// done with it before it returns to its caller's stream order, and the C ABI is synchronous per context.
struct EmTcWorkspace {
	DevBuf cs;                              // [2][D] centre, scale
	DevBuf img2;                            // scorer layout: [blocks][2 ranks][head | tail][64 * Kp]
	DevBuf img128;                          // this file's GEMM 1: [blocks][head | tail][128 * Kp], 128-row K-major canonical layout
	DevBuf hard;                            // [blocks][TC_HSTRIDE]: headers only (valid-column masks), no entries
	DevBuf flag;                            // != 0: do not use the tensor route for these parameters
	DevBuf partials;                        // [n_split][blocks][128][Kp] float64
	PinBuf h_flag;
};
static EmTcWorkspace g_ws;

void em_tc_release() {
	g_ws.cs.release();
	g_ws.img2.release();
	g_ws.img128.release();
	g_ws.hard.release();
	g_ws.flag.release();
	g_ws.partials.release();
	g_ws.h_flag.release();
}

// ---------------------------------------------------------------------------------------------------------------------------
// operand images from the float64 master: params = w[K] | mean[K*D] | sigma[K*D]
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void em_tc_cs_kernel(const double *__restrict__ params, int K, int D, float *__restrict__ cs, int *__restrict__ flag) {
	// one block per dimension: centre = weighted mean of the means, scale = 2^-round(log2 of the overall spread) (pack_models_tc_host)
	__shared__ double sh[3][128];
	const int d = blockIdx.x, tid = threadIdx.x;
	if (d == 0 && tid == 0) *flag = 0;
	const double *w = params, *mu = params + K, *sg = params + K + (size_t)K * D;
	auto block_sum = [&](int slot, double v) {
		sh[slot][tid] = v;
		__syncthreads();
		for (int s = 64; s > 0; s >>= 1) {
			if (tid < s) sh[slot][tid] += sh[slot][tid + s];
			__syncthreads();
		}
		return sh[slot][0];
	};
	double ws = 0.0, cw = 0.0;
	for (int k = tid; k < K; k += 128) {
		const double wk = w[k] > 0.0 ? w[k] : 0.0;
		ws += wk;
		cw += wk * mu[(size_t)k * D + d];
	}
	double wsum = block_sum(0, ws);
	double c = block_sum(1, cw);
	if (!(wsum > 0.0)) wsum = 1.0;
	c = (double)(float)(c / wsum);                           // the kernels subtract the fp32 value
	double sp = 0.0;
	for (int k = tid; k < K; k += 128) {
		const double wk = w[k] > 0.0 ? w[k] : 0.0;
		const double dm = mu[(size_t)k * D + d] - c, s = sg[(size_t)k * D + d];
		sp += wk * (dm * dm + s * s);
	}
	const double spread = block_sum(2, sp);
	if (tid == 0) {
		const double sd = sqrt(spread / wsum);
		int e = (sd > 0.0 && sd == sd) ? (int)lrint(-log2(sd)) : 0;
		e = max(-60, min(60, e));
		cs[d] = (float)c;
		cs[D + d] = (float)ldexp(1.0, e);
	}
}

// k: the scorer's column (x''^2 block, x'' block, constant); k128: this file's own column order (see em_accumulate_tc_kernel)
__device__ __forceinline__ void et_put(__half *img2, __half *img128, int Kp, int blk, int col, int k, int k128, double w) {
	const double wc = fmax(-T2_W_MAX, fmin(T2_W_MAX, w));
	const __half h = __float2half_rn((float)wc);
	const __half l = __float2half_rn((float)(wc - (double)__half2float(h)));
	const size_t b2 = (size_t)blk * t2_block_halfs(Kp) + (size_t)(col >> 6) * 2 * T2_NH * Kp + t2_img_off(col & 63, k);
	img2[b2] = h;
	img2[b2 + (size_t)T2_NH * Kp] = l;
	const size_t b1 = ((size_t)blk * TC_NB + col) * 2 * Kp + k128;      // pass B: row-major [block][mixture][head Kp | tail Kp]
	img128[b1] = h;
	img128[b1 + Kp] = l;
}

// one warp per mixture row (incl. the padding rows of the last block): lanes stride over the dimensions
__global__ void em_tc_pack_kernel(const double *__restrict__ params, int K, int D, int Kp, int blocks, const float *__restrict__ cs, __half *__restrict__ img2,
                                  __half *__restrict__ img128, int4 *__restrict__ hard, int *__restrict__ flag) {
	const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (row >= blocks * TC_NB) return;
	const int blk = row / TC_NB, col = row % TC_NB;
	const double A0 = 0.84932180028801904272;
	if (row >= K) {      // padding: its column is masked out by the valid mask; keep the images finite
		for (int k = lane; k < Kp; k += 32) et_put(img2, img128, Kp, blk, col, k, k, k == 2 * D ? -T2_W_MAX : 0.0);
		return;
	}
	const double *w = params, *mu = params + K + (size_t)row * D, *sg = params + K + (size_t)K * D + (size_t)row * D;
	double nb2 = 0.0, lg = 0.0;
	bool bad = false;
	for (int d = lane; d < D; d += 32) {
		const double c = (double)cs[d], al = (double)cs[D + d];
		const double a = A0 / (sg[d] * al), nb = -(mu[d] - c) * al * a;
		et_put(img2, img128, Kp, blk, col, d, 2 * d, -a * a);
		et_put(img2, img128, Kp, blk, col, D + d, 2 * d + 1, -2.0 * a * nb);
		nb2 += nb * nb;
		lg += log2(2.5066282746310002 * sg[d]);
		bad |= (a * a > T2_W_MAX) || (fabs(2.0 * a * nb) > T2_W_MAX) || !(sg[d] > 0.0);
	}
	for (int k = 2 * D + 1 + lane; k < Kp; k += 32) et_put(img2, img128, Kp, blk, col, k, k, 0.0);
#pragma unroll
	for (int off = 16; off >= 1; off >>= 1) {
		nb2 += __shfl_xor_sync(0xffffffffu, nb2, off);
		lg += __shfl_xor_sync(0xffffffffu, lg, off);
	}
	bad = __any_sync(0xffffffffu, bad);
	if (lane == 0) {
		double C = (w[row] > 0.0 ? log2(w[row]) : -1.0e30) - lg - nb2;
		if (C < -1.0e30 || C != C) C = -1.0e30;
		et_put(img2, img128, Kp, blk, col, 2 * D, 2 * D, C);
		atomicOr(reinterpret_cast<int *>(hard + (size_t)blk * TC_HSTRIDE + (col >> 5)) + 3, (int)(1u << (col & 31)));      // header word g, .w: valid columns
		if (bad || !(nb2 <= ET_SAFE_NB2)) atomicExch(flag, 1);
	}
}

// ---------------------------------------------------------------------------------------------------------------------------
// pass B
// ---------------------------------------------------------------------------------------------------------------------------
struct EtArgs {
	const float *X;
	int ld, D, n, Kp;
	const float *cs;
	const __half *img128;      // [blocks][128 mixtures][head Kp | tail Kp], this file's column order
	const float *lse2;         // [n] log2 p(x_t) from pass A
	double *partials;          // [n_split][blocks][128][Kp]
	int blocks, n_split, clamp;
};

namespace et {
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity) {
	long long spins = 0;
	while (!mbar_try_wait(bar, parity)) {
		if (++spins > 400000000LL) __trap();
	}
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&v)[8]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
	             "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
	             : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
	asm volatile(
	    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
	    "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
	    "r"(v[13]), "r"(v[14]), "r"(v[15])
	    : "memory");
}
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&r)[8]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
	             : "r"(taddr));
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
	      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
	      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
	      "=r"(r[31])
	    : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]^T, single CTA.  These three run inside a region that ONE elected lane executes (one election per
// GEMM of a tile, not per instruction): the operands then stay in uniform registers and the MMAs issue back to back -- the
// per-instruction form (elect + 2 VOTEU + 4 R2UR + 2 UMOV + UTCHMMA) cost ten instructions and a dependent chain per MMA.
template <int ACC>
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
	asm volatile(
	    "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
	    "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
	    : "memory");
}
// the same with the accumulate flag as a run-time predicate
__device__ __forceinline__ void mma_ts_p(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
	asm volatile(
	    "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
	    "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
	    : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
	uint32_t ok;
	asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(ok));
	return ok != 0;
}
__device__ __forceinline__ void split_pack(float v0, float v1, uint32_t &hi, uint32_t &lo) {      // v0 in the low half
	const __half2 h = __floats2half2_rn(v0, v1);
	const float2 hf = __half22float2(h);
	const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
	hi = *reinterpret_cast<const uint32_t *>(&h);
	lo = *reinterpret_cast<const uint32_t *>(&l);
}
}  // namespace et

// shared memory: ET_NS PHI stages [head | tail][64 frames x Kp] + ET_NX raw frame tiles [64][ld] (+32 bytes: the last chunk of the last
// row may read up to two 16-byte units past it) + lse2[ET_NS][64] + centre/scale [2][Kp/2] + barriers / TMEM slot
static size_t et_smem_bytes(int Kp, int ld) {
	return (size_t)ET_NS * 2 * ET_FT * Kp * 2 + (size_t)ET_NX * ((size_t)ET_FT * ld * 4 + 32) + (size_t)ET_NS * ET_FT * 4 + (size_t)Kp * 4 + 512;
}

// Pass B, transposed so that two frame tiles are in flight and nothing but PHI goes through shared memory:
//   GEMM 1   L^T[mixture][frame] = W . PHI^T     A = W in tensor memory (written once per CTA), B = the PHI tile [64 frames][Kp], K-major
//   gamma    thread = mixture row: gamma = 2^(L^T - lse2[frame]) -> fp16 head + tail -> tensor memory (the A operand of GEMM 2)
//   GEMM 2   ACC[mixture][col] += gamma . Y      B = the SAME shared-memory tile read MN-major (Y[frame][col] = PHI[frame][col]:
//            the 8 x 16-byte core matrices of the K-major PHI tile are the core matrices of the MN-major Y tile; only the
//            descriptor's strides differ: LBO = 128 (8 frames on), SBO = 1024 (8 columns on))
// Column order of PHI / W (any order works as long as the operands agree): k = 2 d -> x''_d^2, k = 2 d + 1 -> x''_d, k = 2 D -> 1,
// rest 0 -- four dimensions (one 16-byte read of the frame row) make one 16-byte piece of the operand row.
// Roles: warp 0 issues the MMAs in the order G1(t+1), G2(t); warps 1-8 turn L^T into gamma (double-buffered in tensor memory on both
// sides) and keep the float64 running sums; warps 9-10 build PHI (thread = frame) from raw frame tiles that arrive by bulk copy.
template <int NCH>      // Kp / 16
__global__ void __launch_bounds__(ET_THREADS, 1) em_accumulate_tc_kernel(const EtArgs A) {
	extern __shared__ __align__(1024) unsigned char et_smem[];
	constexpr int Kp = NCH * 16;
	constexpr uint32_t p_part = (uint32_t)ET_FT * Kp * 2, p_stage = 2 * p_part;      // bytes: one PHI part (head or tail), one stage
	const int D = A.D, ld = A.ld;
	const uint32_t x_buf = (uint32_t)ET_FT * ld * 4 + 32;
	unsigned char *sP = et_smem;                                 // [ET_NS][head | tail] (k >> 3) * 1024 + (f >> 3) * 128 + (f & 7) * 16 + (k & 7) * 2
	unsigned char *sXb = sP + ET_NS * p_stage;                   // [ET_NX] raw frame tiles as they lie in HBM
	float *sL = reinterpret_cast<float *>(sXb + ET_NX * x_buf);  // [ET_NS][64] lse2 of the tile's frames (+inf: the frame contributes nothing)
	float *sC = sL + ET_NS * ET_FT;                              // [2][Kp / 2] alpha_d, -c_d alpha_d (zero-padded)
	uint64_t *bars = reinterpret_cast<uint64_t *>(sC + Kp);
	uint64_t *x_full = bars, *x_empty = bars + 4, *phi_full = bars + 8, *phi_empty = bars + 12, *l_full = bars + 16, *l_empty = bars + 18,
	         *g_full = bars + 20, *g_empty = bars + 22;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 26);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int mb = blockIdx.x, split = blockIdx.y;
	for (int i = tid; i < Kp; i += ET_THREADS) {             // x'' = x * alpha + (-c * alpha); both 0 beyond D
		const int d = i % (Kp / 2);
		const float al = d < D ? A.cs[D + d] : 0.f, ce = d < D ? A.cs[d] : 0.f;
		sC[i] = i < Kp / 2 ? al : -ce * al;
	}
	if (tid == 0) {
		for (int i = 0; i < ET_NX; ++i) mbar_init(&x_full[i], 1), mbar_init(&x_empty[i], 10);
		for (int i = 0; i < ET_NS; ++i) mbar_init(&phi_full[i], 10), mbar_init(&phi_empty[i], 1);
		for (int i = 0; i < 2; ++i) mbar_init(&l_full[i], 1), mbar_init(&l_empty[i], 8), mbar_init(&g_full[i], 8), mbar_init(&g_empty[i], 1);
		fence_mbar_init();
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	et::fence_before();
	__syncthreads();
	et::fence_after();
	const uint32_t tmem_base = *tmem_slot;

	const int n_tiles = (A.n + ET_FT - 1) / ET_FT;
	const int per = (n_tiles + A.n_split - 1) / A.n_split;
	const int t0 = min(n_tiles, split * per), t1 = min(n_tiles, t0 + per);
	const int nt = t1 - t0;

	// ---- W of this block into tensor memory: lane = mixture row, the two warps of a lane quarter write head and tail ----
	if (warp >= 1 && warp <= 8) {
		const int q = warp & 3, h = (warp - 1) >> 2, row = q * 32 + lane;
		const uint4 *src = reinterpret_cast<const uint4 *>(A.img128 + (((size_t)mb * TC_NB + row) * 2 + h) * Kp);
#pragma unroll
		for (int c = 0; c < NCH; ++c) {
			const uint4 a = __ldg(src + 2 * c), b = __ldg(src + 2 * c + 1);
			const uint32_t r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
			et::st8(tmem_base + ((uint32_t)(q * 32) << 16) + ET_W_COL + h * (Kp / 2) + 8 * c, r);
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	et::fence_before();
	__syncthreads();
	et::fence_after();

	if (warp == 0) {
		// ======================= MMA issue =======================
		constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(ET_FT >> 3) << 17) | ((uint32_t)(TC_NB >> 4) << 24);                  // A (TMEM) and B K-major, N = 64
		constexpr uint32_t idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(Kp >> 3) << 17) | ((uint32_t)(TC_NB >> 4) << 24);        // B MN-major, N = Kp
		constexpr uint32_t d1_hi = (uint32_t)(128 >> 4) | (1u << 14), d1_lbo = (uint32_t)(1024 >> 4) << 16;                       // K-major: LBO = K-group stride, SBO = 8 frames on
		constexpr uint32_t d2_hi = (uint32_t)(1024 >> 4) | (1u << 14), d2_lbo = (uint32_t)(128 >> 4) << 16;                       // MN-major: LBO = K-group (8 frames), SBO = 8 columns on
		const uint32_t w_hi = tmem_base + ET_W_COL, w_lo = w_hi + Kp / 2;
		for (int i = 0; i <= nt; ++i) {
			if (i < nt) {
				const int s = i % ET_NS, b = i & 1;
				et::wait(&phi_full[s], (uint32_t)(i / ET_NS) & 1);
				et::wait(&l_empty[b], ((uint32_t)(i >> 1) & 1) ^ 1);
				et::fence_after();
				const uint32_t d1 = tmem_base + ET_L_COL + 64 * b;
				const uint32_t pb = ((smem_u32(sP + (size_t)s * p_stage) >> 4) & 0x3FFF) | d1_lbo;
				if (et::elect_one()) {
					et::mma_ts<0>(d1, w_hi, pb, d1_hi, idesc1);
#pragma unroll
					for (int k = 1; k < NCH; ++k) et::mma_ts<1>(d1, w_hi + 8 * k, pb + 128 * k, d1_hi, idesc1);
#pragma unroll
					for (int k = 0; k < NCH; ++k) et::mma_ts<1>(d1, w_lo + 8 * k, pb + 128 * k, d1_hi, idesc1);
#pragma unroll
					for (int k = 0; k < NCH; ++k) et::mma_ts<1>(d1, w_hi + 8 * k, pb + (p_part >> 4) + 128 * k, d1_hi, idesc1);
					et::commit(&l_full[b]);
				}
				__syncwarp();
			}
			if (i > 0) {
				const int j = i - 1, s = j % ET_NS, b = j & 1;
				et::wait(&g_full[b], (uint32_t)(j >> 1) & 1);            // gamma of tile j is in tensor memory (and a due flush has read ACC)
				et::fence_after();
				const uint32_t d2 = tmem_base + ET_ACC_COL, gh = tmem_base + ET_G_COL + 64 * b, gt = gh + 32;
				const uint32_t yb = ((smem_u32(sP + (size_t)s * p_stage) >> 4) & 0x3FFF) | d2_lbo;
				uint32_t acc = (j % ET_FLUSH_TILES) == 0 ? 0u : 1u;      // first GEMM 2 after a flush overwrites the accumulator
				if (et::elect_one()) {
#pragma unroll
					for (int k = 0; k < ET_FT / 16; ++k) {               // gamma head x Y head, gamma tail x Y head, gamma head x Y tail
						et::mma_ts_p(d2, gh + 8 * k, yb + 16 * k, d2_hi, idesc2, acc);
						acc = 1u;
						et::mma_ts<1>(d2, gt + 8 * k, yb + 16 * k, d2_hi, idesc2);
						et::mma_ts<1>(d2, gh + 8 * k, yb + (p_part >> 4) + 16 * k, d2_hi, idesc2);
					}
					et::commit(&g_empty[b]);
					et::commit(&phi_empty[s]);
				}
				__syncwarp();
			}
		}
	} else {
		// ======================= warps 1-10: PHI tiles (all), gamma (warps 1-8, thread = mixture row) =======================
		const int q = warp & 3, h = (warp - 1) >> 2, row = q * 32 + lane;
		const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
		const int btid = tid - 32;                               // 0..319
		const bool issuer = btid == 0;
		const bool gamma_warp = warp < ET_WARP_PROD0;
		auto load_x = [&](int t, int xb) {      // the tile's frames: rows are contiguous in HBM ([n][ld]), so the tile is ONE bulk copy
			const uint32_t bytes = (uint32_t)min(ET_FT, A.n - t * ET_FT) * (uint32_t)ld * 4u;
			mbar_expect_tx(&x_full[xb], bytes);
			bulk_g2s(sXb + (size_t)xb * x_buf, A.X + (size_t)t * ET_FT * ld, bytes, &x_full[xb]);
		};
		if (issuer)
			for (int j = 0; j < min(nt, ET_NX - 1); ++j) load_x(t0 + j, j);
		__syncwarp();
		const float4 *cen4 = reinterpret_cast<const float4 *>(sC), *nca4 = reinterpret_cast<const float4 *>(sC + Kp / 2);
		// PHI of tile j: a task = one frame x one chunk of 16 operand columns (8 dimensions: two 16-byte reads of the frame row, two
		// 16-byte pieces of the head row and two of the tail row); tasks are dealt to the 320 threads, frames fastest
		auto build = [&](int j) {
			const int t = t0 + j, s = j % ET_NS, xb = j % ET_NX;
			et::wait(&x_full[xb], (uint32_t)(j / ET_NX) & 1);
			et::wait(&phi_empty[s], ((uint32_t)(j / ET_NS) & 1) ^ 1);        // GEMM 2 of tile j - ET_NS has read this stage
			for (int task = btid; task < ET_FT * NCH; task += ET_BUILDERS) {
				const int f = task & (ET_FT - 1), c = task >> 6;
				const int fr = t * ET_FT + f;
				const bool valid = fr < A.n;
				const float4 *xr = reinterpret_cast<const float4 *>(sXb + (size_t)xb * x_buf) + (size_t)f * (ld >> 2);
				uint32_t hi[8], lo[8];
#pragma unroll
				for (int u = 0; u < 2; ++u) {
					const float4 xv = valid ? xr[2 * c + u] : make_float4(0.f, 0.f, 0.f, 0.f);
					const float4 av = cen4[2 * c + u], bv = nca4[2 * c + u];      // x'' = x * alpha - c * alpha
					const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, as4[4] = {av.x, av.y, av.z, av.w}, bs4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
					for (int jj = 0; jj < 4; ++jj) {
						const int d = 8 * c + 4 * u + jj;
						const float xin = d < D ? xs[jj] : 0.f;      // row padding / the next row's data: never enters the arithmetic (it may be NaN)
						float v1 = fminf(fmaxf(fmaf(xin, as4[jj], bs4[jj]), -T2_X_CLAMP), T2_X_CLAMP);      // 0 beyond D (alpha = 0 there)
						float v0 = v1 * v1;
						if (d == D) v0 = 1.f;                      // the constant column; columns beyond stay (0, 0)
						if (!valid) v0 = 0.f, v1 = 0.f;
						et::split_pack(v0, v1, hi[4 * u + jj], lo[4 * u + jj]);
					}
				}
				unsigned char *ph = sP + (size_t)s * p_stage + (size_t)(2 * c) * 1024 + (size_t)(f >> 3) * 128 + (f & 7) * 16;
				*reinterpret_cast<uint4 *>(ph) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
				*reinterpret_cast<uint4 *>(ph + 1024) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
				*reinterpret_cast<uint4 *>(ph + p_part) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
				*reinterpret_cast<uint4 *>(ph + p_part + 1024) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
				if (c == 0) {
					float l2 = __int_as_float(0x7f800000);           // +inf -> gamma = 2^(-inf) = 0
					if (valid) {
						l2 = __ldg(A.lse2 + fr);
						if (A.clamp && !(l2 * TC_LN2 >= -744.4400719213812f)) l2 = __int_as_float(0x7f800000);      // gmm.cc:482-490: underflowed frame
					}
					sL[s * ET_FT + f] = l2;
				}
			}
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the PHI stores become visible to the tensor core
			__syncwarp();
			if (lane == 0) {
				mbar_arrive(&phi_full[s]);
				mbar_arrive(&x_empty[xb]);
			}
			if (issuer && j + ET_NX - 1 < nt) {              // tile j + 3 goes into the buffer tile j - 1 used
				if (j >= 1) et::wait(&x_empty[(j - 1) % ET_NX], (uint32_t)((j - 1) / ET_NX) & 1);
				load_x(t + ET_NX - 1, (j + ET_NX - 1) % ET_NX);
			}
			__syncwarp();
		};
		// float64 running sums of this thread's part of the statistics: mixture `row` of the block, its half of the Kp columns.  The
		// tensor-memory accumulator is fp32 and is folded in here every ET_FLUSH_TILES tiles (1024 frames)
		double acc64[Kp / 2];
#pragma unroll
		for (int j = 0; j < Kp / 2; ++j) acc64[j] = 0.0;
		auto flush = [&]() {
#pragma unroll
			for (int c = 0; c < Kp / 2; c += 8) {
				uint32_t r[8];
				et::ld8(lane_base + ET_ACC_COL + h * (Kp / 2) + c, r);
				et::ld_wait();
#pragma unroll
				for (int j = 0; j < 8; ++j) acc64[c + j] += (double)__uint_as_float(r[j]);
			}
		};
		for (int j = 0; j < min(nt, 2); ++j) build(j);
		for (int i = 0; i < nt; ++i) {
			if (i + 2 < nt) build(i + 2);                        // two tiles ahead of the gamma stage: GEMM 1 of tile i + 1 runs under gamma of tile i
			if (!gamma_warp) continue;
			const int s = i % ET_NS, b = i & 1;
			et::wait(&l_full[b], (uint32_t)(i >> 1) & 1);
			et::fence_after();
			uint32_t v[32];
			et::ld32(lane_base + ET_L_COL + 64 * b + 32 * h, v);
			et::ld_wait();
			et::fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&l_empty[b]);                  // GEMM 1 of tile i + 2 may overwrite this buffer
			const float4 *l4 = reinterpret_cast<const float4 *>(sL + s * ET_FT + 32 * h);
			uint32_t hi[16], lo[16];
#pragma unroll
			for (int j4 = 0; j4 < 8; ++j4) {
				const float4 ls = l4[j4];                             // the same address in every lane: a broadcast read
				const float g0 = ex2_approx(__uint_as_float(v[4 * j4]) - ls.x), g1 = ex2_approx(__uint_as_float(v[4 * j4 + 1]) - ls.y);
				const float g2 = ex2_approx(__uint_as_float(v[4 * j4 + 2]) - ls.z), g3 = ex2_approx(__uint_as_float(v[4 * j4 + 3]) - ls.w);
				et::split_pack(g0, g1, hi[2 * j4], lo[2 * j4]);
				et::split_pack(g2, g3, hi[2 * j4 + 1], lo[2 * j4 + 1]);
			}
			if (i > 0 && (i % ET_FLUSH_TILES) == 0) {                 // GEMM 2 of tile i - 1 complete -> fold the accumulator before GEMM 2 of tile i overwrites it
				et::wait(&g_empty[(i - 1) & 1], (uint32_t)((i - 1) >> 1) & 1);
				et::fence_after();
				flush();
			}
			et::wait(&g_empty[b], ((uint32_t)(i >> 1) & 1) ^ 1);      // GEMM 2 of tile i - 2 has read this gamma buffer
			et::fence_after();
			et::st16(lane_base + ET_G_COL + 64 * b + 16 * h, hi);
			et::st16(lane_base + ET_G_COL + 64 * b + 32 + 16 * h, lo);
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
			et::fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&g_full[b]);
		}
		if (gamma_warp) {
			if (nt > 0) {
				et::wait(&g_empty[(nt - 1) & 1], (uint32_t)((nt - 1) >> 1) & 1);
				et::fence_after();
				flush();
			}
			double *p = A.partials + (((size_t)split * A.blocks + mb) * TC_NB + row) * Kp + h * (Kp / 2);
#pragma unroll
			for (int j = 0; j < Kp / 2; ++j) p[j] = acc64[j];
		}
	}
	et::fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// stats[k] = N_k, stats[K + k D + d] = sum gamma x, stats[K + K D + k D + d] = sum gamma x^2 from the centred / scaled moments:
//   S0 = sum gamma, S1 = sum gamma x'', S2 = sum gamma x''^2,  x = x'' / alpha + c
__global__ void em_tc_reduce_kernel(const double *__restrict__ partials, int n_split, int blocks, int Kp, int K, int D, const float *__restrict__ cs,
                                    double *__restrict__ stats) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= K * (D + 1)) return;
	const int k = idx / (D + 1), j = idx % (D + 1);          // j == D: N_k
	const int blk = k / TC_NB, m = k % TC_NB;
	auto col = [&](int c) {
		double s = 0.0;
		for (int sp = 0; sp < n_split; ++sp) s += partials[(((size_t)sp * blocks + blk) * TC_NB + m) * Kp + c];
		return s;
	};
	const double S0 = col(2 * D);                            // column order of gmm_em_tc's own operands: 2 d -> x''^2, 2 d + 1 -> x'', 2 D -> 1
	if (j == D) {
		stats[k] = S0;
		return;
	}
	const double c = (double)cs[j], al = (double)cs[D + j];
	const double S2 = col(2 * j), S1 = col(2 * j + 1);
	stats[K + (size_t)k * D + j] = S1 / al + c * S0;
	stats[K + (size_t)K * D + (size_t)k * D + j] = S2 / (al * al) + 2.0 * c * S1 / al + c * c * S0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------------------
static bool g_et_attr = false;

bool em_tc_eligible(int K, int D, int n) {
	int path = ctx().em_path;
	if (path < 0) {
		const char *e = getenv("SRB200_EM_TC");
		path = (e && e[0] == '0') ? 0 : -1;
	}
	if (path == 0) return false;
	const int Kp = t2_kp(D);
	if (Kp > T2_KP_MAX || K <= 64) return false;
	return path == 1 || (long long)n * K >= (1LL << 24);      // small jobs stay on the FMA kernels (the tensor route reads a flag back every iteration)
}

// Build the operand images of the model whose float64 master is d_params and run the E-step on the tensor cores.
// *used = 0 and nothing written when the packer judged the parameters unsafe for the uncorrected expanded form.
int launch_estep_tc(PackedModels &pm, const double *d_params, const float *dX, int n, int ld, int D, int K, float *d_lse2, double *d_ll_sum, double *d_stats,
                    int clamp, cudaStream_t st, int *used) {
	Context &c = ctx();
	*used = 0;
	const int Kp = t2_kp(D), blocks = ceil_div(K, TC_NB);
	if (et_smem_bytes(Kp, ld) > 227 * 1024 || (ld & 3) != 0 || (reinterpret_cast<uintptr_t>(dX) & 15) != 0) return 0;      // the frame tile is one 16-byte-granular bulk copy
	(void)pm;
	const int n_split = std::max(1, c.sm_count / blocks);
	EmTcWorkspace &w = g_ws;                                   // ensure() is a no-op once the buffers are large enough
	if (w.cs.ensure((size_t)2 * D * sizeof(float)) || w.img2.ensure((size_t)blocks * t2_block_halfs(Kp) * sizeof(__half)) ||
	    w.img128.ensure((size_t)blocks * 2 * TC_NB * Kp * sizeof(__half)) || w.hard.ensure((size_t)blocks * TC_HSTRIDE * sizeof(int4)) ||
	    w.flag.ensure(sizeof(int)) || w.partials.ensure((size_t)n_split * blocks * TC_NB * Kp * sizeof(double)) || w.h_flag.ensure(sizeof(int)))
		return 1;
	float *d_cs = w.cs.as<float>();
	__half *d_img2 = w.img2.as<__half>(), *d_img128 = w.img128.as<__half>();
	int4 *d_hard = w.hard.as<int4>();
	int *d_flag = w.flag.as<int>();
	double *d_partials = w.partials.as<double>();
	// ---- operand images from the master ----
	SRB_CUDA(cudaMemsetAsync(d_hard, 0, (size_t)blocks * TC_HSTRIDE * sizeof(int4), st));
	em_tc_cs_kernel<<<D, 128, 0, st>>>(d_params, K, D, d_cs, d_flag);
	em_tc_pack_kernel<<<ceil_div(blocks * TC_NB * 32, 256), 256, 0, st>>>(d_params, K, D, Kp, blocks, d_cs, d_img2, d_img128, d_hard, d_flag);
	SRB_CUDA(cudaGetLastError());
	int *h_flag = w.h_flag.as<int>();
	*h_flag = 1;
	SRB_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
	SRB_CUDA(cudaStreamSynchronize(st));
	if (*h_flag != 0) return 0;
	// ---- pass A: the tcgen05 scorer on the device-built images: per-frame log2 p(x) (unclamped) + the total log-likelihood ----
	Tc2Images im;
	im.B = d_img2;
	im.hard = d_hard;
	im.cs = d_cs;
	im.Kp = Kp;
	im.S = 1;
	im.cols_per_spk = blocks * TC_NB;
	im.blocks_total = blocks;
	const int utt_off[2] = {0, n};
	if (launch_score_tc_images(im, dX, ld, D, utt_off, 1, d_ll_sum, d_lse2, n, nullptr, clamp, 1, st)) return 1;
	// ---- pass B ----
	if (!g_et_attr) {
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_tc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_tc_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		SRB_CUDA(cudaFuncSetAttribute(em_accumulate_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		g_et_attr = true;
	}
	EtArgs a;
	a.X = dX;
	a.ld = ld;
	a.D = D;
	a.n = n;
	a.Kp = Kp;
	a.cs = d_cs;
	a.img128 = d_img128;
	a.lse2 = d_lse2;
	a.partials = d_partials;
	a.blocks = blocks;
	a.n_split = n_split;
	a.clamp = clamp;
	const dim3 grid(blocks, n_split);
	const size_t smem = et_smem_bytes(Kp, ld);
	switch (Kp / 16) {
		case 2: em_accumulate_tc_kernel<2><<<grid, ET_THREADS, smem, st>>>(a); break;
		case 3: em_accumulate_tc_kernel<3><<<grid, ET_THREADS, smem, st>>>(a); break;
		case 4: em_accumulate_tc_kernel<4><<<grid, ET_THREADS, smem, st>>>(a); break;
		case 5: em_accumulate_tc_kernel<5><<<grid, ET_THREADS, smem, st>>>(a); break;
		case 6: em_accumulate_tc_kernel<6><<<grid, ET_THREADS, smem, st>>>(a); break;
		case 7: em_accumulate_tc_kernel<7><<<grid, ET_THREADS, smem, st>>>(a); break;
		case 8: em_accumulate_tc_kernel<8><<<grid, ET_THREADS, smem, st>>>(a); break;
		default: return 0;      // Kp = 16 (D <= 7): FMA kernels
	}
	SRB_CUDA(cudaGetLastError());
	em_tc_reduce_kernel<<<ceil_div(K * (D + 1), 256), 256, 0, st>>>(d_partials, n_split, blocks, Kp, K, D, d_cs, d_stats);
	SRB_CUDA(cudaGetLastError());
	if (launch_em_tail(d_ll_sum, d_lse2, n, clamp, d_stats, K, D, st)) return 1;
	*used = 1;
	return 0;
}

}  // namespace srb
