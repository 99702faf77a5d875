// gmm_em_tc.cu -- the EM E-step of large models on the tensor cores (sm_100a, tcgen05).
//
// gmm_em.cu recomputes the exponents and accumulates the sufficient statistics on the fp32 FMA pipe (0.43 of its peak: 7.9 ms
// per iteration at BASELINE config 3, K = 512, N = 1e6, D = 39).  Both halves of that work are GEMMs:
//   pass A   l2[t][k] = PHI[t] . W[k]      the expanded-square scores of gmm_score_tc2.cuh  -> per-frame log2 p(x_t), total LL
//   pass B   gamma[t][k] = 2^(l2[t][k] - lse2[t]);   S[k][:] = sum_t gamma[t][k] * [x''_t^2, x''_t, 1]      (GEMM over the FRAMES)
// Here the fp16 head/tail operand images of the CURRENT parameters are built on the device from the float64 master
// (em_tc_pack_kernel: no host round trip inside an EM iteration), pass A is the tcgen05 scorer reading those images, and pass B is
// em_accumulate_tc_kernel: per (128-mixture block, frame range) a CTA alternates
//     GEMM 1  L[128 frames x 128 mixtures] = PHI (tensor memory) . W^T (shared memory, resident for the CTA's lifetime)
//     epilogue  gamma = 2^(L - lse2), split into fp16 head + tail, written MN-major into shared memory (a thread owns a frame row)
//     GEMM 2  ACC[128 mixtures x Kp] += gamma^T . Y,  Y = [x''^2, x'', 1] (fp16 head + tail, MN-major)   -- 3 passes each
// with ACC in tensor memory, flushed into float64 partials every 1024 frames; em_tc_reduce_kernel sums the partials and maps
// the centred / scaled moments back to sum gamma x, sum gamma x^2.  Exactly corrected entries (gmm_score_tc2's add-back) are NOT
// used here: the packer computes max_k sum_d nb''^2 and raises a flag when the expanded form would not be accurate enough
// (collapsed mixtures far from the centre) -- the caller then runs the FMA kernels of gmm_em.cu for that iteration.
#include <cuda_fp16.h>
#include <algorithm>
#include "gmm_score_tc_layout.h"
#include "kernels.h"

namespace srb {

constexpr double ET_SAFE_NB2 = TC_HARD_BUDGET;      // max over rows of sum nb''^2 the uncorrected form is trusted with: the scorer's own budget
constexpr int ET_THREADS = 288;             // warp 0: loads + MMA issue; warps 1-8: epilogue (TMEM lane quarter = warp % 4, column half = (warp - 1) / 4)
constexpr int ET_FLUSH_TILES = 8;           // tiles (of 128 frames) between float64 flushes of the accumulator
constexpr int ET_L_COL = 0, ET_ACC_COL = 128, ET_A_COL = 256;      // TMEM columns: GEMM-1 accumulator | GEMM-2 accumulator (Kp) | PHI head, tail (Kp)

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
	const size_t o128 = ((size_t)(k128 >> 3) * (TC_NB / 8) + (size_t)(col >> 3)) * 64 + (col & 7) * 8 + (k128 & 7);
	const size_t b1 = (size_t)blk * 2 * TC_NB * Kp + o128;
	img128[b1] = h;
	img128[b1 + (size_t)TC_NB * Kp] = l;
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
	const __half *img128;      // [blocks][head | tail][128 * Kp]
	const float *lse2;         // [n] log2 p(x_t) from pass A
	double *partials;          // [n_split][blocks][128][Kp]
	int blocks, n_split, clamp, nY;
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
// D[tmem] (+)= A[tmem] . B[smem]^T, single CTA, whole converged warp, one lane elected
template <int ACC>
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
	asm volatile(
	    "{\n\t.reg .b64 db;\n\t.reg .pred p, pe;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\telect.sync _|pe, 0xffffffff;\n\t"
	    "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
	    "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
	    : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], both operands through shared-memory descriptors; `acc` is a run-time predicate
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
	asm volatile(
	    "{\n\t.reg .b64 da, db;\n\t.reg .pred p, pe;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\telect.sync _|pe, 0xffffffff;\n\t"
	    "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
	    "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
	    : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
	asm volatile(
	    "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
	    "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
	    : "memory");
}
__device__ __forceinline__ void split_pack(float v0, float v1, uint32_t &hi, uint32_t &lo) {      // v0 in the low half
	const __half2 h = __floats2half2_rn(v0, v1);
	const float2 hf = __half22float2(h);
	const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
	hi = *reinterpret_cast<const uint32_t *>(&h);
	lo = *reinterpret_cast<const uint32_t *>(&l);
}
}  // namespace et

// shared memory: W image (head | tail) + nY Y tiles (head | tail, MN-major) + gamma tile (head | tail, MN-major) + raw frame tile
// [128][ld] (+32 bytes: the last chunk of the last row may read up to two 16-byte units past it) + lse2[2][128] + barriers / TMEM slot
static size_t et_smem_bytes(int Kp, int ld, int nY) {
	return (size_t)2 * TC_NB * Kp * 2 + (size_t)nY * 2 * TC_M * Kp * 2 + (size_t)2 * TC_M * TC_NB * 2 + ((size_t)TC_M * ld * 4 + 32) + 2 * TC_M * 4 + (size_t)Kp * 4 + 256;
}
// Y double-buffered where it fits: the next tile's PHI / Y are then built while GEMM 2 of the current tile runs
static int et_ybufs(int Kp, int ld) {
	for (int nY = 2; nY >= 1; --nY)
		if (et_smem_bytes(Kp, ld, nY) <= 227 * 1024) return nY;
	return 0;
}

// Column order of PHI / W / Y inside this file (any order works as long as the operands agree): k = 2 d -> x''_d^2, k = 2 d + 1 -> x''_d,
// k = 2 D -> 1, rest 0 -- so that four dimensions (one 16-byte read of the frame row) make one 16-byte piece of the operand row.
template <int NCH>      // Kp / 16
__global__ void __launch_bounds__(ET_THREADS, 1) em_accumulate_tc_kernel(const EtArgs A) {
	extern __shared__ __align__(1024) unsigned char et_smem[];
	constexpr int Kp = NCH * 16;
	const int D = A.D, ld = A.ld, nY = A.nY;
	const uint32_t w_part = (uint32_t)TC_NB * Kp * 2, y_part = (uint32_t)TC_M * Kp * 2, g_part = (uint32_t)TC_M * TC_NB * 2;      // bytes of one image part
	unsigned char *sW = et_smem;                         // [head | tail] 128 rows x Kp, K-major canonical (as stored in HBM)
	unsigned char *sY = sW + 2 * w_part;                 // nY x [head | tail] Y[frame][col], MN-major canonical: (f >> 3) * (Kp / 8) * 128 + (c >> 3) * 128 + (f & 7) * 16 + (c & 7) * 2
	unsigned char *sG = sY + (size_t)nY * 2 * y_part;    // [head | tail] gamma[frame][mixture], MN-major canonical: (f >> 3) * 2048 + (m >> 3) * 128 + (f & 7) * 16 + (m & 7) * 2
	float *sX = reinterpret_cast<float *>(sG + 2 * g_part);        // [128][ld] the tile's frames as they lie in HBM (one bulk copy)
	float *sL = sX + (size_t)TC_M * ld + 8;              // [2][128] lse2 of the tile's frames (+inf: the frame contributes nothing), by tile parity
	float *sC = sL + 2 * TC_M;                           // [2][Kp / 2] centre, scale per dimension (zero-padded)
	uint64_t *bars = reinterpret_cast<uint64_t *>(sC + Kp);
	uint64_t *bar_w = bars, *bar_a = bars + 1, *bar_g1 = bars + 2, *bar_g = bars + 3, *bar_g2 = bars + 4, *bar_x = bars + 5;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int mb = blockIdx.x, split = blockIdx.y;
	for (int i = tid; i < Kp; i += ET_THREADS) {
		const int d = i % (Kp / 2);
		sC[i] = d < D ? A.cs[(i / (Kp / 2)) * D + d] : 0.f;
	}
	if (tid == 0) {
		mbar_init(bar_w, 1);
		mbar_init(bar_a, 8);
		mbar_init(bar_g1, 1);
		mbar_init(bar_g, 8);
		mbar_init(bar_g2, 1);
		mbar_init(bar_x, 1);
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

	const int n_tiles = (A.n + TC_M - 1) / TC_M;
	const int per = (n_tiles + A.n_split - 1) / A.n_split;
	const int t0 = split * per, t1 = min(n_tiles, t0 + per);
	double *my_part = A.partials + ((size_t)split * A.blocks + mb) * TC_NB * Kp;

	if (warp == 0) {
		// ======================= loads + MMA issue =======================
		auto load_x = [&](int t) {      // the tile's frames: rows are contiguous in HBM ([n][ld]), so the tile is ONE bulk copy
			if (lane == 0) {
				const uint32_t bytes = (uint32_t)min(TC_M, A.n - t * TC_M) * (uint32_t)ld * 4u;
				mbar_expect_tx(bar_x, bytes);
				bulk_g2s(sX, A.X + (size_t)t * TC_M * ld, bytes, bar_x);
			}
			__syncwarp();
		};
		if (lane == 0) {
			mbar_expect_tx(bar_w, 2 * w_part);
			bulk_g2s(sW, A.img128 + (size_t)mb * 2 * TC_NB * Kp, 2 * w_part, bar_w);
		}
		__syncwarp();
		if (t0 < t1) load_x(t0);
		const uint32_t idesc1 = (1u << 4) | ((uint32_t)(TC_NB >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);                               // K-major A (TMEM) and B
		const uint32_t idesc2 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(Kp >> 3) << 17) | ((uint32_t)(TC_NB >> 4) << 24);      // MN-major A and B, N = Kp
		constexpr uint32_t lbo_w = (TC_NB / 8) * 128;
		const uint32_t dW_hi = (uint32_t)(128 >> 4) | (1u << 14);                                                                        // SBO 128 | version
		const uint32_t dW0 = ((smem_u32(sW) >> 4) & 0x3FFF) | ((lbo_w >> 4) << 16);
		constexpr uint32_t stepW = (2 * lbo_w) >> 4;
		// MN-major tiles: K-groups (8 frames) are (MN / 8) * 128 bytes apart = the descriptor's LBO, MN-adjacent core matrices 128 bytes
		// = its SBO (measured: experiments/tc_mn_major_test.cu, profiles/r2_mn_major_probe.txt)
		const uint32_t kg_G = (TC_NB / 8) * 128, kg_Y = (uint32_t)(Kp / 8) * 128;
		const uint32_t dG_hi = (uint32_t)(128 >> 4) | (1u << 14), dY_hi = dG_hi;
		const uint32_t dG0 = ((smem_u32(sG) >> 4) & 0x3FFF) | ((kg_G >> 4) << 16);
		const uint32_t stepG = (2 * kg_G) >> 4, stepY = (2 * kg_Y) >> 4;
		constexpr int ks1 = NCH;
		const uint32_t a_hi = tmem_base + ET_A_COL, a_lo = a_hi + (uint32_t)(Kp >> 1);
		et::wait(bar_w, 0);
		uint32_t it = 0;
		for (int t = t0; t < t1; ++t, ++it) {
			et::wait(bar_a, it & 1);                     // PHI in tensor memory, Y in shared memory; the raw frame tile has been consumed
			et::fence_after();
			if (t + 1 < t1) load_x(t + 1);
			{
				const uint32_t d1 = tmem_base + ET_L_COL;
				et::mma_ts<0>(d1, a_hi, dW0, dW_hi, idesc1);
#pragma unroll
				for (int k = 1; k < ks1; ++k) et::mma_ts<1>(d1, a_hi + 8 * k, dW0 + stepW * k, dW_hi, idesc1);
				for (int k = 0; k < ks1; ++k) et::mma_ts<1>(d1, a_lo + 8 * k, dW0 + stepW * k, dW_hi, idesc1);
				for (int k = 0; k < ks1; ++k) et::mma_ts<1>(d1, a_hi + 8 * k, dW0 + (w_part >> 4) + stepW * k, dW_hi, idesc1);
				et::commit(bar_g1);
			}
			et::wait(bar_g, it & 1);                     // gamma tile in shared memory (and a due flush has read ACC)
			et::fence_after();
			{
				const uint32_t d2 = tmem_base + ET_ACC_COL;
				const uint32_t dY0 = ((smem_u32(sY + (size_t)(it % nY) * 2 * y_part) >> 4) & 0x3FFF) | ((kg_Y >> 4) << 16);
				uint32_t acc = (it % ET_FLUSH_TILES) == 0 ? 0u : 1u;      // first GEMM 2 after a flush overwrites the accumulator
				for (int k = 0; k < TC_M / 16; ++k) {        // gamma head x Y head, gamma tail x Y head, gamma head x Y tail
					et::mma_ss(d2, dG0 + stepG * k, dG_hi, dY0 + stepY * k, dY_hi, idesc2, acc);
					acc = 1u;
					et::mma_ss(d2, dG0 + (g_part >> 4) + stepG * k, dG_hi, dY0 + stepY * k, dY_hi, idesc2, 1u);
					et::mma_ss(d2, dG0 + stepG * k, dG_hi, dY0 + (y_part >> 4) + stepY * k, dY_hi, idesc2, 1u);
				}
				et::commit(bar_g2);
			}
		}
	} else {
		// ======================= epilogue: thread = frame row (PHI / Y / gamma) or mixture row (flush) =======================
		const int q = warp & 3, half = (warp - 1) >> 2;
		const int row = q * 32 + lane;
		const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
		const float4 *cen4 = reinterpret_cast<const float4 *>(sC), *alp4 = reinterpret_cast<const float4 *>(sC + Kp / 2);
		constexpr int nch = NCH;
		// float64 running sums of this thread's part of the statistics: mixture `row` of the block, its half of the Kp columns.  The
		// tensor-memory accumulator is fp32 and is folded in here every ET_FLUSH_TILES tiles (1024 frames)
		double acc64[Kp / 2];
#pragma unroll
		for (int j = 0; j < Kp / 2; ++j) acc64[j] = 0.0;
		auto flush = [&]() {
#pragma unroll
			for (int c = 0; c < Kp / 2; c += 8) {
				uint32_t r[8];
				et::ld8(lane_base + ET_ACC_COL + half * (Kp / 2) + c, r);
				et::ld_wait();
#pragma unroll
				for (int j = 0; j < 8; ++j) acc64[c + j] += (double)__uint_as_float(r[j]);
			}
		};
		// PHI (tensor memory) and Y (shared memory) of this thread's frame; the two warps of a lane quarter split the 16-value chunks
		auto build = [&](int t, uint32_t itb) {
			et::wait(bar_x, itb & 1);
			const int fr = t * TC_M + row;
			const bool valid = fr < A.n;
			const float4 *xr = reinterpret_cast<const float4 *>(sX + (size_t)row * ld);
			unsigned char *ybuf = sY + (size_t)(itb % nY) * 2 * y_part;
#pragma unroll
			for (int c = half; c < nch; c += 2) {
				uint32_t hi[8], lo[8];
#pragma unroll
				for (int u = 0; u < 2; ++u) {
					const float4 xv = valid ? xr[2 * c + u] : make_float4(0.f, 0.f, 0.f, 0.f);
					const float4 cv = cen4[2 * c + u], av = alp4[2 * c + u];
					const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, cs4[4] = {cv.x, cv.y, cv.z, cv.w}, as4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const int d = 8 * c + 4 * u + j;
						float v0 = 0.f, v1 = 0.f;                  // (x''^2, x'') | (1, 0) at d == D | (0, 0) beyond
						if (d < D) {
							v1 = fminf(fmaxf((xs[j] - cs4[j]) * as4[j], -T2_X_CLAMP), T2_X_CLAMP);
							v0 = v1 * v1;
						} else if (d == D) {
							v0 = 1.f;
						}
						if (!valid) v0 = 0.f, v1 = 0.f;
						et::split_pack(v0, v1, hi[4 * u + j], lo[4 * u + j]);
					}
				}
				et::st8(lane_base + ET_A_COL + 8 * c, hi);
				et::st8(lane_base + ET_A_COL + (Kp >> 1) + 8 * c, lo);
				// the same 16 values as two 16-byte pieces of the Y row (columns 16 c .. 16 c + 15)
				unsigned char *yh = ybuf + (size_t)(row >> 3) * (Kp / 8) * 128 + (size_t)(2 * c) * 128 + (row & 7) * 16;
				*reinterpret_cast<uint4 *>(yh) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
				*reinterpret_cast<uint4 *>(yh + 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
				*reinterpret_cast<uint4 *>(yh + y_part) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
				*reinterpret_cast<uint4 *>(yh + y_part + 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
			}
			if (half == 0) {
				float l2 = __int_as_float(0x7f800000);       // +inf -> gamma = 2^(-inf) = 0
				if (valid) {
					l2 = __ldg(A.lse2 + fr);
					if (A.clamp && !(l2 * TC_LN2 >= -744.4400719213812f)) l2 = __int_as_float(0x7f800000);      // gmm.cc:482-490: underflowed frame
				}
				sL[(itb & 1) * TC_M + row] = l2;
			}
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the Y stores become visible to the tensor core
			et::fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_a);
		};
		uint32_t it = 0;
		if (t0 < t1) build(t0, 0);
		for (int t = t0; t < t1; ++t, ++it) {
			// ---- gamma = 2^(L - lse2) for this thread's frame and its 64 mixtures, fp16 head + tail, MN-major.  GEMM 1 of this tile
			// complete implies GEMM 2 of the previous one complete (the tensor pipe is in order): the gamma tile, the accumulator and
			// the other Y buffer are free
			et::wait(bar_g1, it & 1);
			et::fence_after();
			if (it > 0 && (it % ET_FLUSH_TILES) == 0) flush();
			const float lse = sL[(it & 1) * TC_M + row];
#pragma unroll 1
			for (int cc = 0; cc < 2; ++cc) {
				uint32_t v[32];
				et::ld32(lane_base + ET_L_COL + 64 * half + 32 * cc, v);
				et::ld_wait();
				unsigned char *gh = sG + (size_t)(row >> 3) * 2048 + (size_t)(8 * half + 4 * cc) * 128 + (row & 7) * 16;
#pragma unroll
				for (int g8 = 0; g8 < 4; ++g8) {             // 8 mixtures = one 16-byte piece
					uint32_t hi[4], lo[4];
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const float g0 = ex2_approx(__uint_as_float(v[8 * g8 + 2 * j]) - lse), g1 = ex2_approx(__uint_as_float(v[8 * g8 + 2 * j + 1]) - lse);
						et::split_pack(g0, g1, hi[j], lo[j]);
					}
					*reinterpret_cast<uint4 *>(gh + g8 * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
					*reinterpret_cast<uint4 *>(gh + g_part + g8 * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
				}
			}
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			et::fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_g);
			// ---- the next tile's operands, while GEMM 2 of this one runs (two Y buffers) or after it (one)
			if (t + 1 < t1) {
				if (nY == 1) {
					et::wait(bar_g2, it & 1);
					et::fence_after();
				}
				build(t + 1, it + 1);
			}
		}
		if (t1 > t0) {
			et::wait(bar_g2, (it - 1) & 1);
			et::fence_after();
			flush();
		}
		double *p = my_part + (size_t)row * Kp + half * (Kp / 2);
#pragma unroll
		for (int j = 0; j < Kp / 2; ++j) p[j] = acc64[j];
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
	const int nY = et_ybufs(Kp, ld);
	if (nY == 0 || (ld & 3) != 0 || (reinterpret_cast<uintptr_t>(dX) & 15) != 0) return 0;      // the frame tile is one 16-byte-granular bulk copy
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
	a.nY = nY;
	const dim3 grid(blocks, n_split);
	const size_t smem = et_smem_bytes(Kp, ld, nY);
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
