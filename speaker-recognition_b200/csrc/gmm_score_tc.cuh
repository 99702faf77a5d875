// gmm_score_tc.cuh -- tensor-core (tcgen05, kind::tf32) diagonal-GMM scorer for sm_100a: kernel + operand layouts.
//
// Same quantity as gmm_score.cu (log2 p(x) per frame and speaker, per-utterance sums), different arithmetic route:
// the Mahalanobis term is expanded,
//     l2_k(x) = C_k + sum_d [ (-a_kd^2) x_d^2 + (-2 a_kd nb_kd) x_d ],     C_k = c2_k - sum_d nb_kd^2,
// i.e. a GEMM  L[frames x rows] = PHI[frames x Kp] . W[rows x Kp]^T  with PHI = [x^2, x, 1, 0-pad], Kp = 8*ceil((2D+1)/8),
// run on the 5th-generation tensor cores as 3xTF32 (hi*hi + lo*hi + hi*lo; fp32 accumulate in TMEM).
// The expanded form loses accuracy on dimensions with a tiny sigma (the trainer's floor is 0.0316): their three terms are
// ~1e3 each and cancel.  Those (row, dim) entries are therefore left out of W and C and added back EXACTLY in the
// epilogue, t = a x + nb, l2 -= t*t, from a per-block sparse list (about 1 % of the entries in the synthetic configs);
// with that the route measures ~1e-6 per frame against float64 (experiments/tf32_numerics*.py, tests/test_gmm_tc_gpu.py).
//
// Structure (one persistent CTA per SM, 576 threads, warp-specialised; a work unit = one 128-frame tile x a chunk of speakers):
//   warp 0  producer: cp.async.bulk (TMA engine) of the pre-imaged weight blocks into a 3-stage shared-memory ring -- a block is
//           128 mixture rows, its tf32-hi and fp32-remainder-lo images are one stage each -- and of the block's hard-entry
//           list into a 4-stage ring; completion on mbarriers
//   warp 1  MMA issuer: one elected lane issues 3 x Kp/8 tcgen05.mma (M=128, N=128, K=8) per block into one of two
//           128-column TMEM accumulator slots; the A operand (PHI hi / lo) is read from TENSOR MEMORY, only B from shared
//           memory (with both operands in shared memory the operand fetch alone needs the whole 128 B/clk of the SM, ncu);
//           tcgen05.commit -> mbarriers (frees the B stage, publishes the slot)
//   warps 2-17 epilogue, thread = frame row = TMEM lane; the four warps of a lane quarter take the four 32-column groups
//           (sixteen warps, not eight: the per-warp chain ld -> max -> ex2 -> sum is latency-bound, ncu):
//           per unit they build PHI from the raw features (global -> registers -> tcgen05.st, plus a raw copy in shared
//           memory for the add-back); per block they (1) while the block's MMAs still run, sum the hard entries' t^2 per
//           column into a per-warp scratch column in shared memory (dynamic column index -> memory, not registers),
//           (2) tcgen05.ld the accumulator and release the slot, (3) subtract the scratch of the flagged columns,
//           (4) online log2-sum-exp2; at the end of a speaker the quarter's warps merge, applies the clamp, writes the per-frame
//           value and the per-(tile, speaker) partial sum.
// B images are K-major in the no-swizzle canonical layout (8-row x 16-byte core matrices; LBO = byte stride between
// K-adjacent core matrices, SBO = 128 B between 8-row groups; verified by experiments/tc_gemm_test.cu) and are stored in HBM
// in exactly the shared-memory order, so every load is ONE contiguous bulk copy.
#pragma once
#include "common.cuh"
#include "gmm_score_tc_layout.h"

namespace srb {

struct TcArgs {
	const float *X;            // [n][ld] features
	int ld;
	const float *cs;           // [2][D] feature transform of the model set: centre c_d, power-of-two scale alpha_d; x'' = (x - c) alpha
	const float *B;            // [n_blocks_total][2][TC_NB*Kp] weight images (speaker-major, rows padded)
	const int4 *ftiles;        // per frame tile: {first frame, frame count, utterance, 0}
	int n_ftiles;
	int Kp;
	int S;                     // speakers
	int blocks_per_spk;        // 128-row blocks per speaker (rows padded to a multiple of 128) -- or, when seg < 128:
	int seg;                   // columns per speaker inside a block (32 or 64) when a speaker is smaller than a block
	int n_blocks_total;
	int n_chunks;              // the block range is cut into n_chunks contiguous chunks (whole speakers each)
	int spk_per_chunk;
	float *frame_ll;           // optional [S][ldf]
	int ldf;
	double *partial;           // [n_ftiles][S][4]
	int clamp, frame_log2;
	int D;
	int bstages;               // B ring depth (TC_BSTAGES, or 2 when three stages do not fit in shared memory)
	const int4 *hard;          // [n_blocks_total][TC_HSTRIDE]: header 0 = bit masks of the columns with entries in the four
	                           // 32-column groups, header 1 = {first entry of group 1, of group 2, of group 3, total}; entries sorted
	                           // by column, every group padded to a multiple of 4 with null entries: {4 * dim, float bits of a, of nb,
	                           // last entry of its column ? 1 + 128 * (column within the group) : 0}
};

namespace tc {

__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ld32(uint32_t taddr, float *v) {
	uint32_t r[32];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
	      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
	      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
	      "=r"(r[31])
	    : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// bounded wait: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity) {
	long long spins = 0;
	while (!mbar_try_wait(bar, parity)) {
		if (++spins > 400000000LL) __trap();
	}
}

}  // namespace tc

// online log2-sum-exp2 update with 32 new values (one chunk of accumulator columns) for one frame
__device__ __forceinline__ void tc_lse_update(const float *v, float &m_run, float &s_run) {
	float t[16];                                   // tree reduction: depth 5 instead of a 31-long dependent chain
#pragma unroll
	for (int i = 0; i < 16; ++i) t[i] = fmaxf(v[2 * i], v[2 * i + 1]);
#pragma unroll
	for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
		for (int i = 0; i < w; ++i) t[i] = fmaxf(t[i], t[i + w]);
	const float mx = t[0];
	const float m_new = fmaxf(m_run, mx);
	// packed fp32x2 adds (SASS FADD2): the 32 subtractions and the 32 sum updates cost 16 instructions each
	const float nm = -m_new;
	unsigned long long nm2, acc01 = 0ull, acc23 = 0ull;
	asm("mov.b64 %0, {%1, %1};" : "=l"(nm2) : "f"(nm));
#pragma unroll
	for (int i = 0; i < 32; i += 4) {
		unsigned long long p01, p23, d01, d23;
		asm("mov.b64 %0, {%1, %2};" : "=l"(p01) : "f"(v[i]), "f"(v[i + 1]));
		asm("mov.b64 %0, {%1, %2};" : "=l"(p23) : "f"(v[i + 2]), "f"(v[i + 3]));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(p01), "l"(nm2));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(p23), "l"(nm2));
		float a, b, c, d;
		asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(d01));
		asm("mov.b64 {%0, %1}, %2;" : "=f"(c), "=f"(d) : "l"(d23));
		unsigned long long e01, e23;
		asm("mov.b64 %0, {%1, %2};" : "=l"(e01) : "f"(ex2_approx(a)), "f"(ex2_approx(b)));
		asm("mov.b64 %0, {%1, %2};" : "=l"(e23) : "f"(ex2_approx(c)), "f"(ex2_approx(d)));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc01) : "l"(acc01), "l"(e01));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc23) : "l"(acc23), "l"(e23));
	}
	float s0, s1, s2, s3;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(acc01));
	asm("mov.b64 {%0, %1}, %2;" : "=f"(s2), "=f"(s3) : "l"(acc23));
	s_run = fmaf(s_run, ex2_approx(m_run - m_new), (s0 + s1) + (s2 + s3));
	m_run = m_new;
}


namespace tc {
// D[tmem] (+)= A[tmem] . B[smem]^T : the A operand is read from tensor memory (lane = row, column = k)
// Issued by the WHOLE (converged) warp; one lane is elected inside the statement (a per-thread predicate instead of
// elect.sync brings the loop back).  (Issuing from an `if (lane == 0)` region
// makes ptxas wrap every MMA in an elect/branch loop: ~10 dependent instructions and ~100 cycles per MMA, ncu.)
template <int ACC>
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
	asm volatile(
	    "{\n\t.reg .b64 db;\n\t.reg .pred p, pe;\n\t"
	    "mov.b64 db, {%2, %3};\n\t"
	    "setp.ne.b32 p, %5, 0;\n\t"
	    "elect.sync _|pe, 0xffffffff;\n\t"
	    "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
	    "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
	    : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t *bar) {
	asm volatile(
	    "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
	    "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
	    : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&v)[8]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
	             "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
	             : "memory");
}
}  // namespace tc

// KSTEPS = Kp / 8 as a compile-time constant (fully unrolled MMA issue: 3-4 instructions per MMA), or 0 for the generic loop
template <int KSTEPS>
__global__ void __launch_bounds__(TC_THREADS, 1) gmm_score_tc_kernel(const TcArgs A) {
	extern __shared__ __align__(1024) unsigned char tc_smem[];
	const int Kp = A.Kp, D = A.D, XS = tc_x_stride(D);
	const uint32_t b_part_bytes = (uint32_t)TC_NB * Kp * 4;
	const size_t b_part_floats = (size_t)TC_NB * Kp;
	float *sB = reinterpret_cast<float *>(tc_smem);                                  // [stages] one part (hi or lo) each
	const int NST = A.bstages;                                                       // 3, or 2 when Kp > 80 (shared-memory budget)
	int4 *hring = reinterpret_cast<int4 *>(sB + NST * b_part_floats);         // [TC_HSTAGES][TC_HSTRIDE] hard-entry lists
	float *scr = reinterpret_cast<float *>(hring + TC_HSTAGES * TC_HSTRIDE);         // [16 warps][32 columns][32 lanes] add-back sums
	float2 *ms_x = reinterpret_cast<float2 *>(scr + 16 * 32 * 32);                   // [2][3][128] (m, s) handed to the merging warp
	uint64_t *bars = reinterpret_cast<uint64_t *>(ms_x + 2 * 3 * TC_M);
	uint64_t *full_b = bars, *empty_b = full_b + TC_BSTAGES;
	uint64_t *full_h = empty_b + TC_BSTAGES, *empty_h = full_h + TC_HSTAGES;
	uint64_t *tm_full = empty_h + TC_HSTAGES, *tm_empty = tm_full + TC_SLOTS;
	uint64_t *a_ready = tm_empty + TC_SLOTS;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_ready + 1);
	float *sX = reinterpret_cast<float *>(bars + 32);                                // [128][XS] raw features of the unit's tile

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (tid == 0) {
		for (int s = 0; s < TC_BSTAGES; ++s) {
			mbar_init(&full_b[s], 1);
			mbar_init(&empty_b[s], 1);
		}
		for (int s = 0; s < TC_HSTAGES; ++s) {
			mbar_init(&full_h[s], 1);
			mbar_init(&empty_h[s], 16);      // one arrival per epilogue warp
		}
		for (int s = 0; s < TC_SLOTS; ++s) {
			mbar_init(&tm_full[s], 1);
			mbar_init(&tm_empty[s], 16);
		}
		mbar_init(a_ready, 16);
		fence_mbar_init();
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	tc::fence_before();
	__syncthreads();
	tc::fence_after();
	const uint32_t tmem_base = *tmem_slot;

	// work units: (frame tile, chunk of whole speakers); unit u -> tile u / n_chunks, chunk u % n_chunks
	const int n_units = A.n_ftiles * A.n_chunks;
	const int blocks_per_chunk_full = (A.seg >= TC_NB) ? A.spk_per_chunk * A.blocks_per_spk : (A.spk_per_chunk * A.seg + TC_NB - 1) / TC_NB;

	if (warp == 0) {
		// ======================= producer =======================
		if (lane == 0) {
			uint32_t it_h = 0, ph_b = 0;
			int st_b = 0;
			for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
				const int tile = u / A.n_chunks, chunk = u - tile * A.n_chunks;
				const int blk0 = chunk * blocks_per_chunk_full;
				const int blk1 = min(A.n_blocks_total, blk0 + blocks_per_chunk_full);
				for (int b = blk0; b < blk1; ++b, ++it_h) {
					{   // the block's hard-entry list (needed first: the epilogue walks it while the block's MMAs run)
						const int st = it_h % TC_HSTAGES;
						tc::wait(&empty_h[st], ((it_h / TC_HSTAGES) & 1) ^ 1);
						mbar_expect_tx(&full_h[st], TC_HSTRIDE * 16);
						bulk_g2s(hring + (size_t)st * TC_HSTRIDE, A.hard + (size_t)b * TC_HSTRIDE, TC_HSTRIDE * 16, &full_h[st]);
					}
					// the block's hi image, then its lo image: one stage each, contiguous in HBM
					const float *src = A.B + (size_t)b * tc_b_block_floats(Kp);
#pragma unroll 1
					for (int p = 0; p < 2; ++p, src += b_part_floats) {
						tc::wait(&empty_b[st_b], ph_b ^ 1);
						mbar_expect_tx(&full_b[st_b], b_part_bytes);
						bulk_g2s(sB + (size_t)st_b * b_part_floats, src, b_part_bytes, &full_b[st_b]);
						if (++st_b == NST) {
							st_b = 0;
							ph_b ^= 1;
						}
					}
				}
			}
		}
	} else if (warp == 1) {
		// ======================= MMA issuer (whole warp converged; one lane elected per instruction) =======================
		{
			// instruction descriptor: D = F32 (bits 4-5 = 1), A/B = TF32 (bits 7-9 / 10-12 = 2), both K-major, N>>3 at bit 17, M>>4 at bit 24
			const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_NB >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
			constexpr uint32_t lbo_b = (TC_NB / 8) * 128, sbo = 128;
			// the B descriptor is two words; only the 14-bit start-address field of the low word advances between k-steps
			constexpr uint32_t d_hi = (uint32_t)(sbo >> 4) | (1u << 14);            // SBO | version bit 46
			constexpr uint32_t stepB = (2 * lbo_b) >> 4;
			const int ksteps = KSTEPS > 0 ? KSTEPS : Kp / 8;
			const uint32_t a_hi = tmem_base + TC_A_COL, a_lo = a_hi + (uint32_t)Kp;
			uint32_t ph_b = 0, it_a = 0, it_s = 0;
			int st_b = 0;
			for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it_a) {
				const int tile = u / A.n_chunks, chunk = u - tile * A.n_chunks;
				const int blk0 = chunk * blocks_per_chunk_full;
				const int blk1 = min(A.n_blocks_total, blk0 + blocks_per_chunk_full);
				tc::wait(a_ready, it_a & 1);            // PHI of this unit's tile is in tensor memory
				tc::fence_after();
				for (int b = blk0; b < blk1; ++b, ++it_s) {
					const int slot = it_s % TC_SLOTS;
					tc::wait(&tm_empty[slot], ((it_s / TC_SLOTS) & 1) ^ 1);
					tc::fence_after();
					const uint32_t d_tmem = tmem_base + (uint32_t)(slot * TC_NB);
					// 3xTF32: (A_hi + A_lo) * B_hi from the first stage, A_hi * B_lo from the second
					{
						const int st = st_b;
						tc::wait(&full_b[st], ph_b);
						tc::fence_after();
						const uint32_t dB0 = ((smem_u32(sB + (size_t)st * b_part_floats) >> 4) & 0x3FFF) | ((lbo_b >> 4) << 16);
						tc::mma_tf32_ts<0>(d_tmem, a_hi, dB0, d_hi, idesc);     // first MMA of the block overwrites the slot
#pragma unroll
						for (int k = 1; k < ksteps; ++k) tc::mma_tf32_ts<1>(d_tmem, a_hi + 8 * k, dB0 + stepB * k, d_hi, idesc);
#pragma unroll
						for (int k = 0; k < ksteps; ++k) tc::mma_tf32_ts<1>(d_tmem, a_lo + 8 * k, dB0 + stepB * k, d_hi, idesc);
						tc::commit_elect(&empty_b[st]);     // the stage may be refilled once these MMAs have read it
						if (++st_b == NST) {
							st_b = 0;
							ph_b ^= 1;
						}
					}
					{
						const int st = st_b;
						tc::wait(&full_b[st], ph_b);
						tc::fence_after();
						const uint32_t dB0 = ((smem_u32(sB + (size_t)st * b_part_floats) >> 4) & 0x3FFF) | ((lbo_b >> 4) << 16);
#pragma unroll
						for (int k = 0; k < ksteps; ++k) tc::mma_tf32_ts<1>(d_tmem, a_hi + 8 * k, dB0 + stepB * k, d_hi, idesc);
						tc::commit_elect(&empty_b[st]);
						if (++st_b == NST) {
							st_b = 0;
							ph_b ^= 1;
						}
					}
					tc::commit_elect(&tm_full[slot]);       // the accumulator slot is complete
				}
			}
		}
	} else {
		// ======================= epilogue (16 warps; thread = frame row; the 4 warps of a lane quarter take the 4 column groups) ====
		const int ew = warp - 2;                      // 0..15
		const int q = warp & 3;                       // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
		const int cg = ew >> 2;                       // 32-column group of every block this warp owns
		const int row = q * 32 + lane;                // frame within the tile
		const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
		float *my_scr = scr + (size_t)ew * 32 * 32 + lane;
		float *xrow = sX + (size_t)row * XS;
		const unsigned char *xrow_b = reinterpret_cast<const unsigned char *>(xrow);
		unsigned char *scr_b = reinterpret_cast<unsigned char *>(my_scr) - 1;      // entry.w = 1 + byte offset of the scratch column
		const int seg = A.seg;
		const int cols_per_spk = (seg >= TC_NB) ? A.blocks_per_spk * TC_NB : seg;
		const int nch = Kp >> 3;
		uint32_t it_s = 0, it_h = 0, n_merge = 0;
#define SRB_TC_QBAR() asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory")      /* the 4 warps of this lane quarter */
		for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
			const int tile = u / A.n_chunks, chunk = u - tile * A.n_chunks;
			const int blk0 = chunk * blocks_per_chunk_full;
			const int blk1 = min(A.n_blocks_total, blk0 + blocks_per_chunk_full);
			const int4 ft = A.ftiles[tile];
			const bool valid = row < ft.y;

			// ---- PHI = [x^2, x, 1, 0-pad] of this thread's frame -> tensor memory (tf32 hi | fp32 remainder lo); the quarter's four
			// warps split the 8-column groups.  All MMAs of the previous unit have completed (this warp consumed its last
			// tm_full), the quarter barrier makes sure no sibling still reads the previous tile's raw features.
			SRB_TC_QBAR();
			{
				const float *xg = A.X + (size_t)(ft.x + (valid ? row : 0)) * A.ld;
				const float *cen = A.cs, *alp = A.cs + D;
				for (int c = cg; c < nch; c += 4) {
					uint32_t hi[8], lo[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) {
						const int k = 8 * c + j;
						float v = 0.f;
						if (k < D) {
							const float x = valid ? (__ldg(xg + k) - __ldg(cen + k)) * __ldg(alp + k) : 0.f;
							v = x * x;
						} else if (k < 2 * D) {
							v = valid ? (__ldg(xg + k - D) - __ldg(cen + k - D)) * __ldg(alp + k - D) : 0.f;
							xrow[k - D] = v;
						} else if (k == 2 * D) {
							v = valid ? 1.f : 0.f;
						}
						hi[j] = __float_as_uint(v) & ~0x1FFFu;
						lo[j] = __float_as_uint(v - __uint_as_float(hi[j]));
					}
					tc::st8(lane_base + TC_A_COL + 8 * c, hi);
					tc::st8(lane_base + TC_A_COL + Kp + 8 * c, lo);
				}
				asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
				tc::fence_before();
				SRB_TC_QBAR();                                // the quarter's raw features are in shared memory
				if (lane == 0) mbar_arrive(a_ready);
			}

			float m_run = -3.0e38f, s_run = 0.f;
			int col_in_spk = 0;                       // columns of the current speaker consumed so far (by the quarter's warps together)
			int spk = chunk * A.spk_per_chunk;        // first speaker of the current block

			// close a speaker whose columns were shared by `width` (1, 2 or 4) adjacent warps of the quarter: merge, clamp,
			// per-frame output, per-(tile, speaker, lane quarter) partial sum
			auto finish = [&](int my_spk, int width) {
				float m = m_run, sacc = s_run;
				const bool leader = (cg & (width - 1)) == 0;
				if (width > 1) {
					// double-buffered hand-over (buffer = parity of the merge count): one barrier per speaker; the buffer written
					// now is not written again before the NEXT merge's barrier, which every reader of it has passed by then
					float2 *buf = ms_x + (size_t)(n_merge & 1) * 3 * TC_M;
					++n_merge;
					if (!leader) buf[(cg - 1) * TC_M + row] = make_float2(m, sacc);
					SRB_TC_QBAR();
					if (leader) {
						for (int o = 0; o < width - 1; ++o) {
							const float2 other = buf[(cg + o) * TC_M + row];
							const float M = fmaxf(m, other.x);
							sacc = sacc * ex2_approx(m - M) + other.y * ex2_approx(other.x - M);
							m = M;
						}
					}
				}
				if (leader && my_spk < A.S) {
					const float l2 = m + lg2_approx(sacc);
					float ll = l2 * TC_LN2;
					if (A.clamp && !(ll >= -744.4400719213812f)) ll = -34.538776394910684f;
					if (valid && A.frame_ll != nullptr) A.frame_ll[(size_t)my_spk * A.ldf + ft.x + row] = A.frame_log2 ? l2 : ll;
					float local = valid ? ll : 0.f;
#pragma unroll
					for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
					if (lane == 0) A.partial[((size_t)tile * A.S + my_spk) * 4 + q] = (double)local;
				}
				m_run = -3.0e38f;
				s_run = 0.f;
			};

			for (int b = blk0; b < blk1; ++b, ++it_s, ++it_h) {
				const int slot = it_s % TC_SLOTS, hst = it_h % TC_HSTAGES;
				// ---- exact add-back, phase 1 (the block's MMAs are still running): sum t^2 of this warp's hard entries per
				// column -- the list is uniform across the warp, sorted by column, `last` flag on a column's final entry -- and
				// park each column's sum in the warp's scratch row.  The column index is dynamic, hence memory, not registers
				// (the first version switched over 32 register cases per entry: 116 SM sub-partition cycles per entry, ncu).
				tc::wait(&full_h[hst], (it_h / TC_HSTAGES) & 1);
				const int4 *hl = hring + (size_t)hst * TC_HSTRIDE;
				const int4 hmask = hl[0], hoff = hl[1];
				const int e0 = cg == 0 ? 0 : (cg == 1 ? hoff.x : (cg == 2 ? hoff.y : hoff.z));
				const int e1 = cg == 0 ? hoff.x : (cg == 1 ? hoff.y : (cg == 2 ? hoff.z : hoff.w));
				const uint32_t mask = (uint32_t)(cg == 0 ? hmask.x : (cg == 1 ? hmask.y : (cg == 2 ? hmask.z : hmask.w)));
				{
					// four entries per step (the host pads every group's list to a multiple of four with null entries): all loads
					// of a step are issued before its first scratch store -- entries, features and scratch share the shared-memory
					// address space, so the compiler cannot move a load above a store on its own.  Tried and rejected (B200, C2):
					// a dedicated "park" warp per lane quarter doing all four groups one block ahead (8.0 ms vs 6.9: one warp cannot
					// issue 48 dependent entry updates per block in time) and prefetching step i+1 before the stores of step i (7.4 ms).
					const int4 *e = hl + TC_HHEAD + e0;
					const int n = e1 - e0;
					float csum = 0.f;
#pragma unroll 1
					for (int i = 0; i < n; i += 4) {
						const int4 a0 = e[i], a1 = e[i + 1], a2 = e[i + 2], a3 = e[i + 3];
						const float x0 = *reinterpret_cast<const float *>(xrow_b + a0.x), x1 = *reinterpret_cast<const float *>(xrow_b + a1.x);
						const float x2 = *reinterpret_cast<const float *>(xrow_b + a2.x), x3 = *reinterpret_cast<const float *>(xrow_b + a3.x);
						const float t0 = fmaf(x0, __int_as_float(a0.y), __int_as_float(a0.z)), t1 = fmaf(x1, __int_as_float(a1.y), __int_as_float(a1.z));
						const float t2 = fmaf(x2, __int_as_float(a2.y), __int_as_float(a2.z)), t3 = fmaf(x3, __int_as_float(a3.y), __int_as_float(a3.z));
						csum = fmaf(t0, t0, csum);
						if (a0.w) {
							*reinterpret_cast<float *>(scr_b + a0.w) = csum;
							csum = 0.f;
						}
						csum = fmaf(t1, t1, csum);
						if (a1.w) {
							*reinterpret_cast<float *>(scr_b + a1.w) = csum;
							csum = 0.f;
						}
						csum = fmaf(t2, t2, csum);
						if (a2.w) {
							*reinterpret_cast<float *>(scr_b + a2.w) = csum;
							csum = 0.f;
						}
						csum = fmaf(t3, t3, csum);
						if (a3.w) {
							*reinterpret_cast<float *>(scr_b + a3.w) = csum;
							csum = 0.f;
						}
					}
				}
				__syncwarp();
				if (lane == 0) mbar_arrive(&empty_h[hst]);
				// ---- accumulator -> registers, slot released right away
				tc::wait(&tm_full[slot], (it_s / TC_SLOTS) & 1);
				tc::fence_after();
				float v[32];
				tc::ld32(lane_base + (uint32_t)(slot * TC_NB + cg * 32), v);
				tc::fence_before();
				__syncwarp();
				if (lane == 0) mbar_arrive(&tm_empty[slot]);
				// ---- add-back, phase 2: only the flagged columns have a (fresh) scratch value.  (16 % of the kernel's instructions;
				// the alternative -- scratch as float4 [8][32] kept zero outside the flagged columns, eight unconditional 128-bit
				// loads + re-zeroing stores -- was slower, 7.5 vs 6.8 ms: its shared-memory traffic competes with the MMA's B fetch.)
				if (mask) {
#pragma unroll
					for (int j = 0; j < 32; ++j)
						if (mask & (1u << j)) v[j] -= my_scr[j * 32];
				}
				// column segmentation: seg >= 128 -> the block belongs to one speaker (quarter-wide merge at its end); seg == 64 ->
				// two speakers per block, two warps each; seg == 32 -> every warp's group is its own speaker
				tc_lse_update(v, m_run, s_run);
				if (seg == 32) {
					finish(spk + cg, 1);
					spk += 4;
				} else if (seg == 64) {
					finish(spk + (cg >> 1), 2);
					spk += 2;
				} else {
					col_in_spk += TC_NB;
					if (col_in_spk == cols_per_spk) {
						finish(spk, 4);
						col_in_spk = 0;
						++spk;
					}
				}
			}
		}
#undef SRB_TC_QBAR
	}
	tc::fence_before();
	__syncthreads();
	if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

}  // namespace srb
